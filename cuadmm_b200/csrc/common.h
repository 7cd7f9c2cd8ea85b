// common.h — error plumbing and small RAII helpers shared by the whole library.
// Runtime-wrapper equivalent of the reference's include/cuadmm/{check,memory,utils}.h,
// but with 64-bit sizes everywhere and errors that propagate instead of being printed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include "../../include/cuadmm_b200.h"

namespace cuadmm {

void set_last_error(const std::string& msg);

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUADMM_CUDA(call)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            char buf__[512];                                                                \
            snprintf(buf__, sizeof buf__, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), \
                     __FILE__, __LINE__, cudaGetErrorString(e__));                          \
            throw ::cuadmm::Error(e__ == cudaErrorMemoryAllocation ? CUADMM_ENOMEM :        \
                                  (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? CUADMM_ENODEVICE : CUADMM_ECUDA, buf__); \
        }                                                                                   \
    } while (0)

#define CUADMM_REQUIRE(cond, msg)                                                           \
    do { if (!(cond)) throw ::cuadmm::Error(CUADMM_EINVAL, std::string("invalid argument: ") + (msg)); } while (0)

// Wraps a C-ABI body: exceptions -> status code + last-error string.
template <class F>
static inline int guarded(F&& f) {
    try { f(); return CUADMM_OK; }
    catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::bad_alloc&) { set_last_error("host out of memory"); return CUADMM_ENOMEM; }
    catch (const std::exception& e) { set_last_error(e.what()); return CUADMM_EINVAL; }
}

// Device buffer with 64-bit size.
template <class T>
struct DevBuf {
    T* p = nullptr;
    int64_t n = 0;
    bool owned = true;      // false: a view of memory owned elsewhere (the peer arena)
    DevBuf() {}
    explicit DevBuf(int64_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = true; }
    void adopt(T* ptr, int64_t n_) { release(); p = ptr; n = n_; owned = false; }
    void alloc(int64_t n_) {
        release();
        n = n_;
        if (n_ > 0) CUADMM_CUDA(cudaMalloc((void**)&p, sizeof(T) * (size_t)n_));
    }
    void zero(cudaStream_t s = 0) { if (n) CUADMM_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, s)); }
    void upload(const T* h, int64_t cnt, cudaStream_t s = 0) {
        if (cnt) CUADMM_CUDA(cudaMemcpyAsync(p, h, sizeof(T) * (size_t)cnt, cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& h, cudaStream_t s = 0) {
        if ((int64_t)h.size() != n) alloc((int64_t)h.size());
        upload(h.data(), n, s);
    }
    void download(T* h, int64_t cnt, cudaStream_t s = 0) const {
        if (cnt) CUADMM_CUDA(cudaMemcpyAsync(h, p, sizeof(T) * (size_t)cnt, cudaMemcpyDeviceToHost, s));
    }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (dev >= 0) { cudaGetDevice(&prev); CUADMM_CUDA(cudaSetDevice(dev)); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static inline int64_t tri(int64_t n) { return n * (n + 1) / 2; }

}  // namespace cuadmm
