// dist.cpp — see dist.h.
#include "dist.h"
#include <dlfcn.h>
#include <string.h>

namespace cuadmm {

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_getid)(ncclUniqueId_t*);
typedef int (*fn_init)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);
struct Api {
    void* lib = nullptr;
    fn_getid getid = nullptr; fn_init init = nullptr; fn_allreduce allreduce = nullptr; fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
};
Api& api() {
    static Api a;
    if (!a.lib) {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
        if (!a.lib) throw Error(CUADMM_ENCCL, std::string("cannot load libnccl: ") + dlerror());
        a.getid = (fn_getid)dlsym(a.lib, "ncclGetUniqueId");
        a.init = (fn_init)dlsym(a.lib, "ncclCommInitRank");
        a.allreduce = (fn_allreduce)dlsym(a.lib, "ncclAllReduce");
        a.destroy = (fn_destroy)dlsym(a.lib, "ncclCommDestroy");
        a.errstr = (fn_errstr)dlsym(a.lib, "ncclGetErrorString");
        if (!a.getid || !a.init || !a.allreduce || !a.destroy) throw Error(CUADMM_ENCCL, "libnccl misses required symbols");
    }
    return a;
}
void check(int rc, const char* what) {
    if (rc != 0) {
        Api& a = api();
        throw Error(CUADMM_ENCCL, std::string("NCCL error in ") + what + ": " + (a.errstr ? a.errstr(rc) : "?"));
    }
}
}  // namespace

void nccl_unique_id(char out[128]) {
    ncclUniqueId_t id;
    check(api().getid(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}

void NcclComm::init(int rank_, int world_, const char id[128], int device) {
    rank = rank_; world = world_;
    CUADMM_CUDA(cudaSetDevice(device));
    ncclUniqueId_t uid;
    memcpy(uid.internal, id, 128);
    check(api().init(&comm, world, uid, rank), "ncclCommInitRank");
}

void NcclComm::allreduce_sum(const double* send, double* recv, int64_t count, cudaStream_t stream) {
    // ncclFloat64 = 8, ncclSum = 0
    check(api().allreduce(send, recv, (size_t)count, 8, 0, comm, stream), "ncclAllReduce");
}

NcclComm::~NcclComm() {
    if (comm) api().destroy(comm);
}

}  // namespace cuadmm

extern "C" int cuadmm_nccl_unique_id(char out[128]) {
    return cuadmm::guarded([&] { CUADMM_REQUIRE(out, "null argument"); cuadmm::nccl_unique_id(out); });
}
