// common.cpp — last-error slot, version string, device probe.
#include "common.h"
#include <mutex>

namespace cuadmm {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace cuadmm

extern "C" {
const char* cuadmm_last_error(void) { return cuadmm::g_last_error.c_str(); }
const char* cuadmm_version(void) { return "cuadmm_b200 0.1 (sm_100a)"; }
int cuadmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
}
