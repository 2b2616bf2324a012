// Error channel, launch accounting and device probe of the C ABI (include/aero_b200.h).
#include "common.cuh"
#include <atomic>

namespace aero {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return AERO_ERR_LAUNCH;
    }
    return AERO_OK;
}
}  // namespace aero

extern "C" {
int aero_abi_version(void) { return 3; }
const char* aero_last_error(void) { return aero::g_err; }
uint64_t aero_launch_count(void) { return aero::g_launches.load(); }
int aero_device_arch(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return AERO_ERR_NO_DEVICE; }
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess) { cudaGetLastError(); return AERO_ERR_NO_DEVICE; }
    return pr.major * 10 + pr.minor;
}
}
