// api.cu -- error reporting and device checks of the C ABI (include/transkun_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace tkb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s failed: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-free error state
    return (int)e;
}

}  // namespace tkb

extern "C" int tkb_version(void) { return TKB_VERSION; }

extern "C" const char *tkb_last_error(void) { return tkb::g_err; }

extern "C" int tkb_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        tkb::cuda_fail(e, "cudaGetDevice");
        return TKB_ENODEV;
    }
    int major = 0, coop = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (major != 10 || !coop) {
        tkb::set_error("device %d is sm_%d0 (cooperative launch %d); this library is built for sm_100a only", dev,
                       major, coop);
        return TKB_ENODEV;
    }
    return 0;
}
