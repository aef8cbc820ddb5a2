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

// Host -> device copy of the part of score[T][T][N] the semi-CRF reads: rows are taken in chunks of `rows_per_chunk`,
// chunk [e0, e1) is one strided 2-D copy of its first e1 cells per row (a superset of b <= e).  Halves the bytes
// of the dense upload; what lies above the copied staircase is left untouched in the destination.
extern "C" int tkb_upload_lower_triangle(const float *host_score, float *dev_score, int T, int N, int rows_per_chunk,
                                         void *stream_) {
    if (!host_score || !dev_score || T < 1 || N < 1 || rows_per_chunk < 1) {
        tkb::set_error("tkb_upload_lower_triangle: invalid argument (T=%d N=%d rows_per_chunk=%d)", T, N, rows_per_chunk);
        return TKB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t pitch = (size_t)T * N * sizeof(float);
    for (int e0 = 0; e0 < T; e0 += rows_per_chunk) {
        const int e1 = e0 + rows_per_chunk < T ? e0 + rows_per_chunk : T;
        TKB_CUDA(cudaMemcpy2DAsync(dev_score + (size_t)e0 * T * N, pitch, host_score + (size_t)e0 * T * N, pitch,
                                   (size_t)e1 * N * sizeof(float), (size_t)(e1 - e0), cudaMemcpyHostToDevice, stream));
    }
    return 0;
}
