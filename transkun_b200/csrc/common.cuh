// common.cuh -- shared helpers for libtranskun_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/transkun_b200.h"

namespace tkb {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);  // records the error, returns (int)e

// experimental strip design of the sweep (semicrf_sweep_strip.cu), selected with TKB_SWEEP=strip
namespace strip {
size_t workspace_bytes(int T, int N);
int sweep(const float *score, long long pitch, const float *noise, int T, int N, int direction, int flags,
          void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit, float *out_lse, void *stream);
void set_timeline(unsigned long long *buf);
}  // namespace strip

#define TKB_CUDA(call)                                          \
    do {                                                        \
        cudaError_t _e = (call);                                \
        if (_e != cudaSuccess) return ::tkb::cuda_fail(_e, #call); \
    } while (0)

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x * (x > 0): what the reference's bool-mask multiply evaluates to for finite x
__device__ __forceinline__ float relu_mask(float x) { return x * (x > 0.0f ? 1.0f : 0.0f); }
// F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_ref(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// ---- cp.async (LDGSTS) -----------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
// variants taking a precomputed 32-bit shared-window address (no generic->shared conversion per call)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_s(unsigned saddr, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ unsigned long long lds64(unsigned saddr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts32(unsigned saddr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- gpu-scope relaxed 64-bit mailbox words --------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

}  // namespace tkb
