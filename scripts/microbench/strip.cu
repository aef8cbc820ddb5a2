// Streaming microbenchmark for the far field: how fast can an SM consume the score triangle when every CTA reads
// whole rows (32 columns x all tracks, contiguous) instead of one 32-byte sector per cell?
// thread <-> (column, 4 tracks); 4 rows per stage; cp.async FIFO of NS stages; Viterbi + log-sum updates with a
// constant q row (no mailbox).  Diagnostics only.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

constexpr int BX = 32, NQD = 22, NTH = BX * NQD;  // 704 consumer threads
#ifndef NS
#define NS 3
#endif
#ifndef DO_MATH
#define DO_MATH 3
#endif
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void cp16(unsigned s, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }

// score [T][T][N], N = 88.  CTA b owns units u = b, b + grid, ...; unit u -> (column block J, row class k of K)
__global__ void __launch_bounds__(NTH, 1) strip_kernel(const float *score, int T, int N, int K, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned fifo = (unsigned)__cvta_generic_to_shared(smem);  // [NS][4 rows][NTH] float4
    const int t = threadIdx.x;
    const int col = t / NQD, quad = t - col * NQD;
    const int nb = T / BX;
    float vmax[4], lM[4], lS[4];
    int vsel[4];
    float accsum = 0.f;
    const int nunits = nb * K;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int J = nb - 1 - u / K, k = u % K;
        const int x0 = J * BX;
        const int R = T - x0;  // rows y = T-1 .. x0 (bench: the whole strip)
        const int nq4 = (R + 3) / 4;
        const int mine = nq4 > k ? (nq4 - k + K - 1) / K : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { vmax[i] = -INFINITY; vsel[i] = -1; lM[i] = -FLT_MAX; lS[i] = 0.f; }
        const float *base = score + (size_t)(x0 + col) * N + quad * 4;
        auto issue = [&](int i) {
            if (i < mine) {
                const int y0 = T - 1 - 4 * (k + i * K);
                const unsigned dst = fifo + (unsigned)(((i % NS) * 4) * NTH + t) * 16u;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int y = y0 - r;
                    if (y >= x0) cp16(dst + (unsigned)(r * NTH) * 16u, base + (size_t)y * T * N);
                }
            }
            commit();
        };
        for (int i = 0; i < NS - 1; ++i) issue(i);
        for (int i = 0; i < mine; ++i) {
            issue(i + NS - 1);
            waitg<NS - 1>();
            const int y0 = T - 1 - 4 * (k + i * K);
            const unsigned src = fifo + (unsigned)(((i % NS) * 4) * NTH + t) * 16u;
            float x[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 s = lds128(src + (unsigned)(r * NTH) * 16u);
                const float sv[4] = {s.x, s.y, s.z, s.w};
                const float q = 0.25f * r;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    if (DO_MATH & 1) {
                        const float xv = q + sv[i4];
                        const bool tk = xv >= vmax[i4];
                        vmax[i4] = tk ? xv : vmax[i4];
                        vsel[i4] = tk ? y0 - r : vsel[i4];
                    }
                    if (DO_MATH & 2) x[r][i4] = fmaf(sv[i4], kLog2e, q);
                    if (DO_MATH == 0) accsum += sv[i4];
                }
            }
            if (DO_MATH & 2) {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float m = fmaxf(fmaxf(x[0][i4], x[1][i4]), fmaxf(x[2][i4], x[3][i4]));
                    const float Mn = fmaxf(lM[i4], m);
                    float acc = lS[i4] * ex2f(lM[i4] - Mn);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc += ex2f(x[r][i4] - Mn);
                    lS[i4] = acc;
                    lM[i4] = Mn;
                }
            }
        }
        waitg<0>();
#pragma unroll
        for (int i = 0; i < 4; ++i) accsum += vmax[i] + vsel[i] + lM[i] + lS[i];
    }
    if (accsum == 1234.5f) out[t] = accsum;
}

int main(int argc, char **argv) {
    const int T = 2048, N = 88;
    const int K = argc > 1 ? atoi(argv[1]) : 8;
    const int grid = argc > 2 ? atoi(argv[2]) : 148;
    float *score, *out;
    const size_t n = (size_t)T * T * N;
    cudaMalloc(&score, n * 4);
    cudaMalloc(&out, 4096);
    cudaMemset(score, 0, n * 4);
    const size_t smem = (size_t)NS * 4 * NTH * 16;
    cudaFuncSetAttribute(strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        strip_kernel<<<grid, NTH, smem>>>(score, T, N, K, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double bytes = 4.0 * N * (double)T * (T + BX) / 2;  // strips include the diagonal blocks
    printf("NS=%d DO_MATH=%d K=%d grid=%d: %.1f us, %.0f GB/s (%.1f GB/s per SM)  smem %zu  err=%s\n", NS, DO_MATH, K, grid,
           best * 1e3, bytes / best / 1e6, bytes / best / 1e6 / grid, smem, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
