// Chain microbenchmark: cycles per column of the Viterbi diagonal solve, one warp per SMSP, data in shared memory.
//   variant 0: lane = column, one shuffle per row (the round-1 solver step), pushes d = 0..ND
//   variant 1: micro-blocks of 4 columns solved redundantly in every lane, pushes d = 0..ND
// Both produce the same tables (checked).  Diagnostics only.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

constexpr int BX = 32;
#ifndef ND
#define ND 2
#endif
constexpr int BANDCOLS = (ND + 1) * BX;
constexpr unsigned kFull = 0xffffffffu;

// band[e][cc] : row e of the block, column cc of the band; the diagonal tile is cc in [ND*32, ND*32+32)
template <int VARIANT>
__global__ void __launch_bounds__(128, 1) chain_kernel(const float *band_g, const float *unary_g, const float *eta_g,
                                                       float *outq, int *outsel, long long *cyc, int nblk) {
    extern __shared__ float smem[];
    float *band = smem;  // [4 warps][BX][BANDCOLS]
    const int warp = threadIdx.x >> 5, c = threadIdx.x & 31;
    float *myband = band + (size_t)warp * BX * BANDCOLS;
    for (int i = c; i < BX * BANDCOLS; i += 32) myband[i] = band_g[i] + 0.001f * warp;
    __syncthreads();
    float best[ND + 1];
    int bsel[ND + 1];
#pragma unroll
    for (int d = 0; d <= ND; ++d) {
        best[d] = -INFINITY;
        bsel[d] = -1;
    }
    float qtop = 0.f;
    long long t0 = clock64();
    for (int j = nblk - 1; j >= 0; --j) {
        const int x0 = j * BX;
        const float s_d = unary_g[(j & 7) * BX + c];
        const float s_eta = eta_g[(j & 7) * BX + c];
        const float dr = s_d * (s_d > 0.f ? 1.f : 0.f);
        if (j < nblk - 1) {
            const float xk = (c == BX - 1) ? qtop + s_eta : -INFINITY;
            bsel[0] = (xk >= best[0]) ? -1 : bsel[0];
            best[0] = fmaxf(best[0], xk);
        } else if (c == BX - 1) {
            best[0] = -0.0f;
            bsel[0] = -1;
        }
        const float *colp = myband + c;
        if (VARIANT == 0) {
#pragma unroll
            for (int e = BX - 1; e >= 0; --e) {
                const int y = x0 + e;
                const float *rowp = colp + e * BANDCOLS;
                float sv[ND + 1];
#pragma unroll
                for (int d = 0; d <= ND; ++d) sv[d] = rowp[(ND - d) * BX];
                const bool below = c < e;
                const float qb = __shfl_sync(kFull, best[0] + dr, e);
                if (e == 0) qtop = qb;
                {
                    const float xi = below ? qb + sv[0] : -INFINITY;
                    const float xk = (c == e - 1) ? qb + s_eta : -INFINITY;
                    const bool tk = xi >= best[0];
                    const float b1 = fmaxf(best[0], xi);
                    bsel[0] = tk ? y : bsel[0];
                    bsel[0] = (xk >= b1) ? -1 : bsel[0];
                    best[0] = fmaxf(b1, xk);
                }
#pragma unroll
                for (int d = 1; d <= ND; ++d) {
                    const float xi = qb + sv[d];
                    const bool tk = xi >= best[d];
                    bsel[d] = tk ? y : bsel[d];
                    best[d] = fmaxf(best[d], xi);
                }
            }
        } else {
            // micro-blocks of 4 columns, descending
#pragma unroll
            for (int k = BX / 4 - 1; k >= 0; --k) {
                // every lane gathers the partials of the 4 columns and their unary terms
                float P[4], U[4], E[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    P[i] = __shfl_sync(kFull, best[0], 4 * k + i);
                    U[i] = __shfl_sync(kFull, dr, 4 * k + i);
                    E[i] = __shfl_sync(kFull, s_eta, 4 * k + i);
                }
                // micro-triangle S[r][i], r > i  (broadcast loads)
                const float *mt = myband + (4 * k) * BANDCOLS + ND * BX + 4 * k;
                const float s32 = mt[3 * BANDCOLS + 2], s31 = mt[3 * BANDCOLS + 1], s30 = mt[3 * BANDCOLS + 0];
                const float s21 = mt[2 * BANDCOLS + 1], s20 = mt[2 * BANDCOLS + 0], s10 = mt[1 * BANDCOLS + 0];
                float q[4];
                q[3] = P[3] + U[3];
                q[2] = fmaxf(fmaxf(P[2], q[3] + s32), q[3] + E[2]) + U[2];
                q[1] = fmaxf(fmaxf(fmaxf(P[1], q[3] + s31), q[2] + s21), q[2] + E[1]) + U[1];
                q[0] = fmaxf(fmaxf(fmaxf(fmaxf(P[0], q[3] + s30), q[2] + s20), q[1] + s10), q[1] + E[0]) + U[0];
                if (k == 0) qtop = q[0];
                // push the 4 rows into my column of every tile
#pragma unroll
                for (int r = 3; r >= 0; --r) {
                    const int e = 4 * k + r, y = x0 + e;
                    const float *rowp = colp + e * BANDCOLS;
                    const float qb = q[r];
                    {
                        const float xi = (c < e) ? qb + rowp[ND * BX] : -INFINITY;
                        const float xk = (c == e - 1) ? qb + s_eta : -INFINITY;
                        const bool tk = xi >= best[0];
                        const float b1 = fmaxf(best[0], xi);
                        bsel[0] = tk ? y : bsel[0];
                        bsel[0] = (xk >= b1) ? -1 : bsel[0];
                        best[0] = fmaxf(b1, xk);
                    }
#pragma unroll
                    for (int d = 1; d <= ND; ++d) {
                        const float xi = qb + rowp[(ND - d) * BX];
                        const bool tk = xi >= best[d];
                        bsel[d] = tk ? y : bsel[d];
                        best[d] = fmaxf(best[d], xi);
                    }
                }
            }
        }
        outq[((size_t)warp * nblk + j) * BX + c] = best[0] + dr;
        outsel[((size_t)warp * nblk + j) * BX + c] = bsel[0];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            best[d] = best[d + 1];
            bsel[d] = bsel[d + 1];
        }
        best[ND] = -INFINITY;
        bsel[ND] = -1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    const int nblk = 64;
    const size_t nband = (size_t)BX * BANDCOLS;
    float *hb = (float *)malloc(nband * 4), *hu = (float *)malloc(8 * BX * 4), *he = (float *)malloc(8 * BX * 4);
    srand(1);
    for (size_t i = 0; i < nband; ++i) hb[i] = (rand() % 2001 - 1000) * 1e-3f;
    for (int i = 0; i < 8 * BX; ++i) {
        hu[i] = (rand() % 2001 - 1000) * 1e-3f;
        he[i] = (rand() % 2001 - 1000) * 1e-3f;
    }
    float *db, *du, *de, *dq[2];
    int *ds[2];
    long long *dc;
    cudaMalloc(&db, nband * 4);
    cudaMalloc(&du, 8 * BX * 4);
    cudaMalloc(&de, 8 * BX * 4);
    cudaMalloc(&dc, 8);
    cudaMemcpy(db, hb, nband * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(du, hu, 8 * BX * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(de, he, 8 * BX * 4, cudaMemcpyHostToDevice);
    const size_t nout = (size_t)4 * nblk * BX;
    for (int v = 0; v < 2; ++v) {
        cudaMalloc(&dq[v], nout * 4);
        cudaMalloc(&ds[v], nout * 4);
    }
    const size_t smem = 4 * nband * 4;
    cudaFuncSetAttribute(chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 3; ++rep) {
            if (v == 0) chain_kernel<0><<<1, 128, smem>>>(db, du, de, dq[v], ds[v], dc, nblk);
            else chain_kernel<1><<<1, 128, smem>>>(db, du, de, dq[v], ds[v], dc, nblk);
            cudaDeviceSynchronize();
        }
        long long h;
        cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("variant %d ND=%d: %.1f cycles per column (%lld cycles, %d columns)  err=%s\n", v, ND,
               (double)h / (nblk * BX), h, nblk * BX, cudaGetErrorString(cudaGetLastError()));
    }
    float *q0 = (float *)malloc(nout * 4), *q1 = (float *)malloc(nout * 4);
    int *s0 = (int *)malloc(nout * 4), *s1 = (int *)malloc(nout * 4);
    cudaMemcpy(q0, dq[0], nout * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(q1, dq[1], nout * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s0, ds[0], nout * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s1, ds[1], nout * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < nout; ++i) bad += (q0[i] != q1[i]) || (s0[i] != s1[i]);
    printf("mismatches between variants: %zu of %zu\n", bad, nout);
    return 0;
}
