// TMA microbenchmarks for the sweep redesign (round 2).  Diagnostics only.
//
//  A. stream: how fast can an SM pull boxes {W tracks, C columns, R rows} of the track-innermost score tensor
//     [T][T][N] through TMA tensor copies (cp.async.bulk.tensor.3d, UTMALDG), for the far-field traversal of one
//     track group?  W = 8 is the round-1 helper's 32-byte sector gather, W = 88 the all-track contiguous read.
//  B. chain: cycles per column of the round-1 chain step (lane = column, shuffle per row, ND = 2, four tracks per CTA
//     interleaved in the band) while the band ring is (0) static, (1) refilled by four LDGSTS gather warps (round 1),
//     (2) refilled by ONE thread with a TMA box {4 tracks, 96 columns, 32 rows}.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_gather tma_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp16(unsigned s, const void *g, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_arrive_noinc(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// A. streaming
// ------------------------------------------------------------------------------------------------------------
struct StreamCfg {
    int T, W, C, R, G, ctas_per_group, near;  // near: rows closer than this to the column block are skipped
    int stages, box_bytes, consume;
};

// one producer thread (thread 0 of warp 0) + consumer warps 1..; ring of `stages` boxes
__global__ void __launch_bounds__(288, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, const StreamCfg cfg, float *out, unsigned long long *bytes_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned base = smem_u32(smem);
    const unsigned full = base, empty = base + 64;  // 8 stages max
    const unsigned ring = base + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncons = blockDim.x / 32 - 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.stages; ++s) {
            mbar_init(full + s * 8, 1);
            mbar_init(empty + s * 8, ncons);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int g = blockIdx.x / cfg.ctas_per_group, h = blockIdx.x % cfg.ctas_per_group;
    if (g >= cfg.G) return;
    const int nbc = cfg.T / cfg.C;
    const unsigned stride = (unsigned)((cfg.box_bytes + 1023) & ~1023);
    if (warp == 0) {
        if (lane == 0) {
            unsigned long long total = 0;
            int it = 0;
            for (int J = nbc - 1; J >= 0; --J) {
                const int ylo = J * cfg.C + cfg.near;  // first far row
                const int nrow = (cfg.T - ylo) / cfg.R;
                for (int i = h; i < nrow; i += cfg.ctas_per_group, ++it) {
                    const int s = it % cfg.stages;
                    if (it >= cfg.stages) mbar_wait(empty + s * 8, ((it / cfg.stages) - 1) & 1);
                    mbar_expect_tx(full + s * 8, (unsigned)cfg.box_bytes);
                    tma_load_3d(ring + s * stride, &map, g * cfg.W, J * cfg.C, cfg.T - (i + 1) * cfg.R, full + s * 8);
                    total += cfg.box_bytes;
                }
            }
            atomicAdd(bytes_out, total);
        }
        return;
    }
    float acc = 0.f;
    int it = 0;
    for (int J = nbc - 1; J >= 0; --J) {
        const int ylo = J * cfg.C + cfg.near;
        const int nrow = (cfg.T - ylo) / cfg.R;
        for (int i = h; i < nrow; i += cfg.ctas_per_group, ++it) {
            const int s = it % cfg.stages;
            mbar_wait(full + s * 8, (it / cfg.stages) & 1);
            if (cfg.consume) {
                for (int o = (threadIdx.x - 32) * 16; o < cfg.box_bytes; o += (blockDim.x - 32) * 16) {
                    const float4 v = lds128(ring + s * stride + o);
                    acc += v.x + v.y + v.z + v.w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s * 8);
        }
    }
    if (acc == 1234.5f) out[threadIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// B. chain with a live band ring
// ------------------------------------------------------------------------------------------------------------
constexpr int BX = 32, ND = 2, NQ = 4, BANDCOLS = (ND + 1) * BX, NBAND = 4;
constexpr int kBandBytes = BX * BANDCOLS * NQ * 4;  // 49152

// warps 0..7: chain (track = warp & 3; two chains per track as in the fused launch); warps 8..11: LDGSTS loaders;
// warp 12: TMA producer.  LOADER 0 = static band, 1 = LDGSTS gather, 2 = TMA.
template <int LOADER>
__global__ void __launch_bounds__(416, 1)
chain_kernel(const __grid_constant__ CUtensorMap map, const float *score, int T, int N, long long *cyc, float *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned base = smem_u32(smem);
    const unsigned full = base, empty = base + 64;
    const unsigned band_s = base + 1024;
    const float *bands = reinterpret_cast<const float *>(smem + 1024);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = T / BX;
    const int n0 = blockIdx.x * NQ;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAND; ++s) {
            mbar_init(full + s * 8, LOADER == 1 ? 128 : 1);
            mbar_init(empty + s * 8, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < NBAND * kBandBytes / 4; i += blockDim.x)
        reinterpret_cast<float *>(smem + 1024)[i] = 0.001f * (i % 977);
    __syncthreads();
    if (warp >= 8 && warp < 12) {
        if (LOADER != 1) return;
        const int lt = threadIdx.x - 256;
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int slot = it % NBAND;
            if (it >= NBAND) mbar_wait(empty + slot * 8, ((it / NBAND) - 1) & 1);
            const int y0 = j * BX, xlo = (j - ND) * BX;
            const unsigned dst0 = band_s + slot * kBandBytes;
            for (int i = lt; i < BX * BANDCOLS; i += 128) {
                const int e = i / BANDCOLS, cc = i - e * BANDCOLS;
                const int y = y0 + e, x = xlo + cc;
                if (x < 0 || x > y || y >= T) continue;
                cp16(dst0 + i * 16, score + ((size_t)y * T + x) * N + n0, 16);
            }
            cp_arrive_noinc(full + slot * 8);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }
    if (warp == 12) {
        if (LOADER != 2 || lane != 0) return;
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int slot = it % NBAND;
            if (it >= NBAND) mbar_wait(empty + slot * 8, ((it / NBAND) - 1) & 1);
            mbar_expect_tx(full + slot * 8, kBandBytes);
            tma_load_3d(band_s + slot * kBandBytes, &map, n0, (j - ND) * BX, j * BX, full + slot * 8);
        }
        return;
    }
    // chain warps
    const int tr = warp & 3, c = lane;
    float best[ND + 1];
    int bsel[ND + 1];
#pragma unroll
    for (int d = 0; d <= ND; ++d) {
        best[d] = -INFINITY;
        bsel[d] = -1;
    }
    float qtop = 0.f;
    const long long t0 = clock64();
    for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
        const int slot = (LOADER == 0) ? 0 : it % NBAND;
        const int x0 = j * BX;
        const float s_d = 0.01f * ((j * 7 + c) % 13) - 0.05f, s_eta = 0.01f * ((j * 5 + c) % 11) - 0.04f;
        const float dr = s_d * (s_d > 0.f ? 1.f : 0.f);
        if (j < nb - 1) {
            const float xk = (c == BX - 1) ? qtop + s_eta : -INFINITY;
            bsel[0] = (xk >= best[0]) ? -1 : bsel[0];
            best[0] = fmaxf(best[0], xk);
        } else if (c == BX - 1) {
            best[0] = -0.0f;
            bsel[0] = -1;
        }
        if (LOADER != 0) mbar_wait(full + slot * 8, (it / NBAND) & 1);
        const float *colp = bands + (size_t)slot * (kBandBytes / 4) + c * NQ + tr;
#pragma unroll
        for (int e = BX - 1; e >= 0; --e) {
            const int y = x0 + e;
            const float *rowp = colp + (size_t)e * (BANDCOLS * NQ);
            float sv[ND + 1];
#pragma unroll
            for (int d = 0; d <= ND; ++d) sv[d] = rowp[(ND - d) * BX * NQ];
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
        __syncwarp();
        if (LOADER != 0 && lane == 0) mbar_arrive(empty + slot * 8);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            best[d] = best[d + 1];
            bsel[d] = bsel[d + 1];
        }
        best[ND] = -INFINITY;
        bsel[ND] = -1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (best[0] == 1234.5f) out[threadIdx.x] = best[0] + bsel[0];
}

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
        printf("no cuTensorMapEncodeTiled\n");
        exit(1);
    }
    return (EncodeFn)fn;
}
static CUtensorMap make_map(EncodeFn enc, float *score, int T, int N, int W, int C, int R, CUtensorMapL2promotion l2) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)T, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)T * N * 4};
    cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)C, (cuuint32_t)R};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, score, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d (W=%d C=%d R=%d)\n", (int)r, W, C, R);
        exit(1);
    }
    return m;
}

int main(int argc, char **argv) {
    const int T = 2048, N = 88;
    float *score, *out;
    unsigned long long *dbytes;
    long long *dcyc;
    const size_t n = (size_t)T * T * N;
    CK(cudaMalloc(&score, n * 4));
    CK(cudaMalloc(&out, 1 << 16));
    CK(cudaMalloc(&dbytes, 8));
    CK(cudaMalloc(&dcyc, 8 * 256));
    CK(cudaMemset(score, 0, n * 4));
    EncodeFn enc = get_encode();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d\n", sms);

    // ---- A. streaming ----
    struct Row {
        int W, C, R, stages, grid, consume;
        CUtensorMapL2promotion l2;
    };
    const CUtensorMapL2promotion L0 = CU_TENSOR_MAP_L2_PROMOTION_NONE, L128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 L256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const Row rows[] = {
        {8, 32, 2, 8, 121, 1, L0},   {8, 32, 8, 6, 121, 1, L0},    {8, 32, 32, 4, 121, 1, L0},  {8, 32, 8, 6, 121, 1, L128},
        {8, 32, 8, 6, 143, 1, L0},   {4, 96, 32, 4, 22, 1, L0},    {4, 32, 32, 6, 132, 1, L0},  {16, 32, 8, 6, 126, 1, L0},
        {32, 32, 4, 6, 126, 1, L0},  {32, 32, 8, 4, 126, 1, L0},   {88, 4, 8, 8, 126, 1, L0},   {88, 4, 32, 4, 126, 1, L0},
        {88, 16, 8, 4, 126, 1, L0},  {88, 32, 4, 4, 126, 1, L0},   {88, 32, 4, 4, 148, 1, L0},  {88, 32, 4, 4, 148, 1, L256},
        {88, 32, 4, 4, 148, 0, L0},  {88, 8, 8, 6, 126, 1, L0},
    };
    for (const Row &r : rows) {
        StreamCfg cfg;
        cfg.T = T;
        cfg.W = r.W;
        cfg.C = r.C;
        cfg.R = r.R;
        cfg.G = (N + r.W - 1) / r.W;
        cfg.ctas_per_group = r.grid / cfg.G;
        cfg.near = r.C;  // skip the diagonal block itself
        cfg.stages = r.stages;
        cfg.box_bytes = r.W * r.C * r.R * 4;
        cfg.consume = r.consume;
        const CUtensorMap map = make_map(enc, score, T, N, r.W, r.C, r.R, r.l2);
        const size_t smem = 1024 + (size_t)r.stages * ((cfg.box_bytes + 1023) & ~1023);
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e9f;
        unsigned long long hb = 0;
        const int grid = cfg.ctas_per_group * cfg.G;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(dbytes, 0, 8));
            CK(cudaEventRecord(e0));
            stream_kernel<<<grid, 288, smem>>>(map, cfg, out, dbytes);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
            CK(cudaMemcpy(&hb, dbytes, 8, cudaMemcpyDeviceToHost));
        }
        printf("stream box{%2d trk,%2d col,%2d row}=%6d B stages %d grid %3d l2promo %d consume %d: %7.1f us  %6.0f GB/s  %5.1f GB/s/SM  (%.0f MB)\n",
               r.W, r.C, r.R, cfg.box_bytes, r.stages, grid, (int)r.l2, r.consume, best * 1e3, hb / best / 1e6,
               hb / best / 1e6 / grid, hb / 1e6);
        CK(cudaGetLastError());
    }

    // ---- B. chain ----
    {
        const CUtensorMap map = make_map(enc, score, T, N, NQ, BANDCOLS, BX, L0);
        const size_t smem = 1024 + (size_t)NBAND * kBandBytes;
        CK(cudaFuncSetAttribute(chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // background stream (all-track boxes on 121 CTAs) to load the memory system, on a second stream
        cudaStream_t sb, sc;
        CK(cudaStreamCreate(&sb));
        CK(cudaStreamCreate(&sc));
        StreamCfg bg;
        bg.T = T; bg.W = 88; bg.C = 32; bg.R = 4; bg.G = 1; bg.ctas_per_group = 121; bg.near = 32; bg.stages = 4;
        bg.box_bytes = 88 * 32 * 4 * 4; bg.consume = 1;
        const CUtensorMap bgmap = make_map(enc, score, T, N, 88, 32, 4, L0);
        const size_t bgsmem = 1024 + (size_t)bg.stages * ((bg.box_bytes + 1023) & ~1023);
        for (int withbg = 0; withbg < 2; ++withbg)
            for (int loader = 0; loader < 3; ++loader) {
                long long h[22];
                double avg = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    if (withbg) {
                        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bgsmem));
                        for (int k = 0; k < 3; ++k) stream_kernel<<<121, 288, bgsmem, sb>>>(bgmap, bg, out, dbytes);
                    }
                    if (loader == 0) chain_kernel<0><<<22, 416, smem, sc>>>(map, score, T, N, dcyc, out);
                    if (loader == 1) chain_kernel<1><<<22, 416, smem, sc>>>(map, score, T, N, dcyc, out);
                    if (loader == 2) chain_kernel<2><<<22, 416, smem, sc>>>(map, score, T, N, dcyc, out);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost));
                    avg = 0;
                    for (int i = 0; i < 22; ++i) avg += (double)h[i] / 22;
                }
                printf("chain loader %d (0 static, 1 LDGSTS gather, 2 TMA box) background %d: %.1f cycles per column\n", loader,
                       withbg, avg / T);
            }
    }
    return 0;
}
