// tcgen05 / TMA microbenchmarks for the interval scorer (round 2).  Diagnostics only.
//
//  A. mma:    cycles per tcgen05.mma.cta_group::1.kind::tf32 (M128, K8, both operands from 128B-swizzled shared memory)
//             as a function of N in {32, 64, 128, 256}, one and two CTAs per SM issuing at the same time.
//  B. ingest: how many bytes per second can every SM pull through TMA 2-D tile copies ({32 floats, R rows} boxes,
//             SWIZZLE_128B) out of an L2-resident operand, 2 CTAs per SM, 4 boxes in flight per CTA?
//  C. scatter: writing [e][b][8 tracks] cells (32-byte pieces, 352 bytes apart) of a track-innermost [T][T][88] tensor:
//             256-bit stores by 8 warps (one cell per lane) against TMA tensor stores of {8 tracks, B begins, 8 ends}
//             boxes staged in shared memory.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                                 \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                          \
        }                                                                                     \
    } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr) {
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- A
template <int N>
__global__ void __launch_bounds__(128, 2) mma_rate(int rounds, long long *cycles) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (smem_u32(raw) + 1023u) & ~1023u;   // A: 128 rows x 128 B, B: up to 256 rows x 128 B
    __shared__ unsigned tmem_s;
    __shared__ unsigned long long bar_s;
    const unsigned bar = smem_u32(&bar_s);
    for (int i = threadIdx.x; i < (128 + 256) * 32; i += blockDim.x) reinterpret_cast<float *>(raw + (base - smem_u32(raw)))[i] = 0.5f;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_s;
    if (threadIdx.x == 0) {
        const unsigned idesc = umma_idesc_tf32(128, N);
        const unsigned long long da = umma_desc_sw128(base), db = umma_desc_sw128(base + 128 * 128);
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_tf32(tmem, da + 2 * kk, db + 2 * kk, idesc, 1u);
            if ((r & 15) == 15) {   // bound the queue depth: wait for completion every 64 MMAs
                umma_commit(bar);
                mbar_wait(bar, (unsigned)((r >> 4) & 1));
            }
        }
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int N>
static void run_mma(int ctas_per_sm, int sms) {
    const int rounds = 1024;
    long long *cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 2));
    const size_t smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = sms * ctas_per_sm;
    mma_rate<N><<<grid, 128, smem>>>(rounds, cyc);
    CK(cudaDeviceSynchronize());
    mma_rate<N><<<grid, 128, smem>>>(rounds, cyc);
    CK(cudaDeviceSynchronize());
    long long h[2 * 256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double per = mean / (rounds * 4.0);
    printf("mma tf32 M128 N%-3d K8, %d CTA/SM: %.1f cycles per MMA per CTA -> %.1f cycles of tensor pipe per MMA, %.0f flop/clk/SM\n", N,
           ctas_per_sm, per, per / ctas_per_sm, 2.0 * 128 * N * 8 / (per / ctas_per_sm));
    CK(cudaFree(cyc));
}

// ---- B
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
template <int ROWS, int DEPTH>
__global__ void __launch_bounds__(64, 2) ingest(const __grid_constant__ CUtensorMap map, int total_rows, int iters) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ unsigned long long bars[DEPTH];
    if (threadIdx.x == 0) {
        for (int i = 0; i < DEPTH; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nboxes = total_rows / ROWS;
        int box = (int)((blockIdx.x * 7919u) % (unsigned)nboxes);
        for (int it = 0; it < iters + DEPTH; ++it) {
            const int slot = it % DEPTH;
            if (it >= DEPTH) mbar_wait(smem_u32(&bars[slot]), (unsigned)(((it / DEPTH) - 1) & 1));
            if (it < iters) {
                mbar_arrive_expect_tx(smem_u32(&bars[slot]), ROWS * 128);
                tma_load_2d(base + slot * ROWS * 128, &map, (it & 7) * 32, box * ROWS, smem_u32(&bars[slot]));
                if ((it & 7) == 7) box = (box + 1) % nboxes;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int DEPTH>
static void run_ingest(EncodeTiledFn enc, float *buf, int rows, int sms, int ctas_per_sm) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {256, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {32, ROWS};
    const cuuint32_t es[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        exit(1);
    }
    const size_t smem = (size_t)DEPTH * ROWS * 128 + 1024;
    CK(cudaFuncSetAttribute(ingest<ROWS, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096 * 16 / ROWS * 4;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    ingest<ROWS, DEPTH><<<grid, 64, smem>>>(map, rows, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    ingest<ROWS, DEPTH><<<grid, 64, smem>>>(map, rows, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)grid * iters * ROWS * 128;
    printf("ingest box {32 fl, %3d rows} depth %d, %d CTA/SM, operand %4d MB: %.2f TB/s total, %.1f GB/s per SM\n", ROWS, DEPTH,
           ctas_per_sm, (int)((size_t)rows * 1024 >> 20), bytes / ms / 1e9, bytes / ms / 1e6 / sms);
}

// ---- C
template <int BEGINS>
__global__ void __launch_bounds__(256, 1) scatter_tma(const __grid_constant__ CUtensorMap map, int T, int iters) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (smem_u32(raw) + 127u) & ~127u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kBoxBytes = 8 * BEGINS * 8 * 4;
    constexpr int kIssuers = 128 / BEGINS * 8;   // BEGINS = 32: every warp issues its own boxes; 128: two warps per 16 ends
    for (int i = threadIdx.x; i < 65536 / 4; i += 256) reinterpret_cast<float *>(raw + (base - smem_u32(raw)))[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (lane == 0 && warp < kIssuers) {
        unsigned seed = blockIdx.x * 9781u + warp * 131u + 7u;
        for (int it = 0; it < iters; ++it) {
            seed = seed * 1664525u + 1013904223u;
            const int e0 = (int)((seed >> 8) % (unsigned)(T / 8)) * 8;
            const int b0 = (int)((seed >> 3) % (unsigned)(T / BEGINS)) * BEGINS;
            const int n0 = (int)(seed % 11u) * 8;
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map), "r"(n0), "r"(b0),
                         "r"(e0), "r"(base + (unsigned)(warp % (65536 / kBoxBytes)) * kBoxBytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void __launch_bounds__(256, 1) scatter_lsu(float *out, int T, int iters) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned seed = blockIdx.x * 9781u + warp * 131u + 7u;
    for (int it = 0; it < iters; ++it) {
        seed = seed * 1664525u + 1013904223u;
        const int e0 = (int)((seed >> 8) % (unsigned)(T / 8)) * 8;
        const int b0 = (int)((seed >> 3) % (unsigned)(T / 32)) * 32;
        const int n0 = (int)(seed % 11u) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float *o = out + ((size_t)(e0 + j) * T + b0 + lane) * 88 + n0;
            const float v = (float)it;
            asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(o), "f"(v) : "memory");
        }
    }
}

// structured like the scorer's epilogue: work item = (tile of 128 begins x 64 ends, group of TRACKS tracks), group fastest,
// items dealt round-robin to the SMs: the SMs complete whole cells / lines at about the same time
template <int TRACKS>
__global__ void __launch_bounds__(256, 1) scatter_tiles(float *out, int T, int items) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int G = 88 / TRACKS;          // groups per cell (88 tracks)
    constexpr int LPC = TRACKS / 8;         // lanes per cell
    const int q = warp & 3, h = warp >> 2;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int tile = w / G, g = w % G;
        const int b0 = (tile % (T / 128)) * 128, e0 = (tile / (T / 128)) * 64 % T;
        for (int j = 0; j < 32 * LPC; ++j) {   // a warp covers 32 begins x 32 ends; with LPC lanes per cell it takes LPC x more instructions
            const int cell = (j * 32 + lane) / LPC, part = (j * 32 + lane) % LPC;
            const int e = e0 + 32 * h + cell / 32, b = b0 + 32 * q + cell % 32;
            float *o = out + ((size_t)e * T + b) * 88 + g * TRACKS + part * 8;
            const float v = (float)w;
            asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(o), "f"(v) : "memory");
        }
    }
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    run_mma<32>(1, sms);
    run_mma<64>(1, sms);
    run_mma<128>(1, sms);
    run_mma<256>(1, sms);
    run_mma<32>(2, sms);
    run_mma<64>(2, sms);
    run_mma<128>(2, sms);

    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)f;
    for (int mb : {48, 1024}) {   // L2-resident and DRAM-resident operands
        const int rows = mb * 1024;
        float *buf;
        CK(cudaMalloc(&buf, (size_t)rows * 1024));
        CK(cudaMemset(buf, 0, (size_t)rows * 1024));
        run_ingest<32, 4>(enc, buf, rows, sms, 2);
        run_ingest<64, 4>(enc, buf, rows, sms, 2);
        run_ingest<128, 4>(enc, buf, rows, sms, 2);
        run_ingest<128, 6>(enc, buf, rows, sms, 2);
        run_ingest<256, 3>(enc, buf, rows, sms, 2);
        run_ingest<128, 4>(enc, buf, rows, sms, 1);
        CK(cudaFree(buf));
    }
    {
        const int T = 2048;
        float *out;
        CK(cudaMalloc(&out, (size_t)T * T * 88 * 4));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float ms;
        const int iters = 2048;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            scatter_lsu<<<sms, 256>>>(out, T, iters);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("scatter 32-byte cells, 256-bit stores, 8 warps/SM: %.1f GB/s per SM, %.2f TB/s total\n",
               (double)iters * 8 * 8192 / ms / 1e6, (double)sms * iters * 8 * 8192 / ms / 1e9);
        for (int begins : {32, 128}) {
            CUtensorMap map;
            const cuuint64_t dims[3] = {88, (cuuint64_t)T, (cuuint64_t)T};
            const cuuint64_t strides[2] = {88 * 4, (cuuint64_t)T * 88 * 4};
            const cuuint32_t box[3] = {8, (cuuint32_t)begins, 8};
            const cuuint32_t es[3] = {1, 1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("encode failed\n");
                return 1;
            }
            const size_t smem = 65536 + 128;
            const int issuers = 128 / begins * 8 > 8 ? 8 : 128 / begins * 8;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (begins == 32) {
                    CK(cudaFuncSetAttribute(scatter_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    scatter_tma<32><<<sms, 256, smem>>>(map, T, iters);
                } else {
                    CK(cudaFuncSetAttribute(scatter_tma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    scatter_tma<128><<<sms, 256, smem>>>(map, T, iters / 4);
                }
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double bytes = begins == 32 ? (double)iters * issuers * 8192 : (double)(iters / 4) * issuers * 32768;
            printf("scatter 32-byte cells, TMA store boxes {8 tracks, %d begins, 8 ends}, %d issuing threads/SM: %.1f GB/s per SM, %.2f TB/s total\n",
                   begins, issuers, bytes / ms / 1e6, sms * bytes / ms / 1e9);
        }
        {
            const int tiles = (T / 128) * (T / 64) * 2;   // two sweeps over the tensor
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                scatter_tiles<8><<<sms, 256>>>(out, T, tiles * 11);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("tile-ordered 32-byte cells ( 8 tracks per item), 256-bit stores: %.2f TB/s total\n", (double)tiles * 11 * 128 * 64 * 32 / ms / 1e9);
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                scatter_tiles<16><<<sms, 256>>>(out, T, tiles * 5);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("tile-ordered 64-byte cells (16 tracks per item), 256-bit stores: %.2f TB/s total\n", (double)tiles * 5 * 128 * 64 * 64 / ms / 1e9);
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                scatter_tiles<88><<<sms, 256>>>(out, T, tiles);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("tile-ordered whole cells (88 tracks per item, contiguous), 256-bit stores: %.2f TB/s total\n", (double)tiles * 128 * 64 * 352 / ms / 1e9);
        }
        CK(cudaFree(out));
    }
    return 0;
}
