// tcgen05 / TMA microbenchmarks for the interval scorer (round 2).  Diagnostics only.
//
//  A. mma:    cycles per tcgen05.mma.cta_group::1.kind::tf32 (M128, K8, both operands from 128B-swizzled shared memory)
//             as a function of N in {32, 64, 128, 256}, one and two CTAs per SM issuing at the same time.
//  B. ingest: how many bytes per second can every SM pull through TMA 2-D tile copies ({32 floats, R rows} boxes,
//             SWIZZLE_128B) out of an L2-resident operand, 2 CTAs per SM, 4 boxes in flight per CTA?
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
    return 0;
}
