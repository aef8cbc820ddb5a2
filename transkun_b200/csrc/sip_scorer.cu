// sip_scorer.cu -- Scaled-Inner-Product interval scorer on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces transkun/LayersTransformer.py:410-440 (ScaledInnerProductIntervalScorer.forward after the
// Linear projection): per track n
//     S[e,b,n] = ( sum_d (q[n,e,d]/sqrt(D)) * k[n,b,d] ) * |e-b|  +  [e==b] * diag[n,e]
// written directly in the layout the CRF reads, score[e][b][n] (track innermost), LOWER TRIANGLE ONLY
// (e >= b is all the semi-CRF ever reads; the reference computes the full square, multiplies it by the
// length matrix, adds diag_embed and permutes -- four more passes over T*T*N).
//
// One CTA computes a (128 ends x 64 begins) tile for FOUR TRACKS (16 bytes of the track-innermost layout): four
// 128x64 fp32 accumulators = 256 of the 512 TMEM columns, so two CTAs share an SM and one CTA's epilogue overlaps
// the other's main loop.  The CTAs of the 22 track quads of a tile are dispatched together (quad = fastest block
// index), so L2 sees both halves of every 32-byte sector before it evicts them.  Warp-specialised: warps 0-5
// stage the operands per (track, 32-wide K chunk) (Q: 128 rows, K: 64 rows, 128 B each) with cp.async into
// 128B-swizzled K-major shared memory (4 stages, full/empty mbarriers, no CTA barrier in the main loop), one
// thread of warp 6 issues four tcgen05.mma.kind::tf32 (M128 N64 K8) per step and tcgen05.commit releases the
// stage; all 8 warps run the epilogue: tcgen05.ld from TMEM, 1/sqrt(D) (a power of two for D=256, exact), the
// length factor, the diagonal and the triangle mask, 16-byte stores [e][b][4 tracks].
//
// Precision: TF32 operands (the reference's own --allow_tf32 regime, train.py:41-43), fp32 accumulate.
#include <stdlib.h>

#include "common.cuh"

namespace tkb {

constexpr int SC_TM = 128;      // begins per tile (= UMMA M, TMEM lanes)
constexpr int SC_TN = 32;       // ends per tile (= UMMA N, TMEM columns per track)
constexpr int SC_NG = 8;        // tracks per CTA = one 32-byte sector of the track-innermost output; 8 x 32 fp32 accumulator
                                // columns = half of TMEM, two CTAs per SM
constexpr int SC_KC = 32;       // tf32 elements per 128-byte swizzled row
constexpr int SC_UMMA_K = 8;    // tf32 elements per tcgen05.mma
constexpr int SC_THREADS = 256;
constexpr int SC_PRODUCERS = 192;  // warps 0-5 stage the operands, warp 6 issues the MMAs, all 8 warps run the epilogue
constexpr int SC_STAGES = 5;
constexpr int SC_LOOKAHEAD = 2;    // cp.async groups a producer thread keeps in flight
constexpr int SC_A_BYTES = SC_TM * 128;
constexpr int SC_B_BYTES = SC_TN * 128;
constexpr int SC_STAGE_BYTES = SC_A_BYTES + SC_B_BYTES;
constexpr size_t kScorerSmem = (size_t)SC_STAGES * SC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// K-major, SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr) {
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) /*LBO (unused with swizzle)*/ |
           ((unsigned long long)(1024 >> 4) << 32) /*SBO*/ | (1ull << 46) /*sm100 descriptor version*/ |
           (2ull << 61) /*SWIZZLE_128B*/;
}
// kind::tf32, fp32 accumulate, A and B K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ | ((unsigned)(N >> 3) << 17) |
           ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

struct ScorerParams {
    const float *diag;          // [NT][T]
    float *out;                 // [T][T][pitch], the first NT tracks of a cell are written
    long long pitch;
    int NT, T, D;
    float qscale;               // 1/sqrt(D)
    int band;                   // tile columns per band of the block order (see tile_of)
    unsigned long long *trace;  // diagnostics build only (TKB_TIMELINE): [grid][8] globaltimer stamps
};

#ifdef TKB_TIMELINE
#define SC_STAMP(i)                                                                                          \
    do {                                                                                                     \
        if (p.trace) {                                                                                       \
            unsigned long long t_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
            p.trace[(size_t)blockIdx.x * 8 + (i)] = t_;                                                      \
        }                                                                                                    \
    } while (0)
#else
#define SC_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(unsigned bar, unsigned cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t"
        "}\n" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(unsigned bar, unsigned short mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
// box {SC_KC floats, rows} of the [NT*T][D] operand -> 128B-swizzled K-major rows at dst
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// the same box delivered to the same offset of every CTA in `mask`, each CTA's barrier credited with the bytes
__device__ __forceinline__ void tma_load_2d_multicast(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar,
                                                      unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Tiles and block order.  A tile is 128 begins (UMMA M, TMEM lanes) x 32 ends (UMMA N, TMEM columns) x 8 tracks: a warp
// of the epilogue holds 32 CONSECUTIVE BEGINS of one end, i.e. its stores walk along a row of the output (352 B apart,
// one DRAM page / TLB entry) instead of down a column (T * 352 B apart).  Column c (begins 128c ..) has the 32-end tile
// rows r >= 4c.  A "tile cluster" is the CX tile rows [R*CX, R*CX + CX) of one column (they share the k tile, which is
// fetched once and multicast); column c has the clusters R >= 4c / CX up to nR = ceil(ceil(T/32) / CX); CTAs of a
// cluster whose tile is entirely above the diagonal or past T only help fetching.  Columns are taken in bands of `band`
// columns, and inside a band cluster row by cluster row (all columns of the band that reach R, then R + 1, ...): the k
// rows of a band (band * 128 begins x all tracks, 11.5 MB per column at 88 tracks) stay in L2 while the q rows stream
// past once per BAND.  idx -> (c, R).
template <int CX>
__device__ __forceinline__ void tile_of(int idx, int T, int band, int &c, int &R) {
    const int ncol = (T + SC_TM - 1) / SC_TM, nR = ((T + SC_TN - 1) / SC_TN + CX - 1) / CX;
    auto first = [&](int col) { return (SC_TM / SC_TN * col) / CX; };
    int c0 = 0;
    for (;; c0 += band) {   // find the band
        int cnt = 0;
        for (int col = c0; col < min(c0 + band, ncol); ++col) cnt += nR - first(col);
        if (idx < cnt) break;
        idx -= cnt;
    }
    // cluster rows [first(c0+i), first(c0+i+1)) are reached by the columns c0 .. c0+i only
    const int c1 = min(c0 + band, ncol);
    for (int i = 0; c0 + i < c1; ++i) {
        const int lo = first(c0 + i), hi = (c0 + i + 1 < c1) ? first(c0 + i + 1) : nR;
        const int cnt = (hi - lo) * (i + 1);
        if (idx < cnt) {
            R = lo + idx / (i + 1);
            c = c0 + idx % (i + 1);
            return;
        }
        idx -= cnt;
    }
    c = c1 - 1;  // not reached
    R = nR - 1;
}

template <int CX>
__global__ void __launch_bounds__(SC_THREADS, 2)
    sip_scorer_kernel(const __grid_constant__ CUtensorMap mapk, const __grid_constant__ CUtensorMap mapq, const ScorerParams p) {
    extern __shared__ unsigned char smem_raw_sc[];
    const unsigned smem_base = (smem_u32(smem_raw_sc) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
    const unsigned bars = smem_base + SC_STAGES * SC_STAGE_BYTES;  // full[SC_STAGES], empty[SC_STAGES], acc_done
    const unsigned full_b = bars, empty_b = bars + 8 * SC_STAGES, acc_b = bars + 16 * SC_STAGES;
    __shared__ unsigned tmem_base_s;
    constexpr int A_ROWS = SC_TM / CX;   // rows of the cluster's shared k tile this CTA fetches (and multicasts)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T, D = p.D, NT = p.NT;
    if (tid == 64) SC_STAMP(0);
    // blockIdx.x = ((tile cluster) * quads + quad) * CX + rank: the CTAs of all track quads of a tile cluster are
    // dispatched together, so L2 assembles whole 352-byte cells / 128-byte lines of the track-innermost output before it
    // evicts them (measured round 2: running the tiles of a few quads back to back instead is 10-50 % slower).
    const int ngq = (NT + SC_NG - 1) / SC_NG;
    const unsigned rank = CX > 1 ? (unsigned)blockIdx.x % CX : 0u;
    const int g = ((int)blockIdx.x / CX) % ngq;
    const int n0 = g * SC_NG;
    int col, R;
    tile_of<CX>((int)blockIdx.x / (CX * ngq), T, p.band, col, R);
    const int b0 = col * SC_TM, e0 = (R * CX + (int)rank) * SC_TN;
    // a padding CTA (tile entirely above the diagonal or past T) still fetches its share of the k tile for its cluster
    const bool dead = (e0 + SC_TN - 1 < b0) || (e0 >= T);
    const int nchunks = D / SC_KC;
    const int ntrk = min(SC_NG, NT - n0);
    const int nsteps = ntrk * nchunks;
    constexpr unsigned short kAll = (unsigned short)((1u << CX) - 1u);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(SC_NG * SC_TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < SC_STAGES; ++s) {
            mbar_init(full_b + 8 * s, 1);      // the producer's arrive.expect_tx; the TMA copies complete the bytes
            mbar_init(empty_b + 8 * s, CX);    // every CTA of the cluster has finished reading the stage
        }
        mbar_init(acc_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CX > 1) cluster_sync_all();   // peers' barriers exist before anything is multicast into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;
    if (tid == 64) SC_STAMP(1);

    const unsigned idesc = umma_idesc_tf32(SC_TM, SC_TN);
    if (tid == 0) {
        // ---- producer: step s = (track t, 32-wide K chunk kc).  This CTA's q tile (64 rows) and its A_ROWS-row share of
        // the cluster's k tile (128 rows, delivered to every CTA of the cluster) go into stage s % SC_STAGES once every
        // CTA's MMAs that read the stage have completed.
        const unsigned tx = (unsigned)SC_A_BYTES + (dead ? 0u : (unsigned)SC_B_BYTES);
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int slot = s % SC_STAGES;
            if (s >= SC_STAGES) mbar_wait(empty_b + 8 * slot, (unsigned)(((s / SC_STAGES) - 1) & 1));
            const int t = s / nchunks, kc = s - t * nchunks;
            const unsigned stage = smem_base + (unsigned)slot * SC_STAGE_BYTES;
            const int row0 = (n0 + t) * T;
            mbar_arrive_expect_tx(full_b + 8 * slot, tx);
            if (CX > 1)
                tma_load_2d_multicast(stage + rank * (A_ROWS * 128), &mapk, kc * SC_KC, row0 + b0 + (int)rank * A_ROWS,
                                      full_b + 8 * slot, kAll);
            else
                tma_load_2d(stage, &mapk, kc * SC_KC, row0 + b0, full_b + 8 * slot);
            if (!dead) tma_load_2d(stage + SC_A_BYTES, &mapq, kc * SC_KC, row0 + e0, full_b + 8 * slot);
        }
    } else if (tid == 32) {
        // ---- MMA issuer: one thread; acc[track][begin (lane)][end (column)] += k_tile . q_tile^T; tcgen05.commit
        // releases the stage in every CTA of the cluster
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int slot = s % SC_STAGES;
            mbar_wait(full_b + 8 * slot, (unsigned)((s / SC_STAGES) & 1));
            if (s == 0) SC_STAMP(2);
            if (s == nsteps - 1) SC_STAMP(3);
            if (dead) {
                for (unsigned c = 0; c < (unsigned)CX; ++c) mbar_arrive_remote(empty_b + 8 * slot, c);
                continue;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int t = s / nchunks, kc = s - t * nchunks;
            const unsigned stage = smem_base + (unsigned)slot * SC_STAGE_BYTES;
            const unsigned long long da = umma_desc_sw128(stage), db = umma_desc_sw128(stage + SC_A_BYTES);
#pragma unroll
            for (int kk = 0; kk < SC_KC / SC_UMMA_K; ++kk)  // +32 bytes along K inside the swizzle atom = +2 in the address field
                umma_tf32(tmem_base + t * SC_TN, da + 2 * kk, db + 2 * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            if (CX > 1)
                umma_commit_multicast(empty_b + 8 * slot, kAll);
            else
                umma_commit(empty_b + 8 * slot);
        }
        if (!dead) umma_commit(acc_b);
    }
    __syncwarp();
    if (!dead) {
        // all accumulators complete
        mbar_wait(acc_b, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) SC_STAMP(4);

        // epilogue: warp w reads TMEM lanes 32*(w%4).. (its 32 consecutive begins); warps 0-3 take ends 0..15 of the tile,
        // warps 4-7 ends 16..31.  A thread gathers the 8 tracks of a cell and writes them as ONE 32-byte sector (a
        // partially written sector makes L2 fetch the rest from DRAM before it can merge); one store instruction = one
        // end x 32 consecutive begins.
        const int b = b0 + 32 * (warp & 3) + lane;
        const int jbase = (warp >> 2) * (SC_TN / 2);
        const int align = ((p.pitch % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0))   ? 32
                          : ((p.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) ? 16
                                                                                                      : 4;
        float dg[SC_NG];
#pragma unroll
        for (int t = 0; t < SC_NG; ++t) dg[t] = (b < T && t < ntrk) ? p.diag[(size_t)(n0 + t) * T + b] : 0.0f;
#pragma unroll 1
        for (int j0 = jbase; j0 < jbase + SC_TN / 2; j0 += 8) {
            float acc[SC_NG][8];
#pragma unroll
            for (int t = 0; t < SC_NG; ++t) {
                if (t < ntrk) {
                    tmem_ld8(tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(t * SC_TN + j0), acc[t]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[t][i] = 0.0f;
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int e = e0 + j0 + jj;
                if (b <= e && e < T) {
                    float v[SC_NG];
                    const float len = (float)(e - b);
#pragma unroll
                    for (int t = 0; t < SC_NG; ++t) v[t] = (b == e) ? dg[t] : (acc[t][jj] * p.qscale) * len;
                    float *o = p.out + ((size_t)e * T + b) * p.pitch + n0;
                    if (align == 32 && ntrk == SC_NG) {
                        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(v[0]), "f"(v[1]),
                                     "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                                     : "memory");
                    } else if (align >= 16 && ntrk >= 4) {
                        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        if (ntrk == SC_NG) {
                            *reinterpret_cast<float4 *>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
#pragma unroll
                            for (int t = 4; t < SC_NG; ++t)
                                if (t < ntrk) o[t] = v[t];
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < SC_NG; ++t)
                            if (t < ntrk) o[t] = v[t];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 64) SC_STAMP(5);
    if (CX > 1) cluster_sync_all();   // no peer may still signal this CTA's barriers once it has gone
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SC_NG * SC_TN) : "memory");
}

// tile clusters of the lower triangle (see tile_of)
static long long tile_clusters(int T, int cx) {
    const int ncol = (T + SC_TM - 1) / SC_TM, nR = ((T + SC_TN - 1) / SC_TN + cx - 1) / cx;
    long long n = 0;
    for (int c = 0; c < ncol; ++c) n += nR - (SC_TM / SC_TN * c) / cx;
    return n;
}

static int encode_operand(CUtensorMap *map, const float *base, long long rows, int D, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("tkb_sip_score: cuTensorMapEncodeTiled is not available from this driver");
        return TKB_ENODEV;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
    const cuuint32_t box[2] = {SC_KC, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tkb_sip_score: cuTensorMapEncodeTiled failed (%d) for rows=%lld D=%d", (int)r, rows, D);
        return TKB_EINVAL;
    }
    return 0;
}

template <int CX>
static int launch_scorer(const CUtensorMap &mk, const CUtensorMap &mq, const ScorerParams &p, cudaStream_t stream) {
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || !configured[dev]) {
        TKB_CUDA(cudaFuncSetAttribute(sip_scorer_kernel<CX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScorerSmem));
        if (dev >= 0) configured[dev] = true;
    }
    const int ngq = (p.NT + SC_NG - 1) / SC_NG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tile_clusters(p.T, CX) * ngq * CX));
    cfg.blockDim = dim3(SC_THREADS);
    cfg.dynamicSmemBytes = kScorerSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CX;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TKB_CUDA(cudaLaunchKernelEx(&cfg, sip_scorer_kernel<CX>, mk, mq, p));
    return 0;
}

}  // namespace tkb

using namespace tkb;

static unsigned long long *g_scorer_trace = nullptr;  // diagnostics build only
#ifdef TKB_TIMELINE
extern "C" void tkb_debug_set_scorer_trace(unsigned long long *buf) { g_scorer_trace = buf; }
#endif

extern "C" int tkb_sip_score(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                             float *out_score, void *stream_) {
    return tkb_sip_score_pitched(q, k, diag, n_tracks, T, D, out_score, n_tracks, stream_);
}

extern "C" int tkb_sip_score_pitched(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                                     float *out_score, int64_t pitch, void *stream_) {
    return tkb_sip_score_scaled(q, k, diag, n_tracks, T, D, D > 0 ? 1.0f / sqrtf((float)D) : 0.0f, out_score, pitch, stream_);
}

extern "C" int tkb_sip_score_scaled(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                                    float scale, float *out_score, int64_t pitch, void *stream_) {
    if (!q || !k || !diag || !out_score || n_tracks < 1 || T < 1 || D < SC_KC || D % SC_KC != 0 || pitch < n_tracks ||
        (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15)) {
        set_error("tkb_sip_score: invalid argument (tracks=%d T=%d D=%d; D must be a multiple of 32, q/k 16-byte aligned)",
                  n_tracks, T, D);
        return TKB_EINVAL;
    }
    static int cx_env = -1, band_env = -1;
    if (cx_env < 0) {   // diagnostics: cluster width (1, 2, 4) and rows per band of the block order
        const char *e = getenv("TKB_SCORER_CX");
        cx_env = e ? atoi(e) : 0;
        e = getenv("TKB_SCORER_BAND");
        band_env = e ? atoi(e) : 0;
    }
    const int nrow = (T + SC_TN - 1) / SC_TN;
    int cx = cx_env > 0 ? cx_env : 2;
    while (cx > 1 && (cx > nrow || (cx != 2 && cx != 4))) cx >>= 1;
    ScorerParams p;
    p.diag = diag;
    p.out = out_score;
    p.NT = n_tracks;
    p.pitch = pitch;
    p.T = T;
    p.D = D;
    p.qscale = scale;
    p.band = band_env > 0 ? band_env : 4;
    p.trace = g_scorer_trace;
    CUtensorMap mq, mk;
    int rc = encode_operand(&mk, k, (long long)n_tracks * T, D, SC_TM / cx);
    if (rc) return rc;
    rc = encode_operand(&mq, q, (long long)n_tracks * T, D, SC_TN);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (cx == 4) return launch_scorer<4>(mk, mq, p, stream);
    if (cx == 2) return launch_scorer<2>(mk, mq, p, stream);
    return launch_scorer<1>(mk, mq, p, stream);
}
