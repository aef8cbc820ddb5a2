// sip_scorer.cu -- Scaled-Inner-Product interval scorer on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces transkun/LayersTransformer.py:410-440 (ScaledInnerProductIntervalScorer.forward after the
// Linear projection): per track n
//     S[e,b,n] = ( sum_d (q[n,e,d]/sqrt(D)) * k[n,b,d] ) * |e-b|  +  [e==b] * diag[n,e]
// written directly in the layout the CRF reads, score[e][b][n] (track innermost), LOWER TRIANGLE ONLY
// (e >= b is all the semi-CRF ever reads; the reference computes the full square, multiplies it by the
// length matrix, adds diag_embed and permutes -- four more passes over T*T*N).
//
// Persistent kernel, one CTA per SM.  A work item is a tile of 128 begins x 64 ends for EIGHT TRACKS: eight 128x64
// fp32 accumulators fill the 512 TMEM columns, and eight tracks are one 32-byte sector of the track-innermost output,
// so every cell is written whole (a partially written sector makes L2 fetch the rest from DRAM before it can merge:
// measured, 16-byte stores made the epilogue 4x slower than the main loop).  Begins sit on the TMEM lanes, so a warp of
// the epilogue holds 32 consecutive begins of one end and its stores walk along a row of the output.
// Warp-specialised, no CTA barrier after the prologue:
//   warp 0   one thread: TMA producer.  Per step = (track, 32-wide K chunk) one k box (128 rows x 128 B) and one q box
//            (64 rows x 128 B), SWIZZLE_128B, into an 8-stage ring (192 KB in flight: the kernel is bound by operand
//            ingest -- 24 bytes per output at this tile shape -- so bytes in flight are what buys throughput); runs
//            ahead across work items while the epilogue drains TMEM.
//   warp 1   one thread: four tcgen05.mma.kind::tf32 (M128 N64 K8) per step, tcgen05.commit frees the stage; after
//            the last step of an item a commit publishes the accumulators.
//   warps 4-11 epilogue (two warpgroups, registers raised with setmaxnreg): tracks 0-3 are complete half an item before
//            tracks 4-7, so each thread parks its 4 x 32 values of them in registers and returns those TMEM columns at
//            once; when tracks 4-7 arrive: tcgen05.ld, 1/sqrt(D) (a power of two for D=256, exact), the length factor,
//            the diagonal and the triangle mask, one 32-byte store per cell.  The MMA thread never waits for stores.
//
// Precision: TF32 operands (the reference's own --allow_tf32 regime, train.py:41-43), fp32 accumulate; the fp32-grade
// "3xTF32" mode is the same kernel on split operands (tkb_sip_score_scaled, see transkun_b200/LayersTransformer.py).
#include <stdlib.h>

#include "common.cuh"

namespace tkb {

constexpr int SC_TM = 128;      // begins per tile (= UMMA M, TMEM lanes)
constexpr int SC_TN = 64;       // ends per tile (= UMMA N, TMEM columns per track)
constexpr int SC_NG = 8;        // tracks per work item = one 32-byte sector of the output; 8 x 64 columns = all of TMEM
constexpr int SC_KC = 32;       // tf32 elements per 128-byte swizzled row
constexpr int SC_UMMA_K = 8;    // tf32 elements per tcgen05.mma
constexpr int SC_EPI_WARPS = 8;
constexpr int SC_THREADS = 128 + 32 * SC_EPI_WARPS;   // warpgroup 0: producer warp, MMA warp, two idle; warpgroups 1-2: epilogue
constexpr int SC_CH = 2;        // K chunks per stage: one TMA box per operand brings SC_CH chunks (the issue cost of a TMA
                                // instruction, ~100 cycles on the issuing thread, is what bounds a one-chunk ring)
constexpr int SC_STAGES = 4;
constexpr int SC_PRODUCERS = 2; // producer threads (warps 0 and 2), alternating stages
constexpr int SC_A_BYTES = SC_TM * 128;   // one chunk of the k tile
constexpr int SC_B_BYTES = SC_TN * 128;   // one chunk of the q tile
constexpr int SC_STAGE_BYTES = SC_CH * (SC_A_BYTES + SC_B_BYTES);
constexpr size_t kScorerSmem = (size_t)SC_STAGES * SC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(SC_NG * SC_TN == 512, "the accumulators of one work item fill TMEM");

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float *v) {   // v: 8 registers (constant indices after unrolling)
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
    int band;                   // tile columns per band of the work order (see tile_of)
    int items;                  // tiles x track groups
    int tiles, gpass;           // tiles of the lower triangle; track groups per pass of the work order (see item_of)
    unsigned long long *trace;  // diagnostics build only (TKB_TIMELINE): [grid][16] globaltimer stamps / cycle counters
    int ablate;                 // diagnostics build only: 1 = no MMAs, 2 = no stores, 4 = q box only, 8 = k box only,
                                // 256 = no cycle counters
};

#ifdef TKB_TIMELINE
#define SC_STAMP(i)                                                                                          \
    do {                                                                                                     \
        if (p.trace) {                                                                                       \
            unsigned long long t_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
            p.trace[(size_t)blockIdx.x * 16 + (i)] = t_;                                                     \
        }                                                                                                    \
    } while (0)
#define SC_ABLATE(bit) (p.ablate & (bit))
#define SC_CLOCK() ((p.ablate & 256) ? 0ll : clock64())
#define SC_COUNT(i, v)                                              \
    do {                                                            \
        if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (i)] = (unsigned long long)(v); \
    } while (0)
#else
#define SC_CLOCK() 0ll
#define SC_COUNT(i, v) do { } while (0)
#define SC_STAMP(i) do { } while (0)
#define SC_ABLATE(bit) false
#endif

__device__ __forceinline__ void mbar_arrive1(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The operands [NT*T][D] are described to TMA as 3-D tensors {32 floats of a K chunk, row, chunk}: one box
// {32, rows, SC_CH} lands as SC_CH consecutive 128B-swizzled K-major tiles (rows x 128 B each) at dst.
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar,
                                                 unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
        : "memory");
}

// Tiles and work order.  Column c (begins 128c ..) has the 64-end tile rows r >= 2c (the rest is above the diagonal).
// Columns are taken in bands of `band` columns, and inside a band row by row (all columns of the band that reach r,
// then r + 1, ...): the k rows of a band (band * 128 begins x all tracks, 11.5 MB per column at 88 tracks) stay in L2
// while the q rows stream past once per BAND.  A work item is (tile, group of 8 tracks), group fastest: the SMs work on
// all groups of a few neighbouring tiles at the same time, so whole 352-byte cells reach L2 together.  idx -> (c, r).
__device__ __forceinline__ void tile_of(int idx, int T, int band, int &c, int &r) {
    const int ncol = (T + SC_TM - 1) / SC_TM, nrow = (T + SC_TN - 1) / SC_TN;
    constexpr int kRatio = SC_TM / SC_TN;
    int c0 = 0;
    for (;; c0 += band) {   // find the band
        int cnt = 0;
        for (int col = c0; col < min(c0 + band, ncol); ++col) cnt += nrow - kRatio * col;
        if (idx < cnt) break;
        idx -= cnt;
    }
    // rows [kRatio*(c0+i), kRatio*(c0+i+1)) are reached by the columns c0 .. c0+i only
    const int c1 = min(c0 + band, ncol);
    for (int i = 0; c0 + i < c1; ++i) {
        const int lo = kRatio * (c0 + i), hi = (c0 + i + 1 < c1) ? kRatio * (c0 + i + 1) : nrow;
        const int cnt = (hi - lo) * (i + 1);
        if (idx < cnt) {
            r = lo + idx / (i + 1);
            c = c0 + idx % (i + 1);
            return;
        }
        idx -= cnt;
    }
    c = c1 - 1;  // not reached
    r = nrow - 1;
}

// work item -> (tile index, first track).  Track groups are taken in passes of p.gpass groups: all tiles for the first
// gpass groups, then the next gpass groups, ... (group fastest inside a pass).
__device__ __forceinline__ void item_of(int w, int ngroups, int tiles, int gpass, int &tile, int &n0) {
    const int per_pass = tiles * gpass;
    const int pass = w / per_pass, r = w - pass * per_pass;
    const int g_in = min(gpass, ngroups - pass * gpass);   // the last pass may be short
    tile = r / g_in;
    n0 = (pass * gpass + r - tile * g_in) * SC_NG;
}

__global__ void __launch_bounds__(SC_THREADS, 1)
    sip_scorer_kernel(const __grid_constant__ CUtensorMap mapk, const __grid_constant__ CUtensorMap mapq, const ScorerParams p) {
    extern __shared__ unsigned char smem_raw_sc[];
    const unsigned smem_base = (smem_u32(smem_raw_sc) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
    const unsigned bars = smem_base + SC_STAGES * SC_STAGE_BYTES;
    // accumulator hand-over in two halves (tracks 0-3 | 4-7): acc_full[h], acc_empty[h]
    const unsigned full_b = bars, empty_b = bars + 8 * SC_STAGES, acc_full_b = bars + 16 * SC_STAGES,
                   acc_empty_b = acc_full_b + 16;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T, NT = p.NT;
    const int ngroups = (NT + SC_NG - 1) / SC_NG;
    const int nchunks = p.D / SC_KC;
    if (tid == 128) {
        SC_STAMP(0);
        SC_COUNT(14, SC_CLOCK());
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(SC_NG * SC_TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < SC_STAGES; ++s) {
            mbar_init(full_b + 8 * s, 1);    // a producer's arrive.expect_tx; the TMA copies complete the bytes
            mbar_init(empty_b + 8 * s, 1);   // tcgen05.commit: the MMAs that read the stage have completed
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(acc_full_b + 8 * h, 1);                // tcgen05.commit after the last step of the half
            mbar_init(acc_empty_b + 8 * h, SC_EPI_WARPS);    // every epilogue warp has read its part of the half
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;
    if (tid == 128) SC_STAMP(1);

    // Register budget per role (the file is split per scheduler: 3 warps x 32 x 168 is all a 12-warp CTA gets at launch):
    // the producer / MMA warpgroup gives most of its registers back, the epilogue warps take them for the parked half
    const int nstages = (nchunks + SC_CH - 1) / SC_CH;   // stages per track (a last odd chunk is zero-filled by TMA)
    if (warp == 0 || warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(80));
        if (lane == 0) {
            // ---- producers: two threads, alternating stages.  Per stage: wait until the MMAs that read the slot have
            // completed, arm the barrier, one box of the k tile (128 rows x SC_CH chunks) and one of the q tile (64 rows).
            const int me = warp >> 1;
            unsigned long long keep;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
            int it = 0;
            long long c_wait = 0, c_issue = 0;
#pragma unroll 1
            for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
                int col, row, tile, n0;
                item_of(w, ngroups, p.tiles, p.gpass, tile, n0);
                tile_of(tile, T, p.band, col, row);
                const int b0 = col * SC_TM, e0 = row * SC_TN;
                const int ntrk = min(SC_NG, NT - n0);
#pragma unroll 1
                for (int t = 0; t < ntrk; ++t) {
                    const int row0 = (n0 + t) * T;
#pragma unroll 1
                    for (int st = 0; st < nstages; ++st, ++it) {
                        if (it % SC_PRODUCERS != me) continue;
                        const int slot = it % SC_STAGES;
                        const long long c0 = SC_CLOCK();
                        if (it >= SC_STAGES) mbar_wait(empty_b + 8 * slot, (unsigned)(((it / SC_STAGES) - 1) & 1));
                        const long long c1 = SC_CLOCK();
                        c_wait += c1 - c0;
                        const unsigned stage = smem_base + (unsigned)slot * SC_STAGE_BYTES;
                        mbar_arrive_expect_tx(full_b + 8 * slot, (unsigned)(SC_CH * ((SC_ABLATE(4) ? 0 : SC_A_BYTES) + (SC_ABLATE(8) ? 0 : SC_B_BYTES))));
                        // the k tile of a column serves every tile row below it: keep it in L2 in preference to the rest
                        if (!SC_ABLATE(4)) {
                            tma_load_3d_hint(stage, &mapk, 0, row0 + b0, st * SC_CH, full_b + 8 * slot, keep);
                        }
                        if (!SC_ABLATE(8)) tma_load_3d(stage + SC_CH * SC_A_BYTES, &mapq, 0, row0 + e0, st * SC_CH, full_b + 8 * slot);
                        c_issue += SC_CLOCK() - c1;
                    }
                }
            }
            if (me == 0) {
                SC_COUNT(8, c_wait);
                SC_COUNT(9, c_issue);
                SC_COUNT(12, it);
            }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(80));
        if (lane == 0) {
            // ---- MMA issuer: acc[track][begin (lane)][end (column)] += k_tile . q_tile^T
            const unsigned idesc = umma_idesc_tf32(SC_TM, SC_TN);
            int it = 0, ni = 0;
            long long c_wait = 0, c_issue = 0, c_tmem = 0;
#pragma unroll 1
            for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++ni) {
                int tile, n0;
                item_of(w, ngroups, p.tiles, p.gpass, tile, n0);
                const int ntrk = min(SC_NG, NT - n0);
#pragma unroll 1
                for (int t = 0; t < ntrk; ++t) {
                    if (ni > 0 && (t == 0 || t == SC_NG / 2)) {
                        // the epilogue of the previous item has read this half of TMEM (tracks 0-3 are copied to
                        // registers as soon as they are complete, so this wait is normally already satisfied)
                        const long long c0 = SC_CLOCK();
                        mbar_wait(acc_empty_b + (t == 0 ? 0 : 8), (unsigned)((ni - 1) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        c_tmem += SC_CLOCK() - c0;
                    }
#pragma unroll 1
                    for (int st = 0; st < nstages; ++st, ++it) {
                        const int slot = it % SC_STAGES;
                        const long long c0 = SC_CLOCK();
                        mbar_wait(full_b + 8 * slot, (unsigned)((it / SC_STAGES) & 1));
                        const long long c1 = SC_CLOCK();
                        c_wait += c1 - c0;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned stage = smem_base + (unsigned)slot * SC_STAGE_BYTES;
#pragma unroll
                        for (int c = 0; c < SC_CH; ++c) {
                            const unsigned long long da = umma_desc_sw128(stage + c * SC_A_BYTES),
                                                     db = umma_desc_sw128(stage + SC_CH * SC_A_BYTES + c * SC_B_BYTES);
#pragma unroll
                            for (int kk = 0; kk < SC_KC / SC_UMMA_K; ++kk)  // +32 bytes along K inside the swizzle atom = +2 in the address field
                                if (!SC_ABLATE(1)) umma_tf32(tmem_base + t * SC_TN, da + 2 * kk, db + 2 * kk, idesc, (st > 0 || c > 0 || kk > 0) ? 1u : 0u);
                        }
                        umma_commit(empty_b + 8 * slot);
                        c_issue += SC_CLOCK() - c1;
                    }
                    if (t == min(SC_NG / 2, ntrk) - 1) umma_commit(acc_full_b);
                }
                umma_commit(acc_full_b + 8);
            }
            SC_COUNT(10, c_wait);
            SC_COUNT(11, c_issue);
            SC_COUNT(13, c_tmem);
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(80));
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(208));
        // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. (its 32 consecutive begins); two warps share a lane quarter
        // and split the tile's 64 ends.  A thread gathers the 8 tracks of a cell and writes them as ONE 32-byte sector;
        // one store instruction = one end x 32 consecutive begins.
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int align = ((p.pitch % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0))   ? 32
                          : ((p.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) ? 16
                                                                                                      : 4;
        int ni = 0;
#pragma unroll 1
        for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++ni) {
            int col, row, tile, n0;
            item_of(w, ngroups, p.tiles, p.gpass, tile, n0);
            tile_of(tile, T, p.band, col, row);
            const int b0 = col * SC_TM, e0 = row * SC_TN;
            const int ntrk = min(SC_NG, NT - n0);
            const int b = b0 + 32 * quarter + lane;
            const unsigned tlane = (unsigned)(32 * quarter) << 16;
            const int jbase = half * (SC_TN / 2);
            // tracks 0-3 are complete half an item before tracks 4-7: park this thread's 4 x 32 values in registers and
            // give those TMEM columns back, so the next item's MMAs never wait for the stores of this one
            float stash[SC_NG / 2][SC_TN / 2];
            if (lane == 0) mbar_wait(acc_full_b, (unsigned)(ni & 1));   // one probe per warp
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 128 && ni == 0) SC_STAMP(2);
#pragma unroll
            for (int t = 0; t < SC_NG / 2; ++t) {
#pragma unroll
                for (int c = 0; c < SC_TN / 16; ++c) {
                    if (t < ntrk) {
                        tmem_ld8(tmem_base + tlane + (unsigned)(t * SC_TN + jbase + 8 * c), &stash[t][8 * c]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) stash[t][8 * c + i] = 0.0f;
                    }
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive1(acc_empty_b);
            if (lane == 0) mbar_wait(acc_full_b + 8, (unsigned)(ni & 1));
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < SC_TN / 16; ++c) {
                float acc[SC_NG / 2][8];
#pragma unroll
                for (int t = 0; t < SC_NG / 2; ++t) {
                    if (SC_NG / 2 + t < ntrk) {
                        tmem_ld8(tmem_base + tlane + (unsigned)((SC_NG / 2 + t) * SC_TN + jbase + 8 * c), acc[t]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[t][i] = 0.0f;
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == SC_TN / 16 - 1) {
                    // this warp's last read of TMEM for the item
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive1(acc_empty_b + 8);
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int e = e0 + jbase + 8 * c + jj;
                    if (b <= e && e < T && !SC_ABLATE(2)) {
                        float v[SC_NG];
                        const float len = (float)(e - b);
                        if (b == e) {   // the diagonal cell of this thread (at most one per item): the skip score
#pragma unroll
                            for (int t = 0; t < SC_NG; ++t) v[t] = t < ntrk ? p.diag[(size_t)(n0 + t) * T + b] : 0.0f;
                        } else {
#pragma unroll
                            for (int t = 0; t < SC_NG / 2; ++t) {
                                v[t] = (stash[t][8 * c + jj] * p.qscale) * len;
                                v[SC_NG / 2 + t] = (acc[t][jj] * p.qscale) * len;
                            }
                        }
                        float *o = p.out + ((size_t)e * T + b) * p.pitch + n0;
                        if (align == 32 && ntrk == SC_NG) {
                            // written once, read by a later kernel: first in line for eviction, the operands stay
                            asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o),
                                         "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
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
            if (tid == 128 && ni == 0) SC_STAMP(3);
        }
        if (tid == 128) SC_STAMP(4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SC_NG * SC_TN) : "memory");
    if (tid == 128) {
        SC_STAMP(5);
        SC_COUNT(15, SC_CLOCK());
    }
}

// 3xTF32 operand preparation: x = hi + lo with hi = x with the 13 low mantissa bits cleared (exact in TF32) and
// lo = x - hi (exact in fp32); q3 = [hi | hi | lo], k3 = [hi | lo | hi] along the feature axis, so that
// q3 . k3 = hi.hi + hi.lo + lo.hi -- one pass over q and k instead of the four elementwise kernels and two
// concatenations the host-side formulation costs.
__global__ void __launch_bounds__(256) sip_split3_kernel(const float4 *__restrict__ q, const float4 *__restrict__ k,
                                                        float4 *__restrict__ q3, float4 *__restrict__ k3, long long rows,
                                                        int d4) {
    const long long n = rows * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d4;
        const int c = (int)(i - r * d4);
        auto split = [](float4 x, float4 &hi, float4 &lo) {
            hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        };
        float4 qh, ql, kh, kl;
        split(q[i], qh, ql);
        split(k[i], kh, kl);
        float4 *qo = q3 + r * 3 * d4 + c, *ko = k3 + r * 3 * d4 + c;
        qo[0] = qh;
        qo[d4] = qh;
        qo[2 * d4] = ql;
        ko[0] = kh;
        ko[d4] = kl;
        ko[2 * d4] = kh;
    }
}

static long long tile_count(int T) {
    const int ncol = (T + SC_TM - 1) / SC_TM, nrow = (T + SC_TN - 1) / SC_TN;
    long long n = 0;
    for (int c = 0; c < ncol; ++c) n += nrow - SC_TM / SC_TN * c;
    return n;
}

static int encode_operand(CUtensorMap *map, const float *base, long long rows, int D, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("tkb_sip_score: cuTensorMapEncodeTiled is not available from this driver");
        return TKB_ENODEV;
    }
    // {32 floats of a chunk, row, chunk}: a row is D floats = D/32 chunks of 128 bytes
    const cuuint64_t dims[3] = {SC_KC, (cuuint64_t)rows, (cuuint64_t)(D / SC_KC)};
    const cuuint64_t strides[2] = {(cuuint64_t)D * 4, 128};
    const cuuint32_t box[3] = {SC_KC, (cuuint32_t)box_rows, SC_CH};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tkb_sip_score: cuTensorMapEncodeTiled failed (%d) for rows=%lld D=%d", (int)r, rows, D);
        return TKB_EINVAL;
    }
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

extern "C" int tkb_sip_split3(const float *q, const float *k, long long rows, int D, float *q3, float *k3, void *stream_) {
    if (!q || !k || !q3 || !k3 || rows < 1 || D < 4 || D % 4 != 0 || ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                                                        reinterpret_cast<uintptr_t>(q3) | reinterpret_cast<uintptr_t>(k3)) & 15)) {
        set_error("tkb_sip_split3: invalid argument (rows=%lld D=%d; D must be a multiple of 4, pointers 16-byte aligned)", rows, D);
        return TKB_EINVAL;
    }
    const long long n = rows * (D / 4);
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    sip_split3_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(reinterpret_cast<const float4 *>(q), reinterpret_cast<const float4 *>(k),
                                                               reinterpret_cast<float4 *>(q3), reinterpret_cast<float4 *>(k3), rows, D / 4);
    TKB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tkb_sip_score_scaled(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                                    float scale, float *out_score, int64_t pitch, void *stream_) {
    if (!q || !k || !diag || !out_score || n_tracks < 1 || T < 1 || D < SC_KC || D % SC_KC != 0 || pitch < n_tracks ||
        (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15)) {
        set_error("tkb_sip_score: invalid argument (tracks=%d T=%d D=%d; D must be a multiple of 32, q/k 16-byte aligned)",
                  n_tracks, T, D);
        return TKB_EINVAL;
    }
    static int band_env = -1, sms[kMaxDevices] = {};
    static bool configured[kMaxDevices] = {};
    if (band_env < 0) {   // diagnostics: tile columns per band of the work order
        const char *e = getenv("TKB_SCORER_BAND");
        band_env = e ? atoi(e) : 0;
    }
    const int dev = current_device();
    int nsm = dev >= 0 ? sms[dev] : 0;
    if (dev < 0 || !configured[dev]) {
        TKB_CUDA(cudaFuncSetAttribute(sip_scorer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScorerSmem));
        int d = 0;
        TKB_CUDA(cudaGetDevice(&d));
        TKB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, d));
        if (dev >= 0) {
            sms[dev] = nsm;
            configured[dev] = true;
        }
    }
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
    p.ablate = 0;
#ifdef TKB_TIMELINE
    if (const char *e = getenv("TKB_SCORER_ABLATE")) p.ablate = atoi(e);
#endif
    const int ngroups = (n_tracks + SC_NG - 1) / SC_NG;
    const long long items = tile_count(T) * ngroups;
    if (items > 0x7fffffffLL || (long long)n_tracks * T > 0x7fffffffLL) {
        set_error("tkb_sip_score: problem too large (tracks=%d T=%d)", n_tracks, T);
        return TKB_EINVAL;
    }
    p.items = (int)items;
    p.tiles = (int)tile_count(T);
    static int gpass_env = -1;
    if (gpass_env < 0) {
        const char *e = getenv("TKB_SCORER_GPASS");
        gpass_env = e ? atoi(e) : 0;
    }
    // default: 6 groups per pass -- with band = 4 the k tiles of a pass (4 columns x 6 groups x 1 MB at D = 256) stay in
    // L2 next to the streaming q rows and the output; measured at T=2048, 88 tracks: DRAM reads 1.37 GB (one pass over
    // all 11 groups) -> 0.84 GB, 435 -> 416 us
    const int gpass = gpass_env > 0 ? gpass_env : 6;
    p.gpass = gpass < ngroups ? gpass : ngroups;
    CUtensorMap mk, mq;
    const long long rows = (long long)n_tracks * T;
    int rc = encode_operand(&mk, k, rows, D, SC_TM);
    if (!rc) rc = encode_operand(&mq, q, rows, D, SC_TN);
    if (rc) return rc;
    const int grid = (int)(items < nsm ? items : nsm);
    sip_scorer_kernel<<<grid, SC_THREADS, kScorerSmem, (cudaStream_t)stream_>>>(mk, mq, p);
    TKB_CUDA(cudaGetLastError());
    return 0;
}
