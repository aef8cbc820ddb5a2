// semicrf_sweep_strip.cu -- the semi-Markov dynamic programme, STRIP design (experimental: TKB_SWEEP=strip).
//
// Same contract, same ABI and same parity ladder as semicrf_sweep.cu (see there for the recurrence, the mirrored
// coordinates and the reference lines it replaces); a different decomposition, built on what round 1 measured
// (profiles/r01_sweep_experiments.txt):
//   * the track-innermost layout gives a CTA that owns 8 tracks one 32-byte sector per cell; a CTA that owns a
//     32-column strip for ALL tracks of a range reads whole contiguous rows.  STRIP CTAs (every SM that is not a
//     solver) process units (range, column block J, row class k): thread <-> (column, 4 tracks), one 16-byte
//     cp.async per row and thread into its own FIFO slot (a warp covers 512 contiguous bytes), 4 rows per stage,
//     {max, argmax, M, S} in registers for the whole unit.  The solved rows of a stage come from a plain-float row
//     table [T][2][Npad] by cp.async.bulk copies into a shared ring behind mbarriers (one row warp per CTA);
//     validity is one flag per 32-row block, written by the publisher after a fence.  Unit partials go to a
//     [block][class][word][track][column] buffer with a flag per unit; the solver's prep warp merges the classes;
//   * the near band (rows within ND blocks of the diagonal) is copied by the same units, transposed through the
//     FIFO, into an L2-resident ring laid out per track, so that a solver loads its bands with coalesced 16-byte
//     copies instead of a sector gather that shares the LSU pipe with its own chain warps;
//   * SOLVER CTAs own 2 tracks: one chain warp per (track, semiring), each alone on an SMSP, lane = column.
//     Columns are solved in micro-blocks of four, redundantly in every lane's registers (values only: the Viterbi
//     argmax stays in the owner lane's push, so the reference's tie order is untouched); the log-sum chain keeps
//     the scale M of a pair on a max-plus recursion of its own, so every exponent is known from the M's alone and
//     the S's follow with fused multiply-adds whose weights are <= 1.  A prep warp stages per-column constants and
//     the merged far partial, publisher warps do every global store: a chain warp touches shared memory only.
// Today this design is slower than the default (425 vs 269 us at T=2048, N=88): each unit's last stage needs the
// chain to have finished block J+ND+1 and its CTA idles there, and the hand-over flag -> ring -> stage -> partial
// -> flag -> merge is ~15 us against a 2 us block.  All CTAs of a launch must be co-resident (cooperative launch).
#include <stdlib.h>

#include "common.cuh"

namespace tkb {
namespace strip {

constexpr int NG = 8;      // tracks per group
constexpr int NQ = 2;      // tracks per solver CTA
constexpr int NSOLV = NG / NQ;  // solver CTAs per group
constexpr int BX = 32;     // columns per block (= lanes of a chain warp)
#ifndef TKB_ND
#define TKB_ND 2
#endif
constexpr int ND = TKB_ND;  // blocks above the diagonal block that the solver pushes itself
#ifndef TKB_NBAND
#define TKB_NBAND 4
#endif
constexpr int NBAND = TKB_NBAND;  // row bands resident in a solver CTA
constexpr int NPREP = 3;          // prep slots (per-column constants + far partial of a block)
constexpr int BANDCOLS = (ND + 1) * BX;
constexpr int NCW = 2 * NQ;  // chain warps: (track, semiring)
constexpr int NLW = 4;       // loader warps
// strip (far-field) CTAs: thread <-> (column of a 32-column block, 4 consecutive tracks) over a RANGE of up to TRK
// tracks, so that a row of the strip is one contiguous run of the score tensor
constexpr int TRK = 88;            // tracks per range (a multiple of NG)
constexpr int NQD = TRK / 4;       // track quads per range
constexpr int NCT = BX * NQD;      // consumer threads (704 = 22 warps)
constexpr int NCONW = NCT / 32;    // consumer warps of a strip CTA
constexpr int NQW = 1;             // + the row warp (bulk copies of solved rows into a shared ring)
constexpr int NW = NCONW + NQW;    // warps per CTA (every role)
constexpr int NT = NW * 32;        // 768 threads
#ifndef TKB_NQF
#define TKB_NQF 16
#endif
constexpr int NQF = TKB_NQF;       // stages of the untagged ring the consumers read: how far the consumer warps may drift apart
constexpr int KCL = 8;             // row classes: unit (J, k) takes the 4-row groups  == k (mod KCL)
constexpr int NSTG = 3;            // strip FIFO stages of 4 rows
constexpr int RB = 16;             // band slots of the global (L2-resident) near-band ring
// prep slot of one track: nine arrays of 32 floats
constexpr int PR_DR = 0, PR_ETA = 32, PR_FARV = 64, PR_FARS = 96, PR_SP2 = 128, PR_COMB = 160, PR_ETA2 = 192,
              PR_FARM = 224, PR_FARL = 256, PR_FLOATS = 288;

// strip shared memory: S FIFO [NSTG][4 rows][NCT] float4 (each thread only ever touches its own slots, except in
// the near-tile and partial transposes) | tagged mailbox rows [NSTG][4][2 kinds][TRK] u64 | untagged rows
// [NSTG][4][2][TRK] float
constexpr size_t kFifoBytes = (size_t)NSTG * 4 * NCT * 16;
constexpr size_t kQTagBytes = 0;  // (the tagged-word staging ring of the mailbox variant is gone)
constexpr size_t kQValBytes = (size_t)NQF * 4 * 2 * TRK * 4;
constexpr size_t kStripSmem = kFifoBytes + kQTagBytes + kQValBytes + 2 * NQF * 8;
// solver shared memory: row bands [NBAND][NQ tracks][BX rows][BANDCOLS] (planar per track: a chain warp reads
// consecutive words) | prep slots [NPREP][NQ][PR_FLOATS] | publish ring [NCW chain warps][2 blocks][BX] 8-byte
// results | mbarriers band full/empty, prep full/empty, pub full[NCW][2 blocks][8 micro-blocks] (one outstanding
// phase each), pub empty[NCW][2]
constexpr size_t kBandBytes = (size_t)NQ * BX * BANDCOLS * 4;
constexpr size_t kSolverSmem = (size_t)NBAND * kBandBytes + (size_t)NPREP * NQ * PR_FLOATS * 4 +
                               (size_t)NCW * 2 * BX * 8 + 2 * NBAND * 8 + 2 * NPREP * 8 + NCW * 16 * 8 + NCW * 2 * 8;
constexpr size_t kSweepSmem = kStripSmem > kSolverSmem ? kStripSmem : kSolverSmem;
static_assert(kFifoBytes >= (size_t)4 * TRK * 33 * 4, "partial transpose must fit the FIFO");
static_assert(kSweepSmem <= 227 * 1024, "shared memory budget");
static_assert(NCW + NLW + 1 + NCW <= NW, "solver warp roles");
static_assert(TRK % NG == 0 && (NG % NQ) == 0, "ranges hold whole groups");

constexpr size_t kHeaderBytes = 256;  // status word lives here

struct SweepParams {
    const float *Sbase;    // &S(0,0) in mirrored coordinates
    const float *etabase;  // &skip weight of x = 0
    long long sx, sy, se;  // element strides
    int T, N, Npad, g0, dir;
    int nsolv;             // solver CTAs of this launch (blocks [0, nsolv)); the rest are strip CTAs
    int nlo, nhi, SN;      // tracks [nlo, nhi) of this launch; SN = track stride of the near-band ring
    int nrange;            // track ranges of this launch: range r = tracks [nlo + r*TRK, ...)
    unsigned btag;         // launch index << 16: upper half of a flag's low word
    unsigned epoch;
    float *qtab;                // [T][2 semirings][Npad]: solved rows (plain floats)
    unsigned *bdone;            // [nb][2][Npad] epoch: block j of (semiring, track) is solved and visible
    float *part;                // [nb][KCL][4 words][Npad][BX]: far partials of unit (J, k): vmax, vsel, lM, lS
    unsigned long long *uflag;  // [nb][KCL][nrange] {btag, epoch}: partial of unit (J, k) of a range is complete
    float *scratch;             // [RB band slots][SN tracks][BX rows][BANDCOLS]: near bands re-laid out per track
    unsigned long long *bflag;  // [nrange][RB][ND+1][KCL] {btag | band + 1, epoch}: this unit's rows of the tile are in
    int *status;
    unsigned *code;  // [N][T]
    float *outv;     // [T][N] or null
    float *outl;     // [T][N] or null
    unsigned long long *timeline;  // diagnostics build only (TKB_TIMELINE): [grid][64][8] stamps
};

// Wait until a mailbox word carries this launch's epoch.  A protocol bug (or a non-co-resident grid) must not
// hang the GPU: after ~4 s the wait gives up, flags the workspace and lets the kernel drain with garbage.
// BACKOFF_NS > 0 is for waits that are NOT close to a deadline (far rows): hundreds of warps spinning on the few
// mailbox lines the chain is currently writing slow the chain's own writer down.
template <int BACKOFF_NS>
__device__ __noinline__ unsigned long long poll_slow(const unsigned long long *w, unsigned epoch, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            unsigned long long v = ld_relaxed_u64(w);
            if ((unsigned)(v >> 32) == epoch) return v;
            if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
        }
        if (*(volatile int *)status != 0) return 0;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 1);
            return 0;
        }
    }
}
__device__ __forceinline__ float ld_cg_f32(const float *p) {  // L2 only: written by another SM during this launch
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __noinline__ void poll_flag_slow(const unsigned long long *w, unsigned long long want, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            if (ld_acquire_u64(w) == want) return;
            __nanosleep(100);
        }
        if (*(volatile int *)status != 0) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 3);
            return;
        }
    }
}
#ifndef TKB_FAR_BACKOFF_NS
#define TKB_FAR_BACKOFF_NS 400
#endif
__device__ __forceinline__ void publish(unsigned long long *w, float val, unsigned epoch) {
    st_relaxed_u64(w, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(val));
}

// ---- mbarrier (shared::cta) -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrival that fires once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// same watchdog as poll_slow: a protocol bug must drain the kernel, not hang the GPU
__device__ __noinline__ void mbar_wait_slow(unsigned bar, unsigned parity, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 1024; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (*(volatile int *)status != 0) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 2);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity, int *status) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, status);
}
// for waiters that are not on the chain (loaders, prep): sleep between probes so that they do not take issue
// slots from the chain warp of the same SMSP
__device__ __noinline__ void mbar_wait_relaxed(unsigned bar, unsigned parity, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 256; ++i) {
            if (mbar_try_wait(bar, parity)) return;
            __nanosleep(200);
        }
        if (*(volatile int *)status != 0) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 2);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bulk (TMA) copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait until a block flag carries this launch's epoch (same watchdog as poll_slow)
template <int BACKOFF_NS>
__device__ __noinline__ void poll_done(const unsigned *w, unsigned epoch, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            if (ld_acquire_u32(w) == epoch) return;
            if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
        }
        if (*(volatile int *)status != 0) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 1);
            return;
        }
    }
}
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gsrc), "r"(src_bytes) : "memory");
}
// Chain -> publisher hand-off.  No "memory" clobber on purpose: volatile asm statements keep their mutual order
// (store, then arrive with release semantics, both by the same lane), while the compiler stays free to hoist the
// band reads of the next micro-block across them.
__device__ __forceinline__ void sts64_nc(unsigned saddr, unsigned lo, unsigned hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(lo), "r"(hi));
}
__device__ __forceinline__ void mbar_arrive_nc(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar));
}
__device__ __forceinline__ void sts64_f(unsigned saddr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float lds32(unsigned saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

#ifdef TKB_TIMELINE
// [grid][64 owned blocks / 64 chain blocks][8] stamps (TKB_STAMP: globaltimer ns; TKB_CSTAMP: SM clock cycles)
#define TKB_STAMP(idx, slot)                                                                    \
    do {                                                                                        \
        if (p.timeline && (idx) < 64) p.timeline[((size_t)blockIdx.x * 64 + (idx)) * 8 + (slot)] = globaltimer_ns(); \
    } while (0)
#define TKB_CSTAMP(idx, slot)                                                                   \
    do {                                                                                        \
        if (p.timeline && (idx) < 64) p.timeline[((size_t)blockIdx.x * 64 + (idx)) * 8 + (slot)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define TKB_STAMP(idx, slot) \
    do {                     \
    } while (0)
#define TKB_CSTAMP(idx, slot) \
    do {                      \
    } while (0)
#endif

// (M, S) <- (M, S) (+) sb * 2^a        value = M + log2(S); one ex2: one of the two exponents is 0
__device__ __forceinline__ void lse_push(float &M, float &S, float a, float sb) {
    const float d = M - a;
    const float e1 = ex2f(-fabsf(d));
    S = (d < 0.0f) ? fmaf(S, e1, sb) : fmaf(sb, e1, S);
    M = fmaxf(M, a);
}

// =================================================================================================
// STRIP CTA: far partials and near-band re-layout of units (range, column block J, row class k)
// =================================================================================================
// Consumer thread t <-> (column x0 + t / nq, tracks 4 * (t % nq) .. +3 of the range): consecutive threads read
// consecutive addresses, so one cp.async instruction of a warp covers 512 contiguous bytes (the round-1 helper
// fetched one 32-byte sector per cell and was bound by L1 wavefronts at ~30 GB/s per SM).  The rows of a unit are the
// 4-row groups == k (mod KCL) of the far field of block J, top down, one group per stage; a thread keeps its
// 4 x {max, argmax, M, S} in registers for the whole unit and only ever waits for its own cp.async groups, so
// the consumer warps drift freely.  Two mailbox warps stage the solved rows of a stage (4 rows x 2 semirings x TRK
// tagged words, tag-checked once per CTA) as plain floats in a ring behind full/empty mbarriers.
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void strip_role(const SweepParams &p, unsigned char *smem_raw, int h, int nstrip) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int nb = (T + BX - 1) / BX;
    const unsigned fifo_s = smem_u32(smem_raw);             // [NSTG*4 rows][NCT] float4
    const unsigned qtag_s = fifo_s + (unsigned)kFifoBytes;  // [NSTG][8 = row*2+kind][TRK] u64
    const unsigned qval_s = qtag_s + (unsigned)kQTagBytes;  // [NSTG][8][TRK] float
    const unsigned full_s = qval_s + (unsigned)kQValBytes;  // + stage*8
    const unsigned empty_s = full_s + NQF * 8;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NQF; ++i) {
            mbar_init(full_s + i * 8, 32);
            mbar_init(empty_s + i * 8, NCONW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int t = threadIdx.x;
    const bool qwarp = warp >= NCONW;
    const int gkind = DO_V ? 0 : 1;  // semiring whose block flags guard the ring
    const unsigned long long tagw = ((unsigned long long)p.epoch << 32) | p.btag;
    unsigned gi = 0;  // far stages processed so far by this CTA (row ring position, same on both sides)
    int wm[2] = {nb, nb};  // row warp: per range, blocks >= wm are known complete

    const int nunits = p.nrange * nb * KCL;
    for (int u = h; u < nunits; u += nstrip) {
        // deadline order: J descending, then range, then row class
        const int k = u % KCL, rg = (u / KCL) % p.nrange, J = nb - 1 - u / (KCL * p.nrange);
        const int x0 = J * BX;
        const int n_lo = p.nlo + rg * TRK;
        const int ntr = min(TRK, p.nhi - n_lo);
        const int nq = (ntr + 3) >> 2;
        const int col = t / nq, quad = t - col * nq;
        const bool act = t < BX * nq && (x0 + col) < T;  // my column exists
        const int ntq = min(4, ntr - 4 * quad);          // tracks of my quad that exist
        const float *pbase = p.Sbase + (long long)(x0 + col) * p.sx + n_lo + 4 * quad;  // + y * sy
        const int uord = (u - h) / nstrip;  // diagnostics: ordinal of this unit in my list
        (void)uord;
        if (t == 0) TKB_STAMP(uord, 0);
        auto load_piece = [&](unsigned dst, int y, bool ok) {  // my 16 bytes of row y (zero-filled if !ok)
            const float *src = ok ? pbase + (long long)y * p.sy : p.Sbase;
            if (ALIGN == 16) {
                cp_async16_s(dst, src, ok ? ntq * 4 : 0);
            } else if (ALIGN == 8) {
                cp_async8_s(dst, src, (ok && ntq > 0) ? 8 : 0);
                cp_async8_s(dst + 8, ok && ntq > 2 ? src + 2 : p.Sbase, (ok && ntq > 2) ? 8 : 0);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) cp_async4_s(dst + 4 * i, ok && i < ntq ? src + i : p.Sbase, (ok && i < ntq) ? 4 : 0);
            }
        };

#ifndef TKB_EXP_NO_NEAR
        if (!qwarp) {
        // ---------------- near tiles: rows 4k..4k+3 of the tiles d = 0..ND above my columns -> scratch ring ---------
        // ring guard: slot jb % RB may be overwritten once every chain is done with band jb + RB
        for (int d = 0; d <= ND; ++d) {
            const int jg = J + d + RB;
            if (jg <= nb - 1)
                for (int n = t; n < ntr; n += NCT) {
                    const unsigned *w = p.bdone + ((size_t)jg * 2 + gkind) * p.Npad + n_lo + n;
                    if (ld_acquire_u32(w) != p.epoch) poll_done<200>(w, p.epoch, p.status);
                }
        }
#pragma unroll 1
        for (int d0 = 0; d0 <= ND; d0 += NSTG) {  // NSTG tiles fit the FIFO at a time
#pragma unroll
            for (int dd = 0; dd < NSTG; ++dd)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int d = d0 + dd, y = (J + d) * BX + 4 * k + r;
                    load_piece(fifo_s + (unsigned)((dd * 4 + r) * NCT + t) * 16u, y,
                               act && d <= ND && y < T && (x0 + col) <= y);
                }
            cp_async_commit();
            cp_async_wait_all();
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            // transposed read: warp <-> quad, lane <-> column; 128-byte stores per (track, row)
            if (warp < nq) {
#pragma unroll
                for (int dd = 0; dd < NSTG; ++dd) {
                    const int d = d0 + dd, jb = J + d;
                    if (d > ND || jb > nb - 1) continue;
                    float *dst0 = p.scratch +
                                  (((size_t)(jb % RB) * p.SN + (n_lo - p.nlo) + 4 * warp) * BX + 4 * k) * BANDCOLS +
                                  (ND - d) * BX + lane;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 v = lds128(fifo_s + (unsigned)((dd * 4 + r) * NCT + lane * nq + warp) * 16u);
                        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (4 * warp + i < ntr) dst0[((size_t)i * BX + r) * BANDCOLS] = vv[i];
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
        }
        if (t == 0) {
            __threadfence();
            for (int d = 0; d <= ND; ++d) {
                const int jb = J + d;
                if (jb <= nb - 1)
                    st_relaxed_u64(p.bflag + (((size_t)rg * RB + (jb % RB)) * (ND + 1) + d) * KCL + k,
                                   tagw | (unsigned)(jb + 1));
            }
        }

        }
#endif
        if (t == 0) TKB_STAMP(uord, 1);
        // ---------------- far field: rows T-1 .. x0 + (ND+1)*BX, my groups of four ---------------------------------
        const int ylow = x0 + (ND + 1) * BX;
        const int R = T - ylow;
        if (R < 1) continue;
        const int nq4 = (R + 3) >> 2;
        const int mine = nq4 > k ? (nq4 - k + KCL - 1) / KCL : 0;  // stages of this unit
#ifdef TKB_EXP_NO_Q
        if (qwarp) continue;
#endif
        if (qwarp) {
            // ---- the row warp: solved rows of a stage (4 rows x 2 semirings x my range) -> ring, by bulk copies -----
            // A row may be read once its whole 32-row block is flagged complete (the publisher fences, then flags), so
            // the ring needs no per-word validation: lane 0 issues up to 8 bulk copies per stage, the consumers wait
            // on the stage's mbarrier (arrival + byte count).
            const unsigned rowbytes = (unsigned)(((ntr + 3) >> 2) * 16);
            int &wmr = wm[rg];  // blocks >= wmr of my range are known complete
            for (int i = 0; i < mine; ++i) {
                const unsigned fs = (unsigned)((gi + i) % NQF);
                const int y0 = T - 1 - 4 * (k + i * KCL);
                const int ylo = max(y0 - 3, ylow);
                if (wmr > ylo / BX) {
                    // advance the watermark: scan the flags of up to 8 blocks below it with relaxed loads (all in
                    // flight together), keep the complete prefix, one acquire fence per scan
                    const int need = ylo / BX;
                    const unsigned long long t0 = globaltimer_ns();
                    for (;;) {
                        const int nscan = min(8, wmr);
                        unsigned bad = 0;
                        for (int b = 0; b < nscan; ++b) {
                            const unsigned *fl = p.bdone + (size_t)(wmr - 1 - b) * 2 * p.Npad + n_lo;
                            for (int n = lane; n < ntr; n += 32) {
                                if (DO_V && ld_relaxed_u32(fl + n) != p.epoch) bad |= 1u << b;
                                if (DO_L && ld_relaxed_u32(fl + p.Npad + n) != p.epoch) bad |= 1u << b;
                            }
                        }
                        bad = __reduce_or_sync(kFull, bad);
                        const int adv = bad ? (__ffs(bad) - 1) : nscan;  // complete blocks right below the watermark
                        if (adv > 0) {
                            asm volatile("fence.acq_rel.gpu;" ::: "memory");
                            wmr -= adv;
                        }
                        if (wmr <= need) break;
                        if (adv == 0 && (wmr - 1) * BX >= ylow + 2 * BX) __nanosleep(TKB_FAR_BACKOFF_NS);
                        if (*(volatile int *)p.status != 0) break;
                        if (globaltimer_ns() - t0 > 4000000000ull) {
                            atomicExch(p.status, 1);
                            break;
                        }
                    }
                }
#ifndef TKB_EXP_Q_NOWAIT
                if (gi + i >= NQF) mbar_wait(empty_s + fs * 8, (((gi + i) / NQF) - 1) & 1, p.status);
#endif
                const unsigned dst0 = qval_s + (unsigned)(fs * 8 * TRK) * 4u;
                int nlive = 0;
#pragma unroll
                for (int rk = 0; rk < 8; ++rk) {
                    const int y = y0 - (rk >> 1);
                    const bool live = y >= ylow && ((rk & 1) ? DO_L : DO_V);
                    nlive += live;
                    if (!live)  // rows (or a semiring) that do not exist
                        for (int n = lane; n < TRK; n += 32) sts32(dst0 + (unsigned)(rk * TRK + n) * 4u, (rk & 1) ? -FLT_MAX : -INFINITY);
                }
                if (lane == 0) {
                    mbar_arrive_expect_tx(full_s + fs * 8, (unsigned)nlive * rowbytes);
#pragma unroll
                    for (int rk = 0; rk < 8; ++rk) {
                        const int y = y0 - (rk >> 1);
                        if (y >= ylow && ((rk & 1) ? DO_L : DO_V))
                            bulk_g2s(dst0 + (unsigned)(rk * TRK) * 4u, p.qtab + ((size_t)y * 2 + (rk & 1)) * p.Npad + n_lo, rowbytes,
                                     full_s + fs * 8);
                    }
                } else {
                    mbar_arrive(full_s + fs * 8);
                }
            }
            gi += mine;
            continue;
        }
        // ---- consumers ----
        float vmax[4], lM[4], lS[4];
        int vsel[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            vmax[i] = -INFINITY;
            vsel[i] = -1;
            lM[i] = -FLT_MAX;
            lS[i] = 0.0f;
        }
        auto issue = [&](int i) {
            if (i < mine) {
                const int y0 = T - 1 - 4 * (k + i * KCL);
                const unsigned dst = fifo_s + (unsigned)(((i % NSTG) * 4) * NCT + t) * 16u;
#pragma unroll
                for (int r = 0; r < 4; ++r) load_piece(dst + (unsigned)(r * NCT) * 16u, y0 - r, act && (y0 - r) >= ylow);
            }
            cp_async_commit();
        };
        for (int i = 0; i < NSTG - 1; ++i) issue(i);
        for (int i = 0; i < mine; ++i) {
            issue(i + NSTG - 1);
            cp_async_wait<NSTG - 1>();
            const unsigned st = (unsigned)(i % NSTG), qs = (unsigned)((gi + i) % NQF);
            const int y0 = T - 1 - 4 * (k + i * KCL);
#if !defined(TKB_EXP_NO_Q) && !defined(TKB_EXP_Q_NOWAIT)
            mbar_wait(full_s + qs * 8, ((gi + i) / NQF) & 1, p.status);
#endif
            const unsigned src = fifo_s + (unsigned)((st * 4) * NCT + t) * 16u;
            const unsigned qsrc = qval_s + (unsigned)((qs * 8) * TRK + 4 * quad) * 4u;
            float xl[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {  // descending y: the reference's candidate order
                const float4 s4 = lds128(src + (unsigned)(r * NCT) * 16u);
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                if (DO_V) {
                    const float4 q4 = lds128(qsrc + (unsigned)((2 * r) * TRK) * 4u);
                    const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float xv = qv[i4] + sv[i4];
                        const bool tk = (DIR == TKB_BACKWARD) ? (xv >= vmax[i4]) : (xv > vmax[i4]);
                        vmax[i4] = tk ? xv : vmax[i4];
                        vsel[i4] = tk ? y0 - r : vsel[i4];
                    }
                }
                if (DO_L) {
                    const float4 q4 = lds128(qsrc + (unsigned)((2 * r + 1) * TRK) * 4u);
                    const float ql[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) xl[r][i4] = fmaf(sv[i4], kLog2e, ql[i4]);
                }
            }
            if (DO_L) {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {  // one rescale per four rows
                    const float m = fmaxf(fmaxf(xl[0][i4], xl[1][i4]), fmaxf(xl[2][i4], xl[3][i4]));
                    const float Mn = fmaxf(lM[i4], m);
                    float acc = lS[i4] * ex2f(lM[i4] - Mn);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc += ex2f(xl[r][i4] - Mn);
                    lS[i4] = acc;
                    lM[i4] = Mn;
                }
            }
#if !defined(TKB_EXP_NO_Q) && !defined(TKB_EXP_Q_NOWAIT)
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_s + qs * 8);
#endif
        }
        cp_async_wait_all();
        gi += mine;
        if (t == 0) TKB_STAMP(uord, 2);
        // ---- partial of the unit: transpose through the (drained) FIFO to [word][track][column], 128-byte stores ----
        asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");  // everyone is done reading the FIFO
        if (t < BX * nq) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const unsigned o = fifo_s + (unsigned)((4 * quad + i4) * 33 + col) * 4u;
                if (DO_V) {
                    sts32(o, vmax[i4]);
                    sts32(o + TRK * 33 * 4, __int_as_float(vsel[i4]));
                }
                if (DO_L) {
                    sts32(o + 2 * TRK * 33 * 4, lM[i4]);
                    sts32(o + 3 * TRK * 33 * 4, lS[i4]);
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
        {
            float *dst = p.part + ((size_t)(J * KCL + k) * 4) * p.Npad * BX;
            for (int idx = warp; idx < 4 * ntr; idx += NCONW) {
                const int w = idx / ntr, n = idx - w * ntr;
                if ((w < 2) ? !DO_V : !DO_L) continue;
                dst[((size_t)w * p.Npad + n_lo + n) * BX + lane] = lds32(fifo_s + (unsigned)((w * TRK + n) * 33 + lane) * 4u);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
        if (t == 0) {
            __threadfence();
            st_relaxed_u64(p.uflag + ((size_t)J * KCL + k) * p.nrange + rg, tagw);
            TKB_STAMP(uord, 3);
        }
    }
}

// =================================================================================================
// SOLVER: the chains of NQ tracks, from the last position to the first, in one SM
// =================================================================================================
// Warp roles of a solver CTA: chain warps (one per track and semiring, each alone on its SMSP for NQ = 2), loader
// warps (cp.async into the planar row bands), one prep warp (per-column constants and the far partial of the
// block two ahead, so the chain only ever reads shared memory).
//
// A chain warp keeps lane = column of the current 32-column block.  Columns are solved in micro-blocks of four:
// every lane gathers the four partial results with shuffles and solves the 4x4 triangle redundantly in its own
// registers (no communication on the dependent path), then pushes the four finished rows into its own column of
// the diagonal tile and of the ND tiles below it.  The Viterbi argmax lives only in the owner lane's push (the
// reference's candidate order); the redundant solve carries values only.
struct SolverSmem {
    unsigned band, prep, pub, band_full, band_empty, prep_full, prep_empty, pub_full, pub_empty;  // shared-window addresses
};
__device__ __forceinline__ SolverSmem solver_smem(unsigned char *smem_raw) {
    SolverSmem s;
    s.band = smem_u32(smem_raw);
    s.prep = s.band + (unsigned)(NBAND * kBandBytes);
    s.pub = s.prep + (unsigned)(NPREP * NQ * PR_FLOATS * 4);
    s.band_full = s.pub + (unsigned)(NCW * 2 * BX * 8);
    s.band_empty = s.band_full + NBAND * 8;
    s.prep_full = s.band_empty + NBAND * 8;
    s.prep_empty = s.prep_full + NPREP * 8;
    s.pub_full = s.prep_empty + NPREP * 8;
    s.pub_empty = s.pub_full + NCW * 16 * 8;
    return s;
}
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// ---- Viterbi chain of one track --------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void chain_viterbi(const SweepParams &p, unsigned char *smem_raw, const SolverSmem &sm,
                                              int g, int ptrk, int tr) {
    const int c = threadIdx.x & 31;
    const int T = p.T;
    const int nb = (T + BX - 1) / BX;
    (void)g;
    (void)ptrk;
    const float *bands = reinterpret_cast<const float *>(smem_raw);
    const float *preps = bands + (size_t)NBAND * (kBandBytes / 4);
    float best[ND + 1];
    int bsel[ND + 1];
#pragma unroll
    for (int d = 0; d <= ND; ++d) {
        best[d] = -INFINITY;
        bsel[d] = -1;
    }
    float qtop = 0.0f;
    const int cwi = 2 * tr;  // chain warp index
    const unsigned pub_s = sm.pub + (unsigned)(cwi * 2 * BX * 8), pubfull_s = sm.pub_full + cwi * 16 * 8;
    for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
        const int slot = it % NBAND, ps = it % NPREP;
        if (it >= 2) mbar_wait(sm.pub_empty + (cwi * 2 + (it & 1)) * 8, ((it >> 1) - 1) & 1, p.status);
        const int x0 = j * BX, x = x0 + c;
        const int ncols = min(BX, T - x0);
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 3);
        mbar_wait(sm.prep_full + ps * 8, (it / NPREP) & 1, p.status);
        const float *pr = preps + (size_t)(ps * NQ + tr) * PR_FLOATS;
        const float dr = pr[PR_DR + c], s_eta = pr[PR_ETA + c];
        if (j <= nb - ND - 2) {  // far partial (rows of blocks > j+ND): larger y than anything accumulated so far
            const float fv = pr[PR_FARV + c];
            const int fs = __float_as_int(pr[PR_FARS + c]);
            const bool tk = (DIR == TKB_BACKWARD) ? (fv > best[0]) : (fv >= best[0]);
            bsel[0] = tk ? fs : bsel[0];
            best[0] = fmaxf(best[0], fv);
        }
        if (j < nb - 1) {  // the skip out of the top column into row 32(j+1): candidate 0 of the reference
            const float xk = (c == BX - 1) ? qtop + s_eta : -INFINITY;
            bsel[0] = (xk >= best[0]) ? -1 : bsel[0];
            best[0] = fmaxf(best[0], xk);
        }
        if (x == T - 1) {  // terminal column: no candidates; q = S*(S>0) (-0 + dr keeps the reference's signed zero)
            best[0] = -0.0f;
            bsel[0] = -1;
        }
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 4);
        mbar_wait(sm.band_full + slot * 8, (it / NBAND) & 1, p.status);
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 5);
        const float *bnd = bands + (size_t)slot * (kBandBytes / 4) + (size_t)tr * (BX * BANDCOLS);  // [e][cc]
        const float *colp = bnd + c;
        auto micro = [&](const int k, const bool full) {
            float P[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) P[i] = __shfl_sync(kFull, best[0], 4 * k + i);
            const float4 U = *reinterpret_cast<const float4 *>(pr + PR_DR + 4 * k);
            const float4 E = *reinterpret_cast<const float4 *>(pr + PR_ETA + 4 * k);
            const float *mt = bnd + (4 * k) * BANDCOLS + ND * BX + 4 * k;  // micro-triangle: S[r][i], r > i
            const float4 r1 = *reinterpret_cast<const float4 *>(mt + BANDCOLS);
            const float4 r2 = *reinterpret_cast<const float4 *>(mt + 2 * BANDCOLS);
            const float4 r3 = *reinterpret_cast<const float4 *>(mt + 3 * BANDCOLS);
            float q[4];
            q[3] = P[3] + U.w;
            q[2] = max3f(P[2], q[3] + r3.z, q[3] + E.z) + U.z;
            q[1] = fmaxf(max3f(P[1], q[3] + r3.y, q[2] + r2.y), q[2] + E.y) + U.y;
            q[0] = max3f(max3f(P[0], q[3] + r3.x, q[2] + r2.x), q[1] + r1.x, q[1] + E.x) + U.x;
            if (k == 0) qtop = q[0];
#pragma unroll
            for (int r = 3; r >= 0; --r) {
                const int e = 4 * k + r, y = x0 + e;
                if (!full && e >= ncols) continue;
                const float *rowp = colp + e * BANDCOLS;
                const float qb = q[r];
                {
                    const float xi = (c < e) ? qb + rowp[ND * BX] : -INFINITY;
                    const float xk = (c == e - 1) ? qb + s_eta : -INFINITY;
                    const bool tk = (DIR == TKB_BACKWARD) ? (xi >= best[0]) : (xi > best[0]);
                    const float b1 = fmaxf(best[0], xi);
                    bsel[0] = tk ? y : bsel[0];
                    bsel[0] = (xk >= b1) ? -1 : bsel[0];
                    best[0] = fmaxf(b1, xk);
                }
#pragma unroll
                for (int d = 1; d <= ND; ++d) {
                    const float xi = qb + rowp[(ND - d) * BX];
                    const bool tk = (DIR == TKB_BACKWARD) ? (xi >= best[d]) : (xi > best[d]);
                    bsel[d] = tk ? y : bsel[d];
                    best[d] = fmaxf(best[d], xi);
                }
            }
            // my column is final: hand {q, code} to the publisher warp (global stores stall a chain warp)
            if ((c >> 2) == k) {
                const float qfin = best[0] + dr;
                const int osel = bsel[0] < 0 ? -1 : ((DIR == TKB_BACKWARD) ? bsel[0] : T - 1 - bsel[0]);
                const unsigned cw = ((unsigned)(osel + 1) << 1) | (dr > 0.0f ? 1u : 0u);
                sts64_nc(pub_s + (unsigned)(((it & 1) * BX + c) * 8), __float_as_uint(qfin), cw);
                mbar_arrive_nc(pubfull_s + (unsigned)(((it & 1) * 8 + k) * 8));
            }
        };
        if (ncols == BX) {
#pragma unroll
            for (int k = BX / 4 - 1; k >= 0; --k) micro(k, true);
        } else {  // the ragged top block (columns >= ncols hold -inf / zero-filled scores)
#pragma unroll 1
            for (int k = (ncols - 1) >> 2; k >= 0; --k) micro(k, false);
        }
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 7);
        __syncwarp();
        if (c == 0) {
            mbar_arrive(sm.band_empty + slot * 8);
            mbar_arrive(sm.prep_empty + ps * 8);
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            best[d] = best[d + 1];
            bsel[d] = bsel[d + 1];
        }
        best[ND] = -INFINITY;
        bsel[ND] = -1;
    }
}

// ---- log-sum chain of one track: every value is a pair (M, S) = M + log2(S) ---------------------------
template <int DIR>
__device__ __forceinline__ void chain_logsum(const SweepParams &p, unsigned char *smem_raw, const SolverSmem &sm,
                                             int g, int ptrk, int tr) {
    const int c = threadIdx.x & 31;
    const int T = p.T;
    const int nb = (T + BX - 1) / BX;
    (void)g;
    (void)ptrk;
    const float *bands = reinterpret_cast<const float *>(smem_raw);
    const float *preps = bands + (size_t)NBAND * (kBandBytes / 4);
    float lM[ND + 1], lS[ND + 1];
#pragma unroll
    for (int d = 0; d <= ND; ++d) {
        lM[d] = -FLT_MAX;
        lS[d] = 0.0f;
    }
    float qtopM = 0.0f, qtopS = 0.0f;
    const int cwi = 2 * tr + 1;  // chain warp index
    const unsigned pub_s = sm.pub + (unsigned)(cwi * 2 * BX * 8), pubfull_s = sm.pub_full + cwi * 16 * 8;
    for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
        const int slot = it % NBAND, ps = it % NPREP;
        if (it >= 2) mbar_wait(sm.pub_empty + (cwi * 2 + (it & 1)) * 8, ((it >> 1) - 1) & 1, p.status);
        const int x0 = j * BX, x = x0 + c;
        const int ncols = min(BX, T - x0);
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 0);
        mbar_wait(sm.prep_full + ps * 8, (it / NPREP) & 1, p.status);
        const float *pr = preps + (size_t)(ps * NQ + tr) * PR_FLOATS;
        const float sp2 = pr[PR_SP2 + c], comb = pr[PR_COMB + c], eta2 = pr[PR_ETA2 + c];
        if (j <= nb - ND - 2) lse_push(lM[0], lS[0], pr[PR_FARM + c], pr[PR_FARL + c]);
        if (j < nb - 1) lse_push(lM[0], lS[0], (c == BX - 1) ? qtopM + eta2 : -INFINITY, qtopS);
        if (x == T - 1) {
            lM[0] = 0.0f;
            lS[0] = 1.0f;
        }
        if (lS[0] > 0.0f) {  // renormalise: S restarts at 1 in every block
            lM[0] += lg2f(lS[0]);
            lS[0] = 1.0f;
        }
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 1);
        mbar_wait(sm.band_full + slot * 8, (it / NBAND) & 1, p.status);
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 2);
        const float *bnd = bands + (size_t)slot * (kBandBytes / 4) + (size_t)tr * (BX * BANDCOLS);
        const float *colp = bnd + c;
        // One micro-block.  The scale M of a pair never depends on any S: the M's solve a max-plus recursion of
        // their own (short dependent adds and maxes), every exponent is known from the M's alone, and the S's
        // follow with fused multiply-adds whose weights 2^(a - M) are all <= 1 (no overflow, no branch).
        auto micro = [&](const int k, const bool full) {
            (void)full;
            float M[4], S[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                M[i] = __shfl_sync(kFull, lM[0], 4 * k + i);
                S[i] = __shfl_sync(kFull, lS[0], 4 * k + i);
            }
            const float4 SP = *reinterpret_cast<const float4 *>(pr + PR_SP2 + 4 * k);
            const float4 CB = *reinterpret_cast<const float4 *>(pr + PR_COMB + 4 * k);
            const float *mt = bnd + (4 * k) * BANDCOLS + ND * BX + 4 * k;
            const float4 r2 = *reinterpret_cast<const float4 *>(mt + 2 * BANDCOLS);
            const float4 r3 = *reinterpret_cast<const float4 *>(mt + 3 * BANDCOLS);
            float Mb[4], Mn[4];
            Mb[3] = M[3] + SP.w;
            const float a32 = Mb[3] + CB.z;
            Mn[2] = fmaxf(M[2], a32);
            Mb[2] = Mn[2] + SP.z;
            const float a31 = fmaf(r3.y, kLog2e, Mb[3]), a21 = Mb[2] + CB.y;
            Mn[1] = max3f(M[1], a31, a21);
            Mb[1] = Mn[1] + SP.y;
            const float a30 = fmaf(r3.x, kLog2e, Mb[3]), a20 = fmaf(r2.x, kLog2e, Mb[2]), a10 = Mb[1] + CB.x;
            Mn[0] = fmaxf(max3f(M[0], a30, a20), a10);
            Mb[0] = Mn[0] + SP.x;
            S[2] = fmaf(S[3], ex2f(a32 - Mn[2]), S[2] * ex2f(M[2] - Mn[2]));
            S[1] = fmaf(S[2], ex2f(a21 - Mn[1]), fmaf(S[3], ex2f(a31 - Mn[1]), S[1] * ex2f(M[1] - Mn[1])));
            S[0] = fmaf(S[1], ex2f(a10 - Mn[0]),
                        fmaf(S[2], ex2f(a20 - Mn[0]), fmaf(S[3], ex2f(a30 - Mn[0]), S[0] * ex2f(M[0] - Mn[0]))));
            if (k == 0) {
                qtopM = Mb[0];
                qtopS = S[0];
            }
            // push the four finished rows into my column of every tile: one common scale per tile
#pragma unroll
            for (int d = 0; d <= ND; ++d) {
                float a[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int e = 4 * k + r;
                    const float sv = colp[e * BANDCOLS + (ND - d) * BX];
                    if (d == 0)
                        a[r] = Mb[r] + ((c == e - 1) ? comb : ((c < e) ? sv * kLog2e : -INFINITY));
                    else
                        a[r] = fmaf(sv, kLog2e, Mb[r]);
                }
                const float Mx = fmaxf(fmaxf(lM[d], a[0]), max3f(a[3], a[2], a[1]));
                float acc = lS[d] * ex2f(lM[d] - Mx);
#pragma unroll
                for (int r = 3; r >= 0; --r) acc = fmaf(S[r], ex2f(a[r] - Mx), acc);
                lS[d] = acc;
                lM[d] = Mx;
            }
            if ((c >> 2) == k) {
                sts64_nc(pub_s + (unsigned)(((it & 1) * BX + c) * 8), __float_as_uint(lM[0] + sp2), __float_as_uint(lS[0]));
                mbar_arrive_nc(pubfull_s + (unsigned)(((it & 1) * 8 + k) * 8));
            }
        };
        if (ncols == BX) {
#pragma unroll
            for (int k = BX / 4 - 1; k >= 0; --k) micro(k, true);
        } else {
#pragma unroll 1
            for (int k = (ncols - 1) >> 2; k >= 0; --k) micro(k, false);
        }
        if (c == 0 && tr == 0) TKB_CSTAMP(it, 6);
        __syncwarp();
        if (c == 0) {
            mbar_arrive(sm.band_empty + slot * 8);
            mbar_arrive(sm.prep_empty + ps * 8);
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            lM[d] = lM[d + 1];
            lS[d] = lS[d + 1];
        }
        lM[ND] = -FLT_MAX;
        lS[ND] = 0.0f;
    }
}

template <int DIR, int MODE>
__device__ __forceinline__ void solver_role(const SweepParams &p, unsigned char *smem_raw, int g, int qd) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    constexpr int NSEMI = (DO_V ? 1 : 0) + (DO_L ? 1 : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n0 = g * NG + qd * NQ;             // first track of this solver
    const int nvalid = min(max(N - n0, 0), NQ);  // tracks that exist
    const SolverSmem sm = solver_smem(smem_raw);
    const int rg = (n0 - p.nlo) / TRK;  // my track range

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAND; ++s) {
            mbar_init(sm.band_full + s * 8, NLW * 32);
            mbar_init(sm.band_empty + s * 8, nvalid * NSEMI);
        }
        for (int s = 0; s < NPREP; ++s) {
            mbar_init(sm.prep_full + s * 8, 32);
            mbar_init(sm.prep_empty + s * 8, nvalid * NSEMI);
        }
        for (int s = 0; s < NCW * 16; ++s) mbar_init(sm.pub_full + s * 8, 4);  // the four owner lanes arrive
        for (int s = 0; s < NCW; ++s) {
            mbar_init(sm.pub_empty + (s * 2) * 8, 1);
            mbar_init(sm.pub_empty + (s * 2 + 1) * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (nvalid == 0) return;

    if (warp < NCW) {
        // ---------------- chain warps: warp = 2*track + semiring ------------------------------------------
        const int tr = warp >> 1, kind = warp & 1;
        if (tr >= nvalid) return;
        if (kind == 0) {
            if (DO_V) chain_viterbi<DIR>(p, smem_raw, sm, g, qd * NQ + tr, tr);
        } else {
            if (DO_L) chain_logsum<DIR>(p, smem_raw, sm, g, qd * NQ + tr, tr);
        }
        return;
    }
    if (warp < NCW + NLW) {
        // ---------------- loader warps: keep the ring of row bands filled ----------------------------------
        // band of row block j: rows y = 32j .. 32j+31, columns x = 32(j-ND) .. 32j+31.  The strip CTAs have re-laid
        // it out per track in the global ring (p.scratch), so a track's band is one contiguous run: coalesced
        // 16-byte cp.async instead of a 32-byte-sector gather from the score tensor (whose L1 wavefronts used to
        // starve the chain warps of the same SM).
        const int lw = warp - NCW;
        constexpr int RUN16 = BANDCOLS / 4;  // 16-byte pieces per (track, row) run
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int slot = it % NBAND;
            if (it >= NBAND) mbar_wait_relaxed(sm.band_empty + slot * 8, ((it / NBAND) - 1) & 1, p.status);
            {   // the (ND+1) x KCL strip units that fill this band: flags {btag | band + 1, epoch}
                const unsigned long long want = ((unsigned long long)p.epoch << 32) | p.btag | (unsigned)(j + 1);
                const int nfl = min(ND, j) + 1;  // tiles left of column 0 do not exist
                const unsigned long long *fl = p.bflag + ((size_t)rg * RB + (j % RB)) * (ND + 1) * KCL + lane;
                for (int f = lane; f < nfl * KCL; f += 32)
                    if (ld_acquire_u64(fl + (f - lane)) != want) poll_flag_slow(fl + (f - lane), want, p.status);
                __syncwarp();
            }
            const float *src0 = p.scratch + ((size_t)(j % RB) * p.SN + (n0 - p.nlo)) * BX * BANDCOLS;
            const unsigned dst0 = sm.band + (unsigned)(slot * kBandBytes);
            // a track's band is one contiguous run of BX * BANDCOLS floats in the ring
            for (int i = lw * 32 + lane; i < nvalid * BX * RUN16; i += NLW * 32)
                cp_async16_s(dst0 + (unsigned)i * 16u, src0 + (size_t)i * 4, 16);
            mbar_arrive_cp_async(sm.band_full + slot * 8);
        }
        cp_async_wait_all();
        return;
    }
    if (warp == NCW + NLW) {
        // ---------------- prep warp: per-column constants + far partial of block j, NPREP blocks ahead ----------
        float *preps = reinterpret_cast<float *>(smem_raw) + (size_t)NBAND * (kBandBytes / 4);
        const int c = lane;
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int ps = it % NPREP;
            const int x0 = j * BX, x = x0 + c;
            const int ncols = min(BX, T - x0);
            // issue every global load of this block before waiting on anything
            float sd[NQ], se[NQ], ssub[NQ];
            const bool has_far = j <= nb - ND - 2;
#pragma unroll
            for (int tr = 0; tr < NQ; ++tr) {
                sd[tr] = se[tr] = ssub[tr] = 0.0f;
                if (tr < nvalid && x < T) {
                    const int n = n0 + tr;
                    sd[tr] = __ldg(p.Sbase + (long long)x * (p.sx + p.sy) + n);
                    if (x < T - 1) {
                        se[tr] = __ldg(p.etabase + (long long)x * p.se + n);
                        if (DO_L) ssub[tr] = __ldg(p.Sbase + (long long)x * p.sx + (long long)(x + 1) * p.sy + n);
                    }
                }
            }
            // far partial of block j: KCL units of my range, merged here.  Their flags first (acquire), then the data.
            float fV[NQ], fM[NQ], fS[NQ];
            int fsel[NQ];
#pragma unroll
            for (int tr = 0; tr < NQ; ++tr) {
                fV[tr] = -INFINITY;
                fsel[tr] = -1;
                fM[tr] = -FLT_MAX;
                fS[tr] = 0.0f;
            }
            if (has_far) {
                const unsigned long long want = ((unsigned long long)p.epoch << 32) | p.btag;
                const unsigned long long *fl = p.uflag + ((size_t)j * KCL + lane) * p.nrange + rg;
                if (lane < KCL && ld_acquire_u64(fl) != want) poll_flag_slow(fl, want, p.status);
                __syncwarp();
#pragma unroll
                for (int tr = 0; tr < NQ; ++tr)
                    if (tr < nvalid) {
                        const float *src = p.part + ((size_t)(j * KCL) * 4 * p.Npad + (n0 + tr)) * BX + c;
                        const size_t ws = (size_t)p.Npad * BX, ks = 4 * ws;  // word and unit strides
                        if (DO_V) {
                            float pv[KCL];
                            int psel[KCL];
#pragma unroll
                            for (int k = 0; k < KCL; ++k) {
                                pv[k] = ld_cg_f32(src + k * ks);
                                psel[k] = __float_as_int(ld_cg_f32(src + k * ks + ws));
                            }
                            // the maximum, then among the units that attain it the row the reference's candidate
                            // order prefers (BACKWARD: smallest y, FORWARD: largest y); empty partials are (-inf, -1)
                            float best = pv[0];
#pragma unroll
                            for (int k = 1; k < KCL; ++k) best = fmaxf(best, pv[k]);
                            if (DIR == TKB_BACKWARD) {
                                unsigned m = 0xffffffffu;
#pragma unroll
                                for (int k = 0; k < KCL; ++k) m = min(m, pv[k] == best ? (unsigned)psel[k] : 0xffffffffu);
                                fsel[tr] = (int)m;
                            } else {
                                int m = -1;
#pragma unroll
                                for (int k = 0; k < KCL; ++k) m = max(m, pv[k] == best ? psel[k] : -1);
                                fsel[tr] = m;
                            }
                            fV[tr] = best;
                        }
                        if (DO_L) {
                            float pm[KCL], pS[KCL];
                            float M = -FLT_MAX;
#pragma unroll
                            for (int k = 0; k < KCL; ++k) {
                                pm[k] = ld_cg_f32(src + k * ks + 2 * ws);
                                pS[k] = ld_cg_f32(src + k * ks + 3 * ws);
                                M = fmaxf(M, pm[k]);
                            }
                            float S = 0.0f;
#pragma unroll
                            for (int k = 0; k < KCL; ++k) S = fmaf(pS[k], ex2f(pm[k] - M), S);
                            fM[tr] = M;
                            fS[tr] = S;
                        }
                    }
            }
            if (it >= NPREP) mbar_wait_relaxed(sm.prep_empty + ps * 8, ((it / NPREP) - 1) & 1, p.status);
#pragma unroll
            for (int tr = 0; tr < NQ; ++tr)
                if (tr < nvalid) {
                    float *pr = preps + (size_t)(ps * NQ + tr) * PR_FLOATS;
                    if (has_far) {
                        if (DO_V) {
                            pr[PR_FARV + c] = fV[tr];
                            pr[PR_FARS + c] = __int_as_float(fsel[tr]);
                        }
                        if (DO_L) {
                            pr[PR_FARM + c] = fM[tr];
                            pr[PR_FARL + c] = fS[tr];
                        }
                    }
                    if (DO_V) {
                        pr[PR_DR + c] = relu_mask(sd[tr]);
                        pr[PR_ETA + c] = se[tr];
                    }
                    if (DO_L) {
                        const float d2 = sd[tr] * kLog2e;
                        const float e2 = se[tr] * kLog2e;
                        pr[PR_SP2 + c] = (x < T) ? fmaxf(d2, 0.0f) + lg2f(1.0f + ex2f(-fabsf(d2))) : 0.0f;
                        pr[PR_ETA2 + c] = e2;
                        float comb = -INFINITY;  // row x+1 into column x: its score and the skip, folded
                        if (c + 1 < ncols) {
                            const float spv = ssub[tr] * kLog2e;
                            comb = fmaxf(spv, e2) + lg2f(1.0f + ex2f(-fabsf(spv - e2)));
                        }
                        pr[PR_COMB + c] = comb;
                    }
                }
            mbar_arrive(sm.prep_full + ps * 8);
        }
        return;
    }
    if (warp > NCW + NLW && warp <= NCW + NLW + NCW) {
        // ---------------- publisher warps: one per chain warp; results go shared memory -> mailbox and tables -----
        const int cwi = warp - (NCW + NLW + 1);
        const int tr = cwi >> 1, kind = cwi & 1;
        if (tr >= nvalid || (kind == 0 ? !DO_V : !DO_L)) return;
        const int n = n0 + tr;
        const unsigned pub_s = sm.pub + (unsigned)(cwi * 2 * BX * 8), pubfull_s = sm.pub_full + cwi * 16 * 8;
        float *qt = p.qtab + (size_t)kind * p.Npad + n;  // + y * 2 * Npad
        const int kmax_top = (T - (nb - 1) * BX - 1) >> 2;
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int x0 = j * BX;
            const int ncols = min(BX, T - x0);
            // micro-blocks the ragged top block skips never arrive: their barriers start one phase behind
            for (int k = (ncols - 1) >> 2; k >= 0; --k) {
                const unsigned npast = (unsigned)(it >> 1) - (((it & 1) == 0 && it > 0 && k > kmax_top) ? 1u : 0u);
                mbar_wait(pubfull_s + (unsigned)(((it & 1) * 8 + k) * 8), npast & 1, p.status);
                const int c = 4 * k + lane, x = x0 + c;
                if (lane < 4 && x < T) {
                    const unsigned long long w = lds64(pub_s + (unsigned)(((it & 1) * BX + c) * 8));
                    const int pos = (DIR == TKB_BACKWARD) ? x : T - 1 - x;
                    if (kind == 0) {
                        const float qfin = __uint_as_float((unsigned)w);
                        qt[(size_t)x * 2 * p.Npad] = qfin;
                        p.code[(size_t)n * T + pos] = (unsigned)(w >> 32);
                        if (p.outv) p.outv[(size_t)pos * N + n] = qfin;
                    } else {
                        const float v2 = __uint_as_float((unsigned)w) + lg2f(__uint_as_float((unsigned)(w >> 32)));
                        qt[(size_t)x * 2 * p.Npad] = v2;
                        if (p.outl) p.outl[(size_t)pos * N + n] = v2 * kLn2;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(sm.pub_empty + (cwi * 2 + (it & 1)) * 8);
                __threadfence();  // the block's rows are visible before its flag
                st_relaxed_u32(p.bdone + ((size_t)j * 2 + kind) * p.Npad + n, p.epoch);
            }
        }
        return;
    }
}

template <int DIR, int ALIGN, int MODE>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if ((int)blockIdx.x < p.nsolv) {
#ifndef TKB_EXP_STRIP_ONLY
        solver_role<DIR, MODE>(p, smem_raw, p.g0 + (int)blockIdx.x / NSOLV, (int)blockIdx.x % NSOLV);
#endif
    } else {
#ifndef TKB_EXP_SOLVER_ONLY
        strip_role<DIR, ALIGN, MODE>(p, smem_raw, (int)blockIdx.x - p.nsolv, (int)gridDim.x - p.nsolv);
#endif
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int DIR, int ALIGN, int MODE>
static int launch_one(const SweepParams &p, int grid, cudaStream_t stream) {
    auto kern = sweep_kernel<DIR, ALIGN, MODE>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        TKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
        configured = true;
    }
    SweepParams pp = p;
    void *args[] = {&pp};
    TKB_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(NT), args, kSweepSmem, stream));
    return 0;
}

template <int DIR, int ALIGN>
static int launch_mode(int mode, const SweepParams &p, int grid, cudaStream_t stream) {
    switch (mode) {
        case TKB_SWEEP_VITERBI: return launch_one<DIR, ALIGN, TKB_SWEEP_VITERBI>(p, grid, stream);
        case TKB_SWEEP_LOGSUM: return launch_one<DIR, ALIGN, TKB_SWEEP_LOGSUM>(p, grid, stream);
        default: return launch_one<DIR, ALIGN, TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM>(p, grid, stream);
    }
}

template <int DIR>
static int launch_align(int align, int mode, const SweepParams &p, int grid, cudaStream_t stream) {
    switch (align) {
        case 16: return launch_mode<DIR, 16>(mode, p, grid, stream);
        case 8: return launch_mode<DIR, 8>(mode, p, grid, stream);
        default: return launch_mode<DIR, 4>(mode, p, grid, stream);
    }
}

static unsigned long long *g_timeline = nullptr;  // diagnostics build only
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t rowtable_bytes(int T, int N) {  // solved rows [T][2][Npad] fp32 + block flags [nb][2][Npad] u32
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG, nb = (size_t)((T + BX - 1) / BX);
    return align256(2 * (size_t)T * npad * sizeof(float)) + align256(2 * nb * npad * sizeof(unsigned));
}
static_assert(true, "");
constexpr int kGroupsPerLaunch = 2 * TRK / NG;  // two track ranges: 4 solver CTAs per group, the other SMs run strips
constexpr int kRangesPerLaunch = kGroupsPerLaunch * NG / TRK;
static size_t partial_bytes(int T, int N) {
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG, nb = (size_t)((T + BX - 1) / BX);
    return align256(nb * KCL * 4 * npad * BX * sizeof(float));
}
static size_t uflag_bytes(int T) {
    const size_t nb = (size_t)((T + BX - 1) / BX);
    return align256(nb * KCL * kRangesPerLaunch * sizeof(unsigned long long));
}
static size_t bflag_bytes() { return align256((size_t)kRangesPerLaunch * RB * (ND + 1) * KCL * sizeof(unsigned long long)); }
static size_t scratch_bytes(int N) {
    const int G = (N + NG - 1) / NG;
    const size_t sn = (size_t)(G < kGroupsPerLaunch ? G : kGroupsPerLaunch) * NG;
    return (size_t)RB * sn * BX * BANDCOLS * sizeof(float);
}


size_t workspace_bytes(int T, int N) {
    if (T < 1 || N < 1) return 0;
    return kHeaderBytes + rowtable_bytes(T, N) + partial_bytes(T, N) + uflag_bytes(T) + bflag_bytes() + scratch_bytes(N);
}

int sweep(const float *score, long long pitch, const float *noise, int T, int N, int direction, int flags,
          void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit, float *out_lse, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!score || !workspace || T < 1 || N < 1 || (T > 1 && !noise) || epoch == 0 ||
        (direction != TKB_BACKWARD && direction != TKB_FORWARD) ||
        (flags & ~(TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM)) || flags == 0 ||
        ((flags & TKB_SWEEP_VITERBI) && !out_code) || (long long)T * T >= (1ll << 40) || pitch < N) {
        set_error("tkb_semicrf_sweep[strip]: invalid argument (T=%d N=%d dir=%d flags=%d epoch=%u)", T, N, direction,
                  flags, epoch);
        return TKB_EINVAL;
    }
    const int sms = num_sms();
    if (sms < NSOLV + 1) {
        set_error("tkb_semicrf_sweep[strip]: no CUDA device");
        return TKB_ENODEV;
    }
    SweepParams p;
    p.T = T;
    p.N = N;
    const int G = (N + NG - 1) / NG;
    p.Npad = G * NG;
    p.dir = direction;
    p.epoch = epoch;
    p.status = reinterpret_cast<int *>(workspace);
    p.qtab = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + kHeaderBytes);
    p.bdone = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(p.qtab) + align256(2 * (size_t)T * p.Npad * sizeof(float)));
    char *wp = reinterpret_cast<char *>(workspace) + kHeaderBytes + rowtable_bytes(T, N);
    p.part = reinterpret_cast<float *>(wp);
    wp += partial_bytes(T, N);
    p.uflag = reinterpret_cast<unsigned long long *>(wp);
    wp += uflag_bytes(T);
    p.bflag = reinterpret_cast<unsigned long long *>(wp);
    wp += bflag_bytes();
    p.scratch = reinterpret_cast<float *>(wp);
    p.code = out_code;
    p.outv = out_vit;
    p.outl = out_lse;
    p.timeline = g_timeline;
    if (direction == TKB_BACKWARD) {
        p.Sbase = score;
        p.sx = pitch;
        p.sy = (long long)T * pitch;
        p.etabase = noise;
        p.se = N;
    } else {
        p.Sbase = score + ((long long)(T - 1) * T + (T - 1)) * pitch;
        p.sx = -(long long)T * pitch;
        p.sy = -(long long)pitch;
        p.etabase = noise ? noise + (long long)(T - 2) * N : nullptr;  // skip weight of x is noise[T-2-x]
        p.se = -(long long)N;
    }
    const uintptr_t addr = reinterpret_cast<uintptr_t>(score);
    // the track pitch, not N, decides the copy width: a padded score tensor (pitch % 4 == 0) takes the 16-byte path
    const int align = (pitch % 4 == 0 && (addr & 15) == 0) ? 16 : ((pitch % 2 == 0 && (addr & 7) == 0) ? 8 : 4);
    // groups are independent pipelines; a launch holds up to two track ranges, its other SMs run strip CTAs
    // (all CTAs of a launch must be co-resident: one per SM)
    p.SN = (G < kGroupsPerLaunch ? G : kGroupsPerLaunch) * NG;
    int launch = 0;
    for (int g0 = 0; g0 < G; g0 += kGroupsPerLaunch, ++launch) {
        const int gcount = (G - g0) < kGroupsPerLaunch ? (G - g0) : kGroupsPerLaunch;
        p.g0 = g0;
        p.nlo = g0 * NG;
        p.nhi = (g0 + gcount) * NG < N ? (g0 + gcount) * NG : N;
        p.nrange = (p.nhi - p.nlo + TRK - 1) / TRK;
        p.nsolv = ((p.nhi - p.nlo + NQ - 1) / NQ);  // solver CTAs with at least one real track
        p.nsolv = ((p.nsolv + NSOLV - 1) / NSOLV) * NSOLV;
        p.btag = (unsigned)launch << 16;
        int nstrip = sms - p.nsolv;
        if (nstrip < 1) {
            set_error("tkb_semicrf_sweep[strip]: device too small (%d SMs)", sms);
            return TKB_ENODEV;
        }
        const int grid = p.nsolv + nstrip;
        const int rc = direction == TKB_BACKWARD ? launch_align<TKB_BACKWARD>(align, flags, p, grid, stream)
                                                 : launch_align<TKB_FORWARD>(align, flags, p, grid, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

void set_timeline(unsigned long long *buf) { g_timeline = buf; }

}  // namespace strip
}  // namespace tkb
