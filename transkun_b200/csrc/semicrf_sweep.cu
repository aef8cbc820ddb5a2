// semicrf_sweep.cu -- the semi-Markov dynamic programme as ONE persistent kernel (solver CTAs + streaming CTAs).
//
// Replaces the TorchScript step loops of the reference
// (transkun/CRF/NeuralSemiCRFInterval.py:31-51, :124-144, :218-234, :303-327):
//     q[x] = ( skip(x)  (+)  (+)_{y>x} q[y] (x) S(y,x) )  (x)  unary(x)
// over the (max,+) semiring (Viterbi, bit-exact fp32: one add per candidate, exact max, the
// reference's tie order) and the (logsumexp,+) semiring (log-partition), both fed by a single
// read of the score triangle.
//
// Mirrored coordinates.  x is the position being solved, y > x a solved one.
//   BACKWARD: x = begin b, y = end e, S(y,x) = score[e][b]      (sx = P,    sy = T*P)
//   FORWARD : x = T-1-end, y = T-1-begin, S(y,x) = score[T-1-x][T-1-y]
//                                                               (sx = -T*P, sy = -P)
// so one kernel serves viterbiBackward/beta and viterbi/alpha.
//
// It is a lower-triangular solve: T strictly sequential steps per track.  The design:
//   * SOLVER CTAs (one per 4 tracks = 16 bytes of the track-innermost layout) keep the chain inside one SM from
//     the last position to the first: one warp per track and semiring, lane = column of the current 32-column
//     block.  A chain step broadcasts the just-finished row with shuffles and pushes it into the current block
//     (the diagonal tile, on the chain) and into the next ND blocks (off the chain).  The score values come from a
//     shared-memory ring of "row bands" (32 rows x (ND+1)*32 columns x 16 B).  In the BACKWARD direction with a
//     16-byte aligned track pitch ONE thread fills a band with ONE TMA tensor copy
//     (cp.async.bulk.tensor.3d, box {4 tracks, 96 columns, 32 rows}, mbarrier complete_tx): the chain's own SM
//     does no address generation for it.  (The per-lane cp.async gather of round 1 kept the SM's L1 pipe busy for
//     ~6000 of the ~7000 cycles of a block and slowed the chain 2x; it remains for FORWARD / unaligned inputs.)
//     Finished rows leave through a shared-memory ring to one publisher warp per track, which does the global
//     stores (mailbox, back-pointer codes, tables).
//   * STREAMING CTAs (all remaining SMs) own the columns round-robin in chunks of two (column x -> CTA
//     (x/2) mod H) over ALL tracks of the launch, so the bytes a CTA reads of one row are contiguous runs of
//     2 x N x 4 bytes.  A thread owns (column, 4 tracks): Viterbi max/argmax and log-sum (M,S) accumulators live in
//     its registers from row T-1 down to the last far row of its column block; nothing is merged across threads.
//     Rows are consumed in lock-step with the chain in batches of 8: two fetch warps validate the mailbox words of
//     a batch once per CTA and put the plain values into shared memory; the score rows were prefetched with
//     cp.async (16 B per thread and row, 3 batches ahead -- they do not depend on the chain).  When a column block's
//     far field is complete its 32 columns (16 CTAs) publish the "far partial" the solver merges in.
//   * rows travel solver -> streaming CTAs through a global-memory mailbox of 64-bit words {fp32 value, epoch},
//     far partials travel back the same way: one relaxed store publishes, one relaxed load observes (no fence, no
//     flag, no reset; the epoch grows with every launch).
// All CTAs of a launch must be co-resident (cooperative launch).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace tkb {

constexpr int NQ = 4;      // tracks per solver CTA
constexpr int BX = 32;     // columns per block (= lanes of a chain warp)
#ifndef TKB_ND
#define TKB_ND 2
#endif
constexpr int ND = TKB_ND;  // blocks above the diagonal block that the solver pushes itself
constexpr int NBAND = (ND == 2) ? 4 : 3;  // row bands resident in a solver CTA
constexpr int BANDCOLS = (ND + 1) * BX;
constexpr int NW = 16;     // warps per CTA
constexpr int NT = NW * 32;
constexpr int NCW = NQ;    // chain warps per semiring
constexpr int NLW = 4;     // loader warps (cp.async path); the first one hosts the TMA thread
constexpr int PB = 8;      // rows per publish batch = rows per streaming batch
constexpr int NREP = 4;    // replicas of the row mailbox: streaming CTA h reads replica h % NREP (see publisher warps)
#ifndef TKB_FARFETCH_BATCH
#define TKB_FARFETCH_BATCH 3
#endif

// ---- streaming CTA geometry ----
constexpr int CW = 2;                  // columns per chunk
constexpr int NPW = 4;                 // mailbox producer warps: (semiring, half of a batch's rows): the last four warps
constexpr int NCONSW = NW - NPW;       // consumer warps
constexpr int NCONS = NCONSW * 32;     // consumer threads
constexpr int IPT = 2;                 // items (column, 4 tracks) per consumer thread, at most
constexpr int QB = 4;                  // batches of validated mailbox rows resident (plain floats); a power of two
#ifndef TKB_QS
#define TKB_QS 3
#endif
constexpr int QS = TKB_QS;             // batches of raw mailbox words in flight (bulk copies)
constexpr size_t kBarBytes = 256;
constexpr int NLMAX = 128;             // tracks per launch, at most
constexpr int MAXSTG = 4;              // score batches in flight, at most
constexpr size_t kSmemMax = 227 * 1024;
// streaming CTA shared memory: barriers | qbuf [QB batches][row][semiring][NLP] floats | raw mailbox words | score FIFO
// [stage][row][NI] 16 bytes
__host__ __device__ constexpr size_t qbuf_bytes(int nlp) { return (size_t)QB * PB * 2 * nlp * 4; }
__host__ __device__ constexpr size_t qraw_bytes(int nlp) { return (size_t)NPW * QS * 4 * nlp * 8; }  // [warp][stage][row][NLP]
__host__ __device__ constexpr size_t fifo_budget(int nlp) { return kSmemMax - kBarBytes - qbuf_bytes(nlp) - qraw_bytes(nlp); }
static_assert(2 * QB * 8 + NPW * QS * 8 <= kBarBytes, "barrier area");
static_assert(NREP * PB == 32, "one publisher store instruction writes a batch to every replica");

// solver shared memory: row bands [NBAND][BX rows][BANDCOLS][NQ tracks] | mbarriers full[NBAND], empty[NBAND]
constexpr size_t kBandBytes = (size_t)BX * BANDCOLS * NQ * 4;
// | publish ring [NCW tracks][2 blocks][BX][2 semirings] 8-byte results | mbarriers pub full[NCW][2 blocks][BX/PB]
// (eight arrivals each, one outstanding phase), pub empty[NCW][2]
constexpr size_t kPubBytes = (size_t)NCW * 2 * BX * 2 * 8;
constexpr size_t kSolverSmem =
    (size_t)NBAND * kBandBytes + 2 * NBAND * 8 + kPubBytes + NCW * 2 * (BX / PB) * 8 + NCW * 2 * 8;
constexpr size_t kSweepSmem = kSmemMax;
static_assert(kSolverSmem <= kSmemMax, "shared memory budget");
static_assert(kBandBytes % 128 == 0, "TMA destination alignment");

constexpr size_t kHeaderBytes = 256;  // status word lives here

struct SweepParams {
    const float *Sbase;    // &S(0,0) in mirrored coordinates
    const float *etabase;  // &skip weight of x = 0
    long long sx, sy, se;  // element strides
    int T, N, Npad, dir;
    int n_lo, Nl;          // tracks of this launch
    int S, H;              // solver CTAs, streaming CTAs
    int nslots, nitems, NI, nstg;  // streaming geometry (host-computed)
    unsigned epoch;
    unsigned long long *mbox;  // [NREP replicas][2 semirings][T][Npad] {value, epoch}
    unsigned long long *part;  // [nb][2 semirings][Npad][BX][2] {value, epoch}: far partials
    int *status;               // 0, or the epoch of a launch whose inter-CTA wait timed out
    unsigned *code;  // [N][T]
    float *outv;     // [T][N] or null
    float *outl;     // [T][N] or null
    unsigned long long *timeline;  // diagnostics build only (TKB_TIMELINE): [grid][256][8] globaltimer stamps
    int dbg;                       // diagnostics build only: 1 = streaming CTAs skip the arithmetic, 2 = skip the score prefetch
};

// Wait until a mailbox word carries this launch's epoch.  A protocol bug (or a non-co-resident grid) must not
// hang the GPU: after ~4 s the wait gives up, flags the workspace with this launch's epoch and lets the kernel
// drain with garbage.  BACKOFF_NS > 0 is for waits that are NOT close to a deadline.
template <int BACKOFF_NS>
__device__ __noinline__ unsigned long long poll_slow(const unsigned long long *w, unsigned epoch, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            unsigned long long v = ld_relaxed_u64(w);
            if ((unsigned)(v >> 32) == epoch) return v;
            if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
        }
        if (*(volatile unsigned *)status == epoch) return 0;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, (int)epoch);
            return 0;
        }
    }
}
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
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrival that fires once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead
// of spinning -- a spinning publisher or consumer warp steals issue slots from the chain warp on its SMSP
// (ncu, round 2: two thirds of all executed instructions were try_wait spins)
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
// non-blocking test (no hardware suspend): for waits whose wake-up latency matters
__device__ __forceinline__ bool mbar_test_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_spin_slow(unsigned bar, unsigned parity, int *status, unsigned epoch) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 4096; ++i)
            if (mbar_test_wait(bar, parity)) return;
        if (*(volatile unsigned *)status == epoch) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, (int)epoch);
            return;
        }
    }
}
// same watchdog as poll_slow: a protocol bug must drain the kernel, not hang the GPU
__device__ __noinline__ void mbar_wait_slow(unsigned bar, unsigned parity, int *status, unsigned epoch) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (*(volatile unsigned *)status == epoch) return;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, (int)epoch);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity, int *status, unsigned epoch) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, status, epoch);
}
__device__ __forceinline__ void mbar_spin(unsigned bar, unsigned parity, int *status, unsigned epoch) {
#ifdef TKB_SPIN
    for (int i = 0; i < 64; ++i)
        if (mbar_test_wait(bar, parity)) return;
    mbar_spin_slow(bar, parity, status, epoch);
#else
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, status, epoch);
#endif
}
// one band: box {NQ tracks, BANDCOLS columns, BX rows} of the [T][T][N] tensor -> [row][column][track]
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gsrc), "r"(src_bytes) : "memory");
}
// Chain -> publisher hand-off.  No "memory" clobber on purpose: volatile asm statements keep their mutual order
// (store, then arrive with release semantics, both by the same lane), while the compiler stays free to move the
// band reads of the following steps across them.
__device__ __forceinline__ void sts64_nc(unsigned saddr, unsigned lo, unsigned hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(lo), "r"(hi));
}
__device__ __forceinline__ void mbar_arrive_nc(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar));
}

#ifdef TKB_TIMELINE
// [grid][256 batches / chain blocks][4] stamps
#define TKB_STAMP(idx, slot)                                                                    \
    do {                                                                                        \
        if (p.timeline && !(p.dbg & 4) && (idx) < 256) p.timeline[((size_t)blockIdx.x * 256 + (idx)) * 8 + (slot)] = globaltimer_ns(); \
    } while (0)
#else
#define TKB_STAMP(idx, slot) \
    do {                     \
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
// STREAMING CTA, mailbox producer warp: (semiring `kind`, half `hf` of every batch's 8 rows)
// =================================================================================================
// The {value, epoch} words of a batch travel mailbox -> shared memory with bulk copies (cp.async.bulk, one per row, QS
// batches deep, completion on an mbarrier): no register is waiting for an L2 round trip -- long-latency loads of
// several batches in flight alias on the six scoreboard slots of a warp, which serialised the register version of
// this loop at one L2 round trip per batch.  The warp then validates the words in shared memory (they must carry this
// launch's epoch) and stores the plain floats for the consumers.  While a word is missing, ONE lane sleeps on the row
// of this half that is solved last and the rows are copied again only then (hundreds of warps re-reading the lines
// the publishers are writing would slow the writers down).
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
template <int NKIND>
__device__ __noinline__ void mailbox_producer(const SweepParams &p, int h, int pw, int kind, int hf, int k, int btop,
                                              int bmin, unsigned qbuf_s, unsigned qraw_s, unsigned qfull_s,
                                              unsigned qempty_s, unsigned rawfull_s) {
#ifdef TKB_TIMELINE
    if (p.dbg & 16) return;
#endif
    const int lane = threadIdx.x & 31;
    const int T = p.T, Nl = p.Nl;
    const unsigned epoch = p.epoch;
    const int NLP = (Nl + 3) & ~3;
    const unsigned qbuf_stride = (unsigned)(PB * NKIND * NLP * 4);
    const unsigned long long *mb = p.mbox + ((size_t)(h % NREP) * 2 + kind) * T * p.Npad + p.n_lo;
    const unsigned raw0 = qraw_s + (unsigned)(pw * QS * 4 * NLP * 8);   // [stage][row][NLP] words of this warp
    const unsigned bar0 = rawfull_s + (unsigned)(pw * QS * 8);
    const unsigned row_bytes = (unsigned)(((Nl + 1) & ~1) * 8);
    auto raw_issue = [&](int bb, int stage) {   // rows 4*hf .. 4*hf+3 of batch bb (lanes 0..3 copy one row each)
        if (bb < bmin) return;
        const int y0 = bb * PB + 4 * hf;
        const int nrows = min(4, T - y0);
        if (nrows <= 0) return;
        // (the stage was last READ by this warp, before the __syncwarp of the caller: no proxy fence needed for
        // the write-after-read; ~300 cycles per batch with one)
        if (lane == 0) mbar_arrive_expect_tx(bar0 + stage * 8, row_bytes * nrows);
        __syncwarp();
        if (lane < nrows)
            bulk_g2s(raw0 + (unsigned)((stage * 4 + lane) * NLP * 8), mb + (size_t)(y0 + lane) * p.Npad, row_bytes,
                     bar0 + stage * 8);
    };
    for (int s = 0; s < QS - 1; ++s) raw_issue(btop - s, s);
    unsigned ph = 0;   // bit s: parity of the phase stage s completes next
    int it = 0;
#ifdef TKB_TIMELINE
    long long ck[6] = {0, 0, 0, 0, 0, 0};
    long long c0 = clock64(), c1;
#define TKB_PK(k) do { c1 = clock64(); ck[k] += c1 - c0; c0 = c1; } while (0)
#else
#define TKB_PK(k) do {} while (0)
#endif
    for (int b = btop; b >= bmin; --b, ++it) {
        const int stage = it % QS, buf = it % QB;
        __syncwarp();  // every lane is done with the stage the next copies overwrite
        if (threadIdx.x == NCONS) TKB_STAMP(it, 4);
        TKB_PK(0);
        raw_issue(b - (QS - 1), (it + QS - 1) % QS);
        TKB_PK(1);
        const int y0 = b * PB + 4 * hf;
        const int nrows = min(4, T - y0);
#ifdef TKB_TIMELINE
        if (!(p.dbg & 8))
#endif
        if (it >= QB) mbar_spin(qempty_s + buf * 8, ((it / QB) - 1) & 1, p.status, epoch);
        TKB_PK(2);
        if (nrows > 0) {
            mbar_spin(bar0 + stage * 8, (ph >> stage) & 1, p.status, epoch);
            ph ^= 1u << stage;
            TKB_PK(3);
            if (threadIdx.x == NCONS) TKB_STAMP(it, 5);
            const unsigned st = raw0 + (unsigned)(stage * 4 * NLP * 8);
            const unsigned dstb = qbuf_s + (unsigned)buf * qbuf_stride;
            unsigned long long t0 = 0;
            for (unsigned tries = 0;; ++tries) {
                bool ok = true;
                // all loads first (their latencies overlap), then the checks and the stores
                unsigned long long w[4][4];
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    const int n = lane + 32 * jn;
                    // (reading a row of the stage that was not copied is harmless: the value is not used)
                    if (n < Nl) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) w[jn][r] = lds64(st + (unsigned)((r * NLP + n) * 8));
                    }
                }
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    const int n = lane + 32 * jn;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (n < Nl && r < nrows) {
                            if ((unsigned)(w[jn][r] >> 32) == epoch)
                                sts32(dstb + (unsigned)((((4 * hf + r) * NKIND + k) * NLP + n) * 4),
                                      __uint_as_float((unsigned)w[jn][r]));
                            else
                                ok = false;
                        }
                }
                if (__all_sync(kFull, ok)) break;
                if (threadIdx.x == NCONS) TKB_STAMP(it, 6);
                // the copy ran ahead of the chain: wait for the last-solved row of this half, then copy it again
                if (lane == 0) {
                    const unsigned long long *cw = mb + (size_t)y0 * p.Npad + (h % Nl);
                    for (int kk = 0; kk < 64 && (unsigned)(ld_relaxed_u64(cw) >> 32) != epoch; ++kk) __nanosleep(100);
                }
                bool dead = false;
                if ((tries & 15) == 15) {   // watchdog
                    if (t0 == 0) t0 = globaltimer_ns();
                    dead = *(volatile unsigned *)p.status == epoch || globaltimer_ns() - t0 > 4000000000ull;
                }
                if (__any_sync(kFull, dead)) {
                    if (lane == 0) atomicExch(p.status, (int)epoch);
                    break;
                }
                // the stage's barrier is reused for the repeat: one more phase
                if (lane == 0) mbar_arrive_expect_tx(bar0 + stage * 8, row_bytes * nrows);
                __syncwarp();
                if (lane < nrows)
                    bulk_g2s(st + (unsigned)(lane * NLP * 8), mb + (size_t)(y0 + lane) * p.Npad, row_bytes, bar0 + stage * 8);
                mbar_spin(bar0 + stage * 8, (ph >> stage) & 1, p.status, epoch);
                ph ^= 1u << stage;
            }
        }
        TKB_PK(4);
        __syncwarp();
        if (lane == 0) mbar_arrive(qfull_s + buf * 8);
        if (threadIdx.x == NCONS) TKB_STAMP(it, 7);
        TKB_PK(5);
    }
    if (threadIdx.x == NCONS) TKB_STAMP(255, 1);
#ifdef TKB_TIMELINE
    if (p.timeline && lane == 0)   // cycle totals: stamp idx 248..253, slots 2..5 = producer warps 0..3
        for (int kk = 0; kk < 6; ++kk) p.timeline[((size_t)blockIdx.x * 256 + 248 + kk) * 8 + 2 + pw] = (unsigned long long)ck[kk];
#endif
}

// =================================================================================================
// STREAMING CTA: far partials of the owned columns, all tracks of the launch, in lock-step with the chain
// =================================================================================================
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void helper_role(const SweepParams &p, unsigned char *smem_raw, int h) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    constexpr int NKIND = (DO_V ? 1 : 0) + (DO_L ? 1 : 0);
    const int T = p.T, Nl = p.Nl;
    const int nb = (T + BX - 1) / BX;
    const unsigned epoch = p.epoch;
    const int NLP = (Nl + 3) & ~3;
    const int nvec = NLP >> 2;
    const int t = threadIdx.x;

    const unsigned bar_s = smem_u32(smem_raw);
    const unsigned qbuf_s = bar_s + (unsigned)kBarBytes;              // [QB][row][semiring slot][NLP] floats
    const unsigned qraw_s = qbuf_s + (unsigned)qbuf_bytes(NLP);       // [producer warp][stage][row][NLP] tagged words
    const unsigned fifo_s = qraw_s + (unsigned)qraw_bytes(NLP);       // [stage][row][NI] float4

    // batches b = btop .. bmin (8 rows each, descending); the CTA's smallest column is 2h
    if (2 * h >= T) return;
    const int J0 = (2 * h) / BX;
    if (J0 > nb - ND - 2) return;                     // no far field at all
    const int btop = (T - 1) / PB;
    const int bmin = (BX / PB) * (J0 + ND + 1);       // batch of the first far row of column block J0

    const int warp = t >> 5, lane = t & 31;
    const unsigned qfull_s = bar_s, qempty_s = bar_s + QB * 8, rawfull_s = bar_s + 2 * QB * 8;
    if (t == 0) {
        for (int s = 0; s < QB; ++s) {
            mbar_init(qfull_s + s * 8, 2 * NKIND);   // the batch's producer warps arrive
            mbar_init(qempty_s + s * 8, NCONSW);
        }
        for (int s = 0; s < NPW * QS; ++s) mbar_init(rawfull_s + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned qbuf_stride = (unsigned)(PB * NKIND * NLP * 4);

    if (warp >= NCONSW) {
        const int pw = warp - NCONSW;            // 0..3 -> (semiring, half of the batch's rows)
        const int kind = pw >> 1;                // 0 Viterbi, 1 log-sum
        if (!(kind ? DO_L : DO_V)) return;
        const int k = (DO_V && DO_L) ? kind : 0;  // semiring slot inside a qbuf row
        mailbox_producer<NKIND>(p, h, pw, kind, pw & 1, k, btop, bmin, qbuf_s, qraw_s, qfull_s, qempty_s, rawfull_s);
        return;
    }

    // ---- items: (column, 4 tracks) ----------------------------------------------------------------------------
    bool valid[IPT];
    int ix[IPT], iv[IPT], ibend[IPT];
    const float *isrc[IPT];   // &S(0, x) + track offset; row y adds y * sy
    int nbytes[IPT];
    float vmax[IPT][4], lM[IPT][4], lS[IPT][4];
    int vsel[IPT][4];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int item = t + i * NCONS;
        const int m = item / (CW * nvec), rem = item - m * (CW * nvec);
        const int c = rem / nvec, v = rem - c * nvec;
        const int x = CW * (h + m * p.H) + c;
        const int J = x / BX;
        valid[i] = item < p.nitems && x < T && J <= nb - ND - 2;
        ix[i] = x;
        iv[i] = v;
        ibend[i] = (BX / PB) * (J + ND + 1);
        nbytes[i] = min(4, Nl - 4 * v) * 4;
        isrc[i] = valid[i] ? p.Sbase + (long long)x * p.sx + p.n_lo + 4 * v : p.Sbase;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            vmax[i][q] = -INFINITY;
            vsel[i][q] = -1;
            lM[i][q] = -FLT_MAX;
            lS[i][q] = 0.0f;
        }
    }
    const int nstg = p.nstg;
    const unsigned stage_bytes = (unsigned)PB * p.NI * 16u;
    // prefetch of batch bb: 8 rows x 16 bytes per item, into stage (btop - bb) % nstg
    int istage = 0;   // stage of the next batch to prefetch: (btop - bb) % nstg without the division
    // running source pointers: row 8*bb of the next batch to prefetch (one pointer step per batch, one per row)
    const long long bstep = (long long)PB * p.sy;
    const float *nsrc[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) nsrc[i] = isrc[i] + (long long)btop * bstep;
    const unsigned rstride_i = (unsigned)p.NI * 16u;
    auto issue = [&](int bb) {
        const unsigned st = fifo_s + (unsigned)istage * stage_bytes;
        istage = istage + 1 == nstg ? 0 : istage + 1;
#ifdef TKB_TIMELINE
        if (p.dbg & 2) { cp_async_commit(); return; }
#endif
        if (bb >= bmin) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                if (valid[i] && bb >= ibend[i]) {
                    unsigned dst = st + (unsigned)(t + i * NCONS) * 16u;
                    const float *src = nsrc[i];
                    if (bb * PB + PB <= T && ALIGN == 16) {   // the common case: eight rows, 16-byte copies
#pragma unroll
                        for (int r = 0; r < PB; ++r) {
                            cp_async16_s(dst, src, nbytes[i]);
                            dst += rstride_i;
                            src += p.sy;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < PB; ++r) {
                            if (bb * PB + r < T) {
                                if (ALIGN == 16) {
                                    cp_async16_s(dst, src, nbytes[i]);
                                } else {
                                    for (int q = 0; q * 4 < nbytes[i]; ++q) cp_async4_s(dst + q * 4, src + q, 4);
                                }
                            }
                            dst += rstride_i;
                            src += p.sy;
                        }
                    }
                }
                nsrc[i] -= bstep;
            }
        }
        cp_async_commit();
    };
    for (int s = 0; s < nstg - 1; ++s) issue(btop - s);

    int cstage = 0;
#ifdef TKB_TIMELINE
    long long ck[6] = {0, 0, 0, 0, 0, 0};
    long long c0 = clock64(), c1;
#define TKB_CK(k) do { c1 = clock64(); ck[k] += c1 - c0; c0 = c1; } while (0)
#else
#define TKB_CK(k) do {} while (0)
#endif
    for (int b = btop, it = 0; b >= bmin; --b, ++it) {
#ifdef TKB_TIMELINE
        if (p.dbg & 32) break;
#endif
        if (t == 0) TKB_STAMP(it, 0);
        TKB_CK(0);
        issue(b - (nstg - 1));
        TKB_CK(1);
        if (nstg == 4) cp_async_wait<3>();
        else if (nstg == 3) cp_async_wait<2>();
        else cp_async_wait<1>();
        TKB_CK(2);
        const int buf = it & (QB - 1);
#ifdef TKB_TIMELINE
        if (!(p.dbg & 8))
#endif
        mbar_spin(qfull_s + buf * 8, (it / QB) & 1, p.status, epoch);
        TKB_CK(3);
        if (t == 0) TKB_STAMP(it, 1);
        const unsigned st = fifo_s + (unsigned)cstage * stage_bytes;
        cstage = cstage + 1 == nstg ? 0 : cstage + 1;
        const unsigned qb = qbuf_s + (unsigned)buf * qbuf_stride;
        const int y0 = b * PB;
#ifdef TKB_TIMELINE
        const bool skip_math = (p.dbg & 1) != 0;
#else
        constexpr bool skip_math = false;
#endif
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            if (valid[i] && b >= ibend[i] && !(skip_math && b != ibend[i])) {
                const unsigned src = st + (unsigned)(t + i * NCONS) * 16u;
                const unsigned qsrc = qb + (unsigned)(iv[i] * 16);
                const unsigned rstride = (unsigned)p.NI * 16u, qstride = (unsigned)(NKIND * NLP * 4);
                const unsigned long long kL2 = pack2(kLog2e, kLog2e);
                // FULL = all eight rows of the batch exist (every batch but a ragged first one): no per-row branch, so
                // the shared-memory loads of a half-batch are issued together and their latencies overlap
                auto rows = [&](auto full_tag) {
                    constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                    for (int half = 1; half >= 0; --half) {   // rows 7..4, then 3..0 (descending y: tie order)
                        if (FULL || y0 + 4 * half < T) {
                            unsigned long long x01[4], x23[4];   // log-sum candidates of four rows, packed by track pair
#pragma unroll
                            for (int rr = 3; rr >= 0; --rr) {
                                const int r = 4 * half + rr, y = y0 + r;
                                if (FULL || y < T) {
                                    unsigned long long s01, s23;
                                    lds128_2(src + (unsigned)r * rstride, s01, s23);
                                    if (DO_V) {
                                        unsigned long long q01, q23;
                                        lds128_2(qsrc + (unsigned)r * qstride, q01, q23);
                                        float xv[4];
                                        unpack2(add2(q01, s01), xv[0], xv[1]);   // one fp32 add per candidate, as the reference
                                        unpack2(add2(q23, s23), xv[2], xv[3]);
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            const bool tk = (DIR == TKB_BACKWARD) ? (xv[q] >= vmax[i][q]) : (xv[q] > vmax[i][q]);
                                            vmax[i][q] = tk ? xv[q] : vmax[i][q];
                                            vsel[i][q] = tk ? y : vsel[i][q];
                                        }
                                    }
                                    if (DO_L) {
                                        unsigned long long q01, q23;
                                        lds128_2(qsrc + (unsigned)r * qstride + (unsigned)((NKIND - 1) * NLP * 4), q01, q23);
                                        x01[rr] = fma2(s01, kL2, q01);
                                        x23[rr] = fma2(s23, kL2, q23);
                                    }
                                } else if (DO_L) {
                                    x01[rr] = x23[rr] = pack2(-FLT_MAX, -FLT_MAX);
                                }
                            }
                            if (DO_L) {
                                float xs[4][4];
#pragma unroll
                                for (int rr = 0; rr < 4; ++rr) {
                                    unpack2(x01[rr], xs[rr][0], xs[rr][1]);
                                    unpack2(x23[rr], xs[rr][2], xs[rr][3]);
                                }
                                float Mn[4], sc[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float m = fmaxf(fmaxf(xs[0][q], xs[1][q]), fmaxf(xs[2][q], xs[3][q]));
                                    Mn[q] = fmaxf(lM[i][q], m);
                                    sc[q] = ex2f(lM[i][q] - Mn[q]);
                                    lM[i][q] = Mn[q];
                                }
                                const unsigned long long nM01 = pack2(-Mn[0], -Mn[1]), nM23 = pack2(-Mn[2], -Mn[3]);
                                unsigned long long a01 = mul2(pack2(lS[i][0], lS[i][1]), pack2(sc[0], sc[1]));
                                unsigned long long a23 = mul2(pack2(lS[i][2], lS[i][3]), pack2(sc[2], sc[3]));
#pragma unroll
                                for (int rr = 0; rr < 4; ++rr) {
                                    float d0, d1, d2, d3;
                                    unpack2(add2(x01[rr], nM01), d0, d1);
                                    unpack2(add2(x23[rr], nM23), d2, d3);
                                    a01 = add2(a01, pack2(ex2f(d0), ex2f(d1)));
                                    a23 = add2(a23, pack2(ex2f(d2), ex2f(d3)));
                                }
                                unpack2(a01, lS[i][0], lS[i][1]);
                                unpack2(a23, lS[i][2], lS[i][3]);
                            }
                        }
                    }
                };
                if (y0 + PB <= T) rows(std::true_type{});
                else rows(std::false_type{});
                if (b == ibend[i]) {
                    // the far field of this column is complete: hand it to the solver of each of my tracks
                    const int J = ix[i] / BX, cx = ix[i] - J * BX;
                    unsigned long long *dst =
                        p.part + (((size_t)J * 2) * p.Npad + p.n_lo + 4 * iv[i]) * (BX * 2) + 2 * cx;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q * 4 < nbytes[i]) {
                            if (DO_V) {
                                publish(dst + (size_t)q * (BX * 2), vmax[i][q], epoch);
                                publish(dst + (size_t)q * (BX * 2) + 1, __int_as_float(vsel[i][q]), epoch);
                            }
                            if (DO_L) {
                                publish(dst + ((size_t)p.Npad + q) * (BX * 2), lM[i][q], epoch);
                                publish(dst + ((size_t)p.Npad + q) * (BX * 2) + 1, lS[i][q], epoch);
                            }
                        }
                    }
                }
            }
        }
        TKB_CK(4);
        if (t == 0) TKB_STAMP(it, 2);
        __syncwarp();
        if (lane == 0) mbar_arrive(qempty_s + buf * 8);
        TKB_CK(5);
    }
    cp_async_wait_all();
#ifdef TKB_TIMELINE
    if (p.timeline && (t & 31) == 0 && t < 64) {   // cycle totals of warps 0 and 1: stamp slots 248..253
        for (int k = 0; k < 6; ++k) p.timeline[((size_t)blockIdx.x * 256 + 248 + k) * 8 + (t >> 5)] = (unsigned long long)ck[k];
    }
#endif
}

// One chain warp: track n0 + tr, lane = column of the current block, semirings CV / CL.  When a launch computes
// both, each track has two chain warps (one per semiring) on the same SMSP: their dependent chains interleave.
struct ChainCtx {
    unsigned full_s, empty_s, pub_s, pubfull_s, pubempty_s;
    int n0, tr;
};
template <int DIR, bool CV, bool CL>
__device__ __forceinline__ void chain_warp(const SweepParams &p, unsigned char *smem_raw, const ChainCtx &cx) {
    constexpr bool DO_V = CV, DO_L = CL;
    const int lane = threadIdx.x & 31;
    const int T = p.T;
    const int nb = (T + BX - 1) / BX;
    const unsigned epoch = p.epoch;
    const unsigned full_s = cx.full_s, empty_s = cx.empty_s, pub_s = cx.pub_s, pubfull_s = cx.pubfull_s,
                   pubempty_s = cx.pubempty_s;
    const int n0 = cx.n0, tr = cx.tr;
    // ---------------- chain warp: track n, lane = column ---------------------------------------------
    const int n = n0 + tr;
    const int c = lane;
    // accumulators: [0] the block on the chain, [d] the block d below it
    float best[ND + 1], lM[ND + 1], lS[ND + 1];
    int bsel[ND + 1];
#pragma unroll
    for (int d = 0; d <= ND; ++d) {
        best[d] = -INFINITY;
        bsel[d] = -1;
        lM[d] = -FLT_MAX;
        lS[d] = 0.0f;
    }
    // unary terms of my column in the block on the chain, and (prefetched) in the next one
    auto load_unary = [&](int j, float &d_out, float &e_out) {
        const int x = j * BX + c;
        d_out = 0.0f;
        e_out = 0.0f;
        if (j >= 0 && x < T) {
            d_out = __ldg(p.Sbase + (long long)x * (p.sx + p.sy) + n);
            if (x < T - 1) e_out = __ldg(p.etabase + (long long)x * p.se + n);
        }
    };
    float nx_d, nx_eta;
    load_unary(nb - 1, nx_d, nx_eta);
    float qtopV = 0.0f, qtopM = 0.0f, qtopS = 0.0f;  // row 32(j+1) (the row right above column 31), broadcast
    // far partial of the NEXT block, fetched while the chain is still in this one
    unsigned long long fw[4] = {0, 0, 0, 0};
    const unsigned long long *fsrc = nullptr;
    auto far_fetch = [&](int jn) {
        if (jn < 0 || jn > nb - ND - 2) return;
        fsrc = p.part + (((size_t)jn * 2) * p.Npad + n) * (BX * 2) + 2 * c;
        if (DO_V) {
            fw[0] = ld_cg_u64(fsrc);
            fw[1] = ld_cg_u64(fsrc + 1);
        }
        if (DO_L) {
            fw[2] = ld_cg_u64(fsrc + (size_t)p.Npad * BX * 2);
            fw[3] = ld_cg_u64(fsrc + (size_t)p.Npad * BX * 2 + 1);
        }
    };
    const float *bands = reinterpret_cast<const float *>(smem_raw);

    for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
        const int slot = it % NBAND;
        const int x0 = j * BX, x = x0 + c;
        const int ncols = min(BX, T - x0);
        if (it >= 2) mbar_wait(pubempty_s + (tr * 2 + (it & 1)) * 8, ((it >> 1) - 1) & 1, p.status, epoch);
        const float s_d = nx_d, s_eta = nx_eta;
        load_unary(j - 1, nx_d, nx_eta);
        const float dr = relu_mask(s_d);
        float sp2 = 0.0f, eta2 = 0.0f;
        if (DO_L) {
            const float d2 = s_d * kLog2e;
            sp2 = fmaxf(d2, 0.0f) + lg2f(1.0f + ex2f(-fabsf(d2)));  // softplus(d)*log2e
            eta2 = s_eta * kLog2e;
        }
        if (threadIdx.x == 0) TKB_STAMP(it, 0);
        if (lane == 0) TKB_STAMP(128 + it, (DO_V ? 0 : 4) + tr);
        // ---- far partial of this block (rows of blocks > j+ND), written by the streaming CTAs --------------
        if (j <= nb - ND - 2) {
            {
                // all words of the partial in one round trip per attempt
                const unsigned long long *srcL = fsrc + (size_t)p.Npad * BX * 2;
                unsigned long long t0 = 0;
                for (unsigned tries = 0;; ++tries) {
                    bool ok = true;
                    if (DO_V) ok &= (unsigned)(fw[0] >> 32) == epoch && (unsigned)(fw[1] >> 32) == epoch;
                    if (DO_L) ok &= (unsigned)(fw[2] >> 32) == epoch && (unsigned)(fw[3] >> 32) == epoch;
                    if (ok) break;
                    if ((tries & 63) == 63) {  // watchdog, off the fast path
                        if (t0 == 0) t0 = globaltimer_ns();
                        if (*(volatile unsigned *)p.status == epoch) break;
                        if (globaltimer_ns() - t0 > 4000000000ull) {
                            atomicExch(p.status, (int)epoch);
                            break;
                        }
                    }
                    __nanosleep(40);
                    if (DO_V) {
                        fw[0] = ld_relaxed_u64(fsrc);
                        fw[1] = ld_relaxed_u64(fsrc + 1);
                    }
                    if (DO_L) {
                        fw[2] = ld_relaxed_u64(srcL);
                        fw[3] = ld_relaxed_u64(srcL + 1);
                    }
                }
            }
            if (DO_V) {
                const float fv = __uint_as_float((unsigned)fw[0]);
                const int fs = (int)(unsigned)fw[1];
                // far rows are larger y than anything accumulated so far: BACKWARD prefers the smaller y on ties
                const bool tk = (DIR == TKB_BACKWARD) ? (fv > best[0]) : (fv >= best[0]);
                bsel[0] = tk ? fs : bsel[0];
                best[0] = fmaxf(best[0], fv);
            }
            if (DO_L) lse_push(lM[0], lS[0], __uint_as_float((unsigned)fw[2]), __uint_as_float((unsigned)fw[3]));
        }
        if (threadIdx.x == 0) TKB_STAMP(it, 1);
        // ---- the skip out of the top column into row 32(j+1): candidate 0 of the reference, wins every tie ----
        if (j < nb - 1) {
            if (DO_V) {
                const float xk = (c == BX - 1) ? qtopV + s_eta : -INFINITY;
                bsel[0] = (xk >= best[0]) ? -1 : bsel[0];
                best[0] = fmaxf(best[0], xk);
            }
            if (DO_L) lse_push(lM[0], lS[0], (c == BX - 1) ? qtopM + eta2 : -INFINITY, qtopS);
        }
        if (x == T - 1) {  // terminal column: no candidates; q = S*(S>0) (-0 + dr keeps the reference's signed zero)
            best[0] = -0.0f;
            bsel[0] = -1;
            lM[0] = 0.0f;
            lS[0] = 1.0f;
        }
        if (DO_L && lS[0] > 0.0f) {  // renormalise: S restarts at 1 in every block (it at most doubles per step)
            lM[0] += lg2f(lS[0]);
            lS[0] = 1.0f;
        }
        // ---- wait for the row band ------------------------------------------------------------------
        mbar_wait(full_s + slot * 8, (it / NBAND) & 1, p.status, epoch);
        if (threadIdx.x == 0) TKB_STAMP(it, 3);
        // my column in the diagonal tile is band column ND*32 + c; in the tile d blocks below, (ND-d)*32 + c
        const float *colp = bands + (size_t)slot * (kBandBytes / 4) + c * NQ + tr;
        // log-sum: the skip x -> x+1 folded into the coefficient of the row right above my column
        float comb = -INFINITY;
        if (DO_L && c + 1 < ncols) {
            const float spv = colp[((c + 1) * BANDCOLS + ND * BX) * NQ] * kLog2e;
            comb = fmaxf(spv, eta2) + lg2f(1.0f + ex2f(-fabsf(spv - eta2)));
        }
        // ---- one chain step: row e of this block is final in lane e; broadcast it and push it -----------
        auto step = [&](const int e) {
            const int y = x0 + e;
            const float *rowp = colp + (size_t)e * (BANDCOLS * NQ);
            float sv[ND + 1];
#pragma unroll
            for (int d = 0; d <= ND; ++d) sv[d] = rowp[(ND - d) * BX * NQ];
            const bool below = c < e;
            if (DO_V) {
                const float qb = __shfl_sync(kFull, best[0] + dr, e);
                if (e == 0) qtopV = qb;
                {
                    const float xi = below ? qb + sv[0] : -INFINITY;
                    const float xk = (c == e - 1) ? qb + s_eta : -INFINITY;
                    const bool tk = (DIR == TKB_BACKWARD) ? (xi >= best[0]) : (xi > best[0]);
                    const float b1 = fmaxf(best[0], xi);
                    bsel[0] = tk ? y : bsel[0];
                    bsel[0] = (xk >= b1) ? -1 : bsel[0];
                    best[0] = fmaxf(b1, xk);
                }
#pragma unroll
                for (int d = 1; d <= ND; ++d) {
                    const float xi = qb + sv[d];
                    const bool tk = (DIR == TKB_BACKWARD) ? (xi >= best[d]) : (xi > best[d]);
                    bsel[d] = tk ? y : bsel[d];
                    best[d] = fmaxf(best[d], xi);
                }
            }
            if (DO_L) {
                const float Mb = __shfl_sync(kFull, lM[0] + sp2, e);
                const float sb = __shfl_sync(kFull, lS[0], e);
                if (e == 0) {
                    qtopM = Mb;
                    qtopS = sb;
                }
                const float coef = (c == e - 1) ? comb : (below ? sv[0] * kLog2e : -INFINITY);
                lse_push(lM[0], lS[0], Mb + coef, sb);
#pragma unroll
                for (int d = 1; d <= ND; ++d) lse_push(lM[d], lS[d], fmaf(sv[d], kLog2e, Mb), sb);
            }
        };
        // rows 8*e8 .. 8*e8+7 are final in their lanes: hand them to the publisher warp of this track
        auto publish_batch = [&](const int e8) {
            if (c >= e8 * PB && c < e8 * PB + PB) {
                const unsigned dst = pub_s + (unsigned)((((tr * 2 + (it & 1)) * BX + c) * 2) * 8);
                if (DO_V) {
                    const float qfin = best[0] + dr;
                    const int osel = bsel[0] < 0 ? -1 : ((DIR == TKB_BACKWARD) ? bsel[0] : T - 1 - bsel[0]);
                    sts64_nc(dst, __float_as_uint(qfin), ((unsigned)(osel + 1) << 1) | (s_d > 0.0f ? 1u : 0u));
                }
                if (DO_L) sts64_nc(dst + 8, __float_as_uint(lM[0] + sp2), __float_as_uint(lS[0]));
                mbar_arrive_nc(pubfull_s + (unsigned)(((tr * 2 + (it & 1)) * (BX / PB) + e8) * 8));
            }
        };
        if (ncols == BX) {
#pragma unroll
            for (int e8 = BX / PB - 1; e8 >= 0; --e8) {
                if (e8 == TKB_FARFETCH_BATCH) far_fetch(j - 1);
#pragma unroll
                for (int i = PB - 1; i >= 0; --i) step(e8 * PB + i);
                publish_batch(e8);
            }
        } else {  // the ragged top block
            far_fetch(j - 1);
            for (int e8 = (ncols - 1) >> 3; e8 >= 0; --e8) {
                for (int e = min(ncols - 1, e8 * PB + PB - 1); e >= e8 * PB; --e) step(e);
                publish_batch(e8);
            }
        }
        if (threadIdx.x == 0) TKB_STAMP(it, 2);
        // ---- next block: release the band, shift the accumulators ------------------------------------------
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_s + slot * 8);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            best[d] = best[d + 1];
            bsel[d] = bsel[d + 1];
            lM[d] = lM[d + 1];
            lS[d] = lS[d + 1];
        }
        best[ND] = -INFINITY;
        bsel[ND] = -1;
        lM[ND] = -FLT_MAX;
        lS[ND] = 0.0f;
    }
    if (threadIdx.x == 0) TKB_STAMP(255, 0);
}

// =================================================================================================
// SOLVER: the chain of NQ tracks, from the last position to the first, in one SM
// =================================================================================================
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void solver_role(const SweepParams &p, const CUtensorMap *map, unsigned char *smem_raw,
                                            int s) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    constexpr bool USE_TMA = (DIR == TKB_BACKWARD) && (ALIGN == 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n0 = p.n_lo + s * NQ;                        // first track of this solver
    const int nvalid = min(max(p.n_lo + p.Nl - n0, 0), NQ);  // chain warps with a real track
    const unsigned epoch = p.epoch;
    const unsigned band_s = smem_u32(smem_raw);
    const unsigned full_s = band_s + (unsigned)(NBAND * kBandBytes);  // + slot*8
    const unsigned empty_s = full_s + NBAND * 8;
    const unsigned pub_s = empty_s + NBAND * 8;                      // [tr][block parity][c][kind] 8 bytes
    const unsigned pubfull_s = pub_s + (unsigned)kPubBytes;          // [tr][block parity][batch]
    const unsigned pubempty_s = pubfull_s + NCW * 2 * (BX / PB) * 8;  // [tr][block parity]

    if (threadIdx.x == 0) {
        for (int sl = 0; sl < NBAND; ++sl) {
            mbar_init(full_s + sl * 8, USE_TMA ? 1 : NLW * 32);
            mbar_init(empty_s + sl * 8, nvalid * ((DO_V && DO_L) ? 2 : 1));
        }
        for (int sl = 0; sl < NCW * 2 * (BX / PB); ++sl)
            mbar_init(pubfull_s + sl * 8, PB * ((DO_V && DO_L) ? 2 : 1));  // the batch's lanes arrive
        for (int sl = 0; sl < NCW * 2; ++sl) mbar_init(pubempty_s + sl * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (nvalid == 0) return;

    if (warp >= NCW && warp < NCW + NLW) {
        // ---------------- band producer ----------------------------------------------------------------------
        // band of row block j: rows y = 32j .. 32j+31, columns x = 32(j-ND) .. 32j+31, this solver's NQ tracks;
        // chunk (e, cc) -> band + (e*BANDCOLS + cc)*16.
        if (USE_TMA) {
            // one thread, one tensor copy per band; cells outside the tensor (x < 0, y >= T, track >= N) arrive as
            // zeros and are never read, like the cells above the diagonal
            if (warp != NCW || lane != 0) return;
            for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
                const int slot = it % NBAND;
                if (it >= NBAND) mbar_wait(empty_s + slot * 8, ((it / NBAND) - 1) & 1, p.status, epoch);
                mbar_arrive_expect_tx(full_s + slot * 8, (unsigned)kBandBytes);
                tma_load_3d(band_s + (unsigned)(slot * kBandBytes), map, n0, (j - ND) * BX, j * BX, full_s + slot * 8);
            }
            TKB_STAMP(255, 2);
            return;
        }
        // per-lane cp.async gather.  Chunks above the diagonal, left of column 0 or below row T-1 are never read
        // and not fetched.
        const int lt = threadIdx.x - NCW * 32;
        const int nbytes = nvalid * 4;
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int slot = it % NBAND;
            if (it >= NBAND) mbar_wait(empty_s + slot * 8, ((it / NBAND) - 1) & 1, p.status, epoch);
            const int y0 = j * BX, xlo = (j - ND) * BX;
            const unsigned dst0 = band_s + (unsigned)(slot * kBandBytes);
            for (int i = lt; i < BX * BANDCOLS; i += NLW * 32) {
                const int e = i / BANDCOLS, cc = i - e * BANDCOLS;
                const int y = y0 + e, x = xlo + cc;
                if (x < 0 || x > y || y >= T) continue;
                const float *src = p.Sbase + (long long)x * p.sx + (long long)y * p.sy + n0;
                const unsigned dst = dst0 + (unsigned)i * 16u;
                if (ALIGN == 16) {
                    cp_async16_s(dst, src, nbytes);
                } else if (ALIGN == 8) {
                    cp_async8_s(dst, src, nvalid > 0 ? 8 : 0);
                    if (nvalid > 2) cp_async8_s(dst + 8, src + 2, 8);
                } else {
                    for (int q = 0; q < nvalid; ++q) cp_async4_s(dst + q * 4, src + q, 4);
                }
            }
            mbar_arrive_cp_async(full_s + slot * 8);
        }
        cp_async_wait_all();
        return;
    }
    if (warp >= NCW + NLW && warp < NCW + NLW + NCW) {
        // ---------------- publisher warps: one per track; results go shared memory -> mailbox and tables -----------
        // (a global store issued by a chain warp costs it ~200 cycles per batch: profiles/r01_sweep_experiments.txt)
        const int tr = warp - (NCW + NLW);
        if (tr >= nvalid) return;
        const int n = n0 + tr;
        // The mailbox is replicated NREP times and the streaming CTAs read "their" replica: every 128-byte mailbox
        // line is wanted by all of them at the same moment, and an L2 slice hands out one line to ~100 requesters per
        // microsecond (measured: 126 readers of one replica cost 1.3 us per batch of 8 rows, more than the arithmetic).
        // One store instruction writes all replicas: lane -> (replica lane / 8, row lane % 8 of the batch).
        const int rep = lane >> 3, lr = lane & (PB - 1);
        unsigned long long *mV = p.mbox + ((size_t)rep * 2) * T * p.Npad + n, *mL = mV + (size_t)T * p.Npad;
        const int bmax_top = (T - (nb - 1) * BX - 1) / PB;  // last batch index of the ragged top block
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int x0 = j * BX;
            const int ncols = min(BX, T - x0);
            for (int e8 = (ncols - 1) / PB; e8 >= 0; --e8) {
                // batches the ragged top block skips never arrive: their barriers are one phase behind
                const unsigned npast = (unsigned)(it >> 1) - (((it & 1) == 0 && it > 0 && e8 > bmax_top) ? 1u : 0u);
                mbar_wait(pubfull_s + (unsigned)(((tr * 2 + (it & 1)) * (BX / PB) + e8) * 8), npast & 1, p.status, epoch);
                const int c = e8 * PB + lr, x = x0 + c;
                if (rep < NREP && x < T) {
                    const int pos = (DIR == TKB_BACKWARD) ? x : T - 1 - x;
                    const unsigned src = pub_s + (unsigned)((((tr * 2 + (it & 1)) * BX + c) * 2) * 8);
                    if (DO_V) {
                        const unsigned long long w = lds64(src);
                        const float qfin = __uint_as_float((unsigned)w);
                        publish(mV + (size_t)x * p.Npad, qfin, epoch);
                        if (rep == 0) {
                            p.code[(size_t)n * T + pos] = (unsigned)(w >> 32);
                            if (p.outv) p.outv[(size_t)pos * N + n] = qfin;
                        }
                    }
                    if (DO_L) {
                        const unsigned long long w = lds64(src + 8);
                        const float v2 = __uint_as_float((unsigned)w) + lg2f(__uint_as_float((unsigned)(w >> 32)));
                        publish(mL + (size_t)x * p.Npad, v2, epoch);
                        if (rep == 0 && p.outl) p.outl[(size_t)pos * N + n] = v2 * kLn2;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pubempty_s + (tr * 2 + (it & 1)) * 8);
            if (lane == 0) TKB_STAMP(64 + it, tr);
        }
        if (tr == 0 && lane == 0) TKB_STAMP(255, 1);
        return;
    }
    // ---------------- chain warps ------------------------------------------------------------------------
    ChainCtx cx;
    cx.full_s = full_s;
    cx.empty_s = empty_s;
    cx.pub_s = pub_s;
    cx.pubfull_s = pubfull_s;
    cx.pubempty_s = pubempty_s;
    cx.n0 = n0;
    if (DO_V && DO_L) {  // split: Viterbi chains on warps 0..NCW-1, log-sum chains on the last NCW warps
        if (warp < NCW) {
            cx.tr = warp;
            if (cx.tr < nvalid) chain_warp<DIR, true, false>(p, smem_raw, cx);
        } else if (warp >= NW - NCW) {
            cx.tr = warp - (NW - NCW);
            if (cx.tr < nvalid) chain_warp<DIR, false, true>(p, smem_raw, cx);
        }
    } else if (warp < NCW) {
        cx.tr = warp;
        if (cx.tr < nvalid) chain_warp<DIR, DO_V, DO_L>(p, smem_raw, cx);
    }
}

template <int DIR, int ALIGN, int MODE>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const __grid_constant__ CUtensorMap map, const SweepParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if ((int)blockIdx.x < p.S)
        solver_role<DIR, ALIGN, MODE>(p, &map, smem_raw, (int)blockIdx.x);
    else
        helper_role<DIR, ALIGN, MODE>(p, smem_raw, (int)blockIdx.x - p.S);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;

template <int DIR, int ALIGN, int MODE>
static int launch_one(const SweepParams &p, const CUtensorMap &map, int grid, cudaStream_t stream) {
    auto kern = sweep_kernel<DIR, ALIGN, MODE>;
    static bool configured[kMaxDevices] = {};  // per instantiation AND per device (function attributes are per device)
    int dev = 0;
    TKB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        TKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    SweepParams pp = p;
    CUtensorMap mm = map;
    void *args[] = {&mm, &pp};
    TKB_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(NT), args, kSweepSmem, stream));
    return 0;
}

template <int DIR, int ALIGN>
static int launch_mode(int mode, const SweepParams &p, const CUtensorMap &map, int grid, cudaStream_t stream) {
    switch (mode) {
        case TKB_SWEEP_VITERBI: return launch_one<DIR, ALIGN, TKB_SWEEP_VITERBI>(p, map, grid, stream);
        case TKB_SWEEP_LOGSUM: return launch_one<DIR, ALIGN, TKB_SWEEP_LOGSUM>(p, map, grid, stream);
        default: return launch_one<DIR, ALIGN, TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM>(p, map, grid, stream);
    }
}

template <int DIR>
static int launch_align(int align, int mode, const SweepParams &p, const CUtensorMap &map, int grid,
                        cudaStream_t stream) {
    switch (align) {
        case 16: return launch_mode<DIR, 16>(mode, p, map, grid, stream);
        case 8: return launch_mode<DIR, 8>(mode, p, map, grid, stream);
        default: return launch_mode<DIR, 4>(mode, p, map, grid, stream);
    }
}

static unsigned long long *g_timeline = nullptr;  // diagnostics build only
extern int g_dbg_flags;
static int num_sms() {
    static int sms[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}
static size_t mailbox_bytes(int T, int N) {
    const size_t npad = (size_t)((N + 7) / 8) * 8;
    return (size_t)NREP * 2 * (size_t)T * npad * sizeof(unsigned long long);
}
static size_t partial_bytes(int T, int N) {
    const size_t npad = (size_t)((N + 7) / 8) * 8, nb = (size_t)((T + BX - 1) / BX);
    return nb * 2 * npad * BX * 2 * sizeof(unsigned long long);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

int g_dbg_flags = 0;

}  // namespace tkb

using namespace tkb;

extern "C" size_t tkb_sweep_workspace_bytes(int T, int N) {
    if (T < 1 || N < 1) return 0;
    return kHeaderBytes + mailbox_bytes(T, N) + partial_bytes(T, N);
}

extern "C" int tkb_semicrf_sweep(const float *score, const float *noise, int T, int N, int direction, int flags,
                                 void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit,
                                 float *out_lse, void *stream_) {
    return tkb_semicrf_sweep_pitched(score, N, noise, T, N, direction, flags, workspace, epoch, out_code, out_vit, out_lse,
                                     stream_);
}

extern "C" int tkb_semicrf_sweep_pitched(const float *score, int64_t pitch, const float *noise, int T, int N,
                                         int direction, int flags, void *workspace, uint32_t epoch,
                                         uint32_t *out_code, float *out_vit, float *out_lse, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!score || !workspace || T < 1 || N < 1 || (T > 1 && !noise) || epoch == 0 ||
        (direction != TKB_BACKWARD && direction != TKB_FORWARD) ||
        (flags & ~(TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM)) || flags == 0 ||
        ((flags & TKB_SWEEP_VITERBI) && !out_code) || (long long)T * T >= (1ll << 40) || pitch < N) {
        set_error("tkb_semicrf_sweep: invalid argument (T=%d N=%d dir=%d flags=%d epoch=%u)", T, N, direction,
                  flags, epoch);
        return TKB_EINVAL;
    }
    const int sms = num_sms();
    if (sms < 2) {
        set_error("tkb_semicrf_sweep: no CUDA device");
        return TKB_ENODEV;
    }
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.T = T;
    p.N = N;
    p.Npad = (N + 7) / 8 * 8;
    p.dir = direction;
    p.epoch = epoch;
    p.status = reinterpret_cast<int *>(workspace);
    p.mbox = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + kHeaderBytes);
    p.part = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + kHeaderBytes +
                                                    mailbox_bytes(T, N));
    p.code = out_code;
    p.outv = out_vit;
    p.outl = out_lse;
    p.timeline = g_timeline;
    p.dbg = g_dbg_flags;
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
    const int nb = (T + BX - 1) / BX;

    // TMA descriptor of the score tensor for the solver's band copies (BACKWARD, 16-byte path)
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (direction == TKB_BACKWARD && align == 16) {
        EncodeTiledFn enc = encode_tiled();
        if (!enc) {
            set_error("tkb_semicrf_sweep: cuTensorMapEncodeTiled is not available from this driver");
            return TKB_ENODEV;
        }
        const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)T, (cuuint64_t)T};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)T * (cuuint64_t)pitch * 4};
        const cuuint32_t box[3] = {NQ, BANDCOLS, BX};
        const cuuint32_t es[3] = {1, 1, 1};
        const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(score), dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tkb_semicrf_sweep: cuTensorMapEncodeTiled failed (%d) for T=%d N=%d pitch=%lld", (int)r, T, N,
                      (long long)pitch);
            return TKB_EINVAL;
        }
    }

    // Launch geometry.  One launch takes Nl <= 128 tracks: S = ceil(Nl/4) solver CTAs and H = SMs - S streaming
    // CTAs; the streaming CTAs own the columns that have a far field in chunks of two, round-robin.  The tracks are
    // split over more launches until a streaming CTA's items (column, 4 tracks) fit its threads and its FIFO.
    const int far_cols = nb - ND - 1 > 0 ? ((nb - ND - 1) * BX < T ? (nb - ND - 1) * BX : T) : 0;
    const int chunks = (far_cols + CW - 1) / CW;
    int launches = (N + NLMAX - 1) / NLMAX;
    int Nl = 0, S = 0, H = 0, nslots = 0, nitems = 0, NI = 8, nstg = 2;
    for (;; ++launches) {
        Nl = ((N + launches - 1) / launches + 3) / 4 * 4;
        if (Nl > NLMAX) continue;
        S = (Nl + NQ - 1) / NQ;
        if (S > sms - 1 && chunks > 0) continue;
        H = sms - S;
        if (H > chunks) H = chunks;
        if (H < 0) H = 0;
        nslots = H > 0 ? (chunks + H - 1) / H : 0;
        nitems = nslots * CW * (Nl / 4);
        NI = nitems > 8 ? (nitems + 7) / 8 * 8 : 8;
        nstg = (int)(fifo_budget(Nl) / ((size_t)PB * NI * 16));
        if (nstg > MAXSTG) nstg = MAXSTG;
        if ((nitems <= NCONS * IPT && nstg >= 2) || Nl <= 4) break;
    }
    if (nitems > NCONS * IPT || nstg < 2) {
        set_error("tkb_semicrf_sweep: T=%d is too long for the streaming CTAs of this device", T);
        return TKB_ELAUNCH;
    }
    p.nslots = nslots;
    p.nitems = nitems;
    p.NI = NI;
    p.nstg = nstg;
    for (int n_lo = 0; n_lo < N; n_lo += Nl) {
        p.n_lo = n_lo;
        p.Nl = (N - n_lo) < Nl ? (N - n_lo) : Nl;
        p.S = (p.Nl + NQ - 1) / NQ;
        p.H = (g_dbg_flags & 64) ? 0 : H;   // diagnostics: solver CTAs only (replay launches)
        p.nitems = nslots * CW * ((p.Nl + 3) / 4);
        const int grid = p.S + p.H;
        const int rc = direction == TKB_BACKWARD ? launch_align<TKB_BACKWARD>(align, flags, p, map, grid, stream)
                                                 : launch_align<TKB_FORWARD>(align, flags, p, map, grid, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

extern "C" int tkb_sweep_status(const void *workspace, int *status_host, void *stream_) {
    if (!workspace || !status_host) {
        set_error("tkb_sweep_status: null pointer");
        return TKB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    TKB_CUDA(cudaMemcpyAsync(status_host, workspace, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TKB_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

// diagnostics build only (compile with -DTKB_TIMELINE): device buffer of [grid][64][4] globaltimer stamps
extern "C" void tkb_debug_set_timeline(unsigned long long *buf) { g_timeline = buf; }
extern "C" void tkb_debug_set_flags(int flags) { tkb::g_dbg_flags = flags; }
