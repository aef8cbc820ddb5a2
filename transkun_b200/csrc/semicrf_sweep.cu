// semicrf_sweep.cu -- the semi-Markov dynamic programme as ONE persistent kernel (solver / helper CTAs).
//
// Replaces the TorchScript step loops of the reference
// (transkun/CRF/NeuralSemiCRFInterval.py:31-51, :124-144, :218-234, :303-327):
//     q[x] = ( skip(x)  (+)  (+)_{y>x} q[y] (x) S(y,x) )  (x)  unary(x)
// over the (max,+) semiring (Viterbi, bit-exact fp32: one add per candidate, exact max, the
// reference's tie order) and the (logsumexp,+) semiring (log-partition), both fed by a single
// read of the score triangle.
//
// Mirrored coordinates.  x is the position being solved, y > x a solved one.
//   BACKWARD: x = begin b, y = end e, S(y,x) = score[e][b]      (sx = N,    sy = T*N)
//   FORWARD : x = T-1-end, y = T-1-begin, S(y,x) = score[T-1-x][T-1-y]
//                                                               (sx = -T*N, sy = -N)
// so one kernel serves viterbiBackward/beta and viterbi/alpha.
//
// It is a lower-triangular solve: T strictly sequential steps per track.  The design keeps that
// chain inside ONE SM from the first to the last position, and lets every other SM stream the
// triangle:
//   * tracks are independent; a GROUP is 8 tracks = one 32-byte sector of the track-innermost layout;
//   * per group two SOLVER CTAs (4 tracks = 16 bytes each) run the chain: one warp per track, lane =
//     column of the current 32-column block, V and L semirings interleaved in the same instruction
//     stream.  A chain step broadcasts the just-finished row with shuffles and pushes it into the
//     current block (the diagonal tile, on the chain) and into the next ND blocks (off the chain); the
//     score values come from a shared-memory ring of "row bands" (32 rows x (ND+1)*32 columns x 16 B)
//     that four loader warps of the same CTA keep filled with cp.async, mbarrier-synchronised;
//   * per group H HELPER CTAs own the column blocks round-robin and stream everything further than ND
//     blocks above the diagonal (the bulk of the bytes): 16 warps, each every 16th pair of rows,
//     cp.async FIFOs, register accumulators, merged once per block and handed to the solvers as a
//     "far partial";
//   * rows travel solver -> helpers through a global-memory mailbox of 64-bit words {fp32 value, epoch},
//     far partials travel helper -> solver the same way: one relaxed store publishes, one relaxed load
//     observes (no fence, no flag, no reset; the epoch grows with every launch).
// All CTAs of a launch must be co-resident (cooperative launch).
#include <stdlib.h>

#include "common.cuh"

namespace tkb {

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
constexpr int NW = 16;     // warps per CTA (helper: 16 row slices; solver: 2*NQ chain + 4 loader + 1 prep warps)
constexpr int NT = NW * 32;
constexpr int NCW = 2 * NQ;  // chain warps: (track, semiring)
constexpr int NLW = 4;       // loader warps
constexpr int SLOTS = 4;   // helper: per-warp FIFO depth in row PAIRS; SLOTS-1 pairs in flight
// prep slot of one track: nine arrays of 32 floats
constexpr int PR_DR = 0, PR_ETA = 32, PR_FARV = 64, PR_FARS = 96, PR_SP2 = 128, PR_COMB = 160, PR_ETA2 = 192,
              PR_FARM = 224, PR_FARL = 256, PR_FLOATS = 288;

// helper shared memory: per-warp S FIFO (1 KB per row) | per-warp mailbox-row FIFO (tagged) | untagged copy.
// After its far field a warp reuses its own (drained) S FIFO for the partial accumulators it hands to the
// merge: [2 semirings][NG][BX] float2 = 4 KB of its 8 KB.
constexpr size_t kRingFloatsPerWarp = (size_t)SLOTS * 2 * 2 * 32 * 4;  // [slot][row][col][lane] float4
constexpr size_t kRingFloats = (size_t)NW * kRingFloatsPerWarp;
constexpr size_t kQWordsPerWarp = (size_t)SLOTS * 32;    // [slot][row][kind][track] tagged words
constexpr size_t kQcFloatsPerWarp = (size_t)SLOTS * 32;  // untagged copy, same layout
constexpr size_t kHelperSmem = kRingFloats * 4 + (size_t)NW * kQWordsPerWarp * 8 + (size_t)NW * kQcFloatsPerWarp * 4;
// solver shared memory: row bands [NBAND][NQ tracks][BX rows][BANDCOLS] (planar per track: a chain warp reads
// consecutive words) | prep slots [NPREP][NQ][PR_FLOATS] | mbarriers band full/empty, prep full/empty
constexpr size_t kBandBytes = (size_t)NQ * BX * BANDCOLS * 4;
// | publish ring [NCW chain warps][2 blocks][BX] 8-byte results | mbarriers pub full[NCW][2 blocks][8 micro-blocks]
// (one outstanding phase each), pub empty[NCW][2]
constexpr size_t kSolverSmem = (size_t)NBAND * kBandBytes + (size_t)NPREP * NQ * PR_FLOATS * 4 +
                               (size_t)NCW * 2 * BX * 8 + 2 * NBAND * 8 + 2 * NPREP * 8 + NCW * 16 * 8 + NCW * 2 * 8;
// feeder: ring of row stages [FST][BANDCOLS cells][cell stride <= FCS floats]
constexpr int RB = 16;    // band slots of the global (L2-resident) ring the feeders fill and the solvers read
constexpr int FST = 4;    // feeder: rows in flight
constexpr int NTC = 96;   // feeder: tracks per pass
constexpr int FCS = NTC + 4;
constexpr size_t kFeederSmem = (size_t)FST * BANDCOLS * FCS * 4;
constexpr size_t kSweepSmem0 = kHelperSmem > kSolverSmem ? kHelperSmem : kSolverSmem;
constexpr size_t kSweepSmem = kSweepSmem0 > kFeederSmem ? kSweepSmem0 : kFeederSmem;
static_assert(kRingFloatsPerWarp * 4 >= 2 * NG * BX * 8, "partials must fit the warp's own FIFO");
static_assert(kSweepSmem <= 227 * 1024, "shared memory budget");
static_assert(NCW + NLW + 1 + NCW <= NW, "solver warp roles");

constexpr size_t kHeaderBytes = 256;  // status word lives here

struct SweepParams {
    const float *Sbase;    // &S(0,0) in mirrored coordinates
    const float *etabase;  // &skip weight of x = 0
    long long sx, sy, se;  // element strides
    int T, N, Npad, G, H, g0, dir;
    int F, gcount;         // feeder CTAs of this launch (they follow the gcount * (NSOLV + H) solver/helper CTAs)
    int nlo, nhi, SN;      // tracks [nlo, nhi) of this launch; SN = track stride of the scratch ring
    unsigned btag;         // launch index << 16: upper half of a band flag's low word
    unsigned epoch;
    unsigned long long *mbox;  // [2 semirings][T][Npad] {value, epoch}
    unsigned long long *part;  // [G][nb][2 semirings][NG][BX][2] {value, epoch}: far partials
    float *scratch;             // [RB band slots][BX rows][SN tracks][BANDCOLS]: near bands re-laid out per track
    unsigned long long *bflag;  // [RB] {btag | band index + 1, epoch}: band slot is complete
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
// HELPER: far partial of the owned column blocks
// =================================================================================================
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void helper_role(const SweepParams &p, unsigned char *smem_raw, int g, int h) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    constexpr bool A16 = ALIGN == 16;
    constexpr int D = SLOTS - 1;

    float *ring = reinterpret_cast<float *>(smem_raw);
    unsigned long long *qring = reinterpret_cast<unsigned long long *>(ring + kRingFloats);
    float *qcomp = reinterpret_cast<float *>(qring + (size_t)NW * kQWordsPerWarp);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n0 = g * NG;
    const unsigned epoch = p.epoch;
    unsigned long long *mboxV = p.mbox;
    unsigned long long *mboxL = p.mbox + (size_t)T * p.Npad;

    // far-field mapping: lane -> (column pair, track quad)
    const int cpair = lane >> 1, quad = lane & 1;
    const int nq = n0 + quad * 4;
    const int nvalid = min(max(N - nq, 0), 4);
    float *my_ring = ring + (size_t)warp * kRingFloatsPerWarp + lane * 4;  // + slot*256 (+128 for column 1)
    unsigned long long *my_q = qring + (size_t)warp * kQWordsPerWarp;
    float *my_qc = qcomp + (size_t)warp * kQcFloatsPerWarp;
    float2 *my_partV = reinterpret_cast<float2 *>(ring + (size_t)warp * kRingFloatsPerWarp);  // [NG][BX]
    float2 *my_partL = my_partV + NG * BX;
    // merge mapping: warp -> (semiring, track), lane -> column
    const int sn = warp & 7;
    const bool s_is_lse = warp >= 8;
    const long long row_step = (long long)NW * p.sy;
    const long long q_step = (long long)NW * p.Npad;

    int owned_idx = 0;
    for (int J = nb - ND - 2 - h; J >= 0; J -= p.H, ++owned_idx) {
        const int x0 = J * BX;
        if (threadIdx.x == 0) TKB_STAMP(owned_idx, 0);
        float vmax[2][4], lM[2][4], lS[2][4];
        int vsel[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                vmax[j][q] = -INFINITY;
                vsel[j][q] = -1;
                lM[j][q] = -FLT_MAX;
                lS[j][q] = 0.0f;
            }
        const int R = T - (x0 + (ND + 1) * BX);   // rows y = T-1 .. x0+(ND+1)*BX, taken in adjacent pairs (R >= 1)
        const int npairs = (R + 1) >> 1;          // pair pr = rows (T-1-2pr, T-2-2pr); the last may be half
        const int mypairs = npairs > warp ? (npairs - warp + NW - 1) / NW : 0;
        {
            // running source pointers of the next pair to issue (all 32 columns are valid here); pairs past
            // the end are issued with src-size 0 (no global access), so the loop body has no branches
            const float *sp0 = p.Sbase;
            if (nvalid > 0)
                sp0 = p.Sbase + (long long)(x0 + 2 * cpair) * p.sx + (long long)(T - 1 - 2 * warp) * p.sy + nq;
            const long long sstep = nvalid > 0 ? 2 * row_step : 0;
            const long long scol = nvalid > 0 ? p.sx : 0;
            const long long srow = nvalid > 0 ? p.sy : 0;
            // mailbox fetch: lane = row*8 + kind*4 + track pair (lanes 0-15); tag check: lane = row*16 + kind*8 + track
            const int f_row = lane >> 3, f_kind = (lane >> 2) & 1;
            const bool qfetch = lane < 16 && (f_kind ? DO_L : DO_V);
            const unsigned long long *qp =
                (f_kind ? mboxL : mboxV) + (long long)(T - 1 - 2 * warp - f_row) * p.Npad + n0 + 2 * (lane & 3);
            const int c_row = lane >> 4, c_kind = (lane >> 3) & 1;
            const bool c_need = (c_kind ? DO_L : DO_V) && (n0 + (lane & 7)) < N;  // padding tracks are never published
            const unsigned long long *cq = (c_kind ? mboxL : mboxV) + n0 + (lane & 7);  // + y * Npad
            const float c_absent = c_kind ? -FLT_MAX : -INFINITY;  // q of a row that does not exist
            const int nbytes = nvalid * 4;
            const unsigned ring_s = smem_u32(my_ring);  // + slot*2048 + row*1024 + col*512
            const unsigned q_s = smem_u32(my_q);        // + slot*256: tagged words [row][kind][track]
            const unsigned qc_s = smem_u32(my_qc);      // + slot*128: untagged values [row][kind][track]
            int ti = 0;  // next pair to issue
            auto issue = [&]() {
                const int live = ti < mypairs;
                const int liveB = live && (2 * (warp + ti * NW) + 1 < R);
                const unsigned so = (unsigned)(ti & (SLOTS - 1)) * 2048u;
                if (A16) {
                    cp_async16_s(ring_s + so, sp0, live ? nbytes : 0);
                    cp_async16_s(ring_s + so + 512, sp0 + scol, live ? nbytes : 0);
                    cp_async16_s(ring_s + so + 1024, sp0 - srow, liveB ? nbytes : 0);
                    cp_async16_s(ring_s + so + 1536, sp0 - srow + scol, liveB ? nbytes : 0);
                } else if (ALIGN == 8) {
#pragma unroll
                    for (int q = 0; q < 4; q += 2) {
                        const int qq = q < nvalid ? q : 0;
                        const int nA = (live && q < nvalid) ? 8 : 0, nB = (liveB && q < nvalid) ? 8 : 0;
                        cp_async8_s(ring_s + so + q * 4, sp0 + qq, nA);
                        cp_async8_s(ring_s + so + 512 + q * 4, sp0 + scol + qq, nA);
                        cp_async8_s(ring_s + so + 1024 + q * 4, sp0 - srow + qq, nB);
                        cp_async8_s(ring_s + so + 1536 + q * 4, sp0 - srow + scol + qq, nB);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int qq = q < nvalid ? q : 0;
                        const int nA = (live && q < nvalid) ? 4 : 0, nB = (liveB && q < nvalid) ? 4 : 0;
                        cp_async4_s(ring_s + so + q * 4, sp0 + qq, nA);
                        cp_async4_s(ring_s + so + 512 + q * 4, sp0 + scol + qq, nA);
                        cp_async4_s(ring_s + so + 1024 + q * 4, sp0 - srow + qq, nB);
                        cp_async4_s(ring_s + so + 1536 + q * 4, sp0 - srow + scol + qq, nB);
                    }
                }
                if (qfetch) cp_async16_s(q_s + (so >> 3) + lane * 16, qp, (f_row ? liveB : live) ? 16 : 0);
                sp0 -= sstep;
                qp -= 2 * q_step;
                ++ti;
            };
#pragma unroll
            for (int t = 0; t < D; ++t) {
                issue();
                cp_async_commit();
            }
            int yA = T - 1 - 2 * warp;
            // one pair of rows: wait for S and the mailbox words, validate the tags, distribute q, Viterbi update,
            // and (log-sum) stage x = S*log2e + q for the pair flush
            auto do_pair = [&](int t, float (&xlA)[2][4], float (&xlB)[2][4]) {
                issue();
                cp_async_commit();
                cp_async_wait<D>();
                __syncwarp();
                const unsigned so = (unsigned)(t & (SLOTS - 1));
                const bool hasB = 2 * (warp + t * NW) + 1 < R;
                unsigned long long word = lds64(q_s + so * 256 + lane * 8);
                const bool need = c_need && (c_row == 0 || hasB);
                const bool ok = !need || (unsigned)(word >> 32) == epoch;
                if (!__all_sync(kFull, ok)) {  // row not published when prefetched: poll it now
                    if (!ok) {
                        const unsigned long long *w = cq + (long long)(yA - c_row) * p.Npad;
                        word = (yA < x0 + (ND + 3) * BX) ? poll_slow<0>(w, epoch, p.status)
                                                         : poll_slow<TKB_FAR_BACKOFF_NS>(w, epoch, p.status);
                    }
                }
                const float qrow = (c_row && !hasB) ? c_absent : __uint_as_float((unsigned)word);
                sts32(qc_s + so * 128 + lane * 4, qrow);
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {  // row A = yA, then row B = yA - 1 (descending y: tie order)
                    float4 qv4, ql4;
                    if (DO_V) qv4 = lds128(qc_s + so * 128 + rr * 64 + quad * 16);
                    if (DO_L) ql4 = lds128(qc_s + so * 128 + rr * 64 + 32 + quad * 16);
                    const float4 a0 = lds128(ring_s + so * 2048 + rr * 1024);
                    const float4 a1 = lds128(ring_s + so * 2048 + rr * 1024 + 512);
                    const float qv[4] = {qv4.x, qv4.y, qv4.z, qv4.w};
                    const float ql[4] = {ql4.x, ql4.y, ql4.z, ql4.w};
                    const float av[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}};
                    const int y = yA - rr;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (DO_V) {
                                const float xv = qv[q] + av[j][q];
                                const bool tk = (DIR == TKB_BACKWARD) ? (xv >= vmax[j][q]) : (xv > vmax[j][q]);
                                vmax[j][q] = tk ? xv : vmax[j][q];
                                vsel[j][q] = tk ? y : vsel[j][q];
                            }
                            if (DO_L) (rr ? xlB : xlA)[j][q] = fmaf(av[j][q], kLog2e, ql[q]);
                        }
                }
                yA -= 2 * NW;
            };
            for (int t = 0; t < mypairs; ++t) {  // one max/rescale per pair of rows
                float xl[2][2][4];
                do_pair(t, xl[0], xl[1]);
                if (DO_L) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float m = fmaxf(xl[0][j][q], xl[1][j][q]);
                            const float Mn = fmaxf(lM[j][q], m);
                            float acc = lS[j][q] * ex2f(lM[j][q] - Mn);
                            acc += ex2f(xl[0][j][q] - Mn);
                            acc += ex2f(xl[1][j][q] - Mn);
                            lS[j][q] = acc;
                            lM[j][q] = Mn;
                        }
                }
            }
        }
        if (threadIdx.x == 0) TKB_STAMP(owned_idx, 1);
        // ---- hand the 16 partials to the merge mapping (via this warp's drained FIFO) --------
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = (quad * 4 + q) * BX + 2 * cpair + j;
                if (DO_V) my_partV[o] = make_float2(vmax[j][q], __int_as_float(vsel[j][q]));
                if (DO_L) my_partL[o] = make_float2(lM[j][q], lS[j][q]);
            }
        __syncthreads();
        if (threadIdx.x == 0) TKB_STAMP(owned_idx, 2);
        const int c = lane;
        unsigned long long *dst =
            p.part + ((((size_t)g * nb + J) * 2 + (s_is_lse ? 1 : 0)) * NG + sn) * (BX * 2) + 2 * c;
        if (!s_is_lse && DO_V) {
            // branch-free 16-way merge: the maximum, then among the partials that attain it the row the
            // reference's candidate order prefers (BACKWARD: smallest y, FORWARD: largest y).  Empty partials
            // are (-inf, -1); (unsigned)-1 is the largest unsigned, so they never win the min.
            float pv[NW];
            int ps[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float2 e = reinterpret_cast<const float2 *>(ring + (size_t)w * kRingFloatsPerWarp)[sn * BX + c];
                pv[w] = e.x;
                ps[w] = __float_as_int(e.y);
            }
            float best = pv[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) best = fmaxf(best, pv[w]);
            int bsel;
            if (DIR == TKB_BACKWARD) {
                unsigned m = 0xffffffffu;
#pragma unroll
                for (int w = 0; w < NW; ++w) m = min(m, pv[w] == best ? (unsigned)ps[w] : 0xffffffffu);
                bsel = (int)m;
            } else {
                int m = -1;
#pragma unroll
                for (int w = 0; w < NW; ++w) m = max(m, pv[w] == best ? ps[w] : -1);
                bsel = m;
            }
            publish(dst, best, epoch);
            publish(dst + 1, __int_as_float(bsel), epoch);
        } else if (s_is_lse && DO_L) {
            float M = -FLT_MAX, S = 0.0f;
            float m[NW], s[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float2 e =
                    reinterpret_cast<const float2 *>(ring + (size_t)w * kRingFloatsPerWarp)[(NG + sn) * BX + c];
                m[w] = e.x;
                s[w] = e.y;
                M = fmaxf(M, e.x);
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) S += s[w] * ex2f(m[w] - M);
            publish(dst, M, epoch);
            publish(dst + 1, S, epoch);
        }
        if (threadIdx.x == 0) TKB_STAMP(owned_idx, 3);
        __syncthreads();  // the partials live in the FIFOs the next owned block refills
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
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n = g * NG + ptrk;
    const unsigned epoch = p.epoch;
    unsigned long long *mV = p.mbox + n;  // + y * Npad
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
        const bool active = x < T;
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
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n = g * NG + ptrk;
    const unsigned epoch = p.epoch;
    unsigned long long *mL = p.mbox + (size_t)T * p.Npad + n;
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
        const bool active = x < T;
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
        // band of row block j: rows y = 32j .. 32j+31, columns x = 32(j-ND) .. 32j+31.  The feeders have re-laid
        // it out per track in the global ring (p.scratch), so a (track, row) is one contiguous run of BANDCOLS
        // floats: a few coalesced 16-byte cp.async per row instead of a 32-byte-sector gather from the score
        // tensor (whose L1 wavefronts used to starve the chain warps of the same SM).
        const int lw = warp - NCW;
        constexpr int RUN16 = BANDCOLS / 4;  // 16-byte pieces per (track, row) run
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int slot = it % NBAND;
            if (it >= NBAND) mbar_wait_relaxed(sm.band_empty + slot * 8, ((it / NBAND) - 1) & 1, p.status);
            {   // the feeder's flag: {band + 1, epoch}
                const unsigned long long want = ((unsigned long long)p.epoch << 32) | p.btag | (unsigned)(j + 1);
                const unsigned long long *fl = p.bflag + (j % RB);
                if (ld_acquire_u64(fl) != want) poll_flag_slow(fl, want, p.status);
            }
            const float *src0 = p.scratch + (size_t)(j % RB) * BX * p.SN * BANDCOLS;
            const unsigned dst0 = sm.band + (unsigned)(slot * kBandBytes);
            for (int i = lw * 32 + lane; i < nvalid * BX * RUN16; i += NLW * 32) {
                const int tr = i / (BX * RUN16), rem = i - tr * (BX * RUN16);
                const int e = rem / RUN16, pc = rem - e * RUN16;
                cp_async16_s(dst0 + (unsigned)((tr * BX + e) * BANDCOLS + pc * 4) * 4u,
                             src0 + ((size_t)e * p.SN + (n0 + tr - p.nlo)) * BANDCOLS + pc * 4, 16);
            }
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
            unsigned long long fw[NQ][4];
            const bool has_far = j <= nb - ND - 2;
#pragma unroll
            for (int tr = 0; tr < NQ; ++tr) {
                sd[tr] = se[tr] = ssub[tr] = 0.0f;
                fw[tr][0] = fw[tr][1] = fw[tr][2] = fw[tr][3] = 0;
                if (tr < nvalid && x < T) {
                    const int n = n0 + tr;
                    sd[tr] = __ldg(p.Sbase + (long long)x * (p.sx + p.sy) + n);
                    if (x < T - 1) {
                        se[tr] = __ldg(p.etabase + (long long)x * p.se + n);
                        if (DO_L) ssub[tr] = __ldg(p.Sbase + (long long)x * p.sx + (long long)(x + 1) * p.sy + n);
                    }
                }
            }
            if (has_far) {
#pragma unroll
                for (int tr = 0; tr < NQ; ++tr)
                    if (tr < nvalid) {
                        const unsigned long long *src =
                            p.part + ((((size_t)g * nb + j) * 2) * NG + (qd * NQ + tr)) * (BX * 2) + 2 * c;
                        if (DO_V) {
                            fw[tr][0] = ld_relaxed_u64(src);
                            fw[tr][1] = ld_relaxed_u64(src + 1);
                        }
                        if (DO_L) {
                            fw[tr][2] = ld_relaxed_u64(src + (size_t)NG * BX * 2);
                            fw[tr][3] = ld_relaxed_u64(src + (size_t)NG * BX * 2 + 1);
                        }
                    }
            }
            if (it >= NPREP) mbar_wait_relaxed(sm.prep_empty + ps * 8, ((it / NPREP) - 1) & 1, p.status);
#pragma unroll
            for (int tr = 0; tr < NQ; ++tr)
                if (tr < nvalid) {
                    float *pr = preps + (size_t)(ps * NQ + tr) * PR_FLOATS;
                    if (has_far) {
                        const unsigned long long *src =
                            p.part + ((((size_t)g * nb + j) * 2) * NG + (qd * NQ + tr)) * (BX * 2) + 2 * c;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            if (!((w < 2) ? DO_V : DO_L)) continue;
                            const unsigned long long *a = src + (w >= 2 ? (size_t)NG * BX * 2 : 0) + (w & 1);
                            if ((unsigned)(fw[tr][w] >> 32) != p.epoch) fw[tr][w] = poll_slow<100>(a, p.epoch, p.status);
                        }
                        if (DO_V) {
                            pr[PR_FARV + c] = __uint_as_float((unsigned)fw[tr][0]);
                            pr[PR_FARS + c] = __uint_as_float((unsigned)fw[tr][1]);
                        }
                        if (DO_L) {
                            pr[PR_FARM + c] = __uint_as_float((unsigned)fw[tr][2]);
                            pr[PR_FARL + c] = __uint_as_float((unsigned)fw[tr][3]);
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
        unsigned long long *mb = p.mbox + (kind ? (size_t)T * p.Npad : 0) + n;
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
                        publish(mb + (size_t)x * p.Npad, qfin, p.epoch);
                        p.code[(size_t)n * T + pos] = (unsigned)(w >> 32);
                        if (p.outv) p.outv[(size_t)pos * N + n] = qfin;
                    } else {
                        const float v2 = __uint_as_float((unsigned)w) + lg2f(__uint_as_float((unsigned)(w >> 32)));
                        publish(mb + (size_t)x * p.Npad, v2, p.epoch);
                        if (p.outl) p.outl[(size_t)pos * N + n] = v2 * kLn2;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.pub_empty + (cwi * 2 + (it & 1)) * 8);
        }
        return;
    }
}

// =================================================================================================
// FEEDER: stream the near band (all tracks, contiguous rows) and re-lay it out per track
// =================================================================================================
// Band j = rows 32j..32j+31 x columns 32(j-ND)..32j+31.  In the score tensor a row of it is one contiguous run
// (BACKWARD) of BANDCOLS * N floats, but one track's share is 4 bytes out of every 4N.  A feeder CTA reads whole
// rows with coalesced cp.async into a shared-memory stage (cell stride padded so that a 16-byte column read is
// conflict-free), and writes scratch[slot][e][track][cc] with 128-byte stores.  Cells that do not exist (x < 0,
// x > y, y >= T) are written as zeros.  Feeders depend on nothing but the ring: slot j % RB is reused once every
// chain has finished block j + RB (its row 32(j+RB) is in the mailbox).
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void feeder_role(const SweepParams &p, unsigned char *smem_raw, int f) {
    constexpr int W = ALIGN / 4;  // tracks per cp.async
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int nb = (T + BX - 1) / BX;
    const int nlo = p.nlo, nhi = p.nhi;
    const unsigned stage_s = smem_u32(smem_raw);
    const unsigned long long *mguard = (MODE & TKB_SWEEP_VITERBI) ? p.mbox : p.mbox + (size_t)T * p.Npad;
    const int nchunk = (nhi - nlo + NTC - 1) / NTC;
    // rows are linearised as r = ((jj * nchunk) + chunk) * BX + e over this feeder's bands jj = 0, 1, ...
    const int nbands = (nb - 1 - f >= 0) ? (nb - 1 - f) / p.F + 1 : 0;
    const int nrows = nbands * nchunk * BX;
    auto decode = [&](int r, int &j, int &n0c, int &ntc, int &e) {
        e = r % BX;
        const int t = r / BX;
        const int ch = t % nchunk;
        j = nb - 1 - f - (t / nchunk) * p.F;
        n0c = nlo + ch * NTC;
        ntc = min(NTC, nhi - n0c);
    };
    auto cell_stride = [&](int ntc) {  // floats; a multiple of 4 whose quarter is odd
        int q = (ntc + 3) >> 2;
        return ((q & 1) ? q : q + 1) * 4;
    };
    auto issue = [&](int r) {
        if (r < nrows) {
            int j, n0c, ntc, e;
            decode(r, j, n0c, ntc, e);
            const int y = j * BX + e, xlo = (j - ND) * BX;
            const int cs = cell_stride(ntc);
            const int npc = (ntc + W - 1) / W;  // pieces per cell
            const unsigned dst0 = stage_s + (unsigned)((r % FST) * BANDCOLS * FCS) * 4u;
            if (y < T) {
                const int cchi = min(BANDCOLS, y - xlo + 1);  // cells with x <= y
                const int cclo = xlo < 0 ? -xlo : 0;
                for (int i = threadIdx.x; i < BANDCOLS * npc; i += NT) {
                    const int cc = i / npc, pc = i - cc * npc;
                    if (cc < cclo || cc >= cchi) continue;
                    const float *src = p.Sbase + (long long)(xlo + cc) * p.sx + (long long)y * p.sy + n0c + pc * W;
                    const unsigned dst = dst0 + (unsigned)(cc * cs + pc * W) * 4u;
                    if (ALIGN == 16) cp_async16_s(dst, src, min(W, ntc - pc * W) * 4);
                    else if (ALIGN == 8) cp_async8_s(dst, src, min(W, ntc - pc * W) * 4);
                    else cp_async4_s(dst, src, 4);
                }
            }
        }
        cp_async_commit();
    };
    for (int r = 0; r < FST - 1; ++r) issue(r);
    int jcur = -1;
    for (int r = 0; r < nrows; ++r) {
        int j, n0c, ntc, e;
        decode(r, j, n0c, ntc, e);
        if (j != jcur) {
            // ring guard: every chain must be done with band j + RB before its slot is overwritten
            if (j + RB <= nb - 1) {
                const int yg = (j + RB) * BX;
                for (int n = nlo + threadIdx.x; n < nhi; n += NT) {
                    const unsigned long long *w = mguard + (size_t)yg * p.Npad + n;
                    if ((unsigned)(ld_relaxed_u64(w) >> 32) != p.epoch) poll_slow<200>(w, p.epoch, p.status);
                }
            }
            jcur = j;
        }
        cp_async_wait<FST - 2>();
        __syncthreads();  // row r has landed for everyone; everyone is done with row r-1 (its stage is refilled next)
        issue(r + FST - 1);
        const int y = j * BX + e, xlo = (j - ND) * BX;
        const int cs = cell_stride(ntc);
        const int cchi = (y < T) ? min(BANDCOLS, y - xlo + 1) : 0;
        const int cclo = xlo < 0 ? -xlo : 0;
        const unsigned src0 = stage_s + (unsigned)((r % FST) * BANDCOLS * FCS) * 4u;
        float *dst0 = p.scratch + ((size_t)((j % RB) * BX + e) * p.SN + (n0c - nlo)) * BANDCOLS;
        const int nquad = (ntc + 3) >> 2;
        for (int i = warp; i < nquad * (ND + 1); i += NW) {
            const int q = i / (ND + 1), cc = (i - q * (ND + 1)) * BX + lane;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cc >= cclo && cc < cchi) v = lds128(src0 + (unsigned)(cc * cs + q * 4) * 4u);
            float *d = dst0 + (size_t)(q * 4) * BANDCOLS + cc;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (q * 4 + t < ntc) d[(size_t)t * BANDCOLS] = vv[t];
        }
        if (e == BX - 1 && n0c + ntc >= nhi) {  // band complete: publish its flag
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                st_relaxed_u64(p.bflag + (j % RB), ((unsigned long long)p.epoch << 32) | p.btag | (unsigned)(j + 1));
            }
        }
    }
    cp_async_wait_all();
}

template <int DIR, int ALIGN, int MODE>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int per = NSOLV + p.H;
    if ((int)blockIdx.x >= p.gcount * per) {
        feeder_role<DIR, ALIGN, MODE>(p, smem_raw, (int)blockIdx.x - p.gcount * per);
        return;
    }
    const int g = p.g0 + (int)blockIdx.x / per, role = (int)blockIdx.x % per;
    if (role < NSOLV)
        solver_role<DIR, MODE>(p, smem_raw, g, role);
    else {
#ifndef TKB_EXP_NOHELPER
        helper_role<DIR, ALIGN, MODE>(p, smem_raw, g, role - NSOLV);
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
static size_t mailbox_bytes(int T, int N) {
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG;
    return 2 * (size_t)T * npad * sizeof(unsigned long long);
}
static size_t partial_bytes(int T, int N) {
    const size_t G = (size_t)((N + NG - 1) / NG), nb = (size_t)((T + BX - 1) / BX);
    return G * nb * 2 * NG * BX * 2 * sizeof(unsigned long long);
}
constexpr int kMinFeeders = 4, kMaxFeeders = 16;
constexpr size_t kFlagBytes = 256;  // RB band flags
static_assert(RB * 8 <= kFlagBytes, "band flags");
// most groups one launch can hold: NSOLV solvers + one helper each, and the feeders
static int groups_per_launch(int sms) { return (sms - kMinFeeders) / (NSOLV + 1); }
static size_t scratch_bytes(int N, int sms) {
    const int G = (N + NG - 1) / NG, gmax = groups_per_launch(sms);
    const size_t sn = (size_t)(G < gmax ? G : gmax) * NG;
    return (size_t)RB * BX * sn * BANDCOLS * sizeof(float);
}

}  // namespace tkb

using namespace tkb;

extern "C" size_t tkb_sweep_workspace_bytes(int T, int N) {
    if (T < 1 || N < 1) return 0;
    int sms = num_sms();
    if (sms < kMinFeeders + NSOLV + 1) sms = 148;  // sized for a B200 when queried without a device
    return kHeaderBytes + mailbox_bytes(T, N) + partial_bytes(T, N) + kFlagBytes + scratch_bytes(N, sms);
}

extern "C" int tkb_semicrf_sweep(const float *score, const float *noise, int T, int N, int direction, int flags,
                                 void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit,
                                 float *out_lse, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!score || !workspace || T < 1 || N < 1 || (T > 1 && !noise) || epoch == 0 ||
        (direction != TKB_BACKWARD && direction != TKB_FORWARD) ||
        (flags & ~(TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM)) || flags == 0 ||
        ((flags & TKB_SWEEP_VITERBI) && !out_code) || (long long)T * T >= (1ll << 40)) {
        set_error("tkb_semicrf_sweep: invalid argument (T=%d N=%d dir=%d flags=%d epoch=%u)", T, N, direction,
                  flags, epoch);
        return TKB_EINVAL;
    }
    const int sms = num_sms();
    if (sms < kMinFeeders + NSOLV + 1) {
        set_error("tkb_semicrf_sweep: no CUDA device");
        return TKB_ENODEV;
    }
    SweepParams p;
    p.T = T;
    p.N = N;
    p.G = (N + NG - 1) / NG;
    p.Npad = p.G * NG;
    p.dir = direction;
    p.epoch = epoch;
    p.status = reinterpret_cast<int *>(workspace);
    p.mbox = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + kHeaderBytes);
    p.part = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + kHeaderBytes +
                                                    mailbox_bytes(T, N));
    p.bflag = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.part) + partial_bytes(T, N));
    p.scratch = reinterpret_cast<float *>(reinterpret_cast<char *>(p.bflag) + kFlagBytes);
    p.code = out_code;
    p.outv = out_vit;
    p.outl = out_lse;
    p.timeline = g_timeline;
    if (direction == TKB_BACKWARD) {
        p.Sbase = score;
        p.sx = N;
        p.sy = (long long)T * N;
        p.etabase = noise;
        p.se = N;
    } else {
        p.Sbase = score + ((long long)(T - 1) * T + (T - 1)) * N;
        p.sx = -(long long)T * N;
        p.sy = -(long long)N;
        p.etabase = noise ? noise + (long long)(T - 2) * N : nullptr;  // skip weight of x is noise[T-2-x]
        p.se = -(long long)N;
    }
    const uintptr_t addr = reinterpret_cast<uintptr_t>(score);
    const int align = (N % 4 == 0 && (addr & 15) == 0) ? 16 : ((N % 2 == 0 && (addr & 7) == 0) ? 8 : 4);
    const int nb = (T + BX - 1) / BX;
    const int hmax = nb - ND - 1 > 1 ? nb - ND - 1 : 1;  // column blocks that have a far field at all
    // groups are independent pipelines; split them over launches if one launch cannot hold them all
    const int gmax = groups_per_launch(sms);
    p.SN = (p.G < gmax ? p.G : gmax) * NG;
    int launch = 0;
    for (int g0 = 0; g0 < p.G; g0 += gmax, ++launch) {
        const int gcount = (p.G - g0) < gmax ? (p.G - g0) : gmax;
        int F = nb < kMaxFeeders ? nb : kMaxFeeders;
        int H = (sms - F) / gcount - NSOLV;
        if (H > hmax) H = hmax;
        if (H < 1) H = 1;
        const int room = sms - gcount * (NSOLV + H);
        if (F > room) F = room;
        if (F < 1) F = 1;
        p.g0 = g0;
        p.H = H;
        p.F = F;
        p.gcount = gcount;
        p.nlo = g0 * NG;
        p.nhi = (g0 + gcount) * NG < N ? (g0 + gcount) * NG : N;
        p.btag = (unsigned)launch << 16;
        const int grid = gcount * (NSOLV + H) + F;
        const int rc = direction == TKB_BACKWARD ? launch_align<TKB_BACKWARD>(align, flags, p, grid, stream)
                                                 : launch_align<TKB_FORWARD>(align, flags, p, grid, stream);
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
