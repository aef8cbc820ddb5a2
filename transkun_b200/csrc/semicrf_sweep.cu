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
//   * per group two SOLVER CTAs (4 tracks = 16 bytes each) run the chain: one warp per track and semiring
//     (the Viterbi and the log-sum chain of a track sit on the same SMSP, their dependent steps interleave),
//     lane = column of the current 32-column block.  A chain step broadcasts the just-finished row with
//     shuffles and pushes it into the current block (the diagonal tile, on the chain) and into the next ND
//     blocks (off the chain); the score values come from a shared-memory ring of "row bands" (32 rows x
//     (ND+1)*32 columns x 16 B) that four loader warps of the same CTA keep filled with cp.async,
//     mbarrier-synchronised; finished rows leave through a shared-memory ring to one publisher warp per
//     track, which does the global stores (mailbox, back-pointer codes, tables);
//   * per group H HELPER CTAs own the column blocks round-robin and stream everything further than ND
//     blocks above the diagonal (the bulk of the bytes): 16 warps, each every 16th pair of rows,
//     cp.async FIFOs, register accumulators, merged once per block and handed to the solvers as a
//     "far partial";
//   * rows travel solver -> helpers through a global-memory mailbox of 64-bit words {fp32 value, epoch},
//     far partials travel helper -> solver the same way: one relaxed store publishes, one relaxed load
//     observes (no fence, no flag, no reset; the epoch grows with every launch).
// All CTAs of a launch must be co-resident (cooperative launch).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tkb {

constexpr int NG = 8;      // tracks per group
constexpr int NQ = 4;      // tracks per solver CTA
constexpr int BX = 32;     // columns per block (= lanes of a chain warp)
#ifndef TKB_ND
#define TKB_ND 2
#endif
constexpr int ND = TKB_ND;  // blocks above the diagonal block that the solver pushes itself
constexpr int NBAND = (ND == 2) ? 4 : 3;  // row bands resident in a solver CTA
constexpr int BANDCOLS = (ND + 1) * BX;
#ifndef TKB_NW
#define TKB_NW 16
#endif
constexpr int NW = TKB_NW;  // warps per CTA (helper: NW row slices; solver: chain, loader, publisher warps)
constexpr int NT = NW * 32;
constexpr int NCW = NQ;    // chain warps
constexpr int NLW = 4;     // loader warps
#ifndef TKB_SLOTS
#define TKB_SLOTS 4
#endif
constexpr int SLOTS = TKB_SLOTS;  // helper: per-warp FIFO depth in row PAIRS; SLOTS-1 pairs in flight
constexpr int PB = 8;      // rows per publish batch
#ifndef TKB_FARFETCH_BATCH
#define TKB_FARFETCH_BATCH 3
#endif

// helper shared memory: per-warp S FIFO (1 KB per row) | per-warp mailbox-row FIFO (tagged) | untagged copy.
// After its far field a warp reuses its own (drained) S FIFO for the partial accumulators it hands to the
// merge: [2 semirings][NG][BX] float2 = 4 KB of its 8 KB.
constexpr size_t kRingFloatsPerWarp = (size_t)SLOTS * 2 * 2 * 32 * 4;  // [slot][row][col][lane] float4
constexpr size_t kRingFloats = (size_t)NW * kRingFloatsPerWarp;
constexpr size_t kQWordsPerWarp = (size_t)SLOTS * 32;    // [slot][row][kind][track] tagged words
constexpr size_t kQcFloatsPerWarp = (size_t)SLOTS * 32;  // untagged copy, same layout
constexpr size_t kHelperSmem = kRingFloats * 4 + (size_t)NW * kQWordsPerWarp * 8 + (size_t)NW * kQcFloatsPerWarp * 4;
// solver shared memory: row bands [NBAND][BX rows][BANDCOLS][NQ tracks] | mbarriers full[NBAND], empty[NBAND]
constexpr size_t kBandBytes = (size_t)BX * BANDCOLS * NQ * 4;
// | publish ring [NCW tracks][2 blocks][BX][2 semirings] 8-byte results | mbarriers pub full[NCW][2 blocks][BX/PB]
// (eight arrivals each, one outstanding phase), pub empty[NCW][2]
constexpr size_t kPubBytes = (size_t)NCW * 2 * BX * 2 * 8;
constexpr size_t kSolverSmem =
    (size_t)NBAND * kBandBytes + 2 * NBAND * 8 + kPubBytes + NCW * 2 * (BX / PB) * 8 + NCW * 2 * 8;
constexpr size_t kSweepSmem = kHelperSmem > kSolverSmem ? kHelperSmem : kSolverSmem;
static_assert(kRingFloatsPerWarp * 4 >= 2 * NG * BX * 8, "partials must fit the warp's own FIFO");
static_assert(kSweepSmem <= 227 * 1024, "shared memory budget");
static_assert(kBandBytes % 128 == 0, "TMA destination alignment");

constexpr size_t kHeaderBytes = 256;  // status word lives here

struct SweepParams {
    const float *Sbase;    // &S(0,0) in mirrored coordinates
    const float *etabase;  // &skip weight of x = 0
    long long sx, sy, se;  // element strides
    int T, N, Npad, G, H, g0, dir;
    unsigned epoch;
    unsigned long long *mbox;  // [2 semirings][T][Npad] {value, epoch}
    unsigned long long *part;  // [G][nb][2 semirings][NG][BX][2] {value, epoch}: far partials
    int *status;               // 0, or the epoch of a launch whose inter-CTA wait timed out
    unsigned *code;  // [N][T]
    float *outv;     // [T][N] or null
    float *outl;     // [T][N] or null
    unsigned long long *timeline;  // diagnostics build only (TKB_TIMELINE): [grid][64][4] globaltimer stamps
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
        if (*(volatile unsigned *)status == epoch) return 0;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, (int)epoch);
            return 0;
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
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ float lds32(unsigned saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

#ifdef TKB_TIMELINE
// [grid][64 owned blocks / 64 chain blocks][4] stamps
#define TKB_STAMP(idx, slot)                                                                    \
    do {                                                                                        \
        if (p.timeline && (idx) < 64) p.timeline[((size_t)blockIdx.x * 64 + (idx)) * 4 + (slot)] = globaltimer_ns(); \
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
                const unsigned so = (unsigned)(ti % SLOTS) * 2048u;
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
                const unsigned so = (unsigned)(t % SLOTS);
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
        if (warp >= 16) {
            // (more than 16 row slices: the extra warps have nothing to merge)
        } else if (!s_is_lse && DO_V) {
            // branch-free NW-way merge: the maximum, then among the partials that attain it the row the
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

// One chain warp: track n0 + tr, lane = column of the current block, semirings CV / CL.  When a launch computes
// both, each track has two chain warps (one per semiring) on the same SMSP: their dependent chains interleave.
struct ChainCtx {
    unsigned full_s, empty_s, pub_s, pubfull_s, pubempty_s;
    int g, qd, n0, tr;
};
template <int DIR, bool CV, bool CL>
__device__ __forceinline__ void chain_warp(const SweepParams &p, unsigned char *smem_raw, const ChainCtx &cx) {
    constexpr bool DO_V = CV, DO_L = CL;
    const int lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const unsigned epoch = p.epoch;
    const unsigned full_s = cx.full_s, empty_s = cx.empty_s, pub_s = cx.pub_s, pubfull_s = cx.pubfull_s,
                   pubempty_s = cx.pubempty_s;
    const int g = cx.g, qd = cx.qd, n0 = cx.n0, tr = cx.tr;
    (void)N;
    // ---------------- chain warp: track n, lane = column ---------------------------------------------
    const int n = n0 + tr;
    const int c = lane;
    const int ptrk = qd * NQ + tr;  // track inside the group
    (void)epoch;
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
        fsrc = p.part + ((((size_t)g * nb + jn) * 2) * NG + ptrk) * (BX * 2) + 2 * c;
        if (DO_V) {
            fw[0] = ld_relaxed_u64(fsrc);
            fw[1] = ld_relaxed_u64(fsrc + 1);
        }
        if (DO_L) {
            fw[2] = ld_relaxed_u64(fsrc + (size_t)NG * BX * 2);
            fw[3] = ld_relaxed_u64(fsrc + (size_t)NG * BX * 2 + 1);
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
        if (lane == 0 && tr == 0) TKB_STAMP(it, 0);
        // ---- far partial of this block (rows of blocks > j+ND), written by the owning helper -------------
        if (j <= nb - ND - 2) {
            if (DO_V) {
                if ((unsigned)(fw[0] >> 32) != epoch) fw[0] = poll_slow<0>(fsrc, epoch, p.status);
                if ((unsigned)(fw[1] >> 32) != epoch) fw[1] = poll_slow<0>(fsrc + 1, epoch, p.status);
                const float fv = __uint_as_float((unsigned)fw[0]);
                const int fs = (int)(unsigned)fw[1];
                // far rows are larger y than anything accumulated so far: BACKWARD prefers the smaller y on ties
                const bool tk = (DIR == TKB_BACKWARD) ? (fv > best[0]) : (fv >= best[0]);
                bsel[0] = tk ? fs : bsel[0];
                best[0] = fmaxf(best[0], fv);
            }
            if (DO_L) {
                const unsigned long long *srcL = fsrc + (size_t)NG * BX * 2;
                if ((unsigned)(fw[2] >> 32) != epoch) fw[2] = poll_slow<0>(srcL, epoch, p.status);
                if ((unsigned)(fw[3] >> 32) != epoch) fw[3] = poll_slow<0>(srcL + 1, epoch, p.status);
                lse_push(lM[0], lS[0], __uint_as_float((unsigned)fw[2]), __uint_as_float((unsigned)fw[3]));
            }
        }
        if (lane == 0 && tr == 0) TKB_STAMP(it, 1);
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
        if (lane == 0 && tr == 0) TKB_STAMP(it, 2);
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
}

// =================================================================================================
// SOLVER: the chain of NQ tracks, from the last position to the first, in one SM
// =================================================================================================
template <int DIR, int ALIGN, int MODE>
__device__ __forceinline__ void solver_role(const SweepParams &p, const CUtensorMap *map, unsigned char *smem_raw,
                                            int g, int qd) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
    // BACKWARD with a 16-byte aligned track pitch: ONE thread fills a band with ONE TMA tensor copy
    // (cp.async.bulk.tensor.3d, box {4 tracks, 96 columns, 32 rows}); the per-lane cp.async gather of round 1 touched
    // 32 different 128-byte lines per instruction and kept this SM's L1 pipe busy ~6000 of the ~7000 cycles of a block,
    // slowing the chain it shares the SM with (profiles/r02_sweep_experiments.txt: 243 -> 131 cycles per column in
    // the chain microbenchmark under load)
    constexpr bool USE_TMA = (DIR == TKB_BACKWARD) && (ALIGN == 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int nb = (T + BX - 1) / BX;
    const int n0 = g * NG + qd * NQ;           // first track of this solver
    const int nvalid = min(max(N - n0, 0), NQ);  // chain warps with a real track
    const unsigned epoch = p.epoch;
    const unsigned band_s = smem_u32(smem_raw);
    const unsigned full_s = band_s + (unsigned)(NBAND * kBandBytes);  // + slot*8
    const unsigned empty_s = full_s + NBAND * 8;
    const unsigned pub_s = empty_s + NBAND * 8;                      // [tr][block parity][c][kind] 8 bytes
    const unsigned pubfull_s = pub_s + (unsigned)kPubBytes;          // [tr][block parity][batch]
    const unsigned pubempty_s = pubfull_s + NCW * 2 * (BX / PB) * 8;  // [tr][block parity]

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAND; ++s) {
            mbar_init(full_s + s * 8, USE_TMA ? 1 : NLW * 32);
            mbar_init(empty_s + s * 8, nvalid * ((DO_V && DO_L) ? 2 : 1));
        }
        for (int s = 0; s < NCW * 2 * (BX / PB); ++s) mbar_init(pubfull_s + s * 8, PB * ((DO_V && DO_L) ? 2 : 1));  // the batch's lanes arrive
        for (int s = 0; s < NCW * 2; ++s) mbar_init(pubempty_s + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (nvalid == 0) return;

    if (warp >= NCW && warp < NCW + NLW) {
        // ---------------- loader warps: keep the ring of row bands filled ----------------------------
        // band of row block j: rows y = 32j .. 32j+31, columns x = 32(j-ND) .. 32j+31, this solver's NQ tracks;
        // chunk (e, cc) -> band + (e*BANDCOLS + cc)*16.  Chunks above the diagonal, left of column 0 or below
        // row T-1 are never read and not fetched.
        if (USE_TMA) {
            // cells outside the tensor (x < 0, y >= T, track >= N) arrive as zeros and are never read, like the cells
            // above the diagonal
            if (warp != NCW || lane != 0) return;
            for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
                const int slot = it % NBAND;
                if (it >= NBAND) mbar_wait(empty_s + slot * 8, ((it / NBAND) - 1) & 1, p.status, epoch);
                mbar_arrive_expect_tx(full_s + slot * 8, (unsigned)kBandBytes);
                tma_load_3d(band_s + (unsigned)(slot * kBandBytes), map, n0, (j - ND) * BX, j * BX, full_s + slot * 8);
            }
            return;
        }
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
        unsigned long long *mV = p.mbox + n, *mL = p.mbox + (size_t)T * p.Npad + n;
        const int bmax_top = (T - (nb - 1) * BX - 1) / PB;  // last batch index of the ragged top block
        for (int j = nb - 1, it = 0; j >= 0; --j, ++it) {
            const int x0 = j * BX;
            const int ncols = min(BX, T - x0);
            for (int e8 = (ncols - 1) / PB; e8 >= 0; --e8) {
                // batches the ragged top block skips never arrive: their barriers are one phase behind
                const unsigned npast = (unsigned)(it >> 1) - (((it & 1) == 0 && it > 0 && e8 > bmax_top) ? 1u : 0u);
                mbar_wait(pubfull_s + (unsigned)(((tr * 2 + (it & 1)) * (BX / PB) + e8) * 8), npast & 1, p.status, epoch);
                const int c = e8 * PB + lane, x = x0 + c;
                if (lane < PB && x < T) {
                    const int pos = (DIR == TKB_BACKWARD) ? x : T - 1 - x;
                    const unsigned src = pub_s + (unsigned)((((tr * 2 + (it & 1)) * BX + c) * 2) * 8);
                    if (DO_V) {
                        const unsigned long long w = lds64(src);
                        const float qfin = __uint_as_float((unsigned)w);
                        publish(mV + (size_t)x * p.Npad, qfin, epoch);
                        p.code[(size_t)n * T + pos] = (unsigned)(w >> 32);
                        if (p.outv) p.outv[(size_t)pos * N + n] = qfin;
                    }
                    if (DO_L) {
                        const unsigned long long w = lds64(src + 8);
                        const float v2 = __uint_as_float((unsigned)w) + lg2f(__uint_as_float((unsigned)(w >> 32)));
                        publish(mL + (size_t)x * p.Npad, v2, epoch);
                        if (p.outl) p.outl[(size_t)pos * N + n] = v2 * kLn2;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pubempty_s + (tr * 2 + (it & 1)) * 8);
        }
        return;
    }
    // ---------------- chain warps ------------------------------------------------------------------------
    ChainCtx cx;
    cx.full_s = full_s;
    cx.empty_s = empty_s;
    cx.pub_s = pub_s;
    cx.pubfull_s = pubfull_s;
    cx.pubempty_s = pubempty_s;
    cx.g = g;
    cx.qd = qd;
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
    const int per = 2 + p.H;
    const int g = p.g0 + (int)blockIdx.x / per, role = (int)blockIdx.x % per;
    if (role < 2)
        solver_role<DIR, ALIGN, MODE>(p, &map, smem_raw, g, role);
    else
        helper_role<DIR, ALIGN, MODE>(p, smem_raw, g, role - 2);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
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
static int num_sms() {
    static int sms[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}
static size_t mailbox_bytes(int T, int N) {
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG;
    return 2 * (size_t)T * npad * sizeof(unsigned long long);
}
static size_t partial_bytes(int T, int N) {
    const size_t G = (size_t)((N + NG - 1) / NG), nb = (size_t)((T + BX - 1) / BX);
    return G * nb * 2 * NG * BX * 2 * sizeof(unsigned long long);
}

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
    if (sms < 3) {
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
    const int hmax = nb - ND - 1 > 1 ? nb - ND - 1 : 1;  // column blocks that have a far field at all
    // groups are independent pipelines; split them over launches if there are more groups than SMs / 3
    const int gmax = sms / 3;
    for (int g0 = 0; g0 < p.G; g0 += gmax) {
        const int gcount = (p.G - g0) < gmax ? (p.G - g0) : gmax;
        int H = sms / gcount - 2;
        if (H > hmax) H = hmax;
        if (H < 1) H = 1;
        p.g0 = g0;
        p.H = H;
        const int grid = gcount * (2 + H);
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
