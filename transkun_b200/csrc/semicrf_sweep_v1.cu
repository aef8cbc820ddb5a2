// semicrf_sweep.cu -- the semi-Markov dynamic programme as ONE persistent kernel.
//
// Replaces the TorchScript step loops of the reference
// (transkun/CRF/NeuralSemiCRFInterval.py:31-51, :124-144, :218-234, :303-327):
//     q[x] = ( skip(x)  (+)  (+)_{y>x} q[y] (x) S(y,x) )  (x)  unary(x)
// over the (max,+) semiring (Viterbi, bit-exact fp32: one add per candidate,
// exact max, the reference's tie order) and the (logsumexp,+) semiring
// (log-partition), both fed by a single read of the score triangle.
//
// Mirrored coordinates.  x is the position being solved, y > x a solved one.
//   BACKWARD: x = begin b, y = end e, S(y,x) = score[e][b]      (sx = N,    sy = T*N)
//   FORWARD : x = T-1-end, y = T-1-begin, S(y,x) = score[T-1-x][T-1-y]
//                                                               (sx = -T*N, sy = -N)
// so one kernel serves viterbiBackward/beta and viterbi/alpha.
//
// It is a lower-triangular solve, not a map: T strictly sequential steps.
// Decomposition:
//   * tracks are independent -> groups of NG=8 tracks (one 32-byte sector of the
//     track-innermost layout) form independent pipelines;
//   * per group, K CTAs own the 32-column blocks round-robin (block J -> CTA
//     (nb-1-J) mod K).  For its block J a CTA runs three phases:
//       A  FAR FIELD: all rows y in later blocks (order-free semiring mat-vec, the bulk
//          of the bytes): 16 warps each own every 16th PAIR of adjacent rows,
//          cp.async-staged into per-lane shared-memory FIFOs together with the
//          mailbox rows, accumulators in registers;
//       B  the 16 partials are merged into the solver mapping: one warp per
//          (track, semiring), lane = column;
//       D  DIAGONAL SOLVE: 31 dependent steps, one shuffle each, branch-free.  The
//          log-sum chain carries (M, S) pairs (value = M + log2 S) so no log sits on
//          the chain, the skip weight is folded into the coefficient of the row right
//          above a column, and rows are published in batches of 8 (one lg2 per batch);
//   * solved rows are broadcast to the other CTAs of the group through a
//     global-memory mailbox of 64-bit words {fp32 value, epoch tag}: one relaxed
//     store publishes, one relaxed load observes (no fence, no flag, no reset).
// All CTAs of a launch must be co-resident (cooperative launch).
#include "common.cuh"

namespace tkb {
namespace v1 {

constexpr int NG = 8;      // tracks per group
constexpr int BX = 32;     // columns per block (= lanes of a solver warp)
constexpr int NW = 16;     // warps per CTA: far field 16 row-slices; solve 8 tracks x 2 semirings
constexpr int NT = NW * 32;
constexpr int SLOTS = 4;   // per-warp FIFO depth in row PAIRS; SLOTS-1 pairs in flight
constexpr int CH = 2;      // rows per log-sum-exp rescale chunk (= one row pair)
constexpr int PB = 8;      // rows per publish batch

// shared memory: per-warp S FIFO (1 KB per row) | per-warp mailbox-row FIFO (128 B per row) |
// diagonal block transposed to [track][row][col] | parked per-thread constants | q of the row above the block.
// After its far field a warp reuses its own (drained) S FIFO for the partial accumulators it hands to the
// solver: [2 semirings][NG][BX] float2 = 4 KB of its 8 KB.
constexpr size_t kRingFloatsPerWarp = (size_t)SLOTS * 2 * 2 * 32 * 4;  // [slot][row][col][lane] float4
constexpr size_t kRingFloats = (size_t)NW * kRingFloatsPerWarp;
constexpr size_t kQWordsPerWarp = (size_t)SLOTS * 32;  // [slot][row][kind][track] tagged words
constexpr size_t kTileFloats = (size_t)NG * BX * BX;
constexpr size_t kQcFloatsPerWarp = (size_t)SLOTS * 32;  // untagged copy: [slot][row][kind][track]
constexpr size_t kSweepSmem =
    kRingFloats * 4 + (size_t)NW * kQWordsPerWarp * 8 + (size_t)NW * kQcFloatsPerWarp * 4 + 2 * kTileFloats * 4 +
    2 * NT * 4 + 2 * NG * 4;
static_assert(kRingFloatsPerWarp * 4 >= 2 * NG * BX * 8, "partials must fit the warp's own FIFO");

constexpr size_t kHeaderBytes = 256;  // status word lives here
#define TKB_TIMELINE_STAMPS 8

struct SweepParams {
    const float *Sbase;    // &S(0,0) in mirrored coordinates
    const float *etabase;  // &skip weight of x = 0
    long long sx, sy, se;  // element strides
    int T, N, Npad, G, K, g0, dir;
    unsigned epoch;
    unsigned long long *mbox;  // [2][T][Npad] {value, epoch}
    int *status;
    unsigned *code;  // [N][T]
    float *outv;     // [T][N] or null
    float *outl;     // [T][N] or null
    unsigned long long *timeline;  // diagnostics build only (TKB_TIMELINE): [grid][64][8] globaltimer stamps
};

// Wait until the mailbox word carries this launch's epoch.  A protocol bug (or a
// non-co-resident grid) must not hang the GPU: after ~4 s the wait gives up,
// flags the workspace and lets the kernel drain with garbage.
// BACKOFF_NS > 0 is for waits that are NOT on the critical path (far-field rows): hundreds of warps spinning
// on the few mailbox lines the chain is currently writing slow the chain's own reader and writer down.
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
#ifndef TKB_FAR_BACKOFF_NS
#define TKB_FAR_BACKOFF_NS 400
#endif
__device__ __forceinline__ void publish(unsigned long long *w, float val, unsigned epoch) {
    st_relaxed_u64(w, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(val));
}

#ifdef TKB_TIMELINE
#define TKB_STAMP(slot)                                                                                     \
    do {                                                                                                    \
        if (threadIdx.x == 0 && p.timeline && owned_idx < 64)                                               \
            p.timeline[((size_t)blockIdx.x * 64 + owned_idx) * TKB_TIMELINE_STAMPS + (slot)] = globaltimer_ns(); \
    } while (0)
// per-warp stamps: timeline + 148*64*8 words, laid out [grid][64][NW][8]
#define TKB_WSTAMP(slot)                                                                                        \
    do {                                                                                                        \
        if (lane == 0 && p.timeline && owned_idx < 64)                                                          \
            p.timeline[(size_t)148 * 64 * 8 + (((size_t)blockIdx.x * 64 + owned_idx) * NW + warp) * 8 + (slot)] = \
                globaltimer_ns();                                                                               \
    } while (0)
// stamp that cannot be taken before `dep` (a register value) is available
#define TKB_WSTAMP_DEP(slot, dep)                                                                               \
    do {                                                                                                        \
        if (lane == 0 && p.timeline && owned_idx < 64)                                                          \
            p.timeline[(size_t)148 * 64 * 8 + (((size_t)blockIdx.x * 64 + owned_idx) * NW + warp) * 8 + (slot)] = \
                globaltimer_ns() + ((__float_as_uint(dep) == 0x7fedcba9u) ? 1ull : 0ull);                       \
    } while (0)
#else
#define TKB_WSTAMP_DEP(slot, dep) \
    do {                          \
    } while (0)
#define TKB_STAMP(slot) \
    do {                \
    } while (0)
#define TKB_WSTAMP(slot) \
    do {                 \
    } while (0)
#endif

// (M, S) <- (M, S) (+) sb * 2^a        value = M + log2(S); one ex2: one of the two exponents is 0
__device__ __forceinline__ void lse_push(float &M, float &S, float a, float sb) {
    const float d = M - a;
    const float e1 = ex2f(-fabsf(d));
    S = (d < 0.0f) ? fmaf(S, e1, sb) : fmaf(sb, e1, S);
    M = fmaxf(M, a);
}

template <int DIR, bool A16, int MODE>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const SweepParams p) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;
#ifdef TKB_PREFETCH_D
    constexpr int D = TKB_PREFETCH_D;  // tuning builds
#else
    constexpr int D = SLOTS - 1;
#endif
    static_assert(D >= 1 && D <= SLOTS - 1, "prefetch distance must leave one FIFO slot free");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    unsigned long long *qring = reinterpret_cast<unsigned long long *>(ring + kRingFloats);
    float *qcomp = reinterpret_cast<float *>(qring + (size_t)NW * kQWordsPerWarp);
    float *diagS = qcomp + (size_t)NW * kQcFloatsPerWarp;  // [NG][BX rows][BX cols]
    float *diagL = diagS + kTileFloats;  // same block for the log-sum warps: *log2e, skip folded into row c+1
    float *park = diagL + kTileFloats;                                              // [2][NT] per-thread constants
    float *qtop = park + 2 * NT;                                                    // [2][NG] q of row x0+BX

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, N = p.N;
    const int g = p.g0 + (int)blockIdx.x / p.K, k = (int)blockIdx.x % p.K;
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
    unsigned long long *my_q = qring + (size_t)warp * kQWordsPerWarp;      // + slot*16 + {0..7 V, 8..15 L}
    float *my_qc = qcomp + (size_t)warp * kQcFloatsPerWarp;
    float2 *my_partV = reinterpret_cast<float2 *>(ring + (size_t)warp * kRingFloatsPerWarp);  // [NG][BX]
    float2 *my_partL = my_partV + NG * BX;
    // solver mapping: warp -> (semiring, track), lane -> column
    const int sn = warp & 7;
    const bool s_is_lse = warp >= 8;
    const bool s_nok = (n0 + sn) < N;
    unsigned long long *s_mbox = (s_is_lse ? mboxL : mboxV) + n0 + sn;  // + row * Npad
    const long long row_step = (long long)NW * p.sy;
    const long long q_step = (long long)NW * p.Npad;

    int owned_idx = 0;
    for (int J = nb - 1 - k; J >= 0; J -= p.K, ++owned_idx) {
        const int x0 = J * BX;
        const int ncols = min(BX, T - x0);
        const int nr = min(BX, max(T - (x0 + BX), 0));  // rows of block J+1
        const int c = lane;       // solver mapping: my column
        const int x = x0 + c;
        TKB_STAMP(0);

        // ---- A. far field: rows y = T-1 .. x0+BX, this warp takes every NW-th pair ------------
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
        const int R = T - (x0 + BX);              // rows y = T-1 .. x0+BX, taken in adjacent pairs
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
                (f_kind ? mboxL : mboxV) + (size_t)(T - 1 - 2 * warp - f_row) * p.Npad + n0 + 2 * (lane & 3);
            const int c_row = lane >> 4, c_kind = (lane >> 3) & 1;
            const bool c_need = c_kind ? DO_L : DO_V;
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
        // ---- 0. prefetch the diagonal block, transposed to [track][row][col]; rows beyond T (only in the
            //         last block) are filled with -inf = "no candidate"
            for (int i = threadIdx.x; i < BX * BX * NG; i += NT) {
                const int n = i & 7, cc = (i >> 3) & 31, r = i >> 8;
                if (r > cc) {
                    if (r < ncols && (n0 + n) < N)
                        cp_async4(&diagS[(n * BX + r) * BX + cc],
                                  p.Sbase + (long long)(x0 + cc) * p.sx + (long long)(x0 + r) * p.sy + n0 + n, 4);
                    else
                        diagS[(n * BX + r) * BX + cc] = -INFINITY;
                }
            }
            cp_async_commit();
            // unary + skip weights of my solver column
            // (parked in shared memory while the far field needs every register: a compiler spill would be
            // re-read from L2 on the critical path, L1 being almost entirely carved out as shared memory)
            {
                const bool has_d = x < T && s_nok, has_e = has_d && x < T - 1;
                cp_async4(&park[threadIdx.x], has_d ? p.Sbase + (long long)x * (p.sx + p.sy) + n0 + sn : p.Sbase,
                          has_d ? 4 : 0);
                cp_async4(&park[NT + threadIdx.x], has_e ? p.etabase + (long long)x * p.se + n0 + sn : p.Sbase,
                          has_e ? 4 : 0);
            }
            cp_async_commit();
            // make the diagonal block and the parked constants visible to every warp now, so that each warp can
            // do its solver set-up right after ITS far field instead of after the slowest warp's
            cp_async_wait_all();
            __syncthreads();
            if (DO_L) {
                // log-sum copy of the diagonal block, prepared cooperatively and off the critical path:
                // S*log2e, and the skip folded into the coefficient of the row right above each column:
                // v[y] + S2(y,x)  (+)  v[y] + eta2(x)  =  v[y] + log2(2^S2 + 2^eta2)
                for (int i = threadIdx.x; i < BX * BX * NG; i += NT) {
                    const int cc = i & 31, r = (i >> 5) & 31, n = i >> 10;
                    if (r > cc) {
                        float v = diagS[i] * kLog2e;
                        if (r == cc + 1 && v != -INFINITY) {
                            const float e2 = park[NT + (NG + n) * 32 + cc] * kLog2e;  // eta of column cc, track n
                            v = fmaxf(v, e2) + lg2f(1.0f + ex2f(-fabsf(v - e2)));
                        }
                        diagL[i] = v;
                    }
                }
                __syncthreads();  // diagL is read by the log-sum warps right after their own far field
            }
            int yA = T - 1 - 2 * warp;
            // one pair of rows: wait for S and the mailbox words, validate the tags, distribute q, Viterbi update,
            // and (log-sum) stage x = S*log2e + q for the chunk flush
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
                        const unsigned long long *w = cq + (size_t)(yA - c_row) * p.Npad;
                        word = (yA < x0 + 3 * BX) ? poll_slow<0>(w, epoch, p.status)
                                                  : poll_slow<TKB_FAR_BACKOFF_NS>(w, epoch, p.status);
                    }
                }
                const float qrow = (c_row && !hasB) ? c_absent : __uint_as_float((unsigned)word);
                sts32(qc_s + so * 128 + lane * 4, qrow);
                if (yA - c_row == x0 + BX) qtop[lane & 15] = qrow;  // [kind][track] of row x0+BX
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
                float xl[CH][2][4];
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
        TKB_WSTAMP(0);
        // ---- B. hand the 16 partials to the solver mapping (via this warp's drained FIFO) --------
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
        // ---- solver set-up that does not depend on the other warps: done BEFORE the barrier ---------------
        // Solver phases are branch-free: coefficients that must not act (rows at or below a column, rows or
        // columns beyond T) are -inf, so their pushes leave (best, sel) / (M, S) untouched.
        const int pos = (DIR == TKB_BACKWARD) ? x : T - 1 - x;
        const bool active = x < T;
        const bool has_next = nr > 0;  // a later block exists: the top column can skip into row x0+BX
        const float s_d = park[threadIdx.x], s_eta = park[NT + threadIdx.x];
        float sreg[BX];  // my column of the diagonal block (log2 domain for the log-sum warps)
        float u0, u1;    // Viterbi: relu(d), unused | log-sum: softplus(d)*log2e, eta*log2e
        if (!s_is_lse) {
#pragma unroll
            for (int r = 1; r < BX; ++r) sreg[r] = (r > c) ? diagS[(sn * BX + r) * BX + c] : -INFINITY;
            u0 = relu_mask(s_d);
            u1 = 0.0f;
        } else {
#pragma unroll
            for (int r = 1; r < BX; ++r) sreg[r] = (r > c) ? diagL[(sn * BX + r) * BX + c] : -INFINITY;
            {  // softplus(d)*log2e = max(d2,0) + log2(1 + 2^-|d2|), d2 = d*log2e
                const float d2 = s_d * kLog2e;
                u0 = fmaxf(d2, 0.0f) + lg2f(1.0f + ex2f(-fabsf(d2)));
            }
            u1 = s_eta * kLog2e;
        }
        TKB_STAMP(1);
        TKB_WSTAMP(1);
        __syncthreads();

        const float qnext = qtop[(s_is_lse ? 8 : 0) + sn];
        if (!s_is_lse && DO_V) {
            // ================= Viterbi: (max,+), bit-exact fp32 =================================
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
            int bsel;  // mirrored y of the best interval so far, -1 = none / skip
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
            const float dr = u0;
            // terminal column: no candidates, q = S*(S>0)  (-0 + dr reproduces the reference's signed zero)
            if (x == T - 1) best = -0.0f;
            // the skip out of the top column: candidate 0 of the reference, so it wins every tie
            if (has_next && c == BX - 1) {
                const float xk = qnext + s_eta;
                bsel = (xk >= best) ? -1 : bsel;
                best = fmaxf(best, xk);
            }
            TKB_STAMP(3);
            TKB_WSTAMP_DEP(4, best);
            // ---- D. diagonal solve: value chain = FADD -> SHFL -> FADD -> FMNMX ------------------
            float qmine = 0.0f;
#pragma unroll
            for (int e = BX - 1; e >= 1; --e) {
                const float qfin = best + dr;
                const float qb = __shfl_sync(kFull, qfin, e);
                qmine = (c == e) ? qfin : qmine;
                const float xi = qb + sreg[e];                              // -inf for lanes c >= e
                const float xk = (c == e - 1) ? qb + s_eta : -INFINITY;    // skip x -> x+1
                const bool tk = (DIR == TKB_BACKWARD) ? (xi >= best) : (xi > best);
                const float b1 = fmaxf(best, xi);
                bsel = tk ? x0 + e : bsel;
                bsel = (xk >= b1) ? -1 : bsel;
                best = fmaxf(b1, xk);
                if ((e & (PB - 1)) == 0 && c >= e && c < e + PB && active)
                    publish(s_mbox + (size_t)x * p.Npad, qmine, epoch);
            }
            qmine = (c == 0) ? best + dr : qmine;
            if (c < PB && active) publish(s_mbox + (size_t)x * p.Npad, qmine, epoch);
            TKB_STAMP(4);
            TKB_WSTAMP(5);
            if (active && s_nok) {
                const int osel = bsel < 0 ? -1 : ((DIR == TKB_BACKWARD) ? bsel : T - 1 - bsel);
                p.code[(size_t)(n0 + sn) * T + pos] = ((unsigned)(osel + 1) << 1) | (s_d > 0.0f ? 1u : 0u);
                if (p.outv) p.outv[(size_t)pos * N + n0 + sn] = qmine;
            }
        } else if (s_is_lse && DO_L) {
            // ================= log-sum: (logsumexp,+) in the log2 domain, (M, S) pairs ===============
            float M = -FLT_MAX, S = 0.0f;
            {
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
            }
            const float sp2 = u0, eta2 = u1;
            if (x == T - 1) {  // terminal column: value = softplus(S[T-1,T-1])
                M = 0.0f;
                S = 1.0f;
            }
            if (has_next && c == BX - 1) lse_push(M, S, qnext + eta2, 1.0f);  // skip out of the top column
            TKB_WSTAMP_DEP(4, S);
            // ---- D. diagonal solve: broadcast (M + sp2, S) of lane e, push to every lane (no-op for c >= e)
#pragma unroll
            for (int e = BX - 1; e >= 1; --e) {
                const float Mb = __shfl_sync(kFull, M + sp2, e);
                const float sb = __shfl_sync(kFull, S, e);
                lse_push(M, S, Mb + sreg[e], sb);
                if ((e & (PB - 1)) == 0 && c >= e && c < e + PB && active) {
                    const float v2 = (M + sp2) + lg2f(S);
                    publish(s_mbox + (size_t)x * p.Npad, v2, epoch);
                    if (s_nok && p.outl) p.outl[(size_t)pos * N + n0 + sn] = v2 * kLn2;
                }
            }
            if (c < PB && active) {
                const float v2 = (M + sp2) + lg2f(S);
                publish(s_mbox + (size_t)x * p.Npad, v2, epoch);
                if (s_nok && p.outl) p.outl[(size_t)pos * N + n0 + sn] = v2 * kLn2;
            }
            TKB_WSTAMP(5);
        }
        __syncthreads();  // partials (in the FIFOs), diagS and qtop are reused by the next owned block
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int DIR, bool A16, int MODE>
static int launch_one(const SweepParams &p, int grid, cudaStream_t stream) {
    auto kern = sweep_kernel<DIR, A16, MODE>;
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

template <int DIR, bool A16>
static int launch_mode(int mode, const SweepParams &p, int grid, cudaStream_t stream) {
    switch (mode) {
        case TKB_SWEEP_VITERBI: return launch_one<DIR, A16, TKB_SWEEP_VITERBI>(p, grid, stream);
        case TKB_SWEEP_LOGSUM: return launch_one<DIR, A16, TKB_SWEEP_LOGSUM>(p, grid, stream);
        default: return launch_one<DIR, A16, TKB_SWEEP_VITERBI | TKB_SWEEP_LOGSUM>(p, grid, stream);
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

}  // namespace v1
}  // namespace tkb

using namespace tkb;
using namespace tkb::v1;

size_t tkb::sweep_workspace_bytes_v1(int T, int N) {
    if (T < 1 || N < 1) return 0;
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG;
    return kHeaderBytes + 2 * (size_t)T * npad * sizeof(unsigned long long);
}

int tkb::semicrf_sweep_v1(const float *score, const float *noise, int T, int N, int direction, int flags,
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
    if (sms <= 0) {
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
    const bool a16 = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(score) & 15) == 0);
    const int nb = (T + BX - 1) / BX;
    // groups are independent pipelines; split them over launches if there are more groups than SMs
    for (int g0 = 0; g0 < p.G; g0 += sms) {
        const int gcount = (p.G - g0) < sms ? (p.G - g0) : sms;
        int K = sms / gcount;
        if (K > nb) K = nb;
        if (K < 1) K = 1;
        p.g0 = g0;
        p.K = K;
        const int grid = gcount * K;
        int rc;
        if (direction == TKB_BACKWARD)
            rc = a16 ? launch_mode<TKB_BACKWARD, true>(flags, p, grid, stream)
                     : launch_mode<TKB_BACKWARD, false>(flags, p, grid, stream);
        else
            rc = a16 ? launch_mode<TKB_FORWARD, true>(flags, p, grid, stream)
                     : launch_mode<TKB_FORWARD, false>(flags, p, grid, stream);
        if (rc != 0) return rc;
    }
    return 0;
}

