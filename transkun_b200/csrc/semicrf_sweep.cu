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
//     (nb-1-J) mod K).  For its block a CTA first streams the FAR FIELD (all rows
//     y in later blocks; order-free semiring mat-vec; 16 warps each own every
//     16th row, cp.async-staged into per-lane shared-memory FIFOs, accumulators
//     in registers), then merges the 16 partials and runs the sequential
//     32-step DIAGONAL SOLVE with one warp per (track, semiring), lane = column,
//     one shuffle per step;
//   * solved rows are broadcast to the other CTAs of the group through a
//     global-memory mailbox of 64-bit words {fp32 value, epoch tag}: one relaxed
//     store publishes, one relaxed load observes (no fence, no flag, no reset).
// All CTAs of a launch must be co-resident (cooperative launch).
#include "common.cuh"

namespace tkb {

constexpr int NG = 8;      // tracks per group
constexpr int BX = 32;     // columns per block (= lanes of a solver warp)
constexpr int NW = 16;     // warps per CTA: far field 16 row-slices; solve 8 tracks x 2 semirings
constexpr int NT = NW * 32;
constexpr int SLOTS = 6;   // per-warp FIFO depth (rows); SLOTS-1 rows in flight
constexpr int CH = 4;      // rows per log-sum-exp rescale chunk

constexpr size_t kRingFloats = (size_t)NW * SLOTS * 2 * 32 * 4;
constexpr size_t kMergeEntries = (size_t)NW * NG * BX;  // float2 each, one array per semiring
constexpr size_t kDiagFloats = (size_t)NG * BX * BX;
constexpr size_t kSweepSmem = kRingFloats * 4 + 2 * kMergeEntries * 8 + kDiagFloats * 4;

constexpr size_t kHeaderBytes = 256;  // status word lives here

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
};

// Wait until the mailbox word carries this launch's epoch.  A protocol bug (or a
// non-co-resident grid) must not hang the GPU: after ~4 s the wait gives up,
// flags the workspace and lets the kernel drain with garbage.
__device__ __noinline__ unsigned long long poll_slow(const unsigned long long *w, unsigned epoch, int *status) {
    unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            unsigned long long v = ld_relaxed_u64(w);
            if ((unsigned)(v >> 32) == epoch) return v;
        }
        if (*(volatile int *)status != 0) return 0;
        if (globaltimer_ns() - t0 > 4000000000ull) {
            atomicExch(status, 1);
            return 0;
        }
    }
}
__device__ __forceinline__ float poll_value(const unsigned long long *w, unsigned epoch, int *status) {
    unsigned long long v = ld_relaxed_u64(w);
    if ((unsigned)(v >> 32) != epoch) v = poll_slow(w, epoch, status);
    return __uint_as_float((unsigned)v);
}
__device__ __forceinline__ void publish(unsigned long long *w, float val, unsigned epoch) {
    st_relaxed_u64(w, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(val));
}

template <int DIR, bool A16, int MODE>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const SweepParams p) {
    constexpr bool DO_V = (MODE & TKB_SWEEP_VITERBI) != 0;
    constexpr bool DO_L = (MODE & TKB_SWEEP_LOGSUM) != 0;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    float2 *mergeV = reinterpret_cast<float2 *>(ring + kRingFloats);
    float2 *mergeL = mergeV + kMergeEntries;
    float *diagS = reinterpret_cast<float *>(mergeL + kMergeEntries);  // [NG][BX rows][BX cols]

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
    float *my_ring = ring + ((size_t)warp * SLOTS * 2 * 32 + lane) * 4;  // + (slot*2+piece)*128 floats
    // solver mapping: warp -> (semiring, track), lane -> column
    const int sn = warp & 7;
    const bool s_is_lse = warp >= 8;
    const bool s_nok = (n0 + sn) < N;

    for (int J = nb - 1 - k; J >= 0; J -= p.K) {
        const int x0 = J * BX;
        const int ncols = min(BX, T - x0);

        // ---- 0. prefetch the diagonal block, transposed to [track][row][col] --------------
        for (int i = threadIdx.x; i < BX * BX * NG; i += NT) {
            const int n = i & 7, c = (i >> 3) & 31, r = i >> 8;
            if (r > c && r < ncols && (n0 + n) < N)
                cp_async4(&diagS[(n * BX + r) * BX + c],
                          p.Sbase + (long long)(x0 + c) * p.sx + (long long)(x0 + r) * p.sy + n0 + n, 4);
        }
        cp_async_commit();
        // unary + skip weights of my solver column (kept in registers across the far field)
        const int sx_ = x0 + lane;
        float s_d = 0.0f, s_eta = 0.0f;
        if (sx_ < T && s_nok) {
            s_d = __ldg(p.Sbase + (long long)sx_ * (p.sx + p.sy) + n0 + sn);
            if (sx_ < T - 1) s_eta = __ldg(p.etabase + (long long)sx_ * p.se + n0 + sn);
        }

        // ---- 1. far field: rows y = T-1 .. x0+BX, this warp takes every NW-th --------------
        float vmax[2][4], lM[2][4], lS[2][4];
        int vsel[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                vmax[j][c] = -INFINITY;
                vsel[j][c] = -1;
                lM[j][c] = -FLT_MAX;
                lS[j][c] = 0.0f;
            }
        const int R = T - (x0 + BX);
        const int myrows = R > warp ? (R - warp + NW - 1) / NW : 0;
        if (myrows > 0) {
            const float *src[2];
            int nbytes[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = x0 + 2 * cpair + j;  // always < T here: a far field exists only below full blocks
                nbytes[j] = nvalid * 4;
                src[j] = nvalid > 0 ? p.Sbase + (long long)col * p.sx + nq : p.Sbase;
            }
            auto issue = [&](int t) {
                const long long yoff = nvalid > 0 ? (long long)(T - 1 - (warp + t * NW)) * p.sy : 0;
                float *dst = my_ring + (size_t)(t % SLOTS) * 2 * 128;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (A16) {
                        cp_async16(dst + j * 128, src[j] + yoff, nbytes[j]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            cp_async4(dst + j * 128 + c, src[j] + yoff + (c < nvalid ? c : 0), c < nvalid ? 4 : 0);
                    }
                }
            };
            constexpr int D = SLOTS - 1;
#pragma unroll
            for (int t = 0; t < D; ++t) {
                if (t < myrows) issue(t);
                cp_async_commit();
            }
            // mailbox: lanes 0-7 fetch the Viterbi row, lanes 8-15 the log-sum row
            const bool poller = (lane < 8 && DO_V) || (lane >= 8 && lane < 16 && DO_L);
            const unsigned long long *wbase = (lane < 8 ? mboxV : mboxL) + n0 + (lane & 7);
            unsigned long long word = 0;
            if (poller) word = ld_relaxed_u64(wbase + (size_t)(T - 1 - warp) * p.Npad);
            for (int tb = 0; tb < myrows; tb += CH) {
                float xl[CH][2][4];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int t = tb + i;
                    if (t < myrows) {  // warp-uniform
                        const int y = T - 1 - (warp + t * NW);
                        if (t + D < myrows) issue(t + D);
                        cp_async_commit();
                        // -- q[y] of my tracks
                        bool ok = !poller || (unsigned)(word >> 32) == epoch;
                        if (!__all_sync(kFull, ok)) {
                            if (!ok) word = poll_slow(wbase + (size_t)y * p.Npad, epoch, p.status);
                        }
                        const float qval = __uint_as_float((unsigned)word);
                        if (poller && t + 1 < myrows) word = ld_relaxed_u64(wbase + (size_t)(y - NW) * p.Npad);
                        float qv[4], ql[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (DO_V) qv[c] = __shfl_sync(kFull, qval, quad * 4 + c);
                            if (DO_L) ql[c] = __shfl_sync(kFull, qval, 8 + quad * 4 + c);
                        }
                        // -- S(y, my columns, my tracks)
                        cp_async_wait<D>();
                        const float *slot = my_ring + (size_t)(t % SLOTS) * 2 * 128;
                        float4 a[2];
                        a[0] = *reinterpret_cast<const float4 *>(slot);
                        a[1] = *reinterpret_cast<const float4 *>(slot + 128);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float av[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (DO_V) {
                                    const float xv = qv[c] + av[c];
                                    const bool tk =
                                        (DIR == TKB_BACKWARD) ? (xv >= vmax[j][c]) : (xv > vmax[j][c]);
                                    vmax[j][c] = tk ? xv : vmax[j][c];
                                    vsel[j][c] = tk ? y : vsel[j][c];
                                }
                                if (DO_L) xl[i][j][c] = fmaf(av[c], kLog2e, ql[c]);
                            }
                        }
                    } else if (DO_L) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int c = 0; c < 4; ++c) xl[i][j][c] = -FLT_MAX;
                    }
                }
                if (DO_L) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float m = xl[0][j][c];
#pragma unroll
                            for (int i = 1; i < CH; ++i) m = fmaxf(m, xl[i][j][c]);
                            const float Mn = fmaxf(lM[j][c], m);
                            float acc = lS[j][c] * ex2f(lM[j][c] - Mn);
#pragma unroll
                            for (int i = 0; i < CH; ++i) acc += ex2f(xl[i][j][c] - Mn);
                            lS[j][c] = acc;
                            lM[j][c] = Mn;
                        }
                }
            }
        }
        // ---- 2. hand the 16 partials to the solver mapping ---------------------------------
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const size_t o = ((size_t)warp * NG + quad * 4 + c) * BX + 2 * cpair + j;
                if (DO_V) mergeV[o] = make_float2(vmax[j][c], __int_as_float(vsel[j][c]));
                if (DO_L) mergeL[o] = make_float2(lM[j][c], lS[j][c]);
            }
        cp_async_wait_all();
        __syncthreads();

        // ---- 3. diagonal solve: warp = (semiring, track), lane = column --------------------
        const int c = lane;
        const int x = x0 + c;
        const int pos = (DIR == TKB_BACKWARD) ? x : T - 1 - x;
        const bool has_next = (x0 + BX) <= T - 1;  // a later block exists -> skip candidate of the top column
        if (!s_is_lse && DO_V) {
            float best = -INFINITY;
            int bsel = -1;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float2 e = mergeV[((size_t)w * NG + sn) * BX + c];
                const int sl = __float_as_int(e.y);
                const bool better = e.x > best ||
                                    (e.x == best && sl >= 0 &&
                                     ((DIR == TKB_BACKWARD) ? (bsel < 0 || sl < bsel) : (sl > bsel)));
                if (better) {
                    best = e.x;
                    bsel = sl;
                }
            }
            float sreg[BX];
#pragma unroll
            for (int r = 1; r < BX; ++r) sreg[r] = diagS[(sn * BX + r) * BX + c];
            const float dr = relu_mask(s_d);
            float qprev = 0.0f, qmine = 0.0f;
            if (has_next) qprev = poll_value(mboxV + (size_t)(x0 + BX) * p.Npad + n0 + sn, epoch, p.status);
#pragma unroll
            for (int r = BX - 1; r >= 0; --r) {
                if (r < ncols) {
                    const float skipc = qprev + s_eta;
                    const bool tk = best > skipc;  // skip is candidate 0: it wins every tie
                    const float m = tk ? best : skipc;
                    const float qfin = (x == T - 1) ? dr : (m + dr);
                    const float qb = __shfl_sync(kFull, qfin, r);
                    if (c == r) {
                        qmine = qfin;
                        if (!tk) bsel = -1;
                        publish(mboxV + (size_t)x * p.Npad + n0 + sn, qb, epoch);
                    }
                    if (c < r) {
                        const float xx = qb + sreg[r];
                        const bool t2 = (DIR == TKB_BACKWARD) ? (xx >= best) : (xx > best);
                        if (t2) {
                            best = xx;
                            bsel = x0 + r;
                        }
                    }
                    if (c == r - 1) qprev = qb;
                }
            }
            if (x < T && s_nok) {
                const int osel = bsel < 0 ? -1 : ((DIR == TKB_BACKWARD) ? bsel : T - 1 - bsel);
                p.code[(size_t)(n0 + sn) * T + pos] = ((unsigned)(osel + 1) << 1) | (s_d > 0.0f ? 1u : 0u);
                if (p.outv) p.outv[(size_t)pos * N + n0 + sn] = qmine;
            }
        } else if (s_is_lse && DO_L) {
            float m[NW], s[NW];
            float M = -FLT_MAX;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float2 e = mergeL[((size_t)w * NG + sn) * BX + c];
                m[w] = e.x;
                s[w] = e.y;
                M = fmaxf(M, e.x);
            }
            float S = 0.0f;
#pragma unroll
            for (int w = 0; w < NW; ++w) S += s[w] * ex2f(m[w] - M);
            float sreg[BX];
#pragma unroll
            for (int r = 1; r < BX; ++r) sreg[r] = diagS[(sn * BX + r) * BX + c] * kLog2e;
            const float sp2 = softplus_ref(s_d) * kLog2e;
            const float eta2 = s_eta * kLog2e;
            float qprev = 0.0f, vmine = 0.0f;
            if (has_next) qprev = poll_value(mboxL + (size_t)(x0 + BX) * p.Npad + n0 + sn, epoch, p.status);
#pragma unroll
            for (int r = BX - 1; r >= 0; --r) {
                if (r < ncols) {
                    float v2 = 0.0f;
                    if (c == r) {
                        if (x == T - 1) {
                            v2 = sp2;
                        } else {
                            const float xs = qprev + eta2;
                            const float e = ex2f(-fabsf(M - xs));
                            const float tot = (xs > M) ? fmaf(S, e, 1.0f) : (S + e);
                            v2 = (fmaxf(M, xs) + lg2f(tot)) + sp2;
                        }
                    }
                    const float vb = __shfl_sync(kFull, v2, r);
                    if (c == r) {
                        vmine = v2;
                        publish(mboxL + (size_t)x * p.Npad + n0 + sn, vb, epoch);
                    }
                    if (c < r) {
                        const float xlv = vb + sreg[r];
                        const float e = ex2f(-fabsf(M - xlv));
                        S = (xlv > M) ? fmaf(S, e, 1.0f) : (S + e);
                        M = fmaxf(M, xlv);
                    }
                    if (c == r - 1) qprev = vb;
                }
            }
            if (x < T && s_nok && p.outl) p.outl[(size_t)pos * N + n0 + sn] = vmine * kLn2;
        }
        __syncthreads();  // merge buffers and diagS are reused by the next owned block
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

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

}  // namespace tkb

using namespace tkb;

extern "C" size_t tkb_sweep_workspace_bytes(int T, int N) {
    if (T < 1 || N < 1) return 0;
    const size_t npad = (size_t)((N + NG - 1) / NG) * NG;
    return kHeaderBytes + 2 * (size_t)T * npad * sizeof(unsigned long long);
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
