// semicrf_backtrack.cu -- Viterbi back-tracking on the device.
//
// Replaces the per-track host loops of the reference
// (transkun/CRF/NeuralSemiCRFInterval.py:61-102 backward, :157-199 forward), which
// copy ptr[T-1,N] to the host and chase pointers in Python (65-70 % of decode time
// on CPU).
//
// One CTA per track.  The walk  u -> next(u)  (next = u+1 on "skip", else the
// chosen partner) is a path in a functional graph, so instead of chasing T
// dependent loads the visited set is found by pointer doubling in shared
// memory (ceil(log2 T) rounds), pair counts are prefix-summed, and every
// visited position writes its (begin,end) pairs straight to its slot -- in
// exactly the order the reference appends them (and reversed for FORWARD, :196).
//
// backtrack_push_kernel is the same walk FUSED WITH THE MULTI-GPU EXCHANGE: the tracks of a sharded
// problem are decoded on different GPUs and every rank needs all decoded intervals.  Instead of decoding into local
// memory and running an all-gather afterwards, each CTA stores its track's record {count, logZ, pairs} straight into
// the symmetric (peer-mapped, NVLink) record buffer of EVERY rank, and the last CTA of the grid publishes a per-rank
// step flag with system-scope release: no collective kernel, no copy-engine traffic, no barrier kernels, and the next
// sweep can start as soon as this kernel has issued its stores.
//
// Mirrored walk coordinate u: BACKWARD u = position; FORWARD u = T-1-position.
// The walk always runs u = start .. T-1 upwards and ends at u = T-1.
#include <string.h>

#include "common.cuh"

namespace tkb {

constexpr int BT_THREADS = 256;

template <typename T>
__device__ __forceinline__ T warp_scan_incl(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

constexpr int kMaxPeers = 16;
struct PushDst {
    int *rec[kMaxPeers];            // record buffers of the ranks, [world * N][stride] int32 each
    unsigned *flag[kMaxPeers];      // step flags of the ranks, [world] uint32 each
    const float *logz;              // [N] or null
    unsigned *ticket;               // local counter, zero between launches
    int *status;                    // local status word (0, or the step whose wait timed out)
    int world, rank;
    unsigned step;                  // 1, 2, ...
    long long stride;               // ints per record: 2 + 4*T at least
};

template <bool PUSH>
__global__ void __launch_bounds__(BT_THREADS) backtrack_kernel(const unsigned *__restrict__ code, int T, int N,
                                                               const int *__restrict__ forced, int dir,
                                                               int *__restrict__ pairs, int *__restrict__ counts,
                                                               long long pair_stride, long long count_stride,
                                                               const PushDst dst) {
    extern __shared__ int sm[];
    int *nxtA = sm;              // [T]
    int *nxtB = sm + T;          // [T]
    unsigned *cw = (unsigned *)(sm + 2 * T);  // [T] code word per u
    unsigned char *mark = (unsigned char *)(sm + 3 * T);  // [T]
    __shared__ int warp_tot[BT_THREADS / 32];

    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned *row = code + (size_t)n * T;

    int start = forced ? forced[n] : (dir == TKB_BACKWARD ? 0 : T - 1);
    int ustart = (dir == TKB_BACKWARD) ? start : T - 1 - start;
    if (ustart > T - 1) ustart = T - 1;  // reference: the walk loop does not run, only the terminal check
    if (ustart < 0) ustart = 0;

    for (int u = tid; u < T; u += BT_THREADS) {
        const int posn = (dir == TKB_BACKWARD) ? u : T - 1 - u;
        const unsigned w = row[posn];
        cw[u] = w;
        int nx;
        if (u == T - 1) {
            nx = T - 1;
        } else {
            const int sel = (int)(w >> 1) - 1;  // partner position or -1
            nx = sel < 0 ? u + 1 : ((dir == TKB_BACKWARD) ? sel : T - 1 - sel);
        }
        nxtA[u] = nx;
        mark[u] = (u == ustart) ? 1 : 0;
    }
    __syncthreads();
    int *cur = nxtA, *oth = nxtB;
    for (int span = 1; span < T; span <<= 1) {
        for (int u = tid; u < T; u += BT_THREADS)
            if (mark[u]) mark[cur[u]] = 1;
        __syncthreads();
        for (int u = tid; u < T; u += BT_THREADS) oth[u] = cur[cur[u]];
        __syncthreads();
        int *t = cur;
        cur = oth;
        oth = t;
    }
    // pairs emitted at u: the singleton (if visited and diag>0) then the interval (if visited, u<T-1, not skip)
    const int per = (T + BT_THREADS - 1) / BT_THREADS;
    const int ubeg = tid * per, uend = min(T, ubeg + per);
    int local = 0;
    for (int u = ubeg; u < uend; ++u)
        if (mark[u]) {
            const unsigned w = cw[u];
            local += (int)(w & 1u) + ((u < T - 1 && (w >> 1) != 0) ? 1 : 0);
        }
    int incl = warp_scan_incl(local, lane);
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BT_THREADS / 32; ++w) {
        const int t = warp_tot[w];
        if (w < wid) woff += t;
        total += t;
    }
    int k = woff + incl - local;
    if (PUSH) {
        // The slot this step writes was last written two steps ago; every rank may have been reading it until it
        // submitted the step in between, which is what its flag of step-1 says (transkun_b200.sharded.FusedPushGather).
        if (tid == 0 && dst.step >= 2) {
            const volatile unsigned *mine = dst.flag[dst.rank];
            unsigned long long t0 = 0;
            for (unsigned tries = 0;; ++tries) {
                bool ok = true;
                for (int r = 0; r < dst.world; ++r) ok &= (int)(mine[r] - (dst.step - 1)) >= 0;
                if (ok) break;
                if ((tries & 255) == 255) {
                    if (t0 == 0) t0 = globaltimer_ns();
                    if (globaltimer_ns() - t0 > 4000000000ull) {
                        atomicExch(dst.status, (int)dst.step);
                        break;
                    }
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    int2 *stage = reinterpret_cast<int2 *>(sm + 4 * T);   // PUSH: [2T] pairs staged in shared memory (8-byte aligned)
    auto emit = [&](int slot, int b, int e) {
        if (PUSH) {
            stage[slot] = make_int2(b, e);
        } else {
            int *out = pairs + (size_t)n * pair_stride;
            out[2 * slot] = b;
            out[2 * slot + 1] = e;
        }
    };
    for (int u = ubeg; u < uend; ++u)
        if (mark[u]) {
            const unsigned w = cw[u];
            const int posn = (dir == TKB_BACKWARD) ? u : T - 1 - u;
            if (w & 1u) {
                emit((dir == TKB_BACKWARD) ? k : total - 1 - k, posn, posn);
                ++k;
            }
            if (u < T - 1 && (w >> 1) != 0) {
                const int sel = (int)(w >> 1) - 1;
                // (begin, end)
                emit((dir == TKB_BACKWARD) ? k : total - 1 - k, (dir == TKB_BACKWARD) ? posn : sel,
                     (dir == TKB_BACKWARD) ? sel : posn);
                ++k;
            }
        }
    if (!PUSH) {
        if (tid == 0) counts[(size_t)n * count_stride] = total;
        return;
    }
    __syncthreads();
    {
        // the record {count, logZ, pairs} goes to every rank with coalesced 8-byte stores (consecutive threads,
        // consecutive pairs): local memory for this rank, NVLink peer stores for the others
        const long long o = ((long long)dst.rank * N + n) * dst.stride;
        const int2 head = make_int2(total, dst.logz ? __float_as_int(dst.logz[n]) : 0);
        for (int r = 0; r < dst.world; ++r) {
            int2 *out = reinterpret_cast<int2 *>(dst.rec[r] + o);
            if (tid == 0) out[0] = head;
            for (int i = tid; i < total; i += BT_THREADS) out[1 + i] = stage[i];
        }
    }
    // every store of this CTA is performed system-wide before its ticket; the CTA that takes the last ticket
    // publishes this rank's step flag on every rank
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(dst.ticket, 1u);
        if (t == (unsigned)N - 1) {
            *dst.ticket = 0;   // ready for the next launch (stream order)
            __threadfence_system();
            for (int r = 0; r < dst.world; ++r)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst.flag[r] + dst.rank), "r"(dst.step) : "memory");
        }
    }
}

// waits until every rank's flag has reached `step` (one thread; the consumer side of backtrack_push_kernel)
__global__ void wait_flags_kernel(const unsigned *flags, int world, unsigned step, int *status) {
    const volatile unsigned *f = flags;
    unsigned long long t0 = 0;
    for (unsigned tries = 0;; ++tries) {
        bool ok = true;
        for (int r = 0; r < world; ++r) ok &= (int)(f[r] - step) >= 0;
        if (ok) break;
        if ((tries & 255) == 255) {
            if (t0 == 0) t0 = globaltimer_ns();
            if (globaltimer_ns() - t0 > 4000000000ull) {
                atomicExch(status, (int)step);
                break;
            }
        }
        __nanosleep(100);
    }
    __threadfence_system();   // acquire: the records written before the flags are visible to what follows on the stream
}

}  // namespace tkb

using namespace tkb;

static int launch_backtrack(const uint32_t *code, int T, int N, const int32_t *forced_start, int direction,
                            int32_t *out_pairs, int32_t *out_counts, long long pair_stride, long long count_stride,
                            void *stream_) {
    if (!code || !out_pairs || !out_counts || T < 1 || N < 1 || pair_stride < 4ll * T || count_stride < 1 ||
        (direction != TKB_BACKWARD && direction != TKB_FORWARD)) {
        set_error("tkb_semicrf_backtrack: invalid argument (T=%d N=%d dir=%d)", T, N, direction);
        return TKB_EINVAL;
    }
    const size_t smem = (size_t)T * (3 * sizeof(int) + 1);
    if (smem > 220 * 1024) {
        set_error("tkb_semicrf_backtrack: T=%d exceeds the shared-memory walk (max T ~ 17000)", T);
        return TKB_EINVAL;
    }
    static size_t configured_by_dev[kMaxDevices] = {};
    const int dev_ = current_device();
    size_t dummy_ = 0;
    size_t &configured = dev_ >= 0 ? configured_by_dev[dev_] : dummy_;
    if (smem > 48 * 1024 && smem > configured) {
        TKB_CUDA(cudaFuncSetAttribute(backtrack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TKB_CUDA(cudaFuncSetAttribute(backtrack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    PushDst none;
    memset(&none, 0, sizeof(none));
    backtrack_kernel<false><<<N, BT_THREADS, smem, (cudaStream_t)stream_>>>(code, T, N, forced_start, direction, out_pairs,
                                                                            out_counts, pair_stride, count_stride, none);
    TKB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tkb_semicrf_backtrack_push(const uint32_t *code, int T, int N, const int32_t *forced_start, int direction,
                                          const float *logz, void *const *peer_records, void *const *peer_flags, int world,
                                          int rank, int64_t record_stride, uint32_t step, uint32_t *ticket, int32_t *status,
                                          void *stream_) {
    if (!code || !peer_records || !peer_flags || !ticket || !status || T < 1 || N < 1 || world < 1 || world > kMaxPeers ||
        rank < 0 || rank >= world || record_stride < 2 + 4ll * T || (record_stride & 1) || step == 0 ||
        (direction != TKB_BACKWARD && direction != TKB_FORWARD)) {
        set_error("tkb_semicrf_backtrack_push: invalid argument (T=%d N=%d world=%d rank=%d stride=%lld step=%u)", T, N,
                  world, rank, (long long)record_stride, step);
        return TKB_EINVAL;
    }
    const size_t smem = (size_t)T * 4 * sizeof(int) + (size_t)2 * T * sizeof(int2);   // walk tables + staged pairs
    if (smem > 220 * 1024) {
        set_error("tkb_semicrf_backtrack_push: T=%d exceeds the shared-memory walk (max T ~ 7000)", T);
        return TKB_EINVAL;
    }
    static size_t configured_by_dev[kMaxDevices] = {};
    const int dev_ = current_device();
    if (smem > 48 * 1024 && (dev_ < 0 || smem > configured_by_dev[dev_])) {
        TKB_CUDA(cudaFuncSetAttribute(backtrack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev_ >= 0) configured_by_dev[dev_] = smem;
    }
    PushDst d;
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < world; ++r) {
        if (!peer_records[r] || !peer_flags[r]) {
            set_error("tkb_semicrf_backtrack_push: null peer pointer for rank %d", r);
            return TKB_EINVAL;
        }
        d.rec[r] = reinterpret_cast<int *>(peer_records[r]);
        d.flag[r] = reinterpret_cast<unsigned *>(peer_flags[r]);
    }
    d.logz = logz;
    d.ticket = ticket;
    d.status = status;
    d.world = world;
    d.rank = rank;
    d.step = step;
    d.stride = record_stride;
    backtrack_kernel<true><<<N, BT_THREADS, smem, (cudaStream_t)stream_>>>(code, T, N, forced_start, direction, nullptr,
                                                                           nullptr, 0, 0, d);
    TKB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tkb_wait_flags(const uint32_t *flags, int world, uint32_t step, int32_t *status, void *stream_) {
    if (!flags || !status || world < 1 || world > kMaxPeers) {
        set_error("tkb_wait_flags: invalid argument (world=%d)", world);
        return TKB_EINVAL;
    }
    wait_flags_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(flags, world, step, status);
    TKB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tkb_semicrf_backtrack(const uint32_t *code, int T, int N, const int32_t *forced_start, int direction,
                                     int32_t *out_pairs, int32_t *out_counts, void *stream_) {
    return launch_backtrack(code, T, N, forced_start, direction, out_pairs, out_counts, 4ll * T, 1, stream_);
}

extern "C" int tkb_semicrf_backtrack_strided(const uint32_t *code, int T, int N, const int32_t *forced_start,
                                             int direction, int32_t *out_pairs, int64_t pair_stride,
                                             int32_t *out_counts, int64_t count_stride, void *stream_) {
    return launch_backtrack(code, T, N, forced_start, direction, out_pairs, out_counts, pair_stride, count_stride,
                            stream_);
}
