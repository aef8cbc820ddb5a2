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
// Mirrored walk coordinate u: BACKWARD u = position; FORWARD u = T-1-position.
// The walk always runs u = start .. T-1 upwards and ends at u = T-1.
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

__global__ void __launch_bounds__(BT_THREADS) backtrack_kernel(const unsigned *__restrict__ code, int T, int N,
                                                               const int *__restrict__ forced, int dir,
                                                               int *__restrict__ pairs, int *__restrict__ counts,
                                                               long long pair_stride, long long count_stride) {
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
    int *out = pairs + (size_t)n * pair_stride;
    for (int u = ubeg; u < uend; ++u)
        if (mark[u]) {
            const unsigned w = cw[u];
            const int posn = (dir == TKB_BACKWARD) ? u : T - 1 - u;
            if (w & 1u) {
                const int slot = (dir == TKB_BACKWARD) ? k : total - 1 - k;
                out[2 * slot] = posn;
                out[2 * slot + 1] = posn;
                ++k;
            }
            if (u < T - 1 && (w >> 1) != 0) {
                const int sel = (int)(w >> 1) - 1;
                const int slot = (dir == TKB_BACKWARD) ? k : total - 1 - k;
                out[2 * slot] = (dir == TKB_BACKWARD) ? posn : sel;  // (begin, end)
                out[2 * slot + 1] = (dir == TKB_BACKWARD) ? sel : posn;
                ++k;
            }
        }
    if (tid == 0) counts[(size_t)n * count_stride] = total;
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
        TKB_CUDA(cudaFuncSetAttribute(backtrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    backtrack_kernel<<<N, BT_THREADS, smem, (cudaStream_t)stream_>>>(code, T, N, forced_start, direction, out_pairs,
                                                                     out_counts, pair_stride, count_stride);
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
