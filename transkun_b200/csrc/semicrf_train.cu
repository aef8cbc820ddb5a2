// semicrf_train.cu -- training-side kernels of the semi-CRF: marginals (the custom
// gradient of log Z), the un-normalised path score and its gradient.
//
// Replaces, in transkun/CRF/NeuralSemiCRFInterval.py:
//   :417-447  forward_backward's dense marginal construction (>= 6 full-size temporaries)
//   :469-472  ComputeLogZFasterGrad.backward (one more full pass: grad * grad_output)
//   :508-550  evalPath (host list comprehensions + 4 H2D copies + gather/scatter_add)
// All of these are HBM-bound streaming kernels: the marginal writer reads the lower
// triangle once and writes the dense [T,T,N] gradient once (the API returns a dense
// gradient, zero above the diagonal).
#include "common.cuh"

namespace tkb {

// Marginal writer.  Row e = blockIdx.y; a block covers MB_COLS consecutive begins b.  A thread owns one vector of VEC
// tracks (its beta[e] - logZ and grad_output stay in registers) and walks the begins of its sub-row: no division,
// one score load, one alpha load (L1/L2-resident, [T,N]) and one store per vector; everything above the diagonal
// is a plain zero store (the API returns a dense gradient).
constexpr int MB_THREADS = 256;
constexpr int MB_COLS = 64;
template <int VEC>
__global__ void __launch_bounds__(MB_THREADS) marginals_kernel(const float *__restrict__ score,
                                                              const float *__restrict__ alpha,
                                                              const float *__restrict__ beta,
                                                              const float *__restrict__ gscale, int T, int N,
                                                              float *__restrict__ grad) {
    const int e = blockIdx.y;
    const int nv = (N + VEC - 1) / VEC;            // vectors per cell
    const int rows = MB_THREADS / nv > 0 ? MB_THREADS / nv : 1;  // begins handled side by side
    const int b_begin = blockIdx.x * MB_COLS;
    const int b_end = min(b_begin + MB_COLS, T);
    const float *logZ = alpha + (long long)(T - 1) * N;  // :417
    for (int v0 = threadIdx.x; v0 < rows * nv; v0 += MB_THREADS) {  // (one pass unless nv > MB_THREADS)
        const int sub = v0 / nv, n0 = (v0 - sub * nv) * VEC;
        float ce[VEC], gs[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int n = n0 + v;
            ce[v] = n < N ? (beta[(long long)e * N + n] - logZ[n]) : 0.0f;
            gs[v] = (gscale && n < N) ? gscale[n] : 1.0f;
        }
        for (int b = b_begin + sub; b < b_end; b += rows) {
            const long long base = ((long long)e * T + b) * N + n0;
            float out[VEC];
            if (b > e) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) out[v] = 0.0f;
            } else {
                float sv[VEC], av[VEC];
                if (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(score + base);
                    const float4 a = *reinterpret_cast<const float4 *>(alpha + (long long)b * N + n0);
                    sv[0] = t.x; sv[1 % VEC] = t.y; sv[2 % VEC] = t.z; sv[3 % VEC] = t.w;
                    av[0] = a.x; av[1 % VEC] = a.y; av[2 % VEC] = a.z; av[3 % VEC] = a.w;
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        sv[v] = (n0 + v < N) ? score[base + v] : 0.0f;
                        av[v] = (n0 + v < N) ? alpha[(long long)b * N + n0 + v] : 0.0f;
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float x = av[v] + (ce[v] + sv[v]);               // :424
                    if (b == e) x -= 2.0f * softplus_ref(sv[v]);    // :427
                    out[v] = __expf(x) * gs[v];                     // :438, :472
                }
            }
            if (VEC == 4) {
                *reinterpret_cast<float4 *>(grad + base) = make_float4(out[0], out[1 % VEC], out[2 % VEC], out[3 % VEC]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (n0 + v < N) grad[base + v] = out[v];
            }
        }
    }
}

__global__ void grad_noise_kernel(const float *__restrict__ noise, const float *__restrict__ alpha,
                                  const float *__restrict__ beta, const float *__restrict__ gscale, int T, int N,
                                  float *__restrict__ gn) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)(T - 1) * N) return;
    const int n = (int)(i % N);
    const float logZ = alpha[(long long)(T - 1) * N + n];
    float g = __expf(alpha[i] + beta[i + N] + noise[i] - logZ);  // :445-447
    if (gscale) g *= gscale[n];
    gn[i] = g;
}

// cum[t][n] = sum_{u<t} noise[u][n]   (cumsum(pad(noise)), :524-525); one thread per track
__global__ void noise_cumsum_kernel(const float *__restrict__ noise, int T, int N, float *__restrict__ cum) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.0f;
    cum[n] = 0.0f;
    for (int t = 1; t < T; ++t) {
        acc += noise[(long long)(t - 1) * N + n];
        cum[(long long)t * N + n] = acc;
    }
}

// one warp per track: sum_k S[e_k, b_k, n] - (cum[e_k] - cum[b_k]), + cum[T-1]   (:540-548)
__global__ void evalpath_kernel(const float *__restrict__ score, const float *__restrict__ cum, int T, int N,
                                const int *__restrict__ pairs, const long long *__restrict__ offsets,
                                float *__restrict__ out) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc = 0.0f;
    for (long long k = offsets[n] + lane; k < offsets[n + 1]; k += 32) {
        const int b = pairs[2 * k], e = pairs[2 * k + 1];
        acc += score[((long long)e * T + b) * N + n] - (cum[(long long)e * N + n] - cum[(long long)b * N + n]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (lane == 0) out[n] = acc + cum[(long long)(T - 1) * N + n];
}

// d(evalPath)/d(score, noise) accumulated into dense buffers; one warp per track.
__global__ void evalpath_grad_kernel(int T, int N, const int *__restrict__ pairs, const long long *__restrict__ offsets,
                                     const float *__restrict__ gscale, float sign, float *__restrict__ gscore,
                                     float *__restrict__ gnoise) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float g = sign * (gscale ? gscale[n] : 1.0f);
    if (gnoise)
        for (int t = lane; t < T - 1; t += 32) gnoise[(long long)t * N + n] += g;  // d cum[T-1] / d noise[t]
    __syncwarp();
    for (long long k = offsets[n]; k < offsets[n + 1]; ++k) {
        const int b = pairs[2 * k], e = pairs[2 * k + 1];
        if (lane == 0 && gscore) atomicAdd(&gscore[((long long)e * T + b) * N + n], g);
        if (gnoise)
            for (int t = b + lane; t < e; t += 32) gnoise[(long long)t * N + n] -= g;  // -(cum[e]-cum[b])
        __syncwarp();
    }
}

}  // namespace tkb

using namespace tkb;

extern "C" int tkb_semicrf_marginals(const float *score, const float *noise, int T, int N, const float *alpha,
                                     const float *beta, const float *gscale, float *out_grad, float *out_grad_noise,
                                     void *stream_) {
    if (!score || !alpha || !beta || T < 1 || N < 1 || (T > 1 && !noise)) {
        set_error("tkb_semicrf_marginals: invalid argument (T=%d N=%d)", T, N);
        return TKB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    if (out_grad) {
        const bool v4 = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(score) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(out_grad) & 15) == 0);
        const bool a4 = v4 && ((reinterpret_cast<uintptr_t>(alpha) & 15) == 0);
        dim3 grid((unsigned)((T + MB_COLS - 1) / MB_COLS), (unsigned)T);
        if (a4)
            marginals_kernel<4><<<grid, MB_THREADS, 0, stream>>>(score, alpha, beta, gscale, T, N, out_grad);
        else
            marginals_kernel<1><<<grid, MB_THREADS, 0, stream>>>(score, alpha, beta, gscale, T, N, out_grad);
        TKB_CUDA(cudaGetLastError());
    }
    if (out_grad_noise && T > 1) {
        const long long tot = (long long)(T - 1) * N;
        grad_noise_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(noise, alpha, beta, gscale, T, N,
                                                                             out_grad_noise);
        TKB_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int tkb_semicrf_evalpath(const float *score, const float *noise, int T, int N, const int32_t *pairs,
                                    const int64_t *offsets, float *noise_cum, float *out, void *stream_) {
    if (!score || !offsets || !noise_cum || !out || T < 1 || N < 1 || (T > 1 && !noise)) {
        set_error("tkb_semicrf_evalpath: invalid argument (T=%d N=%d)", T, N);
        return TKB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    noise_cumsum_kernel<<<(N + 127) / 128, 128, 0, stream>>>(noise, T, N, noise_cum);
    TKB_CUDA(cudaGetLastError());
    evalpath_kernel<<<(N * 32 + 255) / 256, 256, 0, stream>>>(score, noise_cum, T, N, pairs,
                                                             reinterpret_cast<const long long *>(offsets), out);
    TKB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tkb_semicrf_evalpath_grad(int T, int N, const int32_t *pairs, const int64_t *offsets,
                                         const float *gscale, float sign, float *grad_score, float *grad_noise,
                                         void *stream_) {
    if (!offsets || T < 1 || N < 1) {
        set_error("tkb_semicrf_evalpath_grad: invalid argument (T=%d N=%d)", T, N);
        return TKB_EINVAL;
    }
    evalpath_grad_kernel<<<(N * 32 + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(
        T, N, pairs, reinterpret_cast<const long long *>(offsets), gscale, sign, grad_score, grad_noise);
    TKB_CUDA(cudaGetLastError());
    return 0;
}
