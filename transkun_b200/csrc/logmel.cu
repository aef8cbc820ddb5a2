// logmel.cu -- STFT / log-mel frontend: fused framing+window kernel -> batched cuFFT R2C -> fused
// power / channel-mean / banded mel filterbank / normalised-log kernel.
//
// Replaces transkun/Util.py:104-113 (Spectrum.forward: the [.,nWin,win] windowed-frame temporary and
// torch.fft.rfft(norm="ortho")) and :156-167 (MelSpectrum.forward: abs().pow(2), mean over the audio
// channels, the dense [nFreq x nMel] matmul, the log normalisation).  The input may be the strided
// `unfold` view makeFrame (:21-43) returns: frames are read through their strides, the 4x-overlapped
// copy is never materialised.  The mel filterbank is triangular, i.e. banded: every mel bin reads only its
// own [lo, lo+cnt) range of frequency bins (1472 non-zeros instead of 2049*229 for the shipped config).
#include <cufft.h>

#include "common.cuh"

namespace tkb {

// out[(item, w), i] = frames[item][i] * win[w][i];   item = (b, c, f)
__global__ void __launch_bounds__(256) window_frames_kernel(const float *__restrict__ x, long long sb, long long sc,
                                                            long long sf, int C, int F, int W, int nWin,
                                                            const float *__restrict__ wins, float *__restrict__ out) {
    const int item = blockIdx.x;
    const int f = item % F, c = (item / F) % C, b = item / (F * C);
    const float *src = x + (long long)b * sb + (long long)c * sc + (long long)f * sf;
    float *dst = out + (size_t)item * nWin * W;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const float v = src[i];
        for (int w = 0; w < nWin; ++w) dst[(size_t)w * W + i] = v * wins[(size_t)w * W + i];
    }
}

// one CTA per (b, f): power spectrum of every (channel, window), mean over channels (or not), banded mel, log
__global__ void __launch_bounds__(256) mel_log_kernel(const float2 *__restrict__ spec, int C, int F, int nFreq, int nWin,
                                                      int to_mono, const int *__restrict__ band_lo,
                                                      const int *__restrict__ band_cnt, const float *__restrict__ melfb,
                                                      int nMel, float inv_n, float eps, float log_eps,
                                                      float *__restrict__ out) {
    extern __shared__ float pw[];  // [nWin][nFreq] (mono) -- channels are processed one after the other otherwise
    const int f = blockIdx.x % F, b = blockIdx.x / F;
    const int Cout = to_mono ? 1 : C;
    for (int co = 0; co < Cout; ++co) {
        for (int i = threadIdx.x; i < nWin * nFreq; i += blockDim.x) {
            float acc = 0.0f;
            const int c0 = to_mono ? 0 : co, c1 = to_mono ? C : co + 1;
            for (int c = c0; c < c1; ++c) {
                const float2 z = spec[((size_t)((b * C + c) * F + f) * nWin) * nFreq + i];
                acc += (z.x * z.x + z.y * z.y) * inv_n;  // |rfft(norm="ortho")|^2
            }
            pw[i] = to_mono ? acc / (float)C : acc;
        }
        __syncthreads();
        float *o = out + ((size_t)(b * Cout + co) * F + f) * nMel * nWin;  // [nMel][nWin]
        for (int j = threadIdx.x; j < nMel * nWin; j += blockDim.x) {
            const int m = j / nWin, w = j - m * nWin;
            const int lo = band_lo[m], cnt = band_cnt[m];
            float acc = 0.0f;
            for (int t = 0; t < cnt; ++t) acc = fmaf(pw[w * nFreq + lo + t], melfb[(size_t)(lo + t) * nMel + m], acc);
            o[j] = (logf(acc + eps) - log_eps) / (-log_eps);
        }
        __syncthreads();
    }
}

struct PlanEntry {
    int n, batch, dev;  // a cuFFT plan belongs to the device it was made on
    cufftHandle plan;
    size_t work;
};
static PlanEntry g_plans[8];
static int g_nplans = 0;

static int get_plan(int n, int batch, PlanEntry **out) {
    const int dev = current_device();
    for (int i = 0; i < g_nplans; ++i)
        if (g_plans[i].n == n && g_plans[i].batch == batch && g_plans[i].dev == dev) {
            *out = &g_plans[i];
            return 0;
        }
    if (g_nplans == 8) {  // recycle the oldest
        cufftDestroy(g_plans[0].plan);
        for (int i = 1; i < 8; ++i) g_plans[i - 1] = g_plans[i];
        g_nplans = 7;
    }
    PlanEntry e;
    e.n = n;
    e.batch = batch;
    e.dev = dev;
    if (cufftCreate(&e.plan) != CUFFT_SUCCESS || cufftSetAutoAllocation(e.plan, 0) != CUFFT_SUCCESS) {
        set_error("cufftCreate failed");
        return TKB_ENODEV;
    }
    int nn[1] = {n};
    if (cufftMakePlanMany(e.plan, 1, nn, nullptr, 1, n, nullptr, 1, n / 2 + 1, CUFFT_R2C, batch, &e.work) !=
        CUFFT_SUCCESS) {
        set_error("cufftMakePlanMany(n=%d, batch=%d) failed", n, batch);
        cufftDestroy(e.plan);
        return TKB_EINVAL;
    }
    g_plans[g_nplans] = e;
    *out = &g_plans[g_nplans++];
    return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace tkb

using namespace tkb;

extern "C" size_t tkb_logmel_workspace_bytes(int B, int C, int F, int W, int nWin) {
    if (B < 1 || C < 1 || F < 1 || W < 2 || nWin < 1) return 0;
    PlanEntry *pe;
    if (get_plan(W, B * C * F * nWin, &pe) != 0) return 0;
    const size_t batch = (size_t)B * C * F * nWin;
    return align256(batch * W * sizeof(float)) + align256(batch * (W / 2 + 1) * sizeof(float2)) + align256(pe->work) + 256;
}

extern "C" int tkb_logmel(const float *frames, int64_t stride_b, int64_t stride_c, int64_t stride_f, int B, int C, int F,
                          int W, const float *windows, int nWin, const float *melfb, const int32_t *band_lo,
                          const int32_t *band_cnt, int nMel, int to_mono, float eps, float *out, void *workspace,
                          void *stream_) {
    if (!frames || !windows || !melfb || !band_lo || !band_cnt || !out || !workspace || B < 1 || C < 1 || F < 1 || W < 2 ||
        nWin < 1 || nMel < 1 || !(eps > 0.0f)) {
        set_error("tkb_logmel: invalid argument (B=%d C=%d F=%d W=%d nWin=%d nMel=%d)", B, C, F, W, nWin, nMel);
        return TKB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nFreq = W / 2 + 1;
    const size_t batch = (size_t)B * C * F * nWin;
    PlanEntry *pe;
    int rc = get_plan(W, (int)batch, &pe);
    if (rc != 0) return rc;
    char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float *wf = reinterpret_cast<float *>(ws);
    float2 *spec = reinterpret_cast<float2 *>(ws + align256(batch * W * sizeof(float)));
    void *work = ws + align256(batch * W * sizeof(float)) + align256(batch * nFreq * sizeof(float2));
    window_frames_kernel<<<B * C * F, 256, 0, stream>>>(frames, stride_b, stride_c, stride_f, C, F, W, nWin, windows, wf);
    TKB_CUDA(cudaGetLastError());
    if (cufftSetStream(pe->plan, stream) != CUFFT_SUCCESS || cufftSetWorkArea(pe->plan, work) != CUFFT_SUCCESS ||
        cufftExecR2C(pe->plan, wf, reinterpret_cast<cufftComplex *>(spec)) != CUFFT_SUCCESS) {
        set_error("cuFFT R2C (n=%d, batch=%zu) failed", W, batch);
        return TKB_ENODEV;
    }
    const size_t smem = (size_t)nWin * nFreq * sizeof(float);
    static size_t configured_by_dev[kMaxDevices] = {};
    const int dev_ = current_device();
    size_t dummy_ = 0;
    size_t &configured = dev_ >= 0 ? configured_by_dev[dev_] : dummy_;
    if (smem > 48 * 1024 && smem > configured) {
        TKB_CUDA(cudaFuncSetAttribute(mel_log_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    mel_log_kernel<<<B * F, 256, smem, stream>>>(spec, C, F, nFreq, nWin, to_mono, band_lo, band_cnt, melfb, nMel,
                                                 1.0f / (float)W, eps, logf(eps), out);
    TKB_CUDA(cudaGetLastError());
    return 0;
}
