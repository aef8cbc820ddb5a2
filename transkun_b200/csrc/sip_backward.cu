// sip_backward.cu -- adjoint of the interval scorer's epilogue (training path, config 4).
//
// Forward (sip_scorer.cu; reference LayersTransformer.py:410-440):
//     S[e][b][n] = (q[n,e,:] . k[n,b,:]) * scale * |e-b|  +  [e==b] * diag[n,e]        (e >= b; zero above)
// so with G = dL/dS (what the semi-CRF's marginals kernel writes, layout [e][b][n], track innermost):
//     Gl[n][e][b] = G[e][b][n] * scale * (e-b)   for b < e, else 0
//     dq[n] = Gl[n] @ k[n]          dk[n] = Gl[n]^T @ q[n]          ddiag[n][e] = G[e][e][n]
// This kernel produces Gl (track-major, ready for two plain library batched GEMMs) and ddiag in ONE pass: a tiled
// transpose through shared memory with the length factor and the triangle mask fused.  The torch formulation it
// replaces (permute + tril + arange outer difference + multiply) made four passes over N*T*T*4 bytes with three
// temporaries (4 ms of a 6 ms training step at T=691, N=360).
#include "common.cuh"

namespace tkb {

constexpr int SB_TILE = 32;

// grid: (ceil(T/32) begin tiles, T ends, ceil(N/32) track tiles); block (32, 8)
__global__ void __launch_bounds__(256) sip_backward_prep_kernel(const float *__restrict__ g, long long pitch, int N, int T,
                                                                float scale, float *__restrict__ gl,
                                                                float *__restrict__ gdiag) {
    __shared__ float tile[SB_TILE][SB_TILE + 1];   // [begin][track]
    const int e = blockIdx.y, b0 = blockIdx.x * SB_TILE, n0 = blockIdx.z * SB_TILE;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool any_lower = b0 <= e;   // the tile has cells with b <= e
    if (any_lower) {
#pragma unroll
        for (int i = 0; i < SB_TILE; i += 8) {
            const int b = b0 + ty + i, n = n0 + tx;
            float v = 0.0f;
            if (b <= e && n < N) v = g[((long long)e * T + b) * pitch + n];   // coalesced along the track axis
            tile[ty + i][tx] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SB_TILE; i += 8) {
        const int n = n0 + ty + i, b = b0 + tx;
        if (n < N && b < T) {
            float v = 0.0f;
            if (any_lower && b < e) v = tile[tx][ty + i] * scale * (float)(e - b);
            gl[((long long)n * T + e) * T + b] = v;                            // coalesced along the begin axis
            if (b == e) gdiag[(long long)n * T + e] = tile[tx][ty + i];
        }
    }
}

}  // namespace tkb

using namespace tkb;

extern "C" int tkb_sip_backward_prep(const float *grad_score, int64_t pitch, int n_tracks, int T, float scale,
                                     float *out_gl, float *out_gdiag, void *stream_) {
    if (!grad_score || !out_gl || !out_gdiag || n_tracks < 1 || T < 1 || pitch < n_tracks) {
        set_error("tkb_sip_backward_prep: invalid argument (tracks=%d T=%d pitch=%lld)", n_tracks, T, (long long)pitch);
        return TKB_EINVAL;
    }
    if (T > 65535 || (n_tracks + SB_TILE - 1) / SB_TILE > 65535) {
        set_error("tkb_sip_backward_prep: T=%d / tracks=%d exceed the grid limits", T, n_tracks);
        return TKB_EINVAL;
    }
    dim3 grid((T + SB_TILE - 1) / SB_TILE, T, (n_tracks + SB_TILE - 1) / SB_TILE), block(SB_TILE, 8);
    sip_backward_prep_kernel<<<grid, block, 0, (cudaStream_t)stream_>>>(grad_score, pitch, n_tracks, T, scale, out_gl,
                                                                        out_gdiag);
    TKB_CUDA(cudaGetLastError());
    return 0;
}
