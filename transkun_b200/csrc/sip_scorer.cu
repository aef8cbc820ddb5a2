// sip_scorer.cu -- Scaled-Inner-Product interval scorer on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces transkun/LayersTransformer.py:410-440 (ScaledInnerProductIntervalScorer.forward after the
// Linear projection): per track n
//     S[e,b,n] = ( sum_d (q[n,e,d]/sqrt(D)) * k[n,b,d] ) * |e-b|  +  [e==b] * diag[n,e]
// written directly in the layout the CRF reads, score[e][b][n] (track innermost), LOWER TRIANGLE ONLY
// (e >= b is all the semi-CRF ever reads; the reference computes the full square, multiplies it by the
// length matrix, adds diag_embed and permutes -- four more passes over T*T*N).
//
// One CTA computes a (128 ends x 64 begins) tile for FOUR TRACKS (16 bytes of the track-innermost layout): four
// 128x64 fp32 accumulators = 256 of the 512 TMEM columns, so two CTAs share an SM and one CTA's epilogue overlaps
// the other's main loop.  The CTAs of the 22 track quads of a tile are dispatched together (quad = fastest block
// index), so L2 sees both halves of every 32-byte sector before it evicts them.  Warp-specialised: warps 0-5
// stage the operands per (track, 32-wide K chunk) (Q: 128 rows, K: 64 rows, 128 B each) with cp.async into
// 128B-swizzled K-major shared memory (4 stages, full/empty mbarriers, no CTA barrier in the main loop), one
// thread of warp 6 issues four tcgen05.mma.kind::tf32 (M128 N64 K8) per step and tcgen05.commit releases the
// stage; all 8 warps run the epilogue: tcgen05.ld from TMEM, 1/sqrt(D) (a power of two for D=256, exact), the
// length factor, the diagonal and the triangle mask, 16-byte stores [e][b][4 tracks].
//
// Precision: TF32 operands (the reference's own --allow_tf32 regime, train.py:41-43), fp32 accumulate.
#include <stdlib.h>

#include "common.cuh"

namespace tkb {

constexpr int SC_TE = 128;      // ends per tile (= UMMA M)
constexpr int SC_TB = 64;       // begins per tile (= UMMA N)
constexpr int SC_NG = 4;        // tracks per CTA: 4 x 64 fp32 accumulator columns = half of TMEM, two CTAs per SM
constexpr int SC_KC = 32;       // tf32 elements per 128-byte swizzled row
constexpr int SC_UMMA_K = 8;    // tf32 elements per tcgen05.mma
constexpr int SC_THREADS = 256;
constexpr int SC_PRODUCERS = 192;  // warps 0-5 stage the operands, warp 6 issues the MMAs, all 8 warps run the epilogue
constexpr int SC_STAGES = 4;
constexpr int SC_LOOKAHEAD = 2;    // cp.async groups a producer thread keeps in flight
constexpr int SC_A_BYTES = SC_TE * 128;
constexpr int SC_B_BYTES = SC_TB * 128;
constexpr int SC_STAGE_BYTES = SC_A_BYTES + SC_B_BYTES;
constexpr size_t kScorerSmem = (size_t)SC_STAGES * SC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert((SC_TE + SC_TB) * 8 % SC_PRODUCERS == 0, "pieces per producer thread");

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// K-major, SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr) {
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) /*LBO (unused with swizzle)*/ |
           ((unsigned long long)(1024 >> 4) << 32) /*SBO*/ | (1ull << 46) /*sm100 descriptor version*/ |
           (2ull << 61) /*SWIZZLE_128B*/;
}
// kind::tf32, fp32 accumulate, A and B K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ | ((unsigned)(N >> 3) << 17) |
           ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

struct ScorerParams {
    const float *q, *k, *diag;  // [NT][T][D], [NT][T][D], [NT][T]
    float *out;                 // [T][T][pitch], the first NT tracks of a cell are written
    long long pitch;
    int NT, T, D;
    float qscale;               // 1/sqrt(D)
    int tiles_b_total;          // tiles of the lower triangle
    int gq;                     // track quads per block-order group
};

// tile index -> (eb, bb): lower-triangular enumeration; row eb has nb_row(eb) = min(ceil(T/64), 2*eb + 2) tiles
__global__ void __launch_bounds__(SC_THREADS, 2) sip_scorer_kernel(const ScorerParams p) {
    extern __shared__ unsigned char smem_raw_sc[];
    const unsigned smem_base = (smem_u32(smem_raw_sc) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
    const unsigned bars = smem_base + SC_STAGES * SC_STAGE_BYTES;  // full[SC_STAGES], empty[SC_STAGES], acc_done
    const unsigned full_b = bars, empty_b = bars + 8 * SC_STAGES, acc_b = bars + 16 * SC_STAGES;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T, D = p.D, NT = p.NT;
    // Block order: (group of p.gq track quads) > tile > quad in the group; p.gq = all quads by default, i.e. the CTAs of all
    // 22 quads of a tile are dispatched together and L2 assembles whole 352-byte cells / 128-byte lines of the
    // track-innermost output before it evicts them.  Measured round 2 (scripts/debug_scorer.py, TKB_SCORER_GQ): running
    // all tiles of 2 / 4 / 8 quads back to back keeps their q / k rows in L2 (the operands are no longer streamed from
    // DRAM 4.8 times) but is SLOWER -- 1294 / 1045 / 910 us against 835 us at T=2048 -- because the output then reaches
    // DRAM as isolated 32-byte sectors.  The kernel is bound by how its stores assemble, not by operand traffic.
    const int ngq = (NT + SC_NG - 1) / SC_NG;
    const int gq = p.gq;
    const int g = ((int)blockIdx.x / (p.tiles_b_total * gq)) * gq + (int)blockIdx.x % gq;
    if (g >= ngq) return;   // the last group may be short (whole CTA: nothing allocated yet)
    const int n0 = g * SC_NG;
    // decode (eb, bb) from blockIdx.x
    const int nbb = (T + SC_TB - 1) / SC_TB;
    int eb = 0, bb = 0;
    {
        int rem = ((int)blockIdx.x / gq) % p.tiles_b_total;
        for (;;) {
            const int row = min(nbb, 2 * eb + 2);
            if (rem < row) {
                bb = rem;
                break;
            }
            rem -= row;
            ++eb;
        }
    }
    const int e0 = eb * SC_TE, b0 = bb * SC_TB;
    const int nchunks = D / SC_KC;
    const int ntrk = min(SC_NG, NT - n0);
    const int nsteps = ntrk * nchunks;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(SC_NG * SC_TB)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < SC_STAGES; ++s) {
            mbar_init(full_b + 8 * s, SC_PRODUCERS);
            mbar_init(empty_b + 8 * s, 1);
        }
        mbar_init(acc_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;

    // operand loader: step s = (track t, K chunk kc); 1536 16-byte pieces per step, 8 per producer thread
    auto load_step = [&](int s) {
        const int t = s / nchunks, kc = s - t * nchunks;
        const unsigned stage = smem_base + (unsigned)(s % SC_STAGES) * SC_STAGE_BYTES;
        const float *qn = p.q + ((size_t)(n0 + t) * T) * D + kc * SC_KC;
        const float *kn = p.k + ((size_t)(n0 + t) * T) * D + kc * SC_KC;
#pragma unroll
        for (int i = 0; i < (SC_TE + SC_TB) * 8 / SC_PRODUCERS; ++i) {
            const int piece = tid + i * SC_PRODUCERS;
            const int row = piece >> 3, ch = piece & 7;
            const bool isA = row < SC_TE;
            const int r = isA ? row : row - SC_TE;
            const int grow = (isA ? e0 : b0) + r;
            const float *src = (isA ? qn : kn) + (size_t)min(grow, T - 1) * D + ch * 4;
            const unsigned dst = stage + (isA ? 0 : SC_A_BYTES) + r * 128 + ((ch ^ (r & 7)) << 4);
            cp_async16_s(dst, src, grow < T ? 16 : 0);
        }
    };

    const unsigned idesc = umma_idesc_tf32(SC_TE, SC_TB);
    if (tid < SC_PRODUCERS) {
        // ---- producers: fill stage s % SC_STAGES once the MMAs that read it have completed (empty), hand it over
        // (full) when this thread's pieces have landed; SC_LOOKAHEAD groups in flight per thread, no CTA barrier
        auto hand_over = [&](int s) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async writes -> visible to the tensor core
            mbar_arrive1(full_b + 8 * (s % SC_STAGES));
        };
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            if (s >= SC_STAGES) mbar_wait(empty_b + 8 * (s % SC_STAGES), (unsigned)(((s / SC_STAGES) - 1) & 1));
            load_step(s);
            cp_async_commit();
            if (s >= SC_LOOKAHEAD) {
                cp_async_wait<SC_LOOKAHEAD>();
                hand_over(s - SC_LOOKAHEAD);
            }
        }
        cp_async_wait_all();
#pragma unroll 1
        for (int s = max(nsteps - SC_LOOKAHEAD, 0); s < nsteps; ++s) hand_over(s);
    } else if (tid == SC_PRODUCERS) {
        // ---- MMA issuer: one thread; tcgen05.commit releases the stage when its four MMAs are done
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            mbar_wait(full_b + 8 * (s % SC_STAGES), (unsigned)((s / SC_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int t = s / nchunks, kc = s - t * nchunks;
            const unsigned stage = smem_base + (unsigned)(s % SC_STAGES) * SC_STAGE_BYTES;
            const unsigned long long da = umma_desc_sw128(stage), db = umma_desc_sw128(stage + SC_A_BYTES);
#pragma unroll
            for (int kk = 0; kk < SC_KC / SC_UMMA_K; ++kk)  // +32 bytes along K inside the swizzle atom = +2 in the address field
                umma_tf32(tmem_base + t * SC_TB, da + 2 * kk, db + 2 * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            umma_commit(empty_b + 8 * (s % SC_STAGES));
        }
        umma_commit(acc_b);
    }
    __syncwarp();
    // all accumulators complete
    mbar_wait(acc_b, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w reads TMEM lanes 32*(w%4).., warps 0-3 take begins 0..31 of the tile, warps 4-7 begins 32..63
    const int e = e0 + 32 * (warp & 3) + lane;
    const int jbase = (warp >> 2) * 32;
    const bool vec_ok = (p.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    float dg[SC_NG];
#pragma unroll
    for (int t = 0; t < SC_NG; ++t) dg[t] = (e < T && t < ntrk) ? p.diag[(size_t)(n0 + t) * T + e] : 0.0f;
#pragma unroll 1
    for (int j0 = jbase; j0 < jbase + 32; j0 += 8) {
        float acc[SC_NG][8];
#pragma unroll
        for (int t = 0; t < SC_NG; ++t) {
            if (t < ntrk) {
                tmem_ld8(tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(t * SC_TB + j0), acc[t]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[t][i] = 0.0f;
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (e < T) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int b = b0 + j0 + jj;
                if (b <= e) {
                    float v[SC_NG];
                    const float len = (float)(e - b);
#pragma unroll
                    for (int t = 0; t < SC_NG; ++t) v[t] = (b == e) ? dg[t] : (acc[t][jj] * p.qscale) * len;
                    float *o = p.out + ((size_t)e * T + b) * p.pitch + n0;
                    if (vec_ok && ntrk == SC_NG) {
                        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < SC_NG; ++t)
                            if (t < ntrk) o[t] = v[t];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SC_NG * SC_TB) : "memory");
}

}  // namespace tkb

using namespace tkb;

extern "C" int tkb_sip_score(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                             float *out_score, void *stream_) {
    return tkb_sip_score_pitched(q, k, diag, n_tracks, T, D, out_score, n_tracks, stream_);
}

extern "C" int tkb_sip_score_pitched(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                                     float *out_score, int64_t pitch, void *stream_) {
    return tkb_sip_score_scaled(q, k, diag, n_tracks, T, D, D > 0 ? 1.0f / sqrtf((float)D) : 0.0f, out_score, pitch, stream_);
}

extern "C" int tkb_sip_score_scaled(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                                    float scale, float *out_score, int64_t pitch, void *stream_) {
    if (!q || !k || !diag || !out_score || n_tracks < 1 || T < 1 || D < SC_KC || D % SC_KC != 0 || pitch < n_tracks ||
        (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15)) {
        set_error("tkb_sip_score: invalid argument (tracks=%d T=%d D=%d; D must be a multiple of 32, q/k 16-byte aligned)",
                  n_tracks, T, D);
        return TKB_EINVAL;
    }
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || !configured[dev]) {
        TKB_CUDA(cudaFuncSetAttribute(sip_scorer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScorerSmem));
        if (dev >= 0) configured[dev] = true;
    }
    ScorerParams p;
    p.q = q;
    p.k = k;
    p.diag = diag;
    p.out = out_score;
    p.NT = n_tracks;
    p.pitch = pitch;
    p.T = T;
    p.D = D;
    p.qscale = scale;
    const int neb = (T + SC_TE - 1) / SC_TE, nbb = (T + SC_TB - 1) / SC_TB;
    long long tiles = 0;
    for (int eb = 0; eb < neb; ++eb) tiles += (2 * eb + 2 < nbb) ? 2 * eb + 2 : nbb;
    p.tiles_b_total = (int)tiles;
    const int ngq = (n_tracks + SC_NG - 1) / SC_NG;
    static int gq_env = -1;
    if (gq_env < 0) {
        const char *e = getenv("TKB_SCORER_GQ");   // diagnostics: quads per block-order group (default: all)
        gq_env = e ? atoi(e) : 0;
    }
    p.gq = gq_env > 0 ? gq_env : ngq;
    if (p.gq > ngq) p.gq = ngq;
    dim3 grid((unsigned)(tiles * p.gq * ((ngq + p.gq - 1) / p.gq)));
    sip_scorer_kernel<<<grid, SC_THREADS, kScorerSmem, (cudaStream_t)stream_>>>(p);
    TKB_CUDA(cudaGetLastError());
    return 0;
}
