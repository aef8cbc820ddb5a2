/*
 * transkun_b200.h -- C ABI of libtranskun_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for Transkun's neural semi-CRF hot path.  The reference has
 * no native code and therefore no FFI of its own; the
 * interface each entry point replaces is the Python function it is called from,
 * cited per function as file:line into /root/reference/transkun/.  The ctypes
 * binding a maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter is documented as host;
 *   - score is the reference's [T,T,N] fp32 tensor, contiguous, laid out
 *     [end][begin][track] (track innermost); noise is [T-1,N] fp32 contiguous;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     every call only enqueues work on it and never synchronises;
 *   - no allocation happens inside the library: the caller provides outputs and
 *     workspaces (sizes from the *_bytes functions);
 *   - return value: 0 on success, a negative TKB_E* code for argument errors,
 *     a positive cudaError_t for CUDA failures; tkb_last_error() describes the
 *     last failure of the calling thread.
 */
#ifndef TRANSKUN_B200_H
#define TRANSKUN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TKB_VERSION 1

#define TKB_EINVAL (-1)   /* bad argument (null pointer, T < 1, N < 1, unknown flag) */
#define TKB_ENODEV (-2)   /* no sm_100 device / kernel image not loadable */
#define TKB_ELAUNCH (-3)  /* persistent grid does not fit the device */

/* direction of the dynamic programme */
#define TKB_BACKWARD 0    /* viterbiBackward / beta sweep: positions T-1 .. 0 */
#define TKB_FORWARD 1     /* viterbi / alpha sweep: positions 0 .. T-1 */

/* which semirings a sweep evaluates (OR them to read the score tensor once) */
#define TKB_SWEEP_VITERBI 1  /* (max,+): values + back-pointers */
#define TKB_SWEEP_LOGSUM 2   /* (logsumexp,+): log-partition table */

int tkb_version(void);
const char *tkb_last_error(void);

/* Device the library would run on is sm_100?  Returns 0 if usable. */
int tkb_device_check(void);

/*
 * Workspace for tkb_semicrf_sweep: the inter-CTA mailbox through which solved
 * table rows are broadcast, plus a status word.  Must be zero-filled once when
 * allocated; afterwards it is reused across calls, each call passing an `epoch`
 * strictly greater than any epoch used with this workspace before (re-zero it
 * if the 32-bit epoch wraps).  One workspace serves one stream at a time.
 */
size_t tkb_sweep_workspace_bytes(int T, int N);

/*
 * Host -> device upload of the part of score[T][T][N] the semi-CRF reads (end >= begin), for callers whose
 * score tensor lives in (pinned) host memory: the reference moves the dense tensor with `.cuda()` /
 * `.to(device)` before constructing the object (crfMinimalExample.py:13-14, README usage); this moves the
 * staircase of row chunks instead, about half the bytes.  host_score and dev_score are [T][T][N] fp32,
 * contiguous; cells above the staircase keep whatever dev_score held.  Asynchronous on `stream`.
 */
int tkb_upload_lower_triangle(const float *host_score, float *dev_score, int T, int N, int rows_per_chunk,
                              void *stream);


/*
 * The semi-Markov dynamic programme over the lower triangle of score.
 * Replaces the TorchScript loops of
 *   CRF/NeuralSemiCRFInterval.py:31-51   viterbiBackward   (BACKWARD | VITERBI)
 *   CRF/NeuralSemiCRFInterval.py:124-144 viterbi           (FORWARD  | VITERBI)
 *   CRF/NeuralSemiCRFInterval.py:218-234 computeLogZ       (FORWARD  | LOGSUM)
 *   CRF/NeuralSemiCRFInterval.py:303-327 beta recursion    (BACKWARD | LOGSUM)
 * With VITERBI|LOGSUM both tables come out of ONE read of the triangle.
 *
 * out_code [N][T] uint32 (track-major), VITERBI only:
 *     bit 0      = score[t,t,n] > 0 (the singleton (t,t) is emitted when t is visited)
 *     bits 31..1 = 0 for "skip", else 1 + the chosen partner position
 *                  (BACKWARD: the end e of interval (t,e); FORWARD: the begin b of (b,t)).
 *     Ties are resolved exactly as torch.max over the reference's candidate
 *     list does (skip first, then the first interval candidate).
 * out_vit  [T][N] fp32 or NULL: the Viterbi table q (BACKWARD) / v (FORWARD),
 *     bit-identical to the reference's.
 * out_lse  [T][N] fp32 or NULL: beta (BACKWARD) / alpha (FORWARD), natural log.
 *     logZ = out_lse[0] (BACKWARD) or out_lse[T-1] (FORWARD).
 */
int tkb_semicrf_sweep(const float *score, const float *noise, int T, int N, int direction, int flags,
                      void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit,
                      float *out_lse, void *stream);

/*
 * Same, for a score tensor whose track axis is padded: cell (end, begin) starts at
 * score + (end * T + begin) * pitch and holds N <= pitch tracks.  A pitch that is a multiple of 4 (with a
 * 16-byte aligned base) keeps the 16-byte copy path for track counts like the model's N = 90
 * (ModelTransformer.py:97: 88 keys + 2 pedals), which a dense [T,T,90] tensor cannot offer.
 * tkb_semicrf_sweep(...) == tkb_semicrf_sweep_pitched(score, N, ...).
 */
int tkb_semicrf_sweep_pitched(const float *score, int64_t pitch, const float *noise, int T, int N, int direction,
                              int flags, void *workspace, uint32_t epoch, uint32_t *out_code, float *out_vit,
                              float *out_lse, void *stream);

/*
 * Reads the status word of a sweep workspace (synchronises `stream`).
 * *status_host (HOST pointer) = 0 if every sweep that used the workspace ran to
 * completion, non-zero if an inter-CTA wait timed out (results invalid).
 */
int tkb_sweep_status(const void *workspace, int *status_host, void *stream);

/*
 * Back-tracking on the device.  Replaces the per-track host loops
 *   CRF/NeuralSemiCRFInterval.py:61-102  (BACKWARD; forced_start = begin position, default 0)
 *   CRF/NeuralSemiCRFInterval.py:157-199 (FORWARD;  forced_start = end position, default T-1)
 * forced_start: DEVICE int32[N] or NULL for the default.
 * out_pairs  [N][2*T][2] int32: (begin,end) in the order the reference returns them.
 * out_counts [N] int32: number of pairs per track.
 */
int tkb_semicrf_backtrack(const uint32_t *code, int T, int N, const int32_t *forced_start,
                          int direction, int32_t *out_pairs, int32_t *out_counts, void *stream);

/*
 * Same, writing into caller-strided records: track n's pairs start at out_pairs + n*pair_stride
 * (int32 units, pair_stride >= 4*T) and its count goes to out_counts[n*count_stride].  Lets a
 * caller keep count, log-partition and pairs of a track in ONE fixed-size record, so that the
 * multi-GPU gather of decoded intervals is a single zero-copy all-gather.
 */
int tkb_semicrf_backtrack_strided(const uint32_t *code, int T, int N, const int32_t *forced_start,
                                  int direction, int32_t *out_pairs, int64_t pair_stride,
                                  int32_t *out_counts, int64_t count_stride, void *stream);

/*
 * Back-tracking FUSED WITH THE MULTI-GPU EXCHANGE of a track-sharded problem (the reference has
 * no counterpart: it decodes all 90 symbols on one device, ModelTransformer.py:549).  Rank `rank` of `world` decodes its N
 * tracks and stores every track's record
 *     int32 record[record_stride] = { count, logZ bits (0 if logz is NULL), begin_0, end_0, begin_1, end_1, ... }
 * at row rank*N + n of the record buffer of EVERY rank -- peer_records[r] (HOST array of `world` DEVICE pointers, peer-mapped
 * symmetric memory, [world*N][record_stride] int32, 8-byte aligned, record_stride even and >= 2 + 4*T) -- then publishes
 * peer_flags[r][rank] = step (HOST array of `world` DEVICE pointers to uint32[world]) with system-scope release once all
 * its stores are performed.  `step` counts 1, 2, ...; from step 2 on the kernel first waits until every rank has
 * published step-1, which tells it that the buffer it is about to overwrite (callers alternate two buffers) is no longer
 * read.  ticket: DEVICE uint32, zero on first use; status: DEVICE int32, receives `step` if a wait times out (4 s).
 * tkb_wait_flags is the consumer side: one tiny kernel that returns once flags[0..world) >= step.
 */
int tkb_semicrf_backtrack_push(const uint32_t *code, int T, int N, const int32_t *forced_start, int direction,
                               const float *logz, void *const *peer_records, void *const *peer_flags, int world, int rank,
                               int64_t record_stride, uint32_t step, uint32_t *ticket, int32_t *status, void *stream);
int tkb_wait_flags(const uint32_t *flags, int world, uint32_t step, int32_t *status, void *stream);

/*
 * Marginals (the custom gradient of the log-partition).  Replaces
 *   CRF/NeuralSemiCRFInterval.py:417-447 (forward_backward) fused with
 *   CRF/NeuralSemiCRFInterval.py:469-472 (ComputeLogZFasterGrad.backward).
 * alpha, beta: [T][N] natural-log tables from two LOGSUM sweeps; gscale: [N]
 * upstream gradient or NULL for 1.  out_grad [T][T][N] dense (zero above the
 * diagonal), out_grad_noise [T-1][N].
 */
int tkb_semicrf_marginals(const float *score, const float *noise, int T, int N, const float *alpha,
                          const float *beta, const float *gscale, float *out_grad,
                          float *out_grad_noise, void *stream);

/*
 * Un-normalised path score.  Replaces CRF/NeuralSemiCRFInterval.py:508-550.
 * pairs: [total][2] int32 (begin,end), offsets: [N+1] int64 CSR, both DEVICE.
 * noise_cum: [T][N] workspace (filled with cumsum(pad(noise))), out: [N].
 */
int tkb_semicrf_evalpath(const float *score, const float *noise, int T, int N, const int32_t *pairs,
                         const int64_t *offsets, float *noise_cum, float *out, void *stream);

/*
 * Gradient of the path score w.r.t. score and noise, ACCUMULATED into dense
 * buffers (autograd of the gather at :540-548): grad_score[e,b,n] += g[n] per
 * listed interval; grad_noise[t,n] += g[n] * (1 - #intervals covering [t,t+1]).
 */
int tkb_semicrf_evalpath_grad(int T, int N, const int32_t *pairs, const int64_t *offsets,
                              const float *gscale, float sign, float *grad_score, float *grad_noise,
                              void *stream);

/*
 * Scaled-Inner-Product interval scorer (tcgen05 / TMEM, TF32 operands, fp32 accumulate).  Replaces
 * LayersTransformer.py:410-440 (everything of ScaledInnerProductIntervalScorer.forward after the
 * Linear projection :406): per track n
 *     out[e][b][n] = (sum_d q[n][e][d] * k[n][b][d]) / sqrt(D) * |e-b|  +  [e==b] * diag[n][e]
 * q, k: [n_tracks][T][D] fp32 contiguous, 16-byte aligned, D a multiple of 32; diag: [n_tracks][T].
 * out_score: [T][T][n_tracks] (the CRF's layout).  ONLY e >= b is written -- the only part the
 * semi-CRF reads (the reference fills the full square).
 */
int tkb_sip_score(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                  float *out_score, void *stream);

/* Same, writing cell (end, begin) at out_score + (end * T + begin) * pitch (pitch >= n_tracks): a pitch that is a
 * multiple of 4 gives the semi-CRF sweep its 16-byte copy path for n_tracks = 90 (tkb_semicrf_sweep_pitched). */
int tkb_sip_score_pitched(const float *q, const float *k, const float *diag, int n_tracks, int T, int D,
                          float *out_score, int64_t pitch, void *stream);

/*
 * Same contraction with an explicit scale instead of 1/sqrt(D):
 *     out[e][b][n] = (sum_d q[n][e][d] * k[n][b][d]) * scale * |e-b|  +  [e==b] * diag[n][e]
 * This is how the host gets fp32-grade products out of the TF32 tensor cores (the reference's inference path,
 * transcribe.py, never enables TF32): q and k are split into a TF32-exact high part and a residual, and the kernel
 * is called once on the concatenation [q_hi, q_hi, q_lo] x [k_hi, k_lo, k_hi] (D' = 3 D) with scale = 1/sqrt(D)
 * ("3xTF32"; transkun_b200/LayersTransformer.py).
 */
int tkb_sip_score_scaled(const float *q, const float *k, const float *diag, int n_tracks, int T, int D, float scale,
                         float *out_score, int64_t pitch, void *stream);

/*
 * The operand split of the 3xTF32 mode in one pass: q, k [rows][D] fp32 -> q3 = [q_hi | q_hi | q_lo],
 * k3 = [k_hi | k_lo | k_hi], each [rows][3 D], where x_hi is x with the 13 low mantissa bits cleared (exact in TF32)
 * and x_lo = x - x_hi (exact in fp32).  D a multiple of 4, pointers 16-byte aligned.
 */
int tkb_sip_split3(const float *q, const float *k, long long rows, int D, float *q3, float *k3, void *stream);

/*
 * Adjoint of the scorer's epilogue (training): from grad_score = dL/dS, [T][T][pitch] fp32 (track innermost, what
 * tkb_semicrf_marginals writes), produce in one pass
 *     out_gl   [n_tracks][T][T]: Gl[n][e][b] = grad_score[e][b][n] * scale * (e-b) for b < e, else 0
 *     out_gdiag[n_tracks][T]   : grad_score[e][e][n]
 * so that dq[n] = Gl[n] @ k[n] and dk[n] = Gl[n]^T @ q[n] are two plain batched library GEMMs
 * (autograd of LayersTransformer.py:410-440).
 */
int tkb_sip_backward_prep(const float *grad_score, int64_t pitch, int n_tracks, int T, float scale, float *out_gl,
                          float *out_gdiag, void *stream);

/*
 * STFT / log-mel frontend.  Replaces Util.py:104-113 (Spectrum.forward) and :156-167
 * (MelSpectrum.forward with log=True): for every frame, audio channel and window
 *     X = rfft(frame * window, norm="ortho");  P = |X|^2;  (to_mono: mean over channels)
 *     mel[m] = sum_f P[f] * melfb[f][m];  out = (log(mel+eps) - log(eps)) / (-log(eps))
 * frames: element (b,c,f,i) at frames[b*stride_b + c*stride_c + f*stride_f + i] (float units; the
 *   overlapping `unfold` view of makeFrame, Util.py:21-43, is read in place);
 * windows [nWin][W]; melfb [W/2+1][nMel] dense row-major; band_lo/band_cnt [nMel]: first row and
 *   number of rows of each mel filter's support (rows outside it are not read);
 * out [B][to_mono ? 1 : C][F][nMel][nWin] fp32.
 * workspace: tkb_logmel_workspace_bytes(B, C, F, W, nWin) bytes (windowed frames, spectra, cuFFT
 *   work area).  The cuFFT plan for (W, B*C*F*nWin) is created on first use and cached.
 */
size_t tkb_logmel_workspace_bytes(int B, int C, int F, int W, int nWin);
int tkb_logmel(const float *frames, int64_t stride_b, int64_t stride_c, int64_t stride_f, int B, int C,
               int F, int W, const float *windows, int nWin, const float *melfb, const int32_t *band_lo,
               const int32_t *band_cnt, int nMel, int to_mono, float eps, float *out, void *workspace,
               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TRANSKUN_B200_H */
