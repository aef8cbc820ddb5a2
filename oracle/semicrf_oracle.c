/*
 * oracle/semicrf_oracle.c -- CPU restatement of Transkun's neural semi-CRF.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under transkun_b200/ may import, link or
 * execute this file.  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg have an independent CPU answer
 * to compare the CUDA path against.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this file is pinned against outputs of the reference
 * itself, generated in the build container by tests/golden/make_golden.py
 * (imports /root/reference unmodified) and committed under tests/golden/.
 *
 * Every function cites the reference lines it restates; all paths are relative
 * to /root/reference/transkun/CRF/NeuralSemiCRFInterval.py.
 *
 * Layout (same as the reference): score[e][b][n], e = end, b = begin, n = track,
 * n innermost; noise[t][n] scores "no event between t and t+1".  fp32 everywhere
 * (the reference allocates q/v with torch.zeros -> float32, :22, :116).
 *
 * Loop order.  The reference's backward Viterbi pulls a strided column per step
 * (after a transpose().contiguous(), :27).  Here the same candidates are formed
 * row by row ("push": once q[e] is final, every column b<e absorbs q[e]+S[e,b]).
 * Each candidate is still the single fp32 add fl(q[e]+S[e,b]) and max is exact,
 * so values, argmax and tie-breaks are bit-identical to the reference; only the
 * memory walk differs (contiguous rows).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S_AT(e, b, n) score[((size_t)(e) * T + (size_t)(b)) * N + (size_t)(n)]

int tko_version(void) { return 1; }

void tko_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int tko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* x * (x > 0): the reference multiplies by a bool mask (:29, :51, :122, :144). */
static inline float relu_mask(float x) { return x * (x > 0.0f ? 1.0f : 0.0f); }

/* F.softplus, beta=1, threshold=20 (:218, :232, :261, :395). */
static inline float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

/*
 * viterbiBackward DP, :12-52.
 *   q[T-1] = S[T-1,T-1]*(S>0)                                              (:29)
 *   for b = T-2..0: tmp = [q[b+1]+noise[b], q[b+1..T-1]+S[b+1..T-1,b]]     (:36-42)
 *                   curV, sel = tmp.max(0)  (first index wins ties)        (:44)
 *                   ptr = sel-1 ; q[b] = curV + S[b,b]*(S[b,b]>0)          (:46-51)
 * Outputs: q[T][N]; sel[T][N] int32 = absolute end position chosen for begin b,
 * or -1 for "skip" (reference ptr value k>=0 means end = b+1+k, :89).  sel[T-1]
 * is unused (-1).
 * Tie rule: candidate order is skip, e=b+1, e=b+2, ... and torch.max returns the
 * first maximal index, so skip beats everything and a nearer end beats a farther
 * one.  Rows are pushed in descending e with ">=", which yields the smallest e.
 */
void tko_viterbi_backward_dp(const float *score, const float *noise, int T, int N,
                             float *q, int32_t *sel) {
    float *acc = (float *)malloc(sizeof(float) * (size_t)T * N);
    for (size_t i = 0; i < (size_t)T * N; ++i) { acc[i] = -INFINITY; sel[i] = -1; }
#pragma omp parallel
    {
        for (int e = T - 1; e >= 0; --e) {
#pragma omp single
            {
                for (int n = 0; n < N; ++n) {
                    float d = S_AT(e, e, n);
                    if (e == T - 1) {
                        q[(size_t)e * N + n] = relu_mask(d);
                        sel[(size_t)e * N + n] = -1;
                    } else {
                        float skip = q[(size_t)(e + 1) * N + n] + noise[(size_t)e * N + n];
                        float best = acc[(size_t)e * N + n];
                        /* skip is candidate 0: it wins unless an interval is strictly better */
                        if (!(best > skip)) { best = skip; sel[(size_t)e * N + n] = -1; }
                        q[(size_t)e * N + n] = best + relu_mask(d);
                    }
                }
            } /* implicit barrier */
            const float *qe = q + (size_t)e * N;
#pragma omp for schedule(static)
            for (int b = 0; b < e; ++b) {
                const float *srow = &S_AT(e, b, 0);
                float *a = acc + (size_t)b * N;
                int32_t *s = sel + (size_t)b * N;
                for (int n = 0; n < N; ++n) {
                    float x = qe[n] + srow[n];
                    if (x >= a[n]) { a[n] = x; s[n] = e; }
                }
            } /* implicit barrier */
        }
    }
    free(acc);
}

/*
 * viterbi (forward=True) DP, :106-145.
 *   v[0] = S[0,0]*(S>0)                                                    (:122)
 *   for i = 1..T-1: tmp = [v[i-1]+noise[i-1], v[0..i-1]+S[i,0..i-1]]       (:129-135)
 *                   curV, sel = tmp.max(0) ; ptr = sel-1                   (:137-139)
 *                   v[i] = curV + S[i,i]*(S[i,i]>0)                        (:144)
 * sel[i][n] = absolute begin position chosen for end i, or -1 for skip.
 * Candidate order skip, j=0, j=1, ...: the smallest j wins ties.
 */
void tko_viterbi_forward_dp(const float *score, const float *noise, int T, int N,
                            float *v, int32_t *sel) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        v[n] = relu_mask(S_AT(0, 0, n));
        sel[n] = -1;
    }
    for (int i = 1; i < T; ++i) {
#pragma omp parallel for schedule(static)
        for (int n = 0; n < N; ++n) {
            float best = v[(size_t)(i - 1) * N + n] + noise[(size_t)(i - 1) * N + n];
            int32_t bi = -1;
            for (int j = 0; j < i; ++j) {
                float x = v[(size_t)j * N + n] + S_AT(i, j, n);
                if (x > best) { best = x; bi = j; }
            }
            v[(size_t)i * N + n] = best + relu_mask(S_AT(i, i, n));
            sel[(size_t)i * N + n] = bi;
        }
    }
}

/*
 * Backtracking of viterbiBackward, :61-102.  start[n] = forcedStartPos (default
 * 0).  pairs: [N][2*T][2] int32 (begin,end), counts[N].
 */
void tko_backtrack_backward(const float *score, const int32_t *sel, int T, int N,
                            const int32_t *start, int32_t *pairs, int32_t *counts) {
    for (int n = 0; n < N; ++n) {
        int32_t *out = pairs + (size_t)n * 2 * T * 2;
        int c = 0;
        int j = start ? start[n] : 0;
        while (j < T - 1) {
            if (S_AT(j, j, n) > 0.0f) { out[2 * c] = j; out[2 * c + 1] = j; ++c; }   /* :81-82 */
            int32_t e = sel[(size_t)j * N + n];
            if (e < 0) { j += 1; }                                                   /* :85-86 */
            else { out[2 * c] = j; out[2 * c + 1] = e; ++c; j = e; }                 /* :88-94 */
        }
        if (S_AT(T - 1, T - 1, n) > 0.0f) { out[2 * c] = T - 1; out[2 * c + 1] = T - 1; ++c; } /* :97-98 */
        counts[n] = c;
    }
}

/*
 * Backtracking of viterbi (forward=True), :157-199.  start[n] = forcedStartPos
 * (an END position, default T-1).  The list is reversed at the end (:196).
 */
void tko_backtrack_forward(const float *score, const int32_t *sel, int T, int N,
                           const int32_t *start, int32_t *pairs, int32_t *counts) {
    for (int n = 0; n < N; ++n) {
        int32_t *out = pairs + (size_t)n * 2 * T * 2;
        int c = 0;
        int j = start ? start[n] : T - 1;
        while (j > 0) {
            if (S_AT(j, j, n) > 0.0f) { out[2 * c] = j; out[2 * c + 1] = j; ++c; }   /* :177-178 */
            int32_t b = sel[(size_t)j * N + n];
            if (b < 0) { j -= 1; }                                                   /* :181-183 */
            else { out[2 * c] = b; out[2 * c + 1] = j; ++c; j = b; }                 /* :185-190 */
        }
        if (S_AT(0, 0, n) > 0.0f) { out[2 * c] = 0; out[2 * c + 1] = 0; ++c; }       /* :192-193 */
        for (int a = 0, z = c - 1; a < z; ++a, --z) {                                /* :196 */
            int32_t t0 = out[2 * a], t1 = out[2 * a + 1];
            out[2 * a] = out[2 * z]; out[2 * a + 1] = out[2 * z + 1];
            out[2 * z] = t0; out[2 * z + 1] = t1;
        }
        counts[n] = c;
    }
}

/*
 * computeLogZ forward sweep, :206-246 (and the v half of forward_backward,
 * :398-410).  v[0] = softplus(S[0,0]);
 * v[i] = logsumexp([v[i-1]+noise[i-1], v[0..i-1]+S[i,0..i-1]]) + softplus(S[i,i]).
 * torch.logsumexp = max-shifted log(sum(exp())).  alpha: [T][N]; logZ = alpha[T-1].
 */
void tko_logz_forward(const float *score, const float *noise, int T, int N, float *alpha) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) alpha[n] = softplusf(S_AT(0, 0, n));
    for (int i = 1; i < T; ++i) {
#pragma omp parallel for schedule(static)
        for (int n = 0; n < N; ++n) {
            float skip = alpha[(size_t)(i - 1) * N + n] + noise[(size_t)(i - 1) * N + n];
            float m = skip;
            for (int j = 0; j < i; ++j) {
                float x = alpha[(size_t)j * N + n] + S_AT(i, j, n);
                if (x > m) m = x;
            }
            float s = expf(skip - m);
            for (int j = 0; j < i; ++j) s += expf(alpha[(size_t)j * N + n] + S_AT(i, j, n) - m);
            alpha[(size_t)i * N + n] = (logf(s) + m) + softplusf(S_AT(i, i, n));
        }
    }
}

/*
 * Backward (beta) sweep: the q recursion of forward_backwardOld, :303-327, which
 * forward_backward obtains by running the forward sweep on the flipped tensor
 * (:386-414).  beta[T-1] = softplus(S[T-1,T-1]);
 * beta[b] = logaddexp(beta[b+1]+noise[b], logsumexp_e>b(beta[e]+S[e,b])) + softplus(S[b,b]).
 * logZ = beta[0] (equal to alpha[T-1] up to rounding, SURVEY.md section 4 inv. 2).
 */
void tko_logz_backward(const float *score, const float *noise, int T, int N, float *beta) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        beta[(size_t)(T - 1) * N + n] = softplusf(S_AT(T - 1, T - 1, n));
        for (int b = T - 2; b >= 0; --b) {
            float skip = beta[(size_t)(b + 1) * N + n] + noise[(size_t)b * N + n];
            float m = skip;
            for (int e = b + 1; e < T; ++e) {
                float x = beta[(size_t)e * N + n] + S_AT(e, b, n);
                if (x > m) m = x;
            }
            float s = expf(skip - m);
            for (int e = b + 1; e < T; ++e) s += expf(beta[(size_t)e * N + n] + S_AT(e, b, n) - m);
            beta[(size_t)b * N + n] = (logf(s) + m) + softplusf(S_AT(b, b, n));
        }
    }
}

/*
 * Marginals of forward_backward, :417-447.
 *   grad[e,b]   = exp(alpha[b] + beta[e] - logZ + S[e,b])                 b<e   (:424,:438)
 *   grad[t,t]   = exp(alpha[t] + beta[t] - logZ + S[t,t] - 2 softplus(S[t,t]))  (:427)
 *   grad[e,b]   = 0                                                       b>e   (:436,:440)
 *   gradNoise[t]= exp(alpha[t] + beta[t+1] + noise[t] - logZ)                   (:445-447)
 * logZ = alpha[T-1] (:417).
 */
void tko_marginals(const float *score, const float *noise, int T, int N, const float *alpha,
                   const float *beta, float *grad, float *gradNoise) {
    const float *logZ = alpha + (size_t)(T - 1) * N;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < T; ++e) {
        for (int b = 0; b < T; ++b) {
            float *g = grad + ((size_t)e * T + b) * N;
            if (b > e) { memset(g, 0, sizeof(float) * N); continue; }
            for (int n = 0; n < N; ++n) {
                float s = S_AT(e, b, n);
                float x = alpha[(size_t)b * N + n] + ((beta[(size_t)e * N + n] - logZ[n]) + s);
                if (b == e) x = x - 2.0f * softplusf(s);
                g[n] = expf(x);
            }
        }
    }
    for (int t = 0; t + 1 < T; ++t)
        for (int n = 0; n < N; ++n)
            gradNoise[(size_t)t * N + n] = expf(alpha[(size_t)t * N + n] + beta[(size_t)(t + 1) * N + n] +
                                                noise[(size_t)t * N + n] - logZ[n]);
}

/*
 * evalPath, :508-550 (same value as evalPathSlow :478-502).
 *   cum = cumsum(pad(noise)) ; result[n] = cum[T-1,n] + sum_(b,e) (S[e,b,n] - (cum[e,n]-cum[b,n]))
 * Intervals are given CSR-style: pairs[offsets[n] .. offsets[n+1])[2] = (begin,end).
 */
void tko_eval_path(const float *score, const float *noise, int T, int N, const int32_t *pairs,
                   const int64_t *offsets, float *out) {
    float *cum = (float *)calloc((size_t)T * N, sizeof(float));
    for (int t = 1; t < T; ++t)
        for (int n = 0; n < N; ++n)
            cum[(size_t)t * N + n] = cum[(size_t)(t - 1) * N + n] + noise[(size_t)(t - 1) * N + n];
    for (int n = 0; n < N; ++n) {
        float acc = 0.0f;
        for (int64_t k = offsets[n]; k < offsets[n + 1]; ++k) {
            int b = pairs[2 * k], e = pairs[2 * k + 1];
            acc += S_AT(e, b, n) - (cum[(size_t)e * N + n] - cum[(size_t)b * N + n]);
        }
        out[n] = acc + cum[(size_t)(T - 1) * N + n];
    }
    free(cum);
}
