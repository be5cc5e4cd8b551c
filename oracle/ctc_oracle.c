/* ctc_oracle.c -- plain-C CPU restatement of the warp-ctc CPU algorithm (single thread).
 * TEST INFRASTRUCTURE ONLY: used by tests/, smoke() and bench.py's cpu_baseline leg as the checker;
 * nothing under kaldi-aslp_b200/ links or calls it.
 *
 * Restates, in order (paths relative to /root/reference/src/warp-ctc/include/detail/):
 *   softmax over (t, n) rows ............ cpu_ctc.h:158-179
 *   blank-interleaved labels, repeats,
 *   start/end increment tables .......... cpu_ctc.h:119-155
 *   alpha recursion on the valid window . cpu_ctc.h:217-262
 *   beta recursion + gradient ........... cpu_ctc.h:269-367
 *   per-utterance driver ................ cpu_ctc.h:181-215, 369-428 (skip when L + repeats > T)
 *   log-add ............................. ctc_helper.h:55-68
 * Pinned against warp-ctc's own known-answer tests (tests/test_cpu.cpp:12-242) in
 * tests/test_oracle_ctc.py, and against the compiled reference (oracle/_ref) on random cases.
 * Layout: acts/grads are (t, n, k) contiguous, row stride == K; blank = 0.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static float logadd(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return log1pf(expf(-fabsf(a - b))) + (a > b ? a : b);
}

/* returns 0 on success */
int ctc_oracle_cost_and_grad(const float* acts, float* grads, const int* flat_labels, const int* label_lengths,
                             const int* input_lengths, int K, int mb, float* costs) {
  int maxT = 0, n, t, k, i;
  for (n = 0; n < mb; ++n) if (input_lengths[n] > maxT) maxT = input_lengths[n];
  float* probs = (float*)calloc((size_t)maxT * mb * K, sizeof(float));
  if (!probs) return 1;
  /* softmax of every valid row */
  for (n = 0; n < mb; ++n) {
    for (t = 0; t < input_lengths[n]; ++t) {
      const float* x = acts + ((size_t)t * mb + n) * K;
      float* p = probs + ((size_t)t * mb + n) * K;
      float mx = -INFINITY, den = 0.f;
      for (k = 0; k < K; ++k) if (x[k] > mx) mx = x[k];
      for (k = 0; k < K; ++k) den += expf(x[k] - mx);
      for (k = 0; k < K; ++k) p[k] = expf(x[k] - mx) / den;
    }
  }
  int label_off = 0;
  for (n = 0; n < mb; ++n) {
    const int T = input_lengths[n], L = label_lengths[n], S = 2 * L + 1;
    const int* lab_in = flat_labels + label_off;
    label_off += L;
    int* lab = (int*)malloc(sizeof(int) * S);
    int* s_inc = (int*)malloc(sizeof(int) * (S + 1));
    int* e_inc = (int*)malloc(sizeof(int) * (S + 1));
    float* alpha = (float*)malloc(sizeof(float) * (size_t)S * (T > 0 ? T : 1));
    float* beta = (float*)malloc(sizeof(float) * S);
    float* out = (float*)malloc(sizeof(float) * K);
    int ns = 0, ne = 0, repeats = 0;
    s_inc[ns++] = 1;
    for (i = 1; i < L; ++i) {
      if (lab_in[i - 1] == lab_in[i]) { s_inc[ns++] = 1; s_inc[ns++] = 1; e_inc[ne++] = 1; e_inc[ne++] = 1; ++repeats; }
      else { s_inc[ns++] = 2; e_inc[ne++] = 2; }
    }
    e_inc[ne++] = 1;
    for (i = 0; i < L; ++i) { lab[2 * i] = 0; lab[2 * i + 1] = lab_in[i]; }
    lab[S - 1] = 0;
    costs[n] = 0.f;
    if (L + repeats <= T) {
      const size_t fs = (size_t)mb * K;                 /* floats between frames of one utterance */
      const float* p0 = probs + (size_t)n * K;
      float* g0 = grads + (size_t)n * K;
      for (i = 0; i < S * T; ++i) alpha[i] = -INFINITY;
      for (i = 0; i < S; ++i) beta[i] = -INFINITY;
      /* ---- alpha */
      int start = (L + repeats - T < 0) ? 0 : 1, end = S > 1 ? 2 : 1;
      for (i = start; i < end; ++i) alpha[i] = logf(p0[lab[i]]);
      for (t = 1; t < T; ++t) {
        const int remain = L + repeats - (T - t);
        if (remain >= 0) start += s_inc[remain];
        if (t <= L + repeats) end += e_inc[t - 1];
        const float* p = p0 + (size_t)t * fs;
        const float* ap = alpha + (size_t)(t - 1) * S;
        float* ac = alpha + (size_t)t * S;
        for (i = start; i < end; ++i) {
          if (i == 0) { ac[0] = ap[0] + logf(p[0]); continue; }
          float v = logadd(ap[i], ap[i - 1]);
          if (lab[i] != 0 && i != 1 && lab[i] != lab[i - 2]) v = logadd(v, ap[i - 2]);
          ac[i] = v + logf(p[lab[i]]);
        }
      }
      float ll = -INFINITY;
      for (i = start; i < end; ++i) ll = logadd(ll, alpha[(size_t)(T - 1) * S + i]);
      costs[n] = -ll;
      /* ---- beta + gradient */
      start = S > 1 ? S - 2 : 0;
      end = (T > L + repeats) ? S : S - 1;
      for (t = T - 1; t >= 0; --t) {
        const float* p = p0 + (size_t)t * fs;
        float* a = alpha + (size_t)t * S;
        for (k = 0; k < K; ++k) out[k] = -INFINITY;
        if (t == T - 1) {
          for (i = start; i < end; ++i) {
            beta[i] = logf(p[lab[i]]);
            a[i] += beta[i];
            out[lab[i]] = logadd(a[i], out[lab[i]]);
          }
        } else {
          const int remain = L + repeats - (T - t);
          if (remain >= -1) start -= s_inc[remain + 1];
          if (t < L + repeats) end -= e_inc[t];
          const int endloop = (end == S) ? end - 1 : end;
          for (i = start; i < endloop; ++i) {          /* ascending, in place: beta[i+1], beta[i+2] are still old */
            float v = logadd(beta[i], beta[i + 1]);
            if (lab[i] != 0 && i != S - 2 && lab[i] != lab[i + 2]) v = logadd(v, beta[i + 2]);
            beta[i] = v + logf(p[lab[i]]);
            a[i] += beta[i];
            out[lab[i]] = logadd(a[i], out[lab[i]]);
          }
          if (end == S) {
            beta[S - 1] = beta[S - 1] + logf(p[0]);
            a[S - 1] += beta[S - 1];
            out[lab[S - 1]] = logadd(a[S - 1], out[lab[S - 1]]);
          }
        }
        float* g = g0 + (size_t)t * fs;
        for (k = 0; k < K; ++k) {
          if (out[k] == 0.0f || out[k] == -INFINITY || p[k] == 0.0f) g[k] = p[k];
          else g[k] = p[k] - expf(out[k] - logf(p[k]) - ll);
        }
      }
    }
    free(lab); free(s_inc); free(e_inc); free(alpha); free(beta); free(out);
  }
  free(probs);
  return 0;
}
