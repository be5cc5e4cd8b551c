// Eesen-style CTC on PROBABILITIES (softmax outputs), errors back-propagated through the softmax.
// Reference: Ctc::EvalParallel (src/aslp-nnet/ctc-loss.cc:115-189); kernels _compute_ctc_alpha_multiple_sequence
// (src/aslp-cudamatrix/cu-kernels.cu:3276-3315), _compute_ctc_beta_multiple_sequence (:3391-3451),
// _compute_ctc_error_multiple_sequence (:3512-3534); sentinel log arithmetic (ctc-utils.h:29-95, log_zero = -1e30).
// The reference launches one kernel per time row for alpha and again for beta (2T launches, labels re-uploaded each
// time) and a serial per-(frame, class) loop over all states for the error.  Here: one block per utterance walks all T
// rows (alpha rows spilled to the workspace), and the beta sweep produces the per-class log-sum, the error, the softmax
// back-propagation and the final diff row in place.
#include "common.cuh"

namespace {

constexpr float LOG_ZERO = -1e30f, LOG_INF = 1e30f, EXP_LIMIT = 88.722839f, FMAX = 3.4028235e+038f;
__device__ __forceinline__ float AddAB(float a, float b) { return (a == LOG_ZERO || b == LOG_ZERO) ? LOG_ZERO : a + b; }
__device__ __forceinline__ float SubAB(float a, float b) { return a == LOG_ZERO ? LOG_ZERO : (b == LOG_ZERO ? LOG_INF : a - b); }
__device__ __forceinline__ float ExpA(float a) { return a <= LOG_ZERO ? 0.f : (a >= EXP_LIMIT ? FMAX : expf(a)); }
__device__ __forceinline__ float LogAPlusB(float a, float b) {
  return (b < a) ? AddAB(a, logf(1.f + ExpA(SubAB(b, a)))) : AddAB(b, logf(1.f + ExpA(SubAB(a, b))));
}

template <int G>
__global__ void __launch_bounds__(G) eesen_kernel(float* diff, int ldd, const float* probs, int ldp, int T, int S, int K,
                                                  const int* labels, int Lexp, const int* seq_len, float* pzx_out, float* alphas_ws) {
  extern __shared__ float sm[];
  const int s = blockIdx.x, tid = threadIdx.x;
  int* lab = reinterpret_cast<int*>(sm);       // [Lexp]
  float* a_prev = sm + Lexp;                   // [Lexp + 2]
  float* a_cur = a_prev + Lexp + 2;            // [Lexp + 2]
  float* ab = a_cur + Lexp + 2;                // [Lexp]
  float* red = ab + Lexp;                      // [32]
  __shared__ int sh_len;
  for (int j = tid; j < Lexp; j += G) lab[j] = labels[(size_t)s * Lexp + j];
  __syncthreads();
  if (tid == 0) { int n = 0; for (int j = 0; j < Lexp; ++j) n += (lab[j] != -1); sh_len = n; }
  __syncthreads();
  const int label_len = sh_len, len = min(seq_len[s], T);
  float* alphas = alphas_ws + (size_t)s * T * Lexp;
  // ---- alpha (rows t < len; rows beyond stay log_zero and are never read)
  for (int t = 0; t < len; ++t) {
    const float* p = probs + ((size_t)t * S + s) * ldp;
    for (int j = tid; j < Lexp; j += G) {
      const int cls = lab[j];
      float v = LOG_ZERO;
      if (cls != -1) {
        const float lp = logf(p[cls]);
        if (t == 0) v = (j < 2) ? lp : LOG_ZERO;
        else if (j > 1) {
          float tmp = LogAPlusB(a_prev[j - 1], a_prev[j]);
          if (!(j % 2 == 0 || lab[j - 2] == cls)) tmp = LogAPlusB(a_prev[j - 2], tmp);
          v = AddAB(lp, tmp);
        } else if (j == 1) v = AddAB(lp, LogAPlusB(a_prev[0], a_prev[1]));
        else v = AddAB(lp, a_prev[0]);
      }
      a_cur[j] = v;
      alphas[(size_t)t * Lexp + j] = v;
    }
    __syncthreads();
    float* tmp = a_prev; a_prev = a_cur; a_cur = tmp;
  }
  // log p(z|x) (ctc-loss.cc:166-173, evaluated in double there)
  float pzx = LOG_ZERO;
  if (len > 0 && label_len >= 1) {
    // LogAPlusB<double>(alpha[last], alpha[last-1]): the float sentinel -1e30 is an ordinary (tiny) value in double
    const double t1 = a_prev[label_len - 1], t2 = label_len >= 2 ? (double)a_prev[label_len - 2] : (double)LOG_ZERO;
    const double hi = t1 > t2 ? t1 : t2, lo = t1 > t2 ? t2 : t1, d = lo - hi;
    const double e = d <= -1e100 ? 0.0 : exp(d);
    pzx = (float)(hi + log(1.0 + e));
  }
  if (tid == 0) pzx_out[s] = pzx;
  __syncthreads();
  // ---- beta sweep + error + softmax back-propagation, row by row
  float* b_next = a_prev;      // reuse
  float* b_cur = a_cur;
  for (int j = tid; j < Lexp + 2; j += G) { b_next[j] = LOG_ZERO; b_cur[j] = LOG_ZERO; }
  __syncthreads();
  for (int t = len - 1; t >= 0; --t) {
    const float* p = probs + ((size_t)t * S + s) * ldp;
    const float* al = alphas + (size_t)t * Lexp;
    for (int j = tid; j < Lexp; j += G) {
      const int cls = lab[j];
      float v = LOG_ZERO;
      if (cls != -1) {
        const float lp = logf(p[cls]);
        if (t == len - 1) v = (j > label_len - 3) ? lp : LOG_ZERO;
        else if (j < label_len - 2) {
          float tmp = LogAPlusB(b_next[j + 1], b_next[j]);
          if (!(j % 2 == 0 || lab[j + 2] == cls)) tmp = LogAPlusB(b_next[j + 2], tmp);
          v = AddAB(lp, tmp);
        } else if (j == label_len - 2) v = AddAB(lp, LogAPlusB(b_next[j + 1], b_next[j]));
        else v = AddAB(lp, b_next[j]);
      }
      b_cur[j] = v;
      ab[j] = (cls != -1) ? AddAB(al[j], v) : LOG_ZERO;
    }
    __syncthreads();
    // error per class, times y, and the row sum for the softmax Jacobian
    float part = 0.f;
    float* drow = diff + ((size_t)t * S + s) * ldd;
    for (int k = tid; k < K; k += G) {
      float err = LOG_ZERO;
      for (int j = 0; j < label_len; ++j) if (lab[j] == k) err = LogAPlusB(err, ab[j]);
      const float y = p[k];
      const float val = ExpA(SubAB(err, AddAB(pzx, y == 0.f ? LOG_ZERO : 2.f * logf(y))));
      const float e = -1.0f * val * y;                    // ctc_err_.MulElements(net_out)
      drow[k] = e;
      part += e;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    float row_sum = 0.f;
    for (int w = 0; w < G / 32; ++w) row_sum += red[w];
    for (int k = tid; k < K; k += G) drow[k] = drow[k] - p[k] * row_sum;     // diff = err.*y - y * sum(err.*y)
    __syncthreads();
    float* tmp = b_next; b_next = b_cur; b_cur = tmp;
  }
  // rows past the utterance end carry no error
  for (int t = len; t < T; ++t) {
    float* drow = diff + ((size_t)t * S + s) * ldd;
    for (int k = tid; k < K; k += G) drow[k] = 0.f;
  }
}

}  // namespace

extern "C" {

size_t aslp_ctc_eesen_workspace_bytes(int T, int S, int K, int Lexp_max) {
  (void)K;
  return (size_t)S * T * Lexp_max * sizeof(float) + 256;
}

int aslp_ctc_eesen(aslp_stream_t s, float* diff, int ldd, const float* probs, int ldp, int T, int S, int K, const int* labels_dev,
                   int Lexp_max, const int* seq_len_dev, float* pzx_dev, void* workspace, size_t workspace_bytes) {
  if (T == 0 || S == 0) return 0;
  ASLP_REQUIRE(diff != nullptr && probs != nullptr && labels_dev != nullptr && seq_len_dev != nullptr && pzx_dev != nullptr);
  ASLP_REQUIRE(workspace != nullptr && workspace_bytes >= aslp_ctc_eesen_workspace_bytes(T, S, K, Lexp_max));
  const size_t smem = ((size_t)4 * Lexp_max + 4 + 32) * sizeof(float);
  ASLP_REQUIRE(smem <= 200 * 1024);
  constexpr int G = 128;
  ASLP_CUDA(cudaFuncSetAttribute(eesen_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eesen_kernel<G><<<S, G, smem, (cudaStream_t)s>>>(diff, ldd, probs, ldp, T, S, K, labels_dev, Lexp_max, seq_len_dev, pzx_dev, (float*)workspace);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
