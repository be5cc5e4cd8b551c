// CTC forward-backward in log space behind the unchanged warp-ctc C API (include/ctc.h).
// Reference: CpuCTC<float> (src/warp-ctc/include/detail/cpu_ctc.h): softmax :158-179,
// setup_labels :119-155, compute_alphas :217-262, compute_betas_and_grad :269-367,
// cost_and_grad :369-428, log_plus (detail/ctc_helper.h:55-68).  The reference's GPU path
// (gpu_ctc_kernels.h) is not followed: it does not compile for sm_70+.
//
// Three launches per minibatch, labels uploaded once:
//   1. ctc_softmax_kernel : warp per (t, n) row, probabilities into the workspace (HBM-bound);
//   2. ctc_dp_kernel      : one thread group per utterance (a single warp when the minibatch is
//                           large enough to fill the chip, a wider CTA when it is latency-bound):
//                           alpha sweep (rows spilled to the workspace), beta sweep with the
//                           per-label log-sum of alpha*beta, exactly the reference's valid-state
//                           window (start/end, s_inc/e_inc) and in-place beta semantics;
//   3. ctc_grad_kernel    : pointwise grad = p - exp(out - log p - logZ) (HBM-bound).
#include "common.cuh"
#include "../../include/ctc.h"
#include <string.h>

namespace {

__device__ __forceinline__ float neg_inf() { return -INFINITY; }
__device__ __forceinline__ float log_plus(float p1, float p2) {
  if (p1 == neg_inf()) return p2;
  if (p2 == neg_inf()) return p1;
  return log1pf(expf(-fabsf(p1 - p2))) + fmaxf(p1, p2);
}

// ---- 1. softmax of every valid (t, n) row
__global__ void ctc_softmax_kernel(float* probs, const float* acts, const int* in_len, int K, int mb, int maxT) {
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)maxT * mb;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const int n = (int)(row % mb), t = (int)(row / mb);
    if (t >= in_len[n]) continue;
    const float* x = acts + row * K;
    float* p = probs + row * K;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, x[k]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int k = lane; k < K; k += 32) den += expf(x[k] - mx);
    den = warp_sum(den);
    for (int k = lane; k < K; k += 32) p[k] = expf(x[k] - mx) / den;
  }
}

// block-wide reductions for a block of G threads (G multiple of 32)
template <int G>
__device__ __forceinline__ float block_log_plus(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = log_plus(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (G == 32) return v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < G / 32; ++w) r = log_plus(r, red[w]);
  return r;
}

// ---- 2. alpha / beta dynamic programme, one block of G threads per utterance
template <int G>
__global__ void __launch_bounds__(G) ctc_dp_kernel(float* grads_out /* receives per-label log-sums */, const float* probs,
                                                   float* alphas_ws, float* costs_dev, int* valid_dev, const int* flat_labels,
                                                   const int* label_off, const int* label_len, const int* in_len, int K, int mb,
                                                   int maxT, int maxS) {
  extern __shared__ int smem_i[];
  const int n = blockIdx.x;
  const int T = in_len[n], L = label_len[n], S = 2 * L + 1;
  const int tid = threadIdx.x;
  int* lab = smem_i;                      // [maxS] labels with blanks
  int* s_inc = lab + maxS;                // [maxS]
  int* e_inc = s_inc + maxS;              // [maxS]
  int* cls_start = e_inc + maxS;          // [K+1] CSR of non-blank states per label
  int* cls_list = cls_start + K + 1;      // [maxS]
  float* a_prev = reinterpret_cast<float*>(cls_list + maxS);   // [maxS]
  float* a_cur = a_prev + maxS;           // [maxS]
  float* ab = a_cur + maxS;               // [maxS] alpha+beta of the current frame
  float* red = ab + maxS;                 // [32]
  __shared__ int sh_repeats;

  const int* labels = flat_labels + label_off[n];
  // ---- setup (cpu_ctc.h:119-155), sequential on one thread: O(L)
  if (tid == 0) {
    int e_counter = 0, s_counter = 0, repeats = 0;
    s_inc[s_counter++] = 1;
    for (int i = 1; i < L; ++i) {
      if (labels[i - 1] == labels[i]) {
        s_inc[s_counter++] = 1; s_inc[s_counter++] = 1;
        e_inc[e_counter++] = 1; e_inc[e_counter++] = 1;
        ++repeats;
      } else {
        s_inc[s_counter++] = 2;
        e_inc[e_counter++] = 2;
      }
    }
    e_inc[e_counter++] = 1;
    sh_repeats = repeats;
    for (int i = 0; i < L; ++i) { lab[2 * i] = 0; lab[2 * i + 1] = labels[i]; }
    lab[S - 1] = 0;
    // CSR: states of each non-blank label in ascending state order (the reference's accumulation order)
    for (int k = 0; k <= K; ++k) cls_start[k] = 0;
    for (int i = 0; i < L; ++i) cls_start[labels[i] + 1]++;
    for (int k = 0; k < K; ++k) cls_start[k + 1] += cls_start[k];
  }
  __syncthreads();
  // parallel CSR fill: position of state (2i+1) among equal labels = number of earlier equal labels
  for (int i = tid; i < L; i += G) {
    const int l = labels[i];
    int rank = 0;
    for (int j = 0; j < i; ++j) rank += (labels[j] == l);
    cls_list[cls_start[l] + rank] = 2 * i + 1;
  }
  const int repeats = sh_repeats;
  __syncthreads();

  if (L + repeats > T) {                  // cpu_ctc.h:193-195: cost 0, gradient untouched
    if (tid == 0) { costs_dev[n] = 0.f; valid_dev[n] = 0; }
    return;
  }

  float* alphas = alphas_ws + (size_t)n * maxT * maxS;
  const size_t fstride = (size_t)mb * K;                 // floats between consecutive frames of one utterance
  const float* pr = probs + (size_t)n * K;

  // ---- alpha sweep (cpu_ctc.h:217-262)
  int start = (((S / 2) + repeats - T) < 0) ? 0 : 1;
  int end = S > 1 ? 2 : 1;
  for (int i = tid; i < S; i += G) {
    const float v = (i >= start && i < end) ? logf(pr[lab[i]]) : neg_inf();
    a_prev[i] = v;
    alphas[i] = v;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    const int remain = (S / 2) + repeats - (T - t);
    if (remain >= 0) start += s_inc[remain];
    if (t <= (S / 2) + repeats) end += e_inc[t - 1];
    const float* p = pr + (size_t)t * fstride;
    for (int i = tid; i < S; i += G) {
      float v = neg_inf();
      if (i >= start && i < end) {
        if (i == 0) {
          v = a_prev[0] + logf(p[0]);
        } else {
          float prev_sum = log_plus(a_prev[i], a_prev[i - 1]);
          const int li = lab[i];
          if (li != 0 && i != 1 && li != lab[i - 2]) prev_sum = log_plus(prev_sum, a_prev[i - 2]);
          v = prev_sum + logf(p[li]);
        }
      }
      a_cur[i] = v;
      alphas[(size_t)t * maxS + i] = v;
    }
    __syncthreads();
    float* tmp = a_prev; a_prev = a_cur; a_cur = tmp;
  }
  // log-likelihood over the final window (sequential order of the reference is ascending i; the
  // block reduction uses a tree -- same terms)
  float ll = neg_inf();
  for (int i = tid; i < S; i += G) if (i >= start && i < end) ll = log_plus(ll, a_prev[i]);
  const float loglike = block_log_plus<G>(ll, red);
  if (tid == 0) { costs_dev[n] = -loglike; valid_dev[n] = 1; }

  // ---- beta sweep + per-label log-sums (cpu_ctc.h:269-367); beta lives in a_cur, in-place semantics kept
  float* betas = a_cur;
  __syncthreads();
  for (int i = tid; i < S; i += G) betas[i] = neg_inf();
  __syncthreads();
  start = S > 1 ? (S - 2) : 0;
  end = (T > (S / 2) + repeats) ? S : S - 1;
  for (int t = T - 1; t >= 0; --t) {
    const float* p = pr + (size_t)t * fstride;
    const float* al = alphas + (size_t)t * maxS;
    if (t == T - 1) {
      for (int i = tid; i < S; i += G) {
        float v = neg_inf();
        if (i >= start && i < end) {
          const float b = logf(p[lab[i]]);
          betas[i] = b;
          v = al[i] + b;
        }
        ab[i] = v;
      }
    } else {
      const int remain = (S / 2) + repeats - (T - t);
      if (remain >= -1) start -= s_inc[remain + 1];
      if (t < (S / 2) + repeats) end -= e_inc[t];
      const int endloop = (end == S) ? end - 1 : end;
      // read the old betas first, then write: reproduces the reference's ascending in-place update
      constexpr int MAXI = 8;             // states per thread per pass, kept in registers
      for (int base = 0; base < S; base += G * MAXI) {
        float nb[MAXI];
#pragma unroll
        for (int j = 0; j < MAXI; ++j) {
          const int i = base + j * G + tid;
          float v = 0.f;
          if (i < S) {
            if (i >= start && i < endloop) {
              float next_sum = log_plus(betas[i], betas[i + 1]);
              const int li = lab[i];
              if (li != 0 && i != (S - 2) && li != lab[i + 2]) next_sum = log_plus(next_sum, betas[i + 2]);
              v = next_sum + logf(p[li]);
            } else if (end == S && i == S - 1) {
              v = betas[S - 1] + logf(p[0]);
            }
          }
          nb[j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < MAXI; ++j) {
          const int i = base + j * G + tid;
          if (i < S) {
            const bool in_win = (i >= start && i < endloop) || (end == S && i == S - 1);
            if (in_win) betas[i] = nb[j];
            ab[i] = in_win ? al[i] + nb[j] : neg_inf();
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    // per-label log-sum of alpha*beta.  blank (label 0): all even states, block-parallel
    float bl = neg_inf();
    for (int i = 2 * tid; i < S; i += 2 * G) bl = log_plus(ab[i], bl);
    const float blank_sum = block_log_plus<G>(bl, red);
    float* out = grads_out + (size_t)n * K + (size_t)t * fstride;
    for (int k = tid; k < K; k += G) {
      float o = (k == 0) ? blank_sum : neg_inf();
      for (int q = cls_start[k]; q < cls_start[k + 1]; ++q) o = log_plus(ab[cls_list[q]], o);
      out[k] = o;
    }
    __syncthreads();
  }
}

// ---- 3. grad = p - exp(out - log p - logZ), with the reference's guards (cpu_ctc.h:296-307)
__global__ void ctc_grad_kernel(float* grads, const float* probs, const float* costs_dev, const int* valid_dev,
                                const int* in_len, int K, int mb, int maxT) {
  const long long total = (long long)maxT * mb * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K;
    const int n = (int)(row % mb), t = (int)(row / mb);
    if (t >= in_len[n] || !valid_dev[n]) continue;
    const float p = probs[i], o = grads[i];
    const float log_partition = -costs_dev[n];
    float g;
    if (o == 0.0f || o == -INFINITY || p == 0.0f) g = p;
    else g = p - expf(o - logf(p) - log_partition);
    grads[i] = g;
  }
}

struct Sizes { size_t alphas, probs, costs, valid, meta, total; int maxT, maxL, maxS, sumL; };
Sizes ctc_sizes(const int* label_lengths, const int* input_lengths, int K, int mb) {
  Sizes z; memset(&z, 0, sizeof(z));
  for (int i = 0; i < mb; ++i) {
    if (input_lengths[i] > z.maxT) z.maxT = input_lengths[i];
    if (label_lengths[i] > z.maxL) z.maxL = label_lengths[i];
    z.sumL += label_lengths[i];
  }
  z.maxS = 2 * z.maxL + 1;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  z.alphas = al((size_t)mb * z.maxT * z.maxS * sizeof(float));
  z.probs = al((size_t)mb * z.maxT * K * sizeof(float));
  z.costs = al((size_t)mb * sizeof(float));
  z.valid = al((size_t)mb * sizeof(int));
  z.meta = al((size_t)(3 * mb + z.sumL + 4) * sizeof(int));
  z.total = z.alphas + z.probs + z.costs + z.valid + z.meta;
  return z;
}

template <int G>
int launch_dp(cudaStream_t st, int mb, size_t smem, float* grads, const float* probs, float* alphas, float* costs, int* valid,
              const int* flat, const int* off, const int* llen, const int* ilen, int K, int maxT, int maxS) {
  ASLP_CUDA(cudaFuncSetAttribute(ctc_dp_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_dp_kernel<G><<<mb, G, smem, st>>>(grads, probs, alphas, costs, valid, flat, off, llen, ilen, K, mb, maxT, maxS);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" {

const char* ctcGetStatusString(ctcStatus_t status) {
  switch (status) {
    case CTC_STATUS_SUCCESS: return "no error";
    case CTC_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
    case CTC_STATUS_INVALID_VALUE: return "invalid value";
    case CTC_STATUS_EXECUTION_FAILED: return "execution failed";
    default: return "unknown error";
  }
}

ctcStatus_t get_workspace_size(const int* const label_lengths, const int* const input_lengths, int alphabet_size, int minibatch,
                               struct ctcComputeInfo info, size_t* size_bytes) {
  if (label_lengths == nullptr || input_lengths == nullptr || size_bytes == nullptr || alphabet_size <= 0 || minibatch <= 0)
    return CTC_STATUS_INVALID_VALUE;
  (void)info;
  *size_bytes = ctc_sizes(label_lengths, input_lengths, alphabet_size, minibatch).total;
  return CTC_STATUS_SUCCESS;
}

ctcStatus_t compute_ctc_loss(const float* const activations, float* gradients, const int* const flat_labels,
                             const int* const label_lengths, const int* const input_lengths, int alphabet_size, int minibatch,
                             float* costs, void* workspace, struct ctcComputeInfo info) {
  if (activations == nullptr || flat_labels == nullptr || label_lengths == nullptr || input_lengths == nullptr ||
      costs == nullptr || workspace == nullptr || alphabet_size <= 0 || minibatch <= 0)
    return CTC_STATUS_INVALID_VALUE;
  if (info.loc != CTC_GPU) return CTC_STATUS_EXECUTION_FAILED;     // no CPU path in this library
  if (gradients == nullptr) return CTC_STATUS_INVALID_VALUE;       // score-only mode is not on the training path
  cudaStream_t st = (cudaStream_t)info.stream;
  const int K = alphabet_size, mb = minibatch;
  const Sizes z = ctc_sizes(label_lengths, input_lengths, K, mb);
  char* ws = (char*)workspace;
  float* alphas = (float*)ws;
  float* probs = (float*)(ws + z.alphas);
  float* costs_dev = (float*)(ws + z.alphas + z.probs);
  int* valid_dev = (int*)(ws + z.alphas + z.probs + z.costs);
  int* meta = (int*)(ws + z.alphas + z.probs + z.costs + z.valid);
  // host staging of the label metadata: [label_len mb][in_len mb][label_off mb][flat sumL]
  static thread_local int* hmeta = nullptr; static thread_local size_t hmeta_cap = 0;
  const size_t nmeta = (size_t)3 * mb + z.sumL;
  if (hmeta_cap < nmeta) {
    if (hmeta) cudaFreeHost(hmeta);
    hmeta_cap = nmeta * 2 + 64;
    if (cudaMallocHost((void**)&hmeta, hmeta_cap * sizeof(int)) != cudaSuccess) { hmeta = nullptr; hmeta_cap = 0; return CTC_STATUS_MEMOPS_FAILED; }
  }   // (every call ends with a stream sync, so the staging buffer is free again here)
  int off = 0;
  for (int i = 0; i < mb; ++i) {
    hmeta[i] = label_lengths[i];
    hmeta[mb + i] = input_lengths[i];
    hmeta[2 * mb + i] = off;
    for (int j = 0; j < label_lengths[i]; ++j) {
      const int l = flat_labels[off + j];
      if (l < 0 || l >= K) return CTC_STATUS_INVALID_VALUE;        // KALDI_ASSERT(l < NumCols), warp-ctc.cc:213
      hmeta[3 * mb + off + j] = l;
    }
    off += label_lengths[i];
  }
  if (cudaMemcpyAsync(meta, hmeta, nmeta * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) return CTC_STATUS_MEMOPS_FAILED;
  const int* d_llen = meta; const int* d_ilen = meta + mb; const int* d_off = meta + 2 * mb; const int* d_flat = meta + 3 * mb;

  {
    const long long rows = (long long)z.maxT * mb;
    int blocks = (int)((rows + 7) / 8);
    if (blocks > aslp_num_sms() * 16) blocks = aslp_num_sms() * 16;
    if (blocks < 1) blocks = 1;
    ctc_softmax_kernel<<<blocks, 256, 0, st>>>(probs, activations, d_ilen, K, mb, z.maxT);
    ++g_aslp_launches;
    if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
  }
  {
    const size_t smem = ((size_t)5 * z.maxS + K + 1) * sizeof(int) + ((size_t)3 * z.maxS + 32) * sizeof(float);
    if (smem > 220 * 1024) return CTC_STATUS_INVALID_VALUE;
    // group width: one warp per utterance once the minibatch alone fills the chip, wider when latency-bound
    int G = 32;
    if (mb < aslp_num_sms() * 8) { G = 64; while (G < z.maxS && G < 256) G <<= 1; }
    int rc;
    if (G == 32) rc = launch_dp<32>(st, mb, smem, gradients, probs, alphas, costs_dev, valid_dev, d_flat, d_off, d_llen, d_ilen, K, z.maxT, z.maxS);
    else if (G == 64) rc = launch_dp<64>(st, mb, smem, gradients, probs, alphas, costs_dev, valid_dev, d_flat, d_off, d_llen, d_ilen, K, z.maxT, z.maxS);
    else if (G == 128) rc = launch_dp<128>(st, mb, smem, gradients, probs, alphas, costs_dev, valid_dev, d_flat, d_off, d_llen, d_ilen, K, z.maxT, z.maxS);
    else rc = launch_dp<256>(st, mb, smem, gradients, probs, alphas, costs_dev, valid_dev, d_flat, d_off, d_llen, d_ilen, K, z.maxT, z.maxS);
    if (rc != 0) return CTC_STATUS_EXECUTION_FAILED;
  }
  {
    const long long total = (long long)z.maxT * mb * K;
    int blocks = (int)((total + 255) / 256);
    if (blocks > aslp_num_sms() * 16) blocks = aslp_num_sms() * 16;
    ctc_grad_kernel<<<blocks, 256, 0, st>>>(gradients, probs, costs_dev, valid_dev, d_ilen, K, mb, z.maxT);
    ++g_aslp_launches;
    if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
  }
  // costs are host memory in the warp-ctc API: one small D2H + sync, as the reference's GPU path does
  if (cudaMemcpyAsync(costs, costs_dev, mb * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) return CTC_STATUS_MEMOPS_FAILED;
  if (cudaStreamSynchronize(st) != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
  return CTC_STATUS_SUCCESS;
}

}  // extern "C"
