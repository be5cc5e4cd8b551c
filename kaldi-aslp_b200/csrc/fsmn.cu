// CompactFsmn memory block as ONE fused temporal filter per direction of the pass.
// The reference (src/aslp-nnet/nnet-cfsmn-component.h:169-258; kernels cu-kernels.cu:754-782)
// materialises a [T*(P+F+1), D] product matrix twice per pass (~40x the algorithmic traffic) and
// row-sums it.  Here a CTA stages a time tile with its halo in shared memory once and every
// thread walks a sliding register window over the taps: per tap 2 shared loads feed 8 FMAs.
//   fwd : out[t,d]      = in[t,d] + sum_c coef[c,d] * in[t+c-P, d]
//   bwd : in_diff[t,d]  = out_diff[t,d] + sum_c coef[C-1-c,d] * out_diff[t+c-F, d]   (same kernel, reversed taps)
//   grad: coef_corr[c,d]= sum_t in[t+c-P,d] * out_diff[t,d]
#include "common.cuh"
#include "scratch.cuh"

namespace {

constexpr int COLS = 128;    // columns per CTA (one per thread)
constexpr int TT = 64;       // time steps per CTA
constexpr int TB = 8;        // outputs per thread per window

// y[t,d] = x[t,d] + sum_c w[c,d] * x[t + c - P, d],  w[c] = rev ? coef[C-1-c] : coef[c]
__global__ void __launch_bounds__(COLS) fsmn_filter_kernel(float* y, int ldy, const float* x, int ldx, int T, int D,
                                                           const float* coef, int ldc, int P, int C, int rev) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [TT + C - 1 + TB][COLS] (extra TB rows: the window prefetch runs ahead)
  float* cs = xs + (size_t)(TT + C - 1 + TB) * COLS;   // [C][COLS]
  const int tx = threadIdx.x;
  const int d = blockIdx.x * COLS + tx;
  const int t0 = blockIdx.y * TT;
  const bool col_ok = d < D;
  const int nrows = TT + C - 1 + TB;
  for (int r = 0; r < nrows; ++r) {
    const int t = t0 - P + r;
    xs[r * COLS + tx] = (col_ok && t >= 0 && t < T) ? x[(size_t)t * ldx + d] : 0.f;
  }
  for (int c = 0; c < C; ++c) cs[c * COLS + tx] = col_ok ? coef[(size_t)(rev ? C - 1 - c : c) * ldc + d] : 0.f;
  __syncthreads();   // (each thread only reads its own column, but keep the tile semantics explicit)
  if (!col_ok) return;
  for (int tb = 0; tb < TT; tb += TB) {
    if (t0 + tb >= T) break;
    float acc[TB], win[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) { acc[j] = xs[(tb + j + P) * COLS + tx]; win[j] = xs[(tb + j) * COLS + tx]; }
    // window invariant at tap c: win[(c + j) % TB] holds x_s[tb + c + j]
    int c = 0;
    for (; c + TB <= C; c += TB) {
#pragma unroll
      for (int cc = 0; cc < TB; ++cc) {
        const float w = cs[(c + cc) * COLS + tx];
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[j] = fmaf(w, win[(cc + j) % TB], acc[j]);
        win[cc % TB] = xs[(tb + c + cc + TB) * COLS + tx];     // slot of x_s[tb+c+cc] is free now
      }
    }
    // remaining taps (C % TB): the window is aligned again (c is a multiple of TB)
    for (int cc = 0; c + cc < C; ++cc) {
      const float w = cs[(c + cc) * COLS + tx];
#pragma unroll
      for (int j = 0; j < TB; ++j) acc[j] = fmaf(w, xs[(tb + c + cc + j) * COLS + tx], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int t = t0 + tb + j;
      if (t < T) y[(size_t)t * ldy + d] = acc[j];
    }
  }
}

// partial[chunk][c][d] = sum_{t in chunk} x[t + c - P, d] * g[t, d]
__global__ void __launch_bounds__(COLS) fsmn_grad_partial_kernel(float* partial, const float* x, int ldx, const float* g, int ldg,
                                                                 int T, int D, int P, int C, int Dpad) {
  extern __shared__ float sm[];
  float* xs = sm;                                    // [TT + C - 1 + TB][COLS]
  float* gs = xs + (size_t)(TT + C - 1 + TB) * COLS; // [TT][COLS]
  const int tx = threadIdx.x;
  const int d = blockIdx.x * COLS + tx;
  const int t0 = blockIdx.y * TT;
  const bool col_ok = d < D;
  const int nrows = TT + C - 1 + TB;
  for (int r = 0; r < nrows; ++r) {
    const int t = t0 - P + r;
    xs[r * COLS + tx] = (col_ok && t >= 0 && t < T) ? x[(size_t)t * ldx + d] : 0.f;
  }
  for (int r = 0; r < TT; ++r) {
    const int t = t0 + r;
    gs[r * COLS + tx] = (col_ok && t < T) ? g[(size_t)t * ldg + d] : 0.f;
  }
  __syncthreads();
  if (!col_ok) return;
  float* dst = partial + (size_t)blockIdx.y * C * Dpad + d;
  for (int c0 = 0; c0 < C; c0 += TB) {
    // acc[j] for tap c0 + j ; window over x_s[t + c0 + j]
    float acc[TB], win[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) { acc[j] = 0.f; win[j] = xs[(c0 + j) * COLS + tx]; }
    for (int t = 0; t < TT; t += TB) {
#pragma unroll
      for (int tt = 0; tt < TB; ++tt) {
        const float gv = gs[(t + tt) * COLS + tx];
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[j] = fmaf(gv, win[(tt + j) % TB], acc[j]);
        win[tt % TB] = xs[min(t + tt + c0 + TB, nrows - 1) * COLS + tx];
      }
    }
#pragma unroll
    for (int j = 0; j < TB; ++j) if (c0 + j < C) dst[(size_t)(c0 + j) * Dpad] = acc[j];
  }
}
__global__ void fsmn_grad_final_kernel(float* corr, int ldc, const float* partial, int chunks, int C, int D, int Dpad, float clip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * D) return;
  const int c = i / D, d = i - c * D;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[((size_t)k * C + c) * Dpad + d];
  if (clip > 0.f) s = fminf(fmaxf(s, -clip), clip);
  corr[(size_t)c * ldc + d] = s;          // beta = 0: no momentum (nnet-cfsmn-component.h:224)
}

int filter(cudaStream_t st, float* y, int ldy, const float* x, int ldx, int T, int D, const float* coef, int ldc, int P, int F, int rev) {
  const int C = P + F + 1;
  const size_t smem = ((size_t)(TT + C - 1 + TB) + C) * COLS * sizeof(float);
  if (smem > 220 * 1024) { aslp_set_last_error_msg("FSMN context too large for one shared-memory tile", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  ASLP_CUDA(cudaFuncSetAttribute(fsmn_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(aslp_div_up(D, COLS), aslp_div_up(T, TT));
  fsmn_filter_kernel<<<grid, COLS, smem, st>>>(y, ldy, x, ldx, T, D, coef, ldc, P, C, rev);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" {

int aslp_fsmn_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int T, int D, const float* coef, int ldc, int past, int future) {
  if (T == 0 || D == 0) return 0;
  ASLP_REQUIRE(past >= 0 && future >= 0);
  return filter((cudaStream_t)s, out, ldo, in, ldi, T, D, coef, ldc, past, future, 0);
}
int aslp_fsmn_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* out_diff, int ldo, int T, int D, const float* coef, int ldc, int past, int future) {
  if (T == 0 || D == 0) return 0;
  ASLP_REQUIRE(past >= 0 && future >= 0);
  // reversed taps, and the roles of past / future swap (nnet-cfsmn-component.h:228-250)
  return filter((cudaStream_t)s, in_diff, ldd, out_diff, ldo, T, D, coef, ldc, future, past, 1);
}
int aslp_fsmn_coef_grad(aslp_stream_t s, float* coef_corr, int ldc, const float* in, int ldi, const float* out_diff, int ldo, int T, int D,
                        int past, int future, float clip) {
  if (D == 0) return 0;
  cudaStream_t st = (cudaStream_t)s;
  const int C = past + future + 1;
  const int chunks = aslp_div_up(T > 0 ? T : 1, TT);
  const int Dpad = (D + 3) / 4 * 4;
  float* partial = (float*)aslp_scratch(st, (size_t)chunks * C * Dpad * sizeof(float));
  if (partial == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  const size_t smem = ((size_t)(TT + C - 1 + TB) + TT) * COLS * sizeof(float);
  if (smem > 220 * 1024) { aslp_set_last_error_msg("FSMN context too large for one shared-memory tile", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  ASLP_CUDA(cudaFuncSetAttribute(fsmn_grad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(aslp_div_up(D, COLS), chunks);
  fsmn_grad_partial_kernel<<<grid, COLS, smem, st>>>(partial, in, ldi, out_diff, ldo, T, D, past, C, Dpad);
  ASLP_CHECK_LAUNCH();
  fsmn_grad_final_kernel<<<aslp_div_up(C * D, 256), 256, 0, st>>>(coef_corr, ldc, partial, chunks, C, D, Dpad, clip);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
