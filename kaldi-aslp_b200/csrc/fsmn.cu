// CompactFsmn memory block as ONE fused temporal filter per direction of the pass.
// The reference (src/aslp-nnet/nnet-cfsmn-component.h:169-258; kernels cu-kernels.cu:754-782)
// materialises a [T*(P+F+1), D] product matrix twice per pass (~40x the algorithmic traffic) and
// row-sums it.  Here a CTA stages a [32 + taps - 1] x 128 tile with its halo in shared memory once (128-bit coalesced
// loads, all in flight at once) and every warp walks a sliding register window of 8 outputs x 4 columns per lane over
// the taps: per tap two 128-bit shared loads feed 32 FMAs.
//   fwd : out[t,d]      = in[t,d] + sum_c coef[c,d] * in[t+c-P, d]
//   bwd : in_diff[t,d]  = out_diff[t,d] + sum_c coef[C-1-c,d] * out_diff[t+c-F, d]   (same kernel, reversed taps)
//   grad: coef_corr[c,d]= sum_t in[t+c-P,d] * out_diff[t,d]
#include "common.cuh"
#include "scratch.cuh"

namespace {

constexpr int COLS = 128;    // columns per CTA = 32 lanes x float4
constexpr int TT = 32;       // time steps per CTA
constexpr int TB = 8;        // outputs (filter) / taps (gradient) per warp: the sliding register window

__device__ __forceinline__ float4 f4_fma(float4 w, float4 x, float4 a) {
  a.x = fmaf(w.x, x.x, a.x); a.y = fmaf(w.y, x.y, a.y); a.z = fmaf(w.z, x.z, a.z); a.w = fmaf(w.w, x.w, a.w);
  return a;
}
// stage rows [t_first, t_first + nrows) x columns [d0, d0 + 128) of a [T, D] matrix into tile[nrows][128]; zero outside.
// Items are 16-byte cp.async copies (zero-filled past the matrix edge through the src-size operand): no registers, no
// scoreboard, so a thread has its whole share in flight at once -- with register loads the compiler kept ONE load
// outstanding per thread and the staging loop cost 18 serial DRAM round trips (ncu: long_scoreboard 5.5 / issue).
// The caller waits with stage_wait() before the barrier.
__device__ __forceinline__ void stage_tile(float* tile, const float* x, int ldx, int T, int D, int t_first, int nrows, int d0,
                                           int row_rev_base) {   // row_rev_base >= 0: tile row r reads source row row_rev_base - r
  const int items = nrows * (COLS / 4);
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int r = i >> 5, q = i & 31;
    const int t = row_rev_base >= 0 ? row_rev_base - r : t_first + r;
    const int d = d0 + q * 4;
    const bool in = t >= 0 && t < T && d < D;
    const float* src = in ? x + (size_t)t * ldx + d : x;
    const int bytes = in ? min(4, D - d) * 4 : 0;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + r * COLS + q * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// y[t,d] = x[t,d] + sum_c w[c,d] * x[t + c - P, d],  w[c] = rev ? coef[C-1-c] : coef[c]
// CTA = TT/TB warps; warp w owns outputs t0 + w*TB .. +TB-1 of 128 columns (lane = column quad).  Per tap: one 128-bit
// shared load of the tap row and one of the next window row feed 8 x 4 FMAs.
__global__ void __launch_bounds__(TT / TB * 32) fsmn_filter_kernel(float* y, int ldy, const float* x, int ldx, int T, int D,
                                                                   const float* coef, int ldc, int P, int C, int rev) {
  extern __shared__ __align__(16) float sm[];
  const int nrows = TT + C - 1;
  float* xs = sm;                                    // [nrows + TB][COLS] (TB slack rows: the window load runs ahead)
  float* cs = xs + (size_t)(nrows + TB) * COLS;      // [C][COLS]
  const int d0 = blockIdx.x * COLS, t0 = blockIdx.y * TT;
  stage_tile(xs, x, ldx, T, D, t0 - P, nrows, d0, -1);
  stage_tile(cs, coef, ldc, C, D, 0, C, d0, rev ? C - 1 : -1);
  stage_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tb = warp * TB;
  const int d = d0 + lane * 4;
  if (t0 + tb >= T || d >= D) return;
  const float4* xq = reinterpret_cast<const float4*>(xs) + lane;     // row stride = 32 float4
  const float4* cq = reinterpret_cast<const float4*>(cs) + lane;
  float4 acc[TB], win[TB];
#pragma unroll
  for (int j = 0; j < TB; ++j) { acc[j] = xq[(tb + j + P) * 32]; win[j] = xq[(tb + j) * 32]; }
  // window invariant at tap c: win[(c + j) % TB] holds x_s[tb + c + j]
  int c = 0;
  for (; c + TB <= C; c += TB) {
#pragma unroll
    for (int cc = 0; cc < TB; ++cc) {
      const float4 w = cq[(c + cc) * 32];
#pragma unroll
      for (int j = 0; j < TB; ++j) acc[j] = f4_fma(w, win[(cc + j) % TB], acc[j]);
      win[cc % TB] = xq[(tb + c + cc + TB) * 32];                    // slot of x_s[tb+c+cc] is free now
    }
  }
  for (int cc = 0; c + cc < C; ++cc) {                               // remaining taps: c is a multiple of TB, read directly
    const float4 w = cq[(c + cc) * 32];
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[j] = f4_fma(w, xq[(tb + c + cc + j) * 32], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < TB; ++j) {
    const int t = t0 + tb + j;
    if (t < T) {
      float* p = y + (size_t)t * ldy + d;
      if (d + 3 < D) *reinterpret_cast<float4*>(p) = acc[j];
      else { p[0] = acc[j].x; if (d + 1 < D) p[1] = acc[j].y; if (d + 2 < D) p[2] = acc[j].z; }
    }
  }
}

// partial[chunk][c][d] = sum_{t in chunk} x[t + c - P, d] * g[t, d];  CTA = ceil(C/TB) warps, warp w owns taps w*TB .. +TB-1
__global__ void fsmn_grad_partial_kernel(float* partial, const float* x, int ldx, const float* g, int ldg,
                                         int T, int D, int P, int C, int Dpad) {
  extern __shared__ __align__(16) float sm[];
  const int nrows = TT + C - 1;
  float* xs = sm;                                    // [nrows + 2*TB][COLS]
  float* gs = xs + (size_t)(nrows + 2 * TB) * COLS;  // [TT][COLS]
  const int d0 = blockIdx.x * COLS, t0 = blockIdx.y * TT;
  stage_tile(xs, x, ldx, T, D, t0 - P, nrows, d0, -1);
  stage_tile(gs, g, ldg, T, D, t0, TT, d0, -1);
  stage_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = warp * TB;
  const int d = d0 + lane * 4;
  if (c0 >= C || d >= D) return;
  const float4* xq = reinterpret_cast<const float4*>(xs) + lane;
  const float4* gq = reinterpret_cast<const float4*>(gs) + lane;
  float4 acc[TB], win[TB];
#pragma unroll
  for (int j = 0; j < TB; ++j) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); win[j] = xq[(c0 + j) * 32]; }
  // acc[j] is tap c0 + j; window invariant at step t: win[(t + j) % TB] holds x_s[t + c0 + j]
#pragma unroll 1
  for (int t = 0; t < TT; t += TB) {
#pragma unroll
    for (int tt = 0; tt < TB; ++tt) {
      const float4 gv = gq[(t + tt) * 32];
#pragma unroll
      for (int j = 0; j < TB; ++j) acc[j] = f4_fma(gv, win[(tt + j) % TB], acc[j]);
      win[tt % TB] = xq[(t + tt + c0 + TB) * 32];                    // rows past nrows-1 (slack) are loaded but never used
    }
  }
  float* dst = partial + (size_t)blockIdx.y * C * Dpad + d;           // Dpad % 4 == 0: float4 stores stay inside the row
#pragma unroll
  for (int j = 0; j < TB; ++j) if (c0 + j < C) *reinterpret_cast<float4*>(dst + (size_t)(c0 + j) * Dpad) = acc[j];
}
__global__ void fsmn_grad_final_kernel(float* corr, int ldc, const float* partial, int chunks, int C, int D, int Dpad, float clip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * D) return;
  const int c = i / D, d = i - c * D;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;                       // four independent chains: the loads overlap
  int k = 0;
  for (; k + 4 <= chunks; k += 4) {
    s0 += partial[((size_t)(k + 0) * C + c) * Dpad + d]; s1 += partial[((size_t)(k + 1) * C + c) * Dpad + d];
    s2 += partial[((size_t)(k + 2) * C + c) * Dpad + d]; s3 += partial[((size_t)(k + 3) * C + c) * Dpad + d];
  }
  for (; k < chunks; ++k) s0 += partial[((size_t)k * C + c) * Dpad + d];
  float s = (s0 + s1) + (s2 + s3);
  if (clip > 0.f) s = fminf(fmaxf(s, -clip), clip);
  corr[(size_t)c * ldc + d] = s;          // beta = 0: no momentum (nnet-cfsmn-component.h:224)
}

int filter(cudaStream_t st, float* y, int ldy, const float* x, int ldx, int T, int D, const float* coef, int ldc, int P, int F, int rev) {
  const int C = P + F + 1;
  const size_t smem = ((size_t)(TT + C - 1 + TB) + C) * COLS * sizeof(float);
  if (smem > 220 * 1024) { aslp_set_last_error_msg("FSMN context too large for one shared-memory tile", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  if (((uintptr_t)y % 16) || ((uintptr_t)x % 16) || ((uintptr_t)coef % 16) || (ldy % 4) || (ldx % 4) || (ldc % 4)) {
    aslp_set_last_error_msg("FSMN operands must be 16-byte aligned with strides that are multiples of 4 floats", __FILE__, __LINE__);
    return ASLP_STATUS_INVALID_VALUE;
  }
  static size_t smem_allowed = 0;                      // raised when a wider context comes along, not on every launch
  if (smem > smem_allowed) {
    ASLP_CUDA(cudaFuncSetAttribute(fsmn_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_allowed = smem;
  }
  dim3 grid(aslp_div_up(D, COLS), aslp_div_up(T, TT));
  fsmn_filter_kernel<<<grid, TT / TB * 32, smem, st>>>(y, ldy, x, ldx, T, D, coef, ldc, P, C, rev);
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
  const int warps = aslp_div_up(C, TB);
  ASLP_REQUIRE(warps <= 32);
  ASLP_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out_diff % 16 == 0) && ldi % 4 == 0 && ldo % 4 == 0);
  float* partial = (float*)aslp_scratch(st, (size_t)chunks * C * Dpad * sizeof(float));
  if (partial == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  const size_t smem = ((size_t)(TT + C - 1 + 2 * TB) + TT) * COLS * sizeof(float);
  if (smem > 220 * 1024) { aslp_set_last_error_msg("FSMN context too large for one shared-memory tile", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  static size_t smem_allowed = 0;
  if (smem > smem_allowed) {
    ASLP_CUDA(cudaFuncSetAttribute(fsmn_grad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_allowed = smem;
  }
  dim3 grid(aslp_div_up(D, COLS), chunks);
  fsmn_grad_partial_kernel<<<grid, warps * 32, smem, st>>>(partial, in, ldi, out_diff, ldo, T, D, past, C, Dpad);
  ASLP_CHECK_LAUNCH();
  fsmn_grad_final_kernel<<<aslp_div_up(C * D, 256), 256, 0, st>>>(coef_corr, ldc, partial, chunks, C, D, Dpad, clip);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
