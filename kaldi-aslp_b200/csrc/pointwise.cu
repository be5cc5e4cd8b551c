// Pointwise, row-wise and column-wise kernels of the aslp-nnet path.  All are HBM-bound: 128-bit
// accesses, grid-stride loops over SM-count multiples, warp-shuffle reductions, no intermediates.
// Reference kernels replaced: src/aslp-cudamatrix/cu-kernels.cu _sigmoid 1802, _tanh 1827,
// _diff_sigmoid 1814, _diff_tanh 1846, _softmax_reduce 1858, _add_vec_to_rows 700, _add_mat 584,
// _apply_floor 1359, _apply_ceiling 1505, _add_diag_mat_mat 992, _regularize_l1 2113,
// _find_row_max_id 2141 and the sgemv-with-ones column sums (cu-vector.cc:1145-1166).
#include "common.cuh"
#include "scratch.cuh"
#include "rowreg.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p, int nv) {
  if (nv == 4) return *reinterpret_cast<const float4*>(p);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nv > 0) v.x = p[0];
  if (nv > 1) v.y = p[1];
  if (nv > 2) v.z = p[2];
  return v;
}
__device__ __forceinline__ void st4(float* p, float4 v, int nv) {
  if (nv == 4) { *reinterpret_cast<float4*>(p) = v; return; }
  if (nv > 0) p[0] = v.x;
  if (nv > 1) p[1] = v.y;
  if (nv > 2) p[2] = v.z;
}

#define EW_LOOP_BEGIN(rows, cols)                                                                   \
  const int n4__ = ((cols) + 3) >> 2;                                                               \
  const long long total__ = (long long)(rows) * n4__;                                               \
  for (long long i__ = blockIdx.x * (long long)blockDim.x + threadIdx.x; i__ < total__;             \
       i__ += (long long)gridDim.x * blockDim.x) {                                                  \
    const int r = (int)(i__ / n4__);                                                                \
    const int c = (int)(i__ - (long long)r * n4__) << 2;                                            \
    const int nv = min(4, (cols) - c);
#define EW_LOOP_END }

inline int ew_blocks(long long rows, long long cols) {
  const long long total = rows * ((cols + 3) / 4);
  long long b = (total + 255) / 256;
  const long long cap = (long long)aslp_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <int KIND>
__device__ __forceinline__ float act_f(float x) {
  if (KIND == ASLP_ACT_SIGMOID) return ref_sigmoid(x);
  if (KIND == ASLP_ACT_TANH) return ref_tanh(x);
  return fmaxf(x, 0.f);   // ReLU: CopyFromMat + ApplyFloor(0)
}
template <int KIND>
__device__ __forceinline__ float dact_f(float y_or_x, float e) {
  if (KIND == ASLP_ACT_SIGMOID) return y_or_x * (1.0f - y_or_x) * e;     // _diff_sigmoid
  if (KIND == ASLP_ACT_TANH) return (1.0f - y_or_x * y_or_x) * e;        // _diff_tanh
  return y_or_x > 0.f ? e : 0.f;                                         // Heaviside(x) * e
}

template <int KIND>
__global__ void act_fwd_kernel(float* out, int ldo, const float* in, int ldi, int rows, int cols) {
  EW_LOOP_BEGIN(rows, cols)
    float4 x = ld4(in + (size_t)r * ldi + c, nv);
    x.x = act_f<KIND>(x.x); x.y = act_f<KIND>(x.y); x.z = act_f<KIND>(x.z); x.w = act_f<KIND>(x.w);
    st4(out + (size_t)r * ldo + c, x, nv);
  EW_LOOP_END
}
template <int KIND>
__global__ void act_bwd_kernel(float* din, int ldd, const float* y, int ldy, const float* e, int lde, int rows, int cols) {
  EW_LOOP_BEGIN(rows, cols)
    const float4 yy = ld4(y + (size_t)r * ldy + c, nv);
    const float4 ee = ld4(e + (size_t)r * lde + c, nv);
    float4 o;
    o.x = dact_f<KIND>(yy.x, ee.x); o.y = dact_f<KIND>(yy.y, ee.y); o.z = dact_f<KIND>(yy.z, ee.z); o.w = dact_f<KIND>(yy.w, ee.w);
    st4(din + (size_t)r * ldd + c, o, nv);
  EW_LOOP_END
}

__global__ void axpby_kernel(float* dst, int ldd, const float* src, int lds, int rows, int cols, float alpha, float beta) {
  EW_LOOP_BEGIN(rows, cols)
    const float4 s = ld4(src + (size_t)r * lds + c, nv);
    float4 o = make_float4(alpha * s.x, alpha * s.y, alpha * s.z, alpha * s.w);
    if (beta != 0.f) {
      const float4 d = ld4(dst + (size_t)r * ldd + c, nv);
      o.x += beta * d.x; o.y += beta * d.y; o.z += beta * d.z; o.w += beta * d.w;
    }
    st4(dst + (size_t)r * ldd + c, o, nv);
  EW_LOOP_END
}

__global__ void add_vec_to_rows_kernel(float* dst, int ldd, int rows, int cols, const float* vec, float alpha, float beta) {
  EW_LOOP_BEGIN(rows, cols)
    const float4 v = ld4(vec + c, nv);
    float4 o = make_float4(alpha * v.x, alpha * v.y, alpha * v.z, alpha * v.w);
    if (beta != 0.f) {
      const float4 d = ld4(dst + (size_t)r * ldd + c, nv);
      o.x += beta * d.x; o.y += beta * d.y; o.z += beta * d.z; o.w += beta * d.w;
    }
    st4(dst + (size_t)r * ldd + c, o, nv);
  EW_LOOP_END
}

__global__ void clamp_kernel(float* dst, int ldd, int rows, int cols, float lo, float hi) {
  EW_LOOP_BEGIN(rows, cols)
    float4 d = ld4(dst + (size_t)r * ldd + c, nv);
    d.x = fminf(fmaxf(d.x, lo), hi); d.y = fminf(fmaxf(d.y, lo), hi);
    d.z = fminf(fmaxf(d.z, lo), hi); d.w = fminf(fmaxf(d.w, lo), hi);
    st4(dst + (size_t)r * ldd + c, d, nv);
  EW_LOOP_END
}

// cu::RegularizeL1 CPU branch (src/aslp-cudamatrix/cu-math.cc:56-75), elementwise
__global__ void l1_kernel(float* w, int ldw, float* g, int ldg, int rows, int cols, float l1, float lr) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    float* wp = w + (size_t)r * ldw + c;
    float* gp = g + (size_t)r * ldg + c;
    const float before = *wp;
    if (before == 0.0f) continue;
    const float l1_signed = before < 0.0f ? -l1 : l1;
    const float after = before - lr * (*gp) - l1_signed;
    if ((after > 0.0f) != (before > 0.0f)) { *wp = 0.0f; *gp = 0.0f; }
    else *wp = before - l1_signed;
  }
}

// one warp per row; scl = 1/max(1, ||w_row||/max_norm)
__global__ void max_norm_kernel(float* w, int ldw, int rows, int cols, float max_norm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = w + (size_t)row * ldw;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += p[c] * p[c];
  s = warp_sum(s);
  float scl = sqrtf(s) * (1.0f / max_norm);
  scl = 1.0f / fmaxf(scl, 1.0f);
  for (int c = lane; c < cols; c += 32) p[c] *= scl;
}

// softmax: one warp per row, the row (<= a few KB) is re-read from L1/L2, HBM sees 1 read + 1 write
__global__ void softmax_rows_kernel(float* out, int ldo, const float* in, int ldi, int rows, int cols) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const float* x = in + (size_t)row * ldi;
    float* y = out + (size_t)row * ldo;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) { const float e = expf(x[c] - mx); y[c] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < cols; c += 32) y[c] *= inv;
  }
}

__global__ void row_argmax_kernel(int* idx, const float* m, int ldm, int rows, int cols) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const float* x = m + (size_t)row * ldm;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) { const float v = x[c]; if (v > best) { best = v; bi = c; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) idx[row] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// ---- column reductions: phase 1 partial sums per row chunk, phase 2 finalize (deterministic) ----
// block = 256 threads = 32 column-quads (128 columns) x 8 row lanes
template <bool DOT>
__global__ void col_reduce_partial_kernel(float* partial, const float* a, int lda, const float* b, int ldb,
                                          int rows, int cols, int rows_per_chunk) {
  __shared__ float4 red[8][32];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cq) * 4;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    const int nv = min(4, cols - c);
    // four rows per iteration into four independent accumulators (fixed combination order: deterministic): the
    // single-accumulator loop kept one load in flight per thread and reached 56 % of HBM bandwidth
    float4 acc1 = acc, acc2 = acc, acc3 = acc;
    int r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      float4 x0 = ld4(a + (size_t)r * lda + c, nv), x1 = ld4(a + (size_t)(r + 8) * lda + c, nv);
      float4 x2 = ld4(a + (size_t)(r + 16) * lda + c, nv), x3 = ld4(a + (size_t)(r + 24) * lda + c, nv);
      if (DOT) {
        const float4 y0 = ld4(b + (size_t)r * ldb + c, nv), y1 = ld4(b + (size_t)(r + 8) * ldb + c, nv);
        const float4 y2 = ld4(b + (size_t)(r + 16) * ldb + c, nv), y3 = ld4(b + (size_t)(r + 24) * ldb + c, nv);
        x0.x *= y0.x; x0.y *= y0.y; x0.z *= y0.z; x0.w *= y0.w;  x1.x *= y1.x; x1.y *= y1.y; x1.z *= y1.z; x1.w *= y1.w;
        x2.x *= y2.x; x2.y *= y2.y; x2.z *= y2.z; x2.w *= y2.w;  x3.x *= y3.x; x3.y *= y3.y; x3.z *= y3.z; x3.w *= y3.w;
      }
      acc.x += x0.x; acc.y += x0.y; acc.z += x0.z; acc.w += x0.w;      acc1.x += x1.x; acc1.y += x1.y; acc1.z += x1.z; acc1.w += x1.w;
      acc2.x += x2.x; acc2.y += x2.y; acc2.z += x2.z; acc2.w += x2.w;  acc3.x += x3.x; acc3.y += x3.y; acc3.z += x3.z; acc3.w += x3.w;
    }
    for (; r < r1; r += 8) {
      float4 x = ld4(a + (size_t)r * lda + c, nv);
      if (DOT) {
        const float4 y = ld4(b + (size_t)r * ldb + c, nv);
        x.x *= y.x; x.y *= y.y; x.z *= y.z; x.w *= y.w;
      }
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    acc.x = (acc.x + acc1.x) + (acc2.x + acc3.x); acc.y = (acc.y + acc1.y) + (acc2.y + acc3.y);
    acc.z = (acc.z + acc1.z) + (acc2.z + acc3.z); acc.w = (acc.w + acc1.w) + (acc2.w + acc3.w);
  }
  red[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && c < cols) {
    for (int k = 1; k < 8; ++k) { const float4 o = red[k][cq]; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    float* dst = partial + (size_t)blockIdx.y * (((cols + 3) >> 2) << 2) + c;
    *reinterpret_cast<float4*>(dst) = acc;
  }
}
// 32 columns x 8 chunk-lanes per block: each thread sums every 8th partial row with 4 independent accumulators (the
// one-thread-per-column form was a serial chain of `chunks` dependent L2 loads), fixed order -> deterministic
__global__ void __launch_bounds__(256) col_reduce_final_kernel(float* vec, const float* partial, int chunks, int cols, float alpha, float beta, float clip) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, kl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int ldp = ((cols + 3) >> 2) << 2;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int k = kl;
    for (; k + 24 < chunks; k += 32) {
      s0 += partial[(size_t)k * ldp + c];
      s1 += partial[(size_t)(k + 8) * ldp + c];
      s2 += partial[(size_t)(k + 16) * ldp + c];
      s3 += partial[(size_t)(k + 24) * ldp + c];
    }
    for (; k < chunks; k += 8) s0 += partial[(size_t)k * ldp + c];
  }
  red[kl][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (kl == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][cl];
    float v = alpha * s;
    if (beta != 0.f) v += beta * vec[c];
    if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
    vec[c] = v;
  }
}

int col_reduce(cudaStream_t st, bool dot, float* vec, const float* a, int lda, const float* b, int ldb, int rows, int cols,
               float alpha, float beta, float clip) {
  if (cols == 0) return 0;
  const int col_blocks = aslp_div_up(cols, 128);
  int chunks = aslp_div_up(aslp_num_sms() * 4, col_blocks);
  if (chunks > aslp_div_up(rows, 64)) chunks = aslp_div_up(rows, 64);
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = aslp_div_up(rows > 0 ? rows : 1, chunks);
  chunks = aslp_div_up(rows > 0 ? rows : 1, rows_per_chunk);
  const size_t ldp = (size_t)((cols + 3) / 4 * 4);
  float* partial = (float*)aslp_scratch(st, (size_t)chunks * ldp * sizeof(float));
  if (partial == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  dim3 grid(col_blocks, chunks);
  if (dot) col_reduce_partial_kernel<true><<<grid, 256, 0, st>>>(partial, a, lda, b, ldb, rows, cols, rows_per_chunk);
  else col_reduce_partial_kernel<false><<<grid, 256, 0, st>>>(partial, a, lda, b, ldb, rows, cols, rows_per_chunk);
  ASLP_CHECK_LAUNCH();
  col_reduce_final_kernel<<<aslp_div_up(cols, 32), 256, 0, st>>>(vec, partial, chunks, cols, alpha, beta, clip);
  ASLP_CHECK_LAUNCH();
  return 0;
}

// corr = momentum * corr + column sums of diff; bias -= lr * corr.  One CTA per 32 columns: 32 row lanes x 32 columns (a
// 1000-frame utterance is 32 rows per thread; with 8 row lanes the 125 dependent loads per thread made the launch 12 us), two
// independent accumulators per thread, a fixed-order shared-memory reduction over the row lanes (deterministic).
__global__ void __launch_bounds__(1024) bias_grad_update_kernel(float* bias, float* corr, const float* __restrict__ diff, int ldd, int rows, int cols,
                                                                float momentum, float lr) {
  __shared__ float part[32][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a0 = 0.f, a1 = 0.f;
  if (c < cols) {
    int r = rl;
    for (; r + 32 < rows; r += 64) { a0 += diff[(size_t)r * ldd + c]; a1 += diff[(size_t)(r + 32) * ldd + c]; }
    for (; r < rows; r += 32) a0 += diff[(size_t)r * ldd + c];
  }
  part[rl][cl] = a0 + a1;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k += 4) { s0 += part[k][cl]; s1 += part[k + 1][cl]; s2 += part[k + 2][cl]; s3 += part[k + 3][cl]; }
    const float g = momentum * corr[c] + ((s0 + s1) + (s2 + s3));
    corr[c] = g;
    bias[c] = bias[c] + (-lr) * g;
  }
}

// dst[r, :] = src[idx[r], :]  (idx < 0 -> zeros), 128-bit lanes along the row
__global__ void copy_rows_kernel(float* dst, int ldd, const float* src, int lds, const int* idx, int rows, int cols) {
  const int c4 = (cols + 3) >> 2;
  const long long total = (long long)rows * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), q = (int)(i - (long long)r * c4);
    const int sr = idx[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sr >= 0) v = *reinterpret_cast<const float4*>(src + (size_t)sr * lds + q * 4);
    float* d = dst + (size_t)r * ldd + q * 4;
    if (q * 4 + 3 < cols) *reinterpret_cast<float4*>(d) = v;
    else { const float t[4] = {v.x, v.y, v.z, v.w}; for (int j = 0; q * 4 + j < cols; ++j) d[j] = t[j]; }
  }
}

// 32x32 tiled transpose through shared memory (coalesced both ways)
__global__ void transpose_kernel(float* dst, int ldd, const float* src, int lds, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * lds + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * ldd + r] = tile[threadIdx.x][j];
  }
}

__global__ void sum_check_kernel(const float* m, int ldm, int rows, int cols, double* out2) {
  double s = 0.0; double bad = 0.0;
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = m[(size_t)(i / cols) * ldm + (i % cols)];
    if (isfinite(v)) s += (double)v; else bad += 1.0;
  }
  s = warp_sum_d(s); bad = warp_sum_d(bad);
  if ((threadIdx.x & 31) == 0) { atomicAdd(out2, s); atomicAdd(out2 + 1, bad); }
}

}  // namespace

extern "C" {

int aslp_copy_rows(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, const int* idx_dev, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldd % 4 == 0 && lds % 4 == 0 && idx_dev != nullptr);
  copy_rows_kernel<<<ew_blocks(rows, cols), 256, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, idx_dev, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_act_fwd(aslp_stream_t s, int kind, float* out, int ldo, const float* in, int ldi, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldo % 4 == 0 && ldi % 4 == 0);
  cudaStream_t st = (cudaStream_t)s;
  const int b = ew_blocks(rows, cols);
  switch (kind) {
    case ASLP_ACT_SIGMOID: act_fwd_kernel<ASLP_ACT_SIGMOID><<<b, 256, 0, st>>>(out, ldo, in, ldi, rows, cols); break;
    case ASLP_ACT_TANH: act_fwd_kernel<ASLP_ACT_TANH><<<b, 256, 0, st>>>(out, ldo, in, ldi, rows, cols); break;
    case ASLP_ACT_RELU: act_fwd_kernel<ASLP_ACT_RELU><<<b, 256, 0, st>>>(out, ldo, in, ldi, rows, cols); break;
    default: ASLP_REQUIRE(!"unknown activation kind");
  }
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_act_bwd(aslp_stream_t s, int kind, float* in_diff, int ldd, const float* y, int ldy, const float* e, int lde, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldd % 4 == 0 && ldy % 4 == 0 && lde % 4 == 0);
  cudaStream_t st = (cudaStream_t)s;
  const int b = ew_blocks(rows, cols);
  switch (kind) {
    case ASLP_ACT_SIGMOID: act_bwd_kernel<ASLP_ACT_SIGMOID><<<b, 256, 0, st>>>(in_diff, ldd, y, ldy, e, lde, rows, cols); break;
    case ASLP_ACT_TANH: act_bwd_kernel<ASLP_ACT_TANH><<<b, 256, 0, st>>>(in_diff, ldd, y, ldy, e, lde, rows, cols); break;
    case ASLP_ACT_RELU: act_bwd_kernel<ASLP_ACT_RELU><<<b, 256, 0, st>>>(in_diff, ldd, y, ldy, e, lde, rows, cols); break;
    default: ASLP_REQUIRE(!"unknown activation kind");
  }
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_softmax_rows(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  if (cols <= rowreg::MAX_COLS && rowreg::aligned16(out, ldo) && rowreg::aligned16(in, ldi)) {
    // whole row in registers: one HBM read + one write, every load of a row in flight at once
#define ASLP_SOFTMAX_CALL(G, NV)                                                                                      \
    rowreg::softmax_reg_kernel<G, NV><<<rowreg::row_grid(rowreg::softmax_reg_kernel<G, NV>, rows, 8 * (32 / G)), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, cols)
    ROWREG_DISPATCH(cols, ASLP_SOFTMAX_CALL);
#undef ASLP_SOFTMAX_CALL
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  softmax_rows_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_axpby(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols, float alpha, float beta) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldd % 4 == 0 && lds % 4 == 0);
  axpby_kernel<<<ew_blocks(rows, cols), 256, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, rows, cols, alpha, beta);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_add_vec_to_rows(aslp_stream_t s, float* dst, int ldd, int rows, int cols, const float* vec, float alpha, float beta) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldd % 4 == 0);
  add_vec_to_rows_kernel<<<ew_blocks(rows, cols), 256, 0, (cudaStream_t)s>>>(dst, ldd, rows, cols, vec, alpha, beta);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_col_sum(aslp_stream_t s, float* vec, const float* mat, int ldm, int rows, int cols, float alpha, float beta, float clip) {
  ASLP_REQUIRE(ldm % 4 == 0);
  return col_reduce((cudaStream_t)s, false, vec, mat, ldm, nullptr, 0, rows, cols, alpha, beta, clip);
}
int aslp_bias_grad_update(aslp_stream_t s, float* bias, float* corr, const float* diff, int ldd, int rows, int cols, float momentum, float lr) {
  if (cols == 0) return 0;
  ASLP_REQUIRE(bias != nullptr && corr != nullptr && diff != nullptr && ldd >= cols);
  if (rows > 4096) {                                      // long chunks: the chunked two-pass column sum fills the chip better
    int rc = col_reduce((cudaStream_t)s, false, corr, diff, ldd, nullptr, 0, rows, cols, 1.0f, momentum, 0.0f);
    if (rc != 0) return rc;
    const int ld = (cols + 3) / 4 * 4;
    return aslp_axpby(s, bias, ld, corr, ld, 1, cols, -lr, 1.0f);
  }
  bias_grad_update_kernel<<<aslp_div_up(cols, 32), 1024, 0, (cudaStream_t)s>>>(bias, corr, diff, ldd, rows, cols, momentum, lr);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_col_dot(aslp_stream_t s, float* vec, const float* a, int lda, const float* b, int ldb, int rows, int cols, float alpha, float beta, float clip) {
  ASLP_REQUIRE(lda % 4 == 0 && ldb % 4 == 0);
  return col_reduce((cudaStream_t)s, true, vec, a, lda, b, ldb, rows, cols, alpha, beta, clip);
}

int aslp_clamp(aslp_stream_t s, float* dst, int ldd, int rows, int cols, float lo, float hi) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldd % 4 == 0);
  clamp_kernel<<<ew_blocks(rows, cols), 256, 0, (cudaStream_t)s>>>(dst, ldd, rows, cols, lo, hi);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_regularize_l1(aslp_stream_t s, float* w, int ldw, float* grad, int ldg, int rows, int cols, float l1, float lr) {
  if (rows == 0 || cols == 0) return 0;
  l1_kernel<<<ew_blocks(rows, (long long)cols * 4), 256, 0, (cudaStream_t)s>>>(w, ldw, grad, ldg, rows, cols, l1, lr);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_max_norm_rows(aslp_stream_t s, float* w, int ldw, int rows, int cols, float max_norm) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(max_norm > 0.f);
  max_norm_kernel<<<aslp_div_up(rows, 8), 256, 0, (cudaStream_t)s>>>(w, ldw, rows, cols, max_norm);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_transpose(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  dim3 grid(aslp_div_up(cols, 32), aslp_div_up(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_sum_check(aslp_stream_t s, const float* m, int ldm, int rows, int cols, double* out2_dev) {
  ASLP_CUDA(cudaMemsetAsync(out2_dev, 0, 2 * sizeof(double), (cudaStream_t)s));
  if (rows == 0 || cols == 0) return 0;
  sum_check_kernel<<<ew_blocks(rows, (long long)cols * 4), 256, 0, (cudaStream_t)s>>>(m, ldm, rows, cols, out2_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_row_argmax(aslp_stream_t s, int* idx, const float* m, int ldm, int rows, int cols) {
  if (rows == 0) return 0;
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  row_argmax_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(idx, m, ldm, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
