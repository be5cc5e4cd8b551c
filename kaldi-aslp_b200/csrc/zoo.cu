// Kernels of the smaller components of the aslp-nnet zoo: Dropout, BlockSoftmax, Pnorm / Maxout, LengthNorm,
// Copy, and the gate re-arrangement of the coupled input-forget LSTM.  All HBM-bound single passes: a thread block walks rows
// with a grid-stride loop, lanes cover contiguous columns (float4 where the row geometry allows).
//   Dropout            src/aslp-nnet/nnet-activation.h:203-273   (CuRand::BinarizeProbs, cu-rand.cc:144-178)
//   BlockSoftmax       src/aslp-nnet/nnet-activation.h:64-146
//   Pnorm / Maxout     src/aslp-nnet/nnet-activation.h:305-377   (MatrixBase::GroupPnorm / GroupMax, kaldi-matrix.cc:1088-1138, :2530-2558)
//   LengthNorm         src/aslp-nnet/nnet-various.h:327-365
//   Copy               src/aslp-nnet/nnet-various.h:186-316      (cu::Copy, cu-math.cc)
//   LstmCifg           src/aslp-nnet/nnet-lstm-couple-if-projected-streams.h:300-640
#include "common.cuh"

namespace {

inline int rows_grid(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)aslp_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- counter-based generator: Philox-4x32-10 (Salmon et al. 2011), one call = four 32-bit words for counter (c0..c3), key (k0, k1)
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

// mask = (u < retention) with u uniform in (0, 1); out = in * mask / retention.  Element (r, c) always draws word ((r * cols + c) & 3)
// of counter ((r * cols + c) >> 2, call), so the mask does not depend on the launch geometry.
__global__ void __launch_bounds__(256) dropout_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, float* __restrict__ mask,
                                                          int ldm, int rows, int cols, float retention, unsigned long long seed, unsigned long long call) {
  const long long quads = ((long long)rows * cols + 3) / 4;
  const float inv = 1.0f / retention;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const uint4 w = philox4x32(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)call, (uint32_t)(call >> 32)),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = q * 4 + j;
      if (e < (long long)rows * cols) {
        const int r = (int)(e / cols), c = (int)(e - (long long)r * cols);
        const float u = ((float)(ws[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float m = u < retention ? 1.0f : 0.0f;
        mask[(size_t)r * ldm + c] = m;
        out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c] * m * inv;
      }
    }
  }
}

// out = a * b * scale (Dropout with a host-drawn mask; its backward pass)
__global__ void __launch_bounds__(256) mul_elements_kernel(float* __restrict__ out, int ldo, const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                           int ldb, int rows, int cols, float scale) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    out[(size_t)r * ldo + c] = a[(size_t)r * lda + c] * b[(size_t)r * ldb + c] * scale;
  }
}

// BlockSoftmax backward of one block: dst = src * (1 - sum_cols src)   (rows whose targets lie in another block sum to 1 -> zeroed)
__global__ void __launch_bounds__(256) rows_one_minus_sum_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += src[(size_t)r * lds + c];
    s = warp_sum(s);
    const float m = 1.0f - s;
    for (int c = lane; c < cols; c += 32) dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c] * m;
  }
}

// ---- group p-norm / max: out [rows, groups], in [rows, groups * G]; thread per output element
__device__ __forceinline__ float group_norm(const float* x, int G, float p) {
  float sum = 0.f;
  if (p == 1.0f) { for (int k = 0; k < G; ++k) sum += fabsf(x[k]); return sum; }
  if (p == 2.0f) { for (int k = 0; k < G; ++k) sum += x[k] * x[k]; return sqrtf(sum); }
  if (p == 0.0f) { for (int k = 0; k < G; ++k) if (x[k] != 0.f) sum += 1.f; return sum; }
  for (int k = 0; k < G; ++k) sum += powf(fabsf(x[k]), p);
  float v = powf(sum, 1.0f / p);
  if (isinf(v)) {                                   // the reference's overflow path: normalise by the largest magnitude first
    float mx = 0.f;
    for (int k = 0; k < G; ++k) mx = fmaxf(mx, fabsf(x[k]));
    sum = 0.f;
    for (int k = 0; k < G; ++k) sum += powf(fabsf(x[k]) / mx, p);
    v = powf(sum, 1.0f / p) * mx;
  }
  return v;
}
__global__ void __launch_bounds__(256) group_pnorm_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, int rows, int groups, int G, float p) {
  const long long total = (long long)rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups), g = (int)(i - (long long)r * groups);
    out[(size_t)r * ldo + g] = group_norm(in + (size_t)r * ldi + (size_t)g * G, G, p);
  }
}
// in_diff(i, j) = d|x|_p / dx_j * out_diff(i, j / G)
__global__ void __launch_bounds__(256) group_pnorm_bwd_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ in, int ldi, const float* __restrict__ out,
                                                              int ldo, const float* __restrict__ out_diff, int ldod, int rows, int cols, int G, float p) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float x = in[(size_t)r * ldi + c], y = out[(size_t)r * ldo + c / G];
    float d;
    if (p == 1.0f) d = x == 0.f ? 0.f : (x > 0.f ? 1.f : -1.f);
    else if (y == 0.f) d = 0.f;
    else d = powf(fabsf(x), p - 1.0f) * powf(y, 1.0f - p) * (x >= 0.f ? 1.f : -1.f);
    in_diff[(size_t)r * ldd + c] = d * out_diff[(size_t)r * ldod + c / G];
  }
}
__global__ void __launch_bounds__(256) group_max_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, int rows, int groups, int G) {
  const long long total = (long long)rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups), g = (int)(i - (long long)r * groups);
    const float* x = in + (size_t)r * ldi + (size_t)g * G;
    float m = -1e20f;
    for (int k = 0; k < G; ++k) if (x[k] > m) m = x[k];
    out[(size_t)r * ldo + g] = m;
  }
}
__global__ void __launch_bounds__(256) group_max_bwd_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ in, int ldi, const float* __restrict__ out,
                                                            int ldo, const float* __restrict__ out_diff, int ldod, int rows, int cols, int G) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float d = in[(size_t)r * ldi + c] == out[(size_t)r * ldo + c / G] ? 1.f : 0.f;
    in_diff[(size_t)r * ldd + c] = d * out_diff[(size_t)r * ldod + c / G];
  }
}

// LengthNorm: scale[r] = 1 / sqrt(sum_c x^2); out = x * scale.  Warp per row; the row is read twice (second time from L1/L2).
__global__ void __launch_bounds__(256) length_norm_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, float* __restrict__ scales, int rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) { const float x = in[(size_t)r * ldi + c]; s += x * x; }
    s = warp_sum(s);
    const float sc = 1.0f / sqrtf(s);               // InvertElements of ApplyPow(0.5): IEEE division and square root as on the CPU
    if (lane == 0) scales[r] = sc;
    for (int c = lane; c < cols; c += 32) out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c] * sc;
  }
}
__global__ void __launch_bounds__(256) mul_rows_vec_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, const float* __restrict__ v, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c] * v[r];
  }
}
__global__ void __launch_bounds__(256) copy_cols_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, const int* __restrict__ idx, int rows, int cols_out) {
  const long long total = (long long)rows * cols_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_out), c = (int)(i - (long long)r * cols_out);
    out[(size_t)r * ldo + c] = in[(size_t)r * ldi + idx[c]];
  }
}

// coupled input-forget LSTM: the recurrence kernels run the four-gate cell with i = sigmoid(-(pre-activation of f)) = 1 - f.
// expand: dst rows [g | -f | f | o] (4C) from src rows [g | f | o] (3C); works for matrices (cols >= 1) and for vectors (cols == 1, ld 1)
__global__ void __launch_bounds__(256) cifg_expand_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int C, int cols) {
  const long long total = (long long)4 * C * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const int gate = r / C, cell = r - gate * C;
    const int srow = gate == 0 ? cell : (gate == 3 ? 2 * C + cell : C + cell);
    const float v = src[(size_t)srow * lds + c];
    dst[(size_t)r * ldd + c] = gate == 1 ? -v : v;
  }
}
// compact: dst columns [dg | df - di | do] (3C) from src columns [dg | di | df | do] (4C)
__global__ void __launch_bounds__(256) cifg_compact_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int rows, int C) {
  const long long total = (long long)rows * 3 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / (3 * C)), c = (int)(i - (long long)r * 3 * C);
    const float* s = src + (size_t)r * lds;
    float v;
    if (c < C) v = s[c];
    else if (c < 2 * C) v = s[C + c] - s[c];             // df (column 2C + cell) - di (column C + cell), cell = c - C
    else v = s[C + c];                                   // do: column 3C + cell, cell = c - 2C
    dst[(size_t)r * ldd + c] = v;
  }
}

}  // namespace

extern "C" {

int aslp_dropout_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, float* mask, int ldm, int rows, int cols, float retention,
                     unsigned long long seed, unsigned long long call) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(retention > 0.f && retention <= 1.f);
  dropout_fwd_kernel<<<rows_grid(((long long)rows * cols + 3) / 4, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, mask, ldm, rows, cols, retention, seed, call);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_mul_elements(aslp_stream_t s, float* out, int ldo, const float* a, int lda, const float* b, int ldb, int rows, int cols, float scale) {
  if (rows == 0 || cols == 0) return 0;
  mul_elements_kernel<<<rows_grid((long long)rows * cols, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, a, lda, b, ldb, rows, cols, scale);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_rows_one_minus_sum(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  rows_one_minus_sum_kernel<<<rows_grid(rows, 8), 256, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_group_pnorm_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int groups, int group_size, float p) {
  if (rows == 0 || groups == 0) return 0;
  ASLP_REQUIRE(group_size > 0 && p >= 0.f);
  group_pnorm_fwd_kernel<<<rows_grid((long long)rows * groups, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, groups, group_size, p);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_group_pnorm_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff, int ldod,
                         int rows, int groups, int group_size, float p) {
  if (rows == 0 || groups == 0) return 0;
  group_pnorm_bwd_kernel<<<rows_grid((long long)rows * groups * group_size, 256), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows,
                                                                                                           groups * group_size, group_size, p);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_group_max_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int groups, int group_size) {
  if (rows == 0 || groups == 0) return 0;
  group_max_fwd_kernel<<<rows_grid((long long)rows * groups, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, groups, group_size);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_group_max_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff, int ldod,
                       int rows, int groups, int group_size) {
  if (rows == 0 || groups == 0) return 0;
  group_max_bwd_kernel<<<rows_grid((long long)rows * groups * group_size, 256), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows,
                                                                                                         groups * group_size, group_size);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_length_norm_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, float* row_scales, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  length_norm_fwd_kernel<<<rows_grid(rows, 8), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, row_scales, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_mul_rows_vec(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, const float* v, int rows, int cols) {
  if (rows == 0 || cols == 0) return 0;
  mul_rows_vec_kernel<<<rows_grid((long long)rows * cols, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, v, rows, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_copy_cols(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, const int* idx_dev, int rows, int cols_out) {
  if (rows == 0 || cols_out == 0) return 0;
  copy_cols_kernel<<<rows_grid((long long)rows * cols_out, 256), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, idx_dev, rows, cols_out);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_cifg_expand(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int C, int cols) {
  if (C == 0 || cols == 0) return 0;
  cifg_expand_kernel<<<rows_grid((long long)4 * C * cols, 256), 256, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, C, cols);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_cifg_compact(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int C) {
  if (rows == 0 || C == 0) return 0;
  cifg_compact_kernel<<<rows_grid((long long)rows * 3 * C, 256), 256, 0, (cudaStream_t)s>>>(dst, ldd, src, lds, rows, C);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
