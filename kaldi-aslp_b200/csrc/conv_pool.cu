// ConvolutionalComponent / MaxPoolingComponent device side (the CNN front end of the CTC recipes).
// Reference: src/aslp-nnet/nnet-convolutional-component.h:263-421 (column_map + CopyCols, one AddMatMat per patch position,
// AddCols over a rearranged reverse map) and nnet-max-pooling-component.h:100-156 (Set(-1e20) + pool_size Max calls per pool;
// backward: per (pool, member) an EqualElementMask, a MulElements and an AddMat on freshly allocated matrices, then a Scale).
// Here the convolution is im2col + ONE GEMM per pass: patches are laid out [frame * num_patches + p][filter_dim], so the
// [frames, num_patches * num_filters] output IS the row-major [frames * num_patches, num_filters] product (aslp_gemm, bias in
// the epilogue); the two kernels below are the gather into that layout and the inverse gather-sum for the input derivative
// (each input column sums its patch positions in ascending p -- the order the reference's AddCols passes add them in).
// Max pooling is one pass each way.  All four kernels are HBM-bound: a thread owns one column of the written matrix and
// walks rows, coalesced along the column index.
#include "common.cuh"

namespace {

// Launch shape shared by the four kernels: threadIdx / blockIdx.x walk the COLUMN index of the written matrix, blockIdx.y
// strides over rows.  A thread decomposes its column once (the integer divisions by patch / pool geometry) and then
// only adds row pitches -- the first version redid 64-bit divisions per element and reached 17-27 % of HBM bandwidth.
inline dim3 col_row_grid(int ncols, int rows) {
  const int gx = (ncols + 255) / 256;
  int gy = (aslp_num_sms() * 8 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  return dim3(gx, gy);
}

// patches[(b*np + p)*ldp + s*pd + d] = in[b*ldi + p*step + s*stride + d]
__global__ void __launch_bounds__(256) conv_gather_kernel(float* __restrict__ patches, int ldp, const float* __restrict__ in, int ldi, int rows,
                                                          int np, int ns, int pd, int step, int stride) {
  const int fd = ns * pd;
  const int jj = blockIdx.x * 256 + threadIdx.x;
  if (jj >= np * fd) return;
  const int p = jj / fd, j = jj - p * fd;
  const int s = j / pd, d = j - s * pd;
  const int src = p * step + s * stride + d;
  const size_t dst = (size_t)p * ldp + j, row_pitch = (size_t)np * ldp;
#pragma unroll 4
  for (int b = blockIdx.y; b < rows; b += gridDim.y) patches[b * row_pitch + dst] = __ldg(in + (size_t)b * ldi + src);
}

// in_diff[b, c] = sum over patch positions p (ascending) with 0 <= c % stride - p*step < pd of diffs[(b*np + p)*ldp + (c / stride)*pd + d]
__global__ void __launch_bounds__(256) conv_scatter_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ diffs, int ldp, int rows,
                                                           int in_dim, int np, int pd, int step, int stride) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= in_dim) return;
  const int s = c / stride, off = c - s * stride;
  // p*step <= off  and  off - p*step < pd   <=>   ceil((off - pd + 1) / step) (>= 0) <= p <= off / step
  int p_lo = off - pd + 1;
  p_lo = p_lo <= 0 ? 0 : (p_lo + step - 1) / step;
  int p_hi = off / step;
  if (p_hi > np - 1) p_hi = np - 1;
  const size_t row_pitch = (size_t)np * ldp;
  const int k0 = s * pd + off, dp = ldp - step;          // element (p): k0 + p*ldp - p*step
  for (int b = blockIdx.y; b < rows; b += gridDim.y) {
    const float* src = diffs + b * row_pitch + k0;
    float sum = 0.f;
    for (int p = p_lo; p <= p_hi; ++p) sum += __ldg(src + (size_t)p * dp);
    in_diff[(size_t)b * ldd + c] = sum;
  }
}

// out[b, q*ps + j] = max(-1e20, max_r in[b, (q*step + r)*ps + j])
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, int rows, int pools,
                                                          int size, int step, int ps) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= pools * ps) return;
  const int q = c / ps, j = c - q * ps;
  const int src = q * step * ps + j;
#pragma unroll 2
  for (int b = blockIdx.y; b < rows; b += gridDim.y) {
    const float* x = in + (size_t)b * ldi + src;
    float m = -1e20f;
    for (int r = 0; r < size; ++r) m = fmaxf(m, __ldg(x + (size_t)r * ps));
    out[(size_t)b * ldo + c] = m;
  }
}

// in_diff[b, p*ps + j] = (sum over pools q (ascending) containing p of [in == out_q] * out_diff_q) * (1 / #pools containing p)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ in, int ldi,
                                                          const float* __restrict__ out, int ldo, const float* __restrict__ out_diff, int ldod,
                                                          int rows, int patches, int pools, int size, int step, int ps) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= patches * ps) return;
  const int p = c / ps, j = c - p * ps;
  // pools with q*step <= p < q*step + size
  int q_lo = p - size + 1;
  q_lo = q_lo <= 0 ? 0 : (q_lo + step - 1) / step;
  int q_hi = p / step;
  if (q_hi > pools - 1) q_hi = pools - 1;
  const int n = q_hi - q_lo + 1;
  // the reference scales by BaseFloat(1.0 / patch_summands[p]); a patch outside every pool cannot occur (it asserts)
  const float scale = n > 0 ? (float)(1.0 / (double)n) : 0.f;
#pragma unroll 2
  for (int b = blockIdx.y; b < rows; b += gridDim.y) {
    const float x = __ldg(in + (size_t)b * ldi + c);
    const float* o = out + (size_t)b * ldo + j;
    const float* od = out_diff + (size_t)b * ldod + j;
    float sum = 0.f;
    for (int q = q_lo; q <= q_hi; ++q) {
      const float mask = (x == __ldg(o + (size_t)q * ps)) ? 1.0f : 0.0f;
      sum += __ldg(od + (size_t)q * ps) * mask;
    }
    in_diff[(size_t)b * ldd + c] = sum * scale;
  }
}

}  // namespace

extern "C" {

int aslp_conv_gather_patches(aslp_stream_t s, float* patches, int ldp, const float* in, int ldi, int rows, int num_patches, int num_splice,
                             int patch_dim, int patch_step, int patch_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_splice > 0 && patch_dim > 0 && patch_step > 0 && patch_stride >= patch_dim);
  ASLP_REQUIRE(ldp >= num_splice * patch_dim && (num_patches - 1) * patch_step + patch_dim <= patch_stride);
  if (rows == 0) return 0;
  ASLP_REQUIRE(patches != nullptr && in != nullptr);
  conv_gather_kernel<<<col_row_grid(num_patches * num_splice * patch_dim, rows), 256, 0, (cudaStream_t)s>>>(
      patches, ldp, in, ldi, rows, num_patches, num_splice, patch_dim, patch_step, patch_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_conv_scatter_patch_diffs(aslp_stream_t s, float* in_diff, int ldd, const float* patch_diffs, int ldp, int rows, int num_patches,
                                  int num_splice, int patch_dim, int patch_step, int patch_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_splice > 0 && patch_dim > 0 && patch_step > 0 && patch_stride >= patch_dim);
  ASLP_REQUIRE(ldp >= num_splice * patch_dim && ldd >= num_splice * patch_stride);
  if (rows == 0) return 0;
  ASLP_REQUIRE(in_diff != nullptr && patch_diffs != nullptr);
  conv_scatter_kernel<<<col_row_grid(num_splice * patch_stride, rows), 256, 0, (cudaStream_t)s>>>(
      in_diff, ldd, patch_diffs, ldp, rows, num_splice * patch_stride, num_patches, patch_dim, patch_step, patch_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_maxpool_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int num_pools, int pool_size, int pool_step,
                     int pool_stride) {
  ASLP_REQUIRE(rows >= 0 && num_pools > 0 && pool_size > 0 && pool_step > 0 && pool_stride > 0);
  if (rows == 0) return 0;
  ASLP_REQUIRE(out != nullptr && in != nullptr);
  maxpool_fwd_kernel<<<col_row_grid(num_pools * pool_stride, rows), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, num_pools, pool_size, pool_step,
                                                                                               pool_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_maxpool_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff,
                     int ldod, int rows, int num_patches, int num_pools, int pool_size, int pool_step, int pool_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_pools > 0 && pool_size > 0 && pool_step > 0 && pool_stride > 0);
  ASLP_REQUIRE((num_pools - 1) * pool_step + pool_size <= num_patches);
  if (rows == 0) return 0;
  ASLP_REQUIRE(in_diff != nullptr && in != nullptr && out != nullptr && out_diff != nullptr);
  maxpool_bwd_kernel<<<col_row_grid(num_patches * pool_stride, rows), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows,
                                                                                                 num_patches, num_pools, pool_size, pool_step, pool_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
