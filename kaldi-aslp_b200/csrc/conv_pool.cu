// ConvolutionalComponent / MaxPoolingComponent device side (the CNN front end of the CTC recipes).
// Reference: src/aslp-nnet/nnet-convolutional-component.h:263-421 (column_map + CopyCols, one AddMatMat per patch position,
// AddCols over a rearranged reverse map) and nnet-max-pooling-component.h:100-156 (Set(-1e20) + pool_size Max calls per pool;
// backward: per (pool, member) an EqualElementMask, a MulElements and an AddMat on freshly allocated matrices, then a Scale).
// Here the convolution is im2col + ONE GEMM per pass: patches are laid out [frame * num_patches + p][filter_dim], so the
// [frames, num_patches * num_filters] output IS the row-major [frames * num_patches, num_filters] product (aslp_gemm, bias in
// the epilogue); the two kernels below are the gather into that layout and the inverse gather-sum for the input derivative
// (each input column sums its patch positions in ascending p -- the order the reference's AddCols passes add them in).
// Max pooling is one pass each way.  All four kernels are HBM-bound: thread per element, coalesced along the column index.
#include "common.cuh"

namespace {

// patches[(b*np + p)*ldp + s*pd + d] = in[b*ldi + p*step + s*stride + d]
__global__ void conv_gather_kernel(float* __restrict__ patches, int ldp, const float* __restrict__ in, int ldi, long long rows, int np, int ns,
                                   int pd, int step, int stride) {
  const int fd = ns * pd;
  const long long total = rows * np * fd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % fd);
    const long long bp = i / fd;
    const int p = (int)(bp % np);
    const long long b = bp / np;
    const int s = j / pd, d = j - s * pd;
    patches[bp * ldp + j] = in[b * ldi + p * step + s * stride + d];
  }
}

// in_diff[b, c] = sum over patch positions p (ascending) with 0 <= c % stride - p*step < pd of diffs[(b*np + p)*ldp + (c / stride)*pd + d]
__global__ void conv_scatter_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ diffs, int ldp, long long rows, int in_dim,
                                    int np, int pd, int step, int stride) {
  const long long total = rows * in_dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % in_dim);
    const long long b = i / in_dim;
    const int s = c / stride, off = c - s * stride;
    // p*step <= off  and  off - p*step < pd   <=>   (off - pd + 1) / step (rounded up, >= 0) <= p <= off / step
    int p_lo = off - pd + 1;
    p_lo = p_lo <= 0 ? 0 : (p_lo + step - 1) / step;
    int p_hi = off / step;
    if (p_hi > np - 1) p_hi = np - 1;
    float sum = 0.f;
    for (int p = p_lo; p <= p_hi; ++p) sum += diffs[(b * np + p) * ldp + s * pd + (off - p * step)];
    in_diff[b * ldd + c] = sum;
  }
}

// out[b, q*ps + j] = max(-1e20, max_r in[b, (q*step + r)*ps + j])
__global__ void maxpool_fwd_kernel(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi, long long rows, int pools, int size,
                                   int step, int ps) {
  const int od = pools * ps;
  const long long total = rows * od;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % od);
    const long long b = i / od;
    const int q = c / ps, j = c - q * ps;
    float m = -1e20f;
    for (int r = 0; r < size; ++r) m = fmaxf(m, in[b * ldi + (q * step + r) * ps + j]);
    out[b * ldo + c] = m;
  }
}

// in_diff[b, p*ps + j] = (sum over pools q (ascending) containing p of [in == out_q] * out_diff_q) * (1 / #pools containing p)
__global__ void maxpool_bwd_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ in, int ldi, const float* __restrict__ out,
                                   int ldo, const float* __restrict__ out_diff, int ldod, long long rows, int patches, int pools, int size,
                                   int step, int ps) {
  const int id = patches * ps;
  const long long total = rows * id;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % id);
    const long long b = i / id;
    const int p = c / ps, j = c - p * ps;
    // pools with q*step <= p < q*step + size
    int q_lo = p - size + 1;
    q_lo = q_lo <= 0 ? 0 : (q_lo + step - 1) / step;
    int q_hi = p / step;
    if (q_hi > pools - 1) q_hi = pools - 1;
    const float x = in[b * ldi + c];
    float sum = 0.f;
    int n = 0;
    for (int q = q_lo; q <= q_hi; ++q) {
      const float mask = (x == out[b * ldo + q * ps + j]) ? 1.0f : 0.0f;
      sum += out_diff[b * ldod + q * ps + j] * mask;
      ++n;
    }
    // the reference scales by BaseFloat(1.0 / patch_summands[p]); a patch outside every pool cannot occur (it asserts)
    in_diff[b * ldd + c] = n > 0 ? sum * (float)(1.0 / (double)n) : 0.f;
  }
}

inline int elem_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)aslp_num_sms() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int aslp_conv_gather_patches(aslp_stream_t s, float* patches, int ldp, const float* in, int ldi, int rows, int num_patches, int num_splice,
                             int patch_dim, int patch_step, int patch_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_splice > 0 && patch_dim > 0 && patch_step > 0 && patch_stride >= patch_dim);
  ASLP_REQUIRE(ldp >= num_splice * patch_dim && (num_patches - 1) * patch_step + patch_dim <= patch_stride);
  if (rows == 0) return 0;
  ASLP_REQUIRE(patches != nullptr && in != nullptr);
  const long long total = (long long)rows * num_patches * num_splice * patch_dim;
  conv_gather_kernel<<<elem_grid(total), 256, 0, (cudaStream_t)s>>>(patches, ldp, in, ldi, rows, num_patches, num_splice, patch_dim, patch_step,
                                                                    patch_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_conv_scatter_patch_diffs(aslp_stream_t s, float* in_diff, int ldd, const float* patch_diffs, int ldp, int rows, int num_patches,
                                  int num_splice, int patch_dim, int patch_step, int patch_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_splice > 0 && patch_dim > 0 && patch_step > 0 && patch_stride >= patch_dim);
  ASLP_REQUIRE(ldp >= num_splice * patch_dim && ldd >= num_splice * patch_stride);
  if (rows == 0) return 0;
  ASLP_REQUIRE(in_diff != nullptr && patch_diffs != nullptr);
  const long long total = (long long)rows * num_splice * patch_stride;
  conv_scatter_kernel<<<elem_grid(total), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, patch_diffs, ldp, rows, num_splice * patch_stride, num_patches,
                                                                     patch_dim, patch_step, patch_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_maxpool_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int num_pools, int pool_size, int pool_step,
                     int pool_stride) {
  ASLP_REQUIRE(rows >= 0 && num_pools > 0 && pool_size > 0 && pool_step > 0 && pool_stride > 0);
  if (rows == 0) return 0;
  ASLP_REQUIRE(out != nullptr && in != nullptr);
  const long long total = (long long)rows * num_pools * pool_stride;
  maxpool_fwd_kernel<<<elem_grid(total), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, num_pools, pool_size, pool_step, pool_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_maxpool_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff,
                     int ldod, int rows, int num_patches, int num_pools, int pool_size, int pool_step, int pool_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_pools > 0 && pool_size > 0 && pool_step > 0 && pool_stride > 0);
  ASLP_REQUIRE((num_pools - 1) * pool_step + pool_size <= num_patches);
  if (rows == 0) return 0;
  ASLP_REQUIRE(in_diff != nullptr && in != nullptr && out != nullptr && out_diff != nullptr);
  const long long total = (long long)rows * num_patches * pool_stride;
  maxpool_bwd_kernel<<<elem_grid(total), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows, num_patches, num_pools,
                                                                    pool_size, pool_step, pool_stride);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
