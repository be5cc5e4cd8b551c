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
#include <stdint.h>

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

// ---- row-per-block forms of the two convolution kernels.  The column-per-thread forms above leave every block writing (reading)
// 1 KB pieces 12.8 KB apart with one 4-byte access in flight per thread and row: 21 % / 28 % of HBM bandwidth at the CTC
// recipes' shape.  Here a block owns a run of consecutive frames: a frame's patch rows are ONE contiguous chunk of
// num_patches * ldp floats, so the block streams a contiguous region; the small side of the copy (the 440-float input row /
// the chunk of patch derivatives) is staged in shared memory with cp.async one frame ahead, and the index arithmetic is done
// once per thread (gather: source column of each of its <= 16 chunk elements, in registers) or once per block (scatter:
// a table in shared memory).  Same values, same summation order.
constexpr int CONV_MAX_ELEMS = 16;                        // chunk elements per thread of the gather kernel: chunks up to 4096 floats

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(256) conv_gather_rows_kernel(float* __restrict__ patches, int ldp, const float* __restrict__ in, int ldi, int rows,
                                                               int in_dim, int np, int ns, int pd, int step, int stride, int rows_per_block) {
  extern __shared__ float conv_smem[];                    // [2][in_pad]: two input rows
  const int in_pad = (in_dim + 4) & ~3, chunk = np * ldp, fd = ns * pd;
  int src[CONV_MAX_ELEMS];
#pragma unroll
  for (int k = 0; k < CONV_MAX_ELEMS; ++k) {
    const int e = threadIdx.x + 256 * k;
    src[k] = -1;                                          // beyond the chunk, or a padding column (never written: it may not be ours)
    if (e < chunk) {
      const int p = e / ldp, j = e - p * ldp;
      if (j < fd) { const int sp = j / pd, d = j - sp * pd; src[k] = p * step + sp * stride + d; }
    }
  }
  const int b0 = blockIdx.x * rows_per_block;
  const int b1 = b0 + rows_per_block < rows ? b0 + rows_per_block : rows;
  if (b0 >= b1) return;
  for (int c = threadIdx.x; c < in_dim; c += 256) cp_async4(conv_smem + c, in + (size_t)b0 * ldi + c);
  for (int b = b0; b < b1; ++b) {
    const int buf = (b - b0) & 1;
    cp_async_commit_wait_all();
    __syncthreads();                                      // row b has landed; the other buffer is no longer being read
    if (b + 1 < b1) for (int c = threadIdx.x; c < in_dim; c += 256) cp_async4(conv_smem + (buf ^ 1) * in_pad + c, in + (size_t)(b + 1) * ldi + c);
    const float* sr = conv_smem + buf * in_pad;
    float* dst = patches + (size_t)b * chunk + threadIdx.x;
#pragma unroll
    for (int k = 0; k < CONV_MAX_ELEMS; ++k) if (src[k] >= 0) dst[256 * k] = sr[src[k]];
  }
}

// 16 bytes per store: a thread owns up to CONV_MAX_VECS float4 of the chunk (ldp % 4 == 0, so a float4 never crosses a patch row;
// the pitch is the filter dimension rounded up to 4, and the rounding columns of the last float4 are written as zeros).  The
// staged input row is kept in shared memory in a 4-way de-interleaved order -- column i at (i & 3) * Q + (i >> 2) -- so that
// the lanes of a warp, whose source columns are 4 apart inside a run of patch_dim, read consecutive words, not every 4th bank.
constexpr int CONV_MAX_VECS = 4;                          // chunks up to 4096 floats
__global__ void __launch_bounds__(256) conv_gather_rows4_kernel(float* __restrict__ patches, int ldp, const float* __restrict__ in, int ldi, int rows,
                                                                int in_dim, int np, int ns, int pd, int step, int stride, int rows_per_block) {
  extern __shared__ float conv_smem[];                    // [2][row_pad]: two de-interleaved input rows, each followed by a zero word
  const int Q = (in_dim + 3) >> 2, row_pad = 4 * Q + 4, chunk = np * ldp, nvec = chunk >> 2, fd = ns * pd;
  int pos[CONV_MAX_VECS][4];
#pragma unroll
  for (int k = 0; k < CONV_MAX_VECS; ++k) {
    const int v = threadIdx.x + 256 * k;
#pragma unroll
    for (int c = 0; c < 4; ++c) pos[k][c] = -1;
    if (v < nvec) {
      const int e0 = 4 * v, p = e0 / ldp, j0 = e0 - p * ldp;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + c;
        pos[k][c] = 4 * Q;                                // the zero word
        if (j < fd) { const int sp = j / pd, d = j - sp * pd, src = p * step + sp * stride + d; pos[k][c] = (src & 3) * Q + (src >> 2); }
      }
    }
  }
  if (threadIdx.x < 2) conv_smem[threadIdx.x * row_pad + 4 * Q] = 0.f;
  const int b0 = blockIdx.x * rows_per_block;
  const int b1 = b0 + rows_per_block < rows ? b0 + rows_per_block : rows;
  if (b0 >= b1) return;
  auto fetch = [&](int b, int buf) {
    for (int c = threadIdx.x; c < in_dim; c += 256) cp_async4(conv_smem + buf * row_pad + (c & 3) * Q + (c >> 2), in + (size_t)b * ldi + c);
  };
  fetch(b0, 0);
  for (int b = b0; b < b1; ++b) {
    const int buf = (b - b0) & 1;
    cp_async_commit_wait_all();
    __syncthreads();
    if (b + 1 < b1) fetch(b + 1, buf ^ 1);
    const float* sr = conv_smem + buf * row_pad;
    float4* dst = reinterpret_cast<float4*>(patches + (size_t)b * chunk) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < CONV_MAX_VECS; ++k)
      if (pos[k][0] >= 0) dst[256 * k] = make_float4(sr[pos[k][0]], sr[pos[k][1]], sr[pos[k][2]], sr[pos[k][3]]);
  }
}

__global__ void __launch_bounds__(256) conv_scatter_rows_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ diffs, int ldp, int rows,
                                                                int in_dim, int np, int pd, int step, int stride, int rows_per_block, int vec16) {
  extern __shared__ float conv_smem[];                    // [2][chunk] patch derivatives of two frames, then int [in_dim][3]: k0, p_lo, p_hi
  const int chunk = np * ldp;
  int* tab = reinterpret_cast<int*>(conv_smem + 2 * chunk);
  for (int c = threadIdx.x; c < in_dim; c += 256) {
    const int sp = c / stride, off = c - sp * stride;
    int p_lo = off - pd + 1;
    p_lo = p_lo <= 0 ? 0 : (p_lo + step - 1) / step;
    int p_hi = off / step;
    if (p_hi > np - 1) p_hi = np - 1;
    tab[3 * c] = sp * pd + off; tab[3 * c + 1] = p_lo; tab[3 * c + 2] = p_hi;
  }
  const int b0 = blockIdx.x * rows_per_block;
  const int b1 = b0 + rows_per_block < rows ? b0 + rows_per_block : rows;
  if (b0 >= b1) return;
  auto fetch = [&](int b, int buf) {
    const float* g = diffs + (size_t)b * chunk;
    float* sdst = conv_smem + buf * chunk;
    if (vec16) { for (int v = threadIdx.x * 4; v < chunk; v += 1024) cp_async16(sdst + v, g + v); }
    else { for (int v = threadIdx.x; v < chunk; v += 256) cp_async4(sdst + v, g + v); }
  };
  fetch(b0, 0);
  const int dp = ldp - step;                              // element (p) of a column: k0 + p * ldp - p * step
  for (int b = b0; b < b1; ++b) {
    const int buf = (b - b0) & 1;
    cp_async_commit_wait_all();
    __syncthreads();
    if (b + 1 < b1) fetch(b + 1, buf ^ 1);
    const float* sd = conv_smem + buf * chunk;
    for (int c = threadIdx.x; c < in_dim; c += 256) {
      const int k0 = tab[3 * c], p_lo = tab[3 * c + 1], p_hi = tab[3 * c + 2];
      float sum = 0.f;
      for (int p = p_lo; p <= p_hi; ++p) sum += sd[k0 + p * dp];
      in_diff[(size_t)b * ldd + c] = sum;
    }
  }
}

// rows per block of the two kernels above: a few blocks per SM for overlap, long enough runs to amortise the per-block tables
inline int conv_rows_per_block(int rows) {
  const int blocks = aslp_num_sms() * 6;
  int rpb = (rows + blocks - 1) / blocks;
  return rpb < 4 ? 4 : rpb;
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

// four columns per thread (pool_stride % 4 == 0, 16-byte aligned operands): the four share their patch, hence their pools
__global__ void __launch_bounds__(256) maxpool_bwd4_kernel(float* __restrict__ in_diff, int ldd, const float* __restrict__ in, int ldi,
                                                           const float* __restrict__ out, int ldo, const float* __restrict__ out_diff, int ldod,
                                                           int rows, int patches, int pools, int size, int step, int ps) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= patches * ps) return;
  const int p = c / ps, j = c - p * ps;
  int q_lo = p - size + 1;
  q_lo = q_lo <= 0 ? 0 : (q_lo + step - 1) / step;
  int q_hi = p / step;
  if (q_hi > pools - 1) q_hi = pools - 1;
  const int n = q_hi - q_lo + 1;
  const float scale = n > 0 ? (float)(1.0 / (double)n) : 0.f;
#pragma unroll 2
  for (int b = blockIdx.y; b < rows; b += gridDim.y) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(in + (size_t)b * ldi + c));
    const float* o = out + (size_t)b * ldo + j;
    const float* od = out_diff + (size_t)b * ldod + j;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q_lo; q <= q_hi; ++q) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(o + (size_t)q * ps));
      const float4 g = __ldg(reinterpret_cast<const float4*>(od + (size_t)q * ps));
      sum.x += g.x * (x.x == m.x ? 1.0f : 0.0f); sum.y += g.y * (x.y == m.y ? 1.0f : 0.0f);
      sum.z += g.z * (x.z == m.z ? 1.0f : 0.0f); sum.w += g.w * (x.w == m.w ? 1.0f : 0.0f);
    }
    *reinterpret_cast<float4*>(in_diff + (size_t)b * ldd + c) = make_float4(sum.x * scale, sum.y * scale, sum.z * scale, sum.w * scale);
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
  const int in_dim = (num_patches - 1) * patch_step + (num_splice - 1) * patch_stride + patch_dim;      // columns of `in` the patches touch
  const long long chunk = (long long)num_patches * ldp;
  const size_t smem = sizeof(float) * 2 * ((in_dim + 4) & ~3);
  const int fd = num_splice * patch_dim;
  if (ldp % 4 == 0 && ldp - fd < 4 && ((uintptr_t)patches & 15) == 0 && chunk <= 1024 * CONV_MAX_VECS && smem <= 40 * 1024 && ldi >= in_dim) {
    const int rpb = conv_rows_per_block(rows);
    const size_t smem4 = sizeof(float) * 2 * (4 * ((in_dim + 3) / 4) + 4);
    conv_gather_rows4_kernel<<<(rows + rpb - 1) / rpb, 256, smem4, (cudaStream_t)s>>>(patches, ldp, in, ldi, rows, in_dim, num_patches, num_splice,
                                                                                      patch_dim, patch_step, patch_stride, rpb);
  } else if (chunk <= 256 * CONV_MAX_ELEMS && smem <= 40 * 1024 && ldi >= in_dim) {
    const int rpb = conv_rows_per_block(rows);
    conv_gather_rows_kernel<<<(rows + rpb - 1) / rpb, 256, smem, (cudaStream_t)s>>>(patches, ldp, in, ldi, rows, in_dim, num_patches, num_splice,
                                                                                    patch_dim, patch_step, patch_stride, rpb);
  } else {
    conv_gather_kernel<<<col_row_grid(num_patches * num_splice * patch_dim, rows), 256, 0, (cudaStream_t)s>>>(
        patches, ldp, in, ldi, rows, num_patches, num_splice, patch_dim, patch_step, patch_stride);
  }
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_conv_scatter_patch_diffs(aslp_stream_t s, float* in_diff, int ldd, const float* patch_diffs, int ldp, int rows, int num_patches,
                                  int num_splice, int patch_dim, int patch_step, int patch_stride) {
  ASLP_REQUIRE(rows >= 0 && num_patches > 0 && num_splice > 0 && patch_dim > 0 && patch_step > 0 && patch_stride >= patch_dim);
  ASLP_REQUIRE(ldp >= num_splice * patch_dim && ldd >= num_splice * patch_stride);
  if (rows == 0) return 0;
  ASLP_REQUIRE(in_diff != nullptr && patch_diffs != nullptr);
  const int in_dim = num_splice * patch_stride;
  const long long chunk = (long long)num_patches * ldp;
  const size_t smem = sizeof(float) * 2 * (size_t)chunk + sizeof(int) * 3 * (size_t)in_dim;
  if (smem <= 40 * 1024) {
    const int rpb = conv_rows_per_block(rows);
    const int vec16 = (chunk % 4 == 0 && ((uintptr_t)patch_diffs & 15) == 0) ? 1 : 0;
    conv_scatter_rows_kernel<<<(rows + rpb - 1) / rpb, 256, smem, (cudaStream_t)s>>>(in_diff, ldd, patch_diffs, ldp, rows, in_dim, num_patches,
                                                                                     patch_dim, patch_step, patch_stride, rpb, vec16);
  } else {
    conv_scatter_kernel<<<col_row_grid(num_splice * patch_stride, rows), 256, 0, (cudaStream_t)s>>>(
        in_diff, ldd, patch_diffs, ldp, rows, num_splice * patch_stride, num_patches, patch_dim, patch_step, patch_stride);
  }
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
  const bool vec4 = pool_stride % 4 == 0 && ldd % 4 == 0 && ldi % 4 == 0 && ldo % 4 == 0 && ldod % 4 == 0 &&
                    (((uintptr_t)in_diff | (uintptr_t)in | (uintptr_t)out | (uintptr_t)out_diff) & 15) == 0;
  if (vec4) {
    maxpool_bwd4_kernel<<<col_row_grid(num_patches * pool_stride / 4, rows), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows,
                                                                                                        num_patches, num_pools, pool_size, pool_step, pool_stride);
  } else {
    maxpool_bwd_kernel<<<col_row_grid(num_patches * pool_stride, rows), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, in, ldi, out, ldo, out_diff, ldod, rows,
                                                                                                   num_patches, num_pools, pool_size, pool_step, pool_stride);
  }
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
