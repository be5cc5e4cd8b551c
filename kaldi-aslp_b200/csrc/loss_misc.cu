// Xent (one fused pass), Splice gather, RowConvolution.  All HBM-bound gathers / row passes.
// Reference: Xent::Eval (src/aslp-nnet/nnet-loss.cc:63-156) = ~15 elementwise/reduction passes
// + 3 D2H copies; Splice (nnet-various.h:139-175, cu-math.cc:153-166); RowConvolution
// (nnet-row-convolution.cc:90-169) = a [D,D,F+1] GEMM + diagonal extract per frame.
#include "common.cuh"
#include "rowreg.cuh"

namespace {

// stats: [0] cross-entropy  [1] entropy  [2] likelihood  [3] correct  [4] frames
__device__ __forceinline__ void flush_stats(double* stats, double ce, double en, double lk, double co, double fr) {
  ce = warp_sum_d(ce); en = warp_sum_d(en); lk = warp_sum_d(lk); co = warp_sum_d(co); fr = warp_sum_d(fr);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats + 0, ce); atomicAdd(stats + 1, en); atomicAdd(stats + 2, lk); atomicAdd(stats + 3, co); atomicAdd(stats + 4, fr);
  }
}

// one warp per row
template <bool DENSE>
__global__ void xent_kernel(float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt, int rows, int cols,
                            const int* tgt_idx, const float* tgt_w, const float* frame_w, double* stats) {
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  double ce = 0, en = 0, lk = 0, co = 0, fr = 0;     // lane 0 carries the per-row scalars
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const float* yr = y + (size_t)row * ldy;
    float* dr = diff + (size_t)row * ldd;
    const float fw = frame_w[row];
    if (!DENSE) {
      const int ti = tgt_idx[row];
      const float tw = tgt_w[row];
      const float w = fw * tw;                        // frame_weights * sum_k t
      float best = -INFINITY; int bi = 0x7fffffff;
      for (int c = lane; c < cols; c += 32) {
        const float v = yr[c];
        if (v > best) { best = v; bi = c; }
        const float t = (c == ti) ? tw : 0.f;
        dr[c] = (v - t) * w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) {
        const int targ = tw > 0.f ? ti : (tw == 0.f ? 0 : (ti == 0 ? 1 : 0));   // FindRowMaxId of the target row
        const float yt = yr[ti];
        ce -= (double)(logf(yt + 1e-20f) * tw * w);
        en -= (double)(logf(tw + 1e-20f) * tw * w);
        lk += (double)(yt * tw * w);
        co += (bi == targ) ? (double)w : 0.0;
        fr += (double)w;
      }
    } else {
      const float* tr = tgt + (size_t)row * ldt;
      float tsum = 0.f;
      for (int c = lane; c < cols; c += 32) tsum += tr[c];
      tsum = warp_sum(tsum);
      const float w = fw * tsum;
      float best = -INFINITY, tbest = -INFINITY; int bi = 0x7fffffff, tbi = 0x7fffffff;
      float pce = 0.f, pen = 0.f, plk = 0.f;
      for (int c = lane; c < cols; c += 32) {
        const float v = yr[c], t = tr[c];
        if (v > best) { best = v; bi = c; }
        if (t > tbest) { tbest = t; tbi = c; }
        dr[c] = (v - t) * w;
        pce += logf(v + 1e-20f) * t * w;
        pen += logf(t + 1e-20f) * t * w;
        plk += v * t * w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        ov = __shfl_xor_sync(0xffffffffu, tbest, o); oi = __shfl_xor_sync(0xffffffffu, tbi, o);
        if (ov > tbest || (ov == tbest && oi < tbi)) { tbest = ov; tbi = oi; }
      }
      pce = warp_sum(pce); pen = warp_sum(pen); plk = warp_sum(plk);
      if (lane == 0) {
        ce -= (double)pce; en -= (double)pen; lk += (double)plk;
        co += (bi == tbi) ? (double)w : 0.0;
        fr += (double)w;
      }
    }
  }
  flush_stats(stats, ce, en, lk, co, fr);
}

__global__ void splice_fwd_kernel(float* out, int ldo, const float* in, int ldi, int rows, int dim, const int* offs, int noff) {
  const long long total = (long long)rows * noff * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % dim);
    const long long rem = i / dim;
    const int c = (int)(rem % noff), r = (int)(rem / noff);
    int rs = r + offs[c];
    rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
    out[(size_t)r * ldo + (size_t)c * dim + j] = in[(size_t)rs * ldi + j];
  }
}
// 128-bit form (dim % 4 == 0, aligned rows).  A thread owns ONE float4 slot of the output row (its offset index and
// column quad never change) and walks a contiguous range of rows: no index division in the loop, loads of four rows in
// flight per thread, an input row is re-read by its noff consumers out of L1/L2 while it is hot.
__global__ void __launch_bounds__(128) splice_fwd4_kernel(float* out, int ldo, const float* in, int ldi, int rows, int dim4,
                                                          const int* offs, int noff, int rows_per_block) {
  const int items = noff * dim4;
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  for (int item = threadIdx.x; item < items; item += blockDim.x) {
    const int c = item / dim4, q = item - c * dim4;
    const int off = offs[c];
    const float* src = in + q * 4;
    float* dst = out + (size_t)c * dim4 * 4 + q * 4;
    int r = r_begin;
    for (; r + 4 <= r_end; r += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int rs = r + u + off;
        rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
        v[u] = *reinterpret_cast<const float4*>(src + (size_t)rs * ldi);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) st_stream4(dst + (size_t)(r + u) * ldo, v[u]);
    }
    for (; r < r_end; ++r) {
      int rs = r + off;
      rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
      st_stream4(dst + (size_t)r * ldo, *reinterpret_cast<const float4*>(src + (size_t)rs * ldi));
    }
  }
}
// backward, 128-bit: thread per (t, column quad), the noff gathered float4 of a group of 4 offsets loaded before they are summed
// (same ascending-c summation order as the scalar kernel)
__global__ void splice_bwd4_kernel(float* din, int ldd, const float* dout, int ldo, int rows, int dim4, const int* offs, int noff) {
  const long long total = (long long)rows * dim4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / dim4), q = (int)(i - (long long)t * dim4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;
    for (; c + 4 <= noff; c += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int rs = t + offs[c + u];
        rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
        v[u] = ld_stream4(dout + (size_t)rs * ldo + (size_t)(c + u) * dim4 * 4 + q * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; c < noff; ++c) {
      int rs = t + offs[c];
      rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
      const float4 v = ld_stream4(dout + (size_t)rs * ldo + (size_t)c * dim4 * 4 + q * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(din + (size_t)t * ldd + q * 4) = acc;
  }
}
// the reference's backward gathers with the SAME clamp(t + off) index (nnet-various.h:151-173)
__global__ void splice_bwd_kernel(float* din, int ldd, const float* dout, int ldo, int rows, int dim, const int* offs, int noff) {
  const long long total = (long long)rows * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % dim), t = (int)(i / dim);
    float acc = 0.f;
    for (int c = 0; c < noff; ++c) {
      int rs = t + offs[c];
      rs = rs < 0 ? 0 : (rs >= rows ? rows - 1 : rs);
      acc += dout[(size_t)rs * ldo + (size_t)c * dim + j];
    }
    din[(size_t)t * ldd + j] = acc;
  }
}

__global__ void rowconv_fwd_kernel(float* out, int ldo, const float* in, int ldi, int T, int S, int dim, const float* w, int ldw,
                                   int future, const int* seq_len) {
  const long long total = (long long)T * S * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dim);
    const long long row = i / dim;
    const int s = (int)(row % S), t = (int)(row / S);
    const int len = seq_len[s];
    float acc = 0.f;
    if (t < len) {
      for (int k = 0; k <= future; ++k) {
        const int tt = min(t + k, len - 1);
        acc = fmaf(w[(size_t)d * ldw + k], in[((size_t)tt * S + s) * ldi + d], acc);
      }
    }
    out[(size_t)row * ldo + d] = acc;
  }
}
__global__ void rowconv_bwd_data_kernel(float* din, int ldd, const float* dout, int ldo, int T, int S, int dim, const float* w,
                                        int ldw, int future, const int* seq_len) {
  const long long total = (long long)T * S * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dim);
    const long long row = i / dim;
    const int s = (int)(row % S), t = (int)(row / S);
    const int len = seq_len[s];
    float acc = 0.f;
    if (t < len) {
      for (int k = 0; k <= future && k <= t; ++k)
        acc = fmaf(w[(size_t)d * ldw + k], dout[((size_t)(t - k) * S + s) * ldo + d], acc);
    }
    din[(size_t)row * ldd + d] = acc;
  }
}
// w_diff[d,k] = sum_{s, t<len_s} in[min(t+k,len-1), s, d] * dout[t, s, d] ; one block per (k, 256 columns)
__global__ void rowconv_bwd_w_kernel(float* wd, int ldwd, const float* in, int ldi, const float* dout, int ldo, int T, int S,
                                     int dim, int future, const int* seq_len) {
  const int k = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) {
    const int len = seq_len[s];
    for (int t = 0; t < len && t < T; ++t) {
      const int tt = min(t + k, len - 1);
      acc = fmaf(in[((size_t)tt * S + s) * ldi + d], dout[((size_t)t * S + s) * ldo + d], acc);
    }
  }
  wd[(size_t)d * ldwd + k] = acc;
}

inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)aslp_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}


// ---- forwarder post-processing (aslp-nnet-forward.cc:184-207): ONE pass instead of Min, Max, Add, ApplyLog, ColRange.Add,
// Min, Max, AddVecToRows, Sum.  stats (sortable-uint encoded floats, decoded by posterior_stats_decode_kernel):
// [0] min / [1] max of the input, [2] min / [3] max after the log and blank stages (what the prior warning looks at),
// [4] number of non-finite outputs.
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void posterior_finalize_kernel(float* m, int ldm, int rows, int cols, int apply_log, float log_add, float blank_shift,
                                          const float* log_priors, float prior_scale, unsigned* stats) {
  float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
  unsigned bad = 0;
  const int n4 = (cols + 3) >> 2;
  const long long total = (long long)rows * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i - (long long)r * n4) << 2;
    float* p = m + (size_t)r * ldm + c;
    const int nv = min(4, cols - c);
    float v[4];
    if (nv == 4) { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else { for (int j = 0; j < 4; ++j) v[j] = j < nv ? p[j] : 0.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nv) {
        float x = v[j];
        mn0 = fminf(mn0, x); mx0 = fmaxf(mx0, x);
        if (apply_log) x = logf(x + log_add);
        if (c + j == 0 && blank_shift > 0.f) x -= blank_shift;
        mn1 = fminf(mn1, x); mx1 = fmaxf(mx1, x);
        if (log_priors != nullptr) x += -prior_scale * log_priors[c + j];
        if (!isfinite(x)) ++bad;
        v[j] = x;
      }
    }
    if (nv == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else { for (int j = 0; j < nv; ++j) p[j] = v[j]; }
  }
  mn0 = -warp_max(-mn0); mx0 = warp_max(mx0); mn1 = -warp_max(-mn1); mx1 = warp_max(mx1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stats + 0, f2ord(mn0)); atomicMax(stats + 1, f2ord(mx0));
    atomicMin(stats + 2, f2ord(mn1)); atomicMax(stats + 3, f2ord(mx1));
    if (bad) atomicAdd(stats + 4, bad);
  }
}
__global__ void posterior_stats_init_kernel(unsigned* stats) {
  stats[0] = 0xffffffffu; stats[1] = 0u; stats[2] = 0xffffffffu; stats[3] = 0u; stats[4] = 0u;
}
__global__ void posterior_stats_decode_kernel(unsigned* stats) {
  float* f = reinterpret_cast<float*>(stats);
  for (int i = 0; i < 4; ++i) f[i] = ord2f(stats[i]);
  f[4] = (float)stats[4];
}

}  // namespace


// Mse::Eval (src/aslp-nnet/nnet-loss.cc:205-258): diff = w (y - t); loss += 0.5 * sum_rows w * diff^2 -- the reference squares the
// ALREADY weighted difference and weights it again (w^3 (y - t)^2), kept.  One pass, warp per row, float4 when aligned; the loss
// is accumulated in double on the device (stats_dev[0]), one atomic per block.
__global__ void __launch_bounds__(256) mse_kernel(float* __restrict__ diff, int ldd, const float* __restrict__ y, int ldy, const float* __restrict__ t,
                                                  int ldt, int rows, int cols, const float* __restrict__ w, double* stats, int vec4) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float fw = w[r];
    float part = 0.f;
    if (vec4) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(y + (size_t)r * ldy + c), b = *reinterpret_cast<const float4*>(t + (size_t)r * ldt + c);
        const float4 d = make_float4((a.x - b.x) * fw, (a.y - b.y) * fw, (a.z - b.z) * fw, (a.w - b.w) * fw);
        *reinterpret_cast<float4*>(diff + (size_t)r * ldd + c) = d;
        part += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      }
    } else {
      for (int c = lane; c < cols; c += 32) {
        const float d = (y[(size_t)r * ldy + c] - t[(size_t)r * ldt + c]) * fw;
        diff[(size_t)r * ldd + c] = d;
        part += d * d;
      }
    }
    acc += (double)(part * fw);
  }
  acc = warp_sum_d(acc);
  __shared__ double sh[8];
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += sh[i];
    atomicAdd(stats, 0.5 * tot);
  }
}

extern "C" {

int aslp_xent_sparse(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, int rows, int cols, const int* tgt_idx,
                     const float* tgt_w, const float* frame_w, double* stats_dev) {
  if (rows == 0) return 0;
  if (cols <= rowreg::MAX_COLS && rowreg::aligned16(diff, ldd) && rowreg::aligned16(y, ldy)) {
#define ASLP_XENT_CALL(G, NV)                                                                                          \
    rowreg::xent_reg_kernel<G, NV, false><<<rowreg::row_grid(rowreg::xent_reg_kernel<G, NV, false>, rows, 8 * (32 / G)), 256, 0, (cudaStream_t)s>>>(       \
        diff, ldd, y, ldy, nullptr, 0, rows, cols, tgt_idx, tgt_w, frame_w, stats_dev)
    ROWREG_DISPATCH(cols, ASLP_XENT_CALL);
#undef ASLP_XENT_CALL
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  xent_kernel<false><<<blocks, 256, 0, (cudaStream_t)s>>>(diff, ldd, y, ldy, nullptr, 0, rows, cols, tgt_idx, tgt_w, frame_w, stats_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_xent_dense(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt, int rows, int cols,
                    const float* frame_w, double* stats_dev) {
  if (rows == 0) return 0;
  if (cols <= 1024 && rowreg::aligned16(diff, ldd) && rowreg::aligned16(y, ldy) && rowreg::aligned16(tgt, ldt)) {
#define ASLP_XENT_CALL(G, NV)                                                                                          \
    rowreg::xent_reg_kernel<G, NV, true><<<rowreg::row_grid(rowreg::xent_reg_kernel<G, NV, true>, rows, 8 * (32 / G)), 256, 0, (cudaStream_t)s>>>(        \
        diff, ldd, y, ldy, tgt, ldt, rows, cols, nullptr, nullptr, frame_w, stats_dev)
    if (cols <= 32) { ASLP_XENT_CALL(8, 1); } else if (cols <= 64) { ASLP_XENT_CALL(8, 2); } else if (cols <= 128) { ASLP_XENT_CALL(8, 4); }
    else if (cols <= 256) { ASLP_XENT_CALL(32, 2); } else if (cols <= 512) { ASLP_XENT_CALL(32, 4); } else { ASLP_XENT_CALL(32, 8); }
#undef ASLP_XENT_CALL
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  xent_kernel<true><<<blocks, 256, 0, (cudaStream_t)s>>>(diff, ldd, y, ldy, tgt, ldt, rows, cols, nullptr, nullptr, frame_w, stats_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_mse(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt, int rows, int cols,
             const float* frame_w, double* stats_dev) {
  if (rows == 0 || cols == 0) return 0;
  const int vec4 = (cols % 4 == 0 && rowreg::aligned16(diff, ldd) && rowreg::aligned16(y, ldy) && rowreg::aligned16(tgt, ldt)) ? 1 : 0;
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  mse_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(diff, ldd, y, ldy, tgt, ldt, rows, cols, frame_w, stats_dev, vec4);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_splice_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int dim, const int* offsets_dev, int n_offsets) {
  if (rows == 0 || dim == 0 || n_offsets == 0) return 0;
  if (dim % 4 == 0 && rowreg::aligned16(out, ldo) && rowreg::aligned16(in, ldi)) {
    int blocks = aslp_num_sms() * 8;
    int rpb = aslp_div_up(rows, blocks);
    if (rpb < 8) rpb = 8;
    blocks = aslp_div_up(rows, rpb);
    splice_fwd4_kernel<<<blocks, 128, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, dim / 4, offsets_dev, n_offsets, rpb);
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  splice_fwd_kernel<<<grid_for((long long)rows * n_offsets * dim), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, dim, offsets_dev, n_offsets);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_splice_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* out_diff, int ldo, int rows, int dim, const int* offsets_dev, int n_offsets) {
  if (rows == 0 || dim == 0) return 0;
  if (dim % 4 == 0 && rowreg::aligned16(in_diff, ldd) && rowreg::aligned16(out_diff, ldo)) {
    splice_bwd4_kernel<<<grid_for((long long)rows * (dim / 4)), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, out_diff, ldo, rows, dim / 4, offsets_dev, n_offsets);
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  splice_bwd_kernel<<<grid_for((long long)rows * dim), 256, 0, (cudaStream_t)s>>>(in_diff, ldd, out_diff, ldo, rows, dim, offsets_dev, n_offsets);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_rowconv_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int T, int S, int dim, const float* w, int ldw,
                     int future, const int* seq_len_dev) {
  if (T == 0 || S == 0 || dim == 0) return 0;
  rowconv_fwd_kernel<<<grid_for((long long)T * S * dim), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, T, S, dim, w, ldw, future, seq_len_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_rowconv_bwd(aslp_stream_t s, float* in_diff, int ldd, float* w_diff, int ldwd, const float* in, int ldi, const float* out_diff,
                     int ldo, int T, int S, int dim, const float* w, int ldw, int future, const int* seq_len_dev) {
  if (T == 0 || S == 0 || dim == 0) return 0;
  cudaStream_t st = (cudaStream_t)s;
  if (in_diff != nullptr) {
    rowconv_bwd_data_kernel<<<grid_for((long long)T * S * dim), 256, 0, st>>>(in_diff, ldd, out_diff, ldo, T, S, dim, w, ldw, future, seq_len_dev);
    ASLP_CHECK_LAUNCH();
  }
  dim3 grid(aslp_div_up(dim, 128), future + 1);
  rowconv_bwd_w_kernel<<<grid, 128, 0, st>>>(w_diff, ldwd, in, ldi, out_diff, ldo, T, S, dim, future, seq_len_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_posterior_finalize(aslp_stream_t s, float* m, int ldm, int rows, int cols, int apply_log, float log_add, float blank_shift,
                            const float* log_priors_dev, float prior_scale, float* stats5_dev) {
  ASLP_REQUIRE(stats5_dev != nullptr && ldm % 4 == 0);
  cudaStream_t st = (cudaStream_t)s;
  unsigned* su = reinterpret_cast<unsigned*>(stats5_dev);
  posterior_stats_init_kernel<<<1, 1, 0, st>>>(su);
  ASLP_CHECK_LAUNCH();
  if (rows > 0 && cols > 0) {
    const long long total = (long long)rows * ((cols + 3) / 4);
    long long b = (total + 255) / 256;
    const long long cap = (long long)aslp_num_sms() * 16;
    if (b > cap) b = cap;
    posterior_finalize_kernel<<<(int)b, 256, 0, st>>>(m, ldm, rows, cols, apply_log, log_add, blank_shift, log_priors_dev, prior_scale, su);
    ASLP_CHECK_LAUNCH();
  }
  posterior_stats_decode_kernel<<<1, 1, 0, st>>>(su);
  ASLP_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
