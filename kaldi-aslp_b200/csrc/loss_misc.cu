// Xent (one fused pass), Splice gather, RowConvolution.  All HBM-bound gathers / row passes.
// Reference: Xent::Eval (src/aslp-nnet/nnet-loss.cc:63-156) = ~15 elementwise/reduction passes
// + 3 D2H copies; Splice (nnet-various.h:139-175, cu-math.cc:153-166); RowConvolution
// (nnet-row-convolution.cc:90-169) = a [D,D,F+1] GEMM + diagonal extract per frame.
#include "common.cuh"

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

}  // namespace

extern "C" {

int aslp_xent_sparse(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, int rows, int cols, const int* tgt_idx,
                     const float* tgt_w, const float* frame_w, double* stats_dev) {
  if (rows == 0) return 0;
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  xent_kernel<false><<<blocks, 256, 0, (cudaStream_t)s>>>(diff, ldd, y, ldy, nullptr, 0, rows, cols, tgt_idx, tgt_w, frame_w, stats_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_xent_dense(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt, int rows, int cols,
                    const float* frame_w, double* stats_dev) {
  if (rows == 0) return 0;
  int blocks = aslp_div_up(rows, 8);
  if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
  xent_kernel<true><<<blocks, 256, 0, (cudaStream_t)s>>>(diff, ldd, y, ldy, tgt, ldt, rows, cols, nullptr, nullptr, frame_w, stats_dev);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_splice_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int dim, const int* offsets_dev, int n_offsets) {
  if (rows == 0 || dim == 0 || n_offsets == 0) return 0;
  splice_fwd_kernel<<<grid_for((long long)rows * n_offsets * dim), 256, 0, (cudaStream_t)s>>>(out, ldo, in, ldi, rows, dim, offsets_dev, n_offsets);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_splice_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* out_diff, int ldo, int rows, int dim, const int* offsets_dev, int n_offsets) {
  if (rows == 0 || dim == 0) return 0;
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

}  // extern "C"
