// rowreg.cuh -- row-wise kernels with the whole row resident in registers (softmax, Xent).
// A row is owned by a group of G lanes (8 or 32); lane lg of the group holds NV float4: elements (i*G + lg)*4 .. +3,
// i < NV, so every load of a row is an independent 128-bit access issued before the first use (the one-load-per-
// iteration loops these replace kept ~1 KB in flight per SM and sat at 7-45 % of HBM bandwidth).  HBM sees exactly one
// read and one write of the matrix.  Rows must be 16-byte aligned; cols <= G*NV*4.
#pragma once
#include "common.cuh"

namespace rowreg {

template <int G> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// first maximal index (ties -> smaller index), as FindRowMaxId
template <int G> __device__ __forceinline__ void group_argmax(float& best, int& bi) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
}

template <int G, int NV>
__device__ __forceinline__ void load_row(float (&v)[NV][4], const float* row, int cols, int lg, bool ok, float fill) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * G + lg) * 4;
    if (ok && c + 3 < cols) {
      const float4 q = ld_stream4(row + c);
      v[i][0] = q.x; v[i][1] = q.y; v[i][2] = q.z; v[i][3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = (ok && c + j < cols) ? row[c + j] : fill;
    }
  }
}
template <int G, int NV>
__device__ __forceinline__ void store_row(const float (&v)[NV][4], float* row, int cols, int lg, bool ok) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * G + lg) * 4;
    if (ok && c + 3 < cols) {
      st_stream4(row + c, make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ok && c + j < cols) row[c + j] = v[i][j];
    }
  }
}

// ---- softmax (Softmax::PropagateFnc; kaldi-vector.cc:852-859: max, exp(x - max), scale by 1 / sum)
template <int G, int NV>
__global__ void __launch_bounds__(256, NV >= 8 ? 2 : 3) softmax_reg_kernel(float* out, int ldo, const float* in, int ldi, int rows, int cols) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, lg = lane % G, sub = lane / G;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  for (long long r0 = (long long)warp * RPW; r0 < rows; r0 += (long long)nwarps * RPW) {
    const long long row = r0 + sub;
    const bool ok = row < rows;
    float v[NV][4];
    load_row<G, NV>(v, in + (size_t)row * ldi, cols, lg, ok, -INFINITY);
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mx = fmaxf(mx, v[i][j]);
    mx = group_max<G>(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float e = aslp_exp(v[i][j] - mx); v[i][j] = e; sum += e; }   // fill = -inf -> e = 0
    sum = group_sum<G>(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] *= inv;
    store_row<G, NV>(v, out + (size_t)row * ldo, cols, lg, ok);
  }
}

// ---- Xent::Eval (nnet-loss.cc:63-156) in one pass; stats: [0] cross-entropy [1] entropy [2] likelihood [3] correct [4] frames
__device__ __forceinline__ void block_flush_stats(double* stats, double (&acc)[5]) {
  __shared__ double sred[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 5; ++q) acc[q] = warp_sum_d(acc[q]);
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < 5; ++q) sred[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w][threadIdx.x];
    if (s != 0.0) atomicAdd(stats + threadIdx.x, s);
  }
}

template <int G, int NV, bool DENSE>
__global__ void __launch_bounds__(256, NV >= 8 ? 2 : 3) xent_reg_kernel(float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt,
                                                       int rows, int cols, const int* tgt_idx, const float* tgt_w,
                                                       const float* frame_w, double* stats) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, lg = lane % G, sub = lane / G;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};            // group leaders carry the per-row scalars
  for (long long r0 = (long long)warp * RPW; r0 < rows; r0 += (long long)nwarps * RPW) {
    const long long row = r0 + sub;
    const bool ok = row < rows;
    float v[NV][4];
    load_row<G, NV>(v, y + (size_t)row * ldy, cols, lg, ok, -INFINITY);
    const float fw = ok ? frame_w[row] : 0.f;
    float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { if (v[i][j] > best) { best = v[i][j]; bi = (i * G + lg) * 4 + j; } }
    group_argmax<G>(best, bi);
    if (!DENSE) {
      const int ti = ok ? tgt_idx[row] : 0;
      const float tw = ok ? tgt_w[row] : 0.f;
      const float w = fw * tw;                          // frame_weights * sum_k t
      float yt = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = (i * G + lg) * 4 + j;
          const bool hit = c == ti;
          if (hit) yt = v[i][j];
          v[i][j] = (v[i][j] - (hit ? tw : 0.f)) * w;
        }
      yt = group_sum<G>(yt);                            // exactly one lane holds it
      store_row<G, NV>(v, diff + (size_t)row * ldd, cols, lg, ok);
      if (ok && lg == 0) {
        const int targ = tw > 0.f ? ti : (tw == 0.f ? 0 : (ti == 0 ? 1 : 0));   // FindRowMaxId of the target row
        acc[0] -= (double)(logf(yt + 1e-20f) * tw * w);
        acc[1] -= (double)(logf(tw + 1e-20f) * tw * w);
        acc[2] += (double)(yt * tw * w);
        acc[3] += (bi == targ) ? (double)w : 0.0;
        acc[4] += (double)w;
      }
    } else {
      float t[NV][4];
      load_row<G, NV>(t, tgt + (size_t)row * ldt, cols, lg, ok, 0.f);
      float tsum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tsum += t[i][j];
      tsum = group_sum<G>(tsum);
      const float w = fw * tsum;
      float tbest = -INFINITY; int tbi = 0x7fffffff;
      float pce = 0.f, pen = 0.f, plk = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = (i * G + lg) * 4 + j;
          if (c < cols) {
            const float yv = v[i][j], tv = t[i][j];
            if (tv > tbest) { tbest = tv; tbi = c; }
            pce += logf(yv + 1e-20f) * tv * w;
            pen += logf(tv + 1e-20f) * tv * w;
            plk += yv * tv * w;
            v[i][j] = (yv - tv) * w;
          }
        }
      group_argmax<G>(tbest, tbi);
      pce = group_sum<G>(pce); pen = group_sum<G>(pen); plk = group_sum<G>(plk);
      store_row<G, NV>(v, diff + (size_t)row * ldd, cols, lg, ok);
      if (ok && lg == 0) {
        acc[0] -= (double)pce; acc[1] -= (double)pen; acc[2] += (double)plk;
        acc[3] += (bi == tbi) ? (double)w : 0.0;
        acc[4] += (double)w;
      }
    }
  }
  block_flush_stats(stats, acc);
}

inline bool aligned16(const void* p, int ld) { return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0); }
// persistent grid: as many blocks as are resident at once (queried per kernel), each walking rows with a grid stride;
// a grid of several register-limited waves paid a launch / drain ramp per wave (ncu: 8 waves for the 1500-column softmax)
template <typename KernelT>
inline int row_grid(KernelT kernel, long long rows, int rows_per_block) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long b = (rows + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)aslp_num_sms() * per_sm;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// picks (G, NV) for the row length; F is a functor template taking <G, NV>
#define ROWREG_DISPATCH(cols, CALL)                                    \
  do {                                                                 \
    if ((cols) <= 32) { CALL(8, 1); }                                  \
    else if ((cols) <= 64) { CALL(8, 2); }                             \
    else if ((cols) <= 128) { CALL(8, 4); }                            \
    else if ((cols) <= 256) { CALL(32, 2); }                           \
    else if ((cols) <= 512) { CALL(32, 4); }                           \
    else if ((cols) <= 1024) { CALL(32, 8); }                          \
    else if ((cols) <= 2048) { CALL(32, 16); }                         \
  } while (0)
constexpr int MAX_COLS = 2048;

}  // namespace rowreg
