// recur.cuh -- device building blocks shared by the persistent recurrence kernels (lstm.cu, gru.cu):
// sentinel-polled exchange staging, the (4 rows) x (16 streams) work-unit contraction and its butterfly reduce.
#pragma once
#include "common.cuh"

namespace recur {

constexpr int NT = 256;                 // threads per CTA
constexpr int NW = NT / 32;
constexpr unsigned SENTINEL = 0xFFFFFFFFu;   // a NaN payload arithmetic never produces
constexpr unsigned POLL_LIMIT = 1u << 24;

__device__ __forceinline__ float4 ld_vol4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool has_sentinel(const float4& v) {
  return __float_as_uint(v.x) == SENTINEL || __float_as_uint(v.y) == SENTINEL ||
         __float_as_uint(v.z) == SENTINEL || __float_as_uint(v.w) == SENTINEL;
}

// publishing store of an exchange word (gpu-scope relaxed: lands in L2 where the pollers read)
__device__ __forceinline__ void st_pub(float* p, float v) {
  asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// stage X[k][s0 .. s0+SG) (global, stream-minor, stride SX) into xT[k][SP]; polls until produced.
// All loads of a batch are issued before any is checked, so one L2 round trip covers the batch.
__device__ __forceinline__ void stage_poll(float* xT, int SP, const float* g, int K, int SX, int s0, int sg4) {
  constexpr int BATCH = 8;
  const int total = K * sg4;
  for (int base = 0; base < total; base += NT * BATCH) {
    float4 v[BATCH];
    const float* src[BATCH];
    int dst[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = base + j * NT + threadIdx.x;
      const int ii = i < total ? i : 0;
      const int k = ii / sg4, q = ii - k * sg4;
      src[j] = g + (size_t)k * SX + s0 + q * 4;
      dst[j] = i < total ? k * SP + q * 4 : -1;
      if (dst[j] >= 0) v[j] = ld_vol4(src[j]);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      if (dst[j] >= 0) {
        unsigned n = 0;
        while (has_sentinel(v[j])) { v[j] = ld_vol4(src[j]); if (++n > POLL_LIMIT) __trap(); }
        *reinterpret_cast<float4*>(xT + dst[j]) = v[j];
      }
    }
  }
}

// acc[r][j] = sum_k w[r][k] * xT[k][sc + j],  lanes stride k.  w rows are ldw apart in smem.
template <int NR>
__device__ __forceinline__ void unit_dot(float (&acc)[NR][16], const float* w, int ldw, int K, const float* xT, int SP, int sc, int lane) {
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[r][j] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float4* xp = reinterpret_cast<const float4*>(xT + k * SP + sc);
    const float4 x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
    const float xs[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float wv = w[r * ldw + k];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[r][j] = fmaf(wv, xs[j], acc[r][j]);
    }
  }
}

// butterfly: 16 stream columns over 32 lanes.  On return lane L holds in out[r] the full sum for
// stream (L >> 1) (both lanes of a pair hold the same value).
template <int NR>
__device__ __forceinline__ void unit_reduce(float (&acc)[NR][16], float (&out)[NR], int lane) {
  // step 1: xor 16 -> keep 8 streams
  float a8[NR][8];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float keep = hi ? acc[r][8 + j] : acc[r][j];
        const float send = hi ? acc[r][j] : acc[r][8 + j];
        a8[r][j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
  }
  float a4[NR][4];
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = hi ? a8[r][4 + j] : a8[r][j];
        const float send = hi ? a8[r][j] : a8[r][4 + j];
        a4[r][j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
  }
  float a2[NR][2];
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float keep = hi ? a4[r][2 + j] : a4[r][j];
        const float send = hi ? a4[r][j] : a4[r][2 + j];
        a2[r][j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
  }
  {
    const bool hi = lane & 2;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float keep = hi ? a2[r][1] : a2[r][0];
      const float send = hi ? a2[r][0] : a2[r][1];
      float v = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      out[r] = v;
    }
  }
}
// stream handled by a lane after unit_reduce: bits (4,3,2,1) of the lane select halves in that order
__device__ __forceinline__ int lane_stream(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ---------------------------------------------------------------- exchange-buffer initialisation
// rows 1..T (all slots except the boundary one) get the sentinel for valid streams and 0 for the
// padding streams; the boundary slot gets the boundary state read from buf/dbuf (or zeros).
static __global__ void xch_init_kernel(float* x, int dim, int T, int S, int SX, int boundary_slot, const float* src, int lds, int col0) {
  const long long total = (long long)(T + 2) * dim * SX;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % SX);
    const long long rem = i / SX;
    const int k = (int)(rem % dim);
    const int slot = (int)(rem / dim);
    float v;
    if (s >= S) v = 0.f;
    else if (slot == boundary_slot) v = (src != nullptr) ? src[((size_t)slot * S + s) * lds + col0 + k] : 0.f;
    else v = __uint_as_float(SENTINEL);
    x[i] = v;
  }
}


}  // namespace recur
