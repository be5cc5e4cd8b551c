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
// The copies are cp.async.cg (L2 -> shared memory, no registers, no L1), so a thread can have its whole share of the
// exchange vector in flight at once; it then inspects what landed in shared memory and re-issues only the items that
// still show the sentinel.  A round therefore costs ONE L2 round trip however many items it covers (checking item by
// item with register loads serialised a round trip per item once the producers were done, and the registers the
// in-flight data needed collided with the register-resident weights of the tensor-core form).
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Pipelined polling (ASLP_PIPE_ROUNDS = R > 1).  A staging round samples L2 about half a round trip after it is issued and
// lands half a round trip later (~700 cycles in all); issued once and early it misses the slowest producer and pays a
// second full round trip, issued late it idles -- so R rounds are issued a few hundred cycles apart, each its own
// cp.async group, all into the same shared-memory destination.  Exchange words only ever change from the sentinel to
// their final value, so overlapping rounds are monotonic: the consumer waits for the groups in order and leaves on the
// first round that shows no sentinel; later rounds rewrite identical values.  A thread drains its stale rounds
// (cp_async_wait_all, long complete by then) before the barrier that releases the destination / the source slot.
// MEASURED AND NOT ADOPTED (profiles/r02_ab_pipe_polling.jsonl, cfg3 geometry, same box): R = 3 is 9 % slower forward
// (2.17 vs 1.99 us per step) and 8 % slower backward (2.65 vs 2.45), R = 2 in between -- issuing a round costs the
// finishing threads ~170 cycles of LSU issue each (forward finish phase 971 -> 1412 cycles) and the single well-timed
// round already samples L2 just as the slowest producer's store lands (rounds = 1.0 in the phase timing).  Default 1.
#ifndef ASLP_PIPE_ROUNDS
#define ASLP_PIPE_ROUNDS 1
#endif
#ifndef ASLP_PIPE_IDLE_NS0
#define ASLP_PIPE_IDLE_NS0 200     // idle (non-finishing) warps: sleep before their first round ...
#endif
#ifndef ASLP_PIPE_IDLE_NS
#define ASLP_PIPE_IDLE_NS 100      // ... and between rounds
#endif

template <int BATCH>
__device__ __forceinline__ void stage_poll(float* xT, int SP, const float* g, int K, int SX, int s0, int sg4,
                                           long long* dbg = nullptr) {   // dbg (timing builds): [0] += rounds, [1] += first-round ticks
  static_assert(BATCH <= 32, "pending mask is 32 bits");
  const int total = K * sg4;
  if (NT % sg4 == 0) {
    // a thread's items are a constant stride apart (same stream quad, every (NT/sg4)-th row): no per-item address state
    const int kstep = NT / sg4;
    const size_t sstride = (size_t)kstep * SX;
    const int dstride = kstep * SP;
    for (int base = 0; base < total; base += NT * BATCH) {
      const int i0 = base + threadIdx.x;
      const int k0 = i0 / sg4, q = i0 - k0 * sg4;
      const float* src0 = g + (size_t)k0 * SX + s0 + q * 4;
      float* dst0 = xT + k0 * SP + q * 4;
      unsigned pending = 0;
#pragma unroll
      for (int j = 0; j < BATCH; ++j)
        if (i0 + j * NT < total) { cp_async16(dst0 + j * dstride, src0 + j * sstride); pending |= 1u << j; }
      unsigned rounds = 0;
      const long long tq0 = dbg != nullptr ? clock64() : 0;
      while (pending != 0) {
        cp_async_wait_all();
        if (dbg != nullptr) { if (rounds == 0) dbg[1] += clock64() - tq0; dbg[0] += 1; }
        unsigned still = 0;
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
          if ((pending >> j) & 1u) {
            const float4 v = *reinterpret_cast<const float4*>(dst0 + j * dstride);
            if (has_sentinel(v)) { still |= 1u << j; cp_async16(dst0 + j * dstride, src0 + j * sstride); }
          }
        }
        pending = still;
        if (++rounds > POLL_LIMIT) __trap();
      }
    }
    return;
  }
  // generic (odd-sized last stream group): one item at a time per thread
  for (int i = threadIdx.x; i < total; i += NT) {
    const int k = i / sg4, q = i - k * sg4;
    const float* src = g + (size_t)k * SX + s0 + q * 4;
    float4 v = ld_vol4(src);
    unsigned rounds = 0;
    while (has_sentinel(v)) { v = ld_vol4(src); if (++rounds > POLL_LIMIT) __trap(); }
    *reinterpret_cast<float4*>(xT + k * SP + q * 4) = v;
  }
}

// ---- hoisted form: everything about a thread's share of a staging pass that does not change from step to step
// (the integer divisions and 64-bit address arithmetic of stage_poll cost ~200 instructions per warp per step)
struct StageDesc {
  unsigned src_off, dst_off;     // floats: first item inside a time slot [K][SX] / inside xT
  unsigned sstride, dstride;     // floats between consecutive items of this thread
  unsigned mask;                 // bit j = item j exists
  bool fast;                     // false: shape needs the generic stage_poll
};
template <int BATCH>
__device__ __forceinline__ StageDesc stage_prepare(int K, int SX, int SP, int s0, int sg4) {
  StageDesc d;
  const int total = K * sg4;
  d.fast = (sg4 > 0) && (NT % sg4 == 0) && (total <= NT * BATCH);
  d.src_off = d.dst_off = d.sstride = d.dstride = d.mask = 0;
  if (!d.fast) return d;
  const int kstep = NT / sg4;
  const int i0 = threadIdx.x;
  const int k0 = i0 / sg4, q = i0 - k0 * sg4;
  d.src_off = (unsigned)(k0 * SX + s0 + q * 4);
  d.dst_off = (unsigned)(k0 * SP + q * 4);
  d.sstride = (unsigned)(kstep * SX);
  d.dstride = (unsigned)(kstep * SP);
#pragma unroll
  for (int j = 0; j < BATCH; ++j)
    if (i0 + j * NT < total) d.mask |= 1u << j;
  if (d.mask == 0) { d.src_off = 0; d.dst_off = 0; }      // a thread without a share still inspects a valid address
  return d;
}
// issue / complete halves of a hoisted staging pass, so that work which does not depend on the exchange (the stores
// of the step just finished) can sit between them while the copies are in flight
template <int BATCH>
__device__ __forceinline__ void stage_issue(const StageDesc& d, float* xT, const float* g) {
  const float* src0 = g + d.src_off;
  float* dst0 = xT + d.dst_off;
#pragma unroll
  for (int j = 0; j < BATCH; ++j)
    if ((d.mask >> j) & 1u) cp_async16(dst0 + j * d.dstride, src0 + (size_t)j * d.sstride);
}
__device__ __forceinline__ unsigned sentinel_in(const float4& v) {
  // 0xFFFFFFFF is the largest unsigned word: one max-reduction instead of four compares
  const unsigned m = max(max(__float_as_uint(v.x), __float_as_uint(v.y)), max(__float_as_uint(v.z), __float_as_uint(v.w)));
  return m == SENTINEL ? 1u : 0u;
}
template <int BATCH>
__device__ __forceinline__ void stage_complete(const StageDesc& d, float* xT, const float* g, long long* dbg = nullptr) {
  const float* src0 = g + d.src_off;
  float* dst0 = xT + d.dst_off;
  unsigned rounds = 0;
  const long long tq0 = dbg != nullptr ? clock64() : 0;
  for (;;) {
    cp_async_wait_all();
    if (dbg != nullptr) { if (rounds == 0) dbg[1] += clock64() - tq0; dbg[0] += 1; }
    // branch-free inspection: all loads first, one mask out
    unsigned still = 0;
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const bool live = (d.mask >> j) & 1u;
      const float4 v = *reinterpret_cast<const float4*>(live ? dst0 + j * d.dstride : dst0);
      still |= (live ? sentinel_in(v) : 0u) << j;
    }
    if (still == 0) break;
#pragma unroll
    for (int j = 0; j < BATCH; ++j)
      if ((still >> j) & 1u) cp_async16(dst0 + j * d.dstride, src0 + (size_t)j * d.sstride);
    if (++rounds > POLL_LIMIT) __trap();
  }
}
// pipelined form: the R rounds were issued (stage_issue + cp_async_commit each) by the caller; inspect them in order
template <int BATCH, int R>
__device__ __forceinline__ void stage_complete_pipe(const StageDesc& d, float* xT, const float* g, long long* dbg = nullptr) {
  float* dst0 = xT + d.dst_off;
  auto clean = [&]() {
    unsigned still = 0;
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const bool live = (d.mask >> j) & 1u;
      const float4 v = *reinterpret_cast<const float4*>(live ? dst0 + j * d.dstride : dst0);
      still |= (live ? sentinel_in(v) : 0u) << j;
    }
    return still == 0;
  };
  const long long tq0 = dbg != nullptr ? clock64() : 0;
  bool ok = false;
  if (R >= 3) { cp_async_wait_group<2>(); if (dbg != nullptr) { dbg[1] += clock64() - tq0; dbg[0] += 1; } ok = clean(); }
  if (R >= 2 && !ok) { cp_async_wait_group<1>(); if (dbg != nullptr) dbg[0] += 1; ok = clean(); }
  if (!ok) { cp_async_wait_group<0>(); if (dbg != nullptr) dbg[0] += 1; ok = clean(); }
  unsigned rounds = 0;
  while (!ok) {
    stage_issue<BATCH>(d, xT, g);
    cp_async_wait_all();
    if (dbg != nullptr) dbg[0] += 1;
    ok = clean();
    if (++rounds > POLL_LIMIT) __trap();
  }
}
template <int BATCH>
__device__ __forceinline__ void stage_poll_desc(const StageDesc& d, float* xT, const float* g, long long* dbg = nullptr) {
  stage_issue<BATCH>(d, xT, g);
  stage_complete<BATCH>(d, xT, g, dbg);
}
// fire-and-forget: pull a line into L2 without tying up a register or a scoreboard
__device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

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
