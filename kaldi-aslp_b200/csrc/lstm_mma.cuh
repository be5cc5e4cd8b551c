// lstm_mma.cuh -- tensor-core form of the persistent LSTM recurrence (included by lstm.cu inside its anonymous
// namespace, after DirDev / Launch).  Same exchange protocol and buffer layout as the SIMT kernels in lstm.cu; what
// changes is the partitioning and the per-step contraction:
//   * the streams of a minibatch are independent, so they are split into `pgroups` parallel stream groups; a CHAIN is
//     (direction, stream group) and owns nblk CTAs that partition the cells.  A CTA therefore gathers only its
//     group's columns of the exchange vector each step: the all-gather volume through L2 -- what bounded the first
//     versions of this kernel (every CTA re-reading the full [dim][S] vector, 2.6 MB fwd / 10.5 MB bwd per step at
//     cfg3) -- drops by the number of groups;
//   * the CTA's slice of the recurrent weights lives in REGISTERS for the whole sequence as fp32 values in
//     mma.m16n8k8 A-fragment layout (rows of the slice = M); 8 streams are the N dimension; the contraction dimension
//     is split over the 8 warps.  Operands are split into tf32 hi/lo halves on the fly (3xTF32: lo*hi + hi*lo + hi*hi,
//     fp32-grade).  The 8 partial tiles meet in shared memory and one thread per (cell, stream) finishes the sum and
//     runs the whole cell update in registers;
//   * the non-recurrent inputs of the NEXT step (x*W_x^T + bias forward; the forward activations backward) are
//     prefetched before the exchange wait, so no HBM latency sits on the recurrent chain.
// tcgen05 does not fit here: a CTA owns 8..48 output rows and 8 streams, and the chain is latency-bound, so the
// warp-level mma.sync with register-resident operands is the shortest path from "exchange landed" to "next value
// published".  Only the single-exchange form is handled (R == 0: plain LSTM, or the folded projection W_gifo_r W_r_m).
#pragma once

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Accumulator scheme of the forward contraction.  ASLP_FWD_ACC3 = 1: three independent accumulators per (m-tile, slot) --
// lo*hi, hi*lo, hi*hi never wait for each other -- with 1 (MT >= 2) or 2 k-interleave slots; 0: two accumulators (the cross
// terms share one, issued with the hi*hi MMA between them) with 2 / 4 slots.  A/B-measured in one run on the cfg3 slice
// (tools/gpu_ab_fwd.sh); the issue floor is 8 cycles per MMA per sub-partition (tools/micro/mma_rate.cu).
#ifndef ASLP_FWD_ACC3
#define ASLP_FWD_ACC3 1
#endif
#ifndef ASLP_ISSUE_EARLY
#define ASLP_ISSUE_EARLY 0
#endif
#ifndef ASLP_FWD_FP16X3
#define ASLP_FWD_FP16X3 1      // 1: forward contraction as fp16 hi/lo split (3 x m16n8k16 per 16 k); 0: 3xTF32 (6 x m16n8k8)
#endif
#ifndef ASLP_WARM_AHEAD
#define ASLP_WARM_AHEAD 2      // items ahead the non-recurrent operands are pulled into L2 (0: no prefetch)
#endif
#if ASLP_FWD_ACC3
// ---- fp16 two-term split (forward contraction, ASLP_FWD_FP16X3): x = h1 + h2 + O(2^-22 |x|) with h1, h2 fp16; the products
// h*h are exact in the fp32 accumulator, so w1*m1 + w2*m1 + w1*m2 has the accuracy of 3xTF32 at HALF the MMA count (the
// legacy m16n8k16.f16 issues at the same 8 cycles as m16n8k8.tf32, tools/micro/mma_rate.cu).  Only for BOUNDED operands:
// the recurrent input m = o * tanh(c) lies in (-1, 1) and weights are O(1), so fp16 range is no issue and the absolute
// representation error is <= 3e-8 (fp16 subnormal step / 2); the backward operand dgifo is unbounded below and stays tf32.
#include <cuda_fp16.h>
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {      // two values -> packed halves (x0 low)
  const __half a0 = __float2half_rn(x0), a1 = __float2half_rn(x1);
  const __half b0 = __float2half_rn(x0 - __half2float(a0)), b1 = __float2half_rn(x1 - __half2float(a1));
  const __half2 h = __halves2half2(a0, a1), l = __halves2half2(b0, b1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// part_w = Wslice * X with pre-split fp16 weight fragments wh / wl [MT_][KH_][4] (a0: row g, k 2tig..+1; a1: row g+8;
// a2: row g, k 2tig+8..+9; a3: row g+8) and the m values split on the fly; KH_ 16-wide k tiles starting at row kbase
template <int MT_, int KH_>
__device__ __forceinline__ void mma_contract_h(const uint32_t (&wh)[MT_][KH_][4], const uint32_t (&wl)[MT_][KH_][4], const float* xT, int SP,
                                               int kbase, int Kdim, int ntiles, float* part_w, int lane) {
  constexpr int NPP = MT_ * 16 + 4;
  const int g = lane >> 2, tig = lane & 3;
  for (int nt = 0; nt < ntiles; ++nt) {
    float acc[MT_][3][4];
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
      for (int h = 0; h < 3; ++h) { acc[mt][h][0] = acc[mt][h][1] = acc[mt][h][2] = acc[mt][h][3] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < KH_; ++kh) {
      // rows past the warp's share carry zero weights; clamp them onto valid (finite) rows
      const int k0 = kbase + kh * 16 + 2 * tig;
      const float* col = xT + nt * 8 + g;
      const float x00 = col[(size_t)min(k0, Kdim - 1) * SP], x01 = col[(size_t)min(k0 + 1, Kdim - 1) * SP];
      const float x10 = col[(size_t)min(k0 + 8, Kdim - 1) * SP], x11 = col[(size_t)min(k0 + 9, Kdim - 1) * SP];
      uint32_t bh0, bl0, bh1, bl1;
      split_h2(x00, x01, bh0, bl0);
      split_h2(x10, x11, bh1, bl1);
#pragma unroll
      for (int mt = 0; mt < MT_; ++mt) {
        mma_f16(acc[mt][1], wl[mt][kh], bh0, bh1);
        mma_f16(acc[mt][2], wh[mt][kh], bl0, bl1);
        mma_f16(acc[mt][0], wh[mt][kh], bh0, bh1);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt) {
      float c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) c[q] = (acc[mt][1][q] + acc[mt][2][q]) + acc[mt][0][q];      // small terms first
      float* p0 = part_w + (size_t)(nt * 8 + 2 * tig) * NPP + mt * 16 + g;
      p0[0] = c[0]; p0[NPP] = c[1]; p0[8] = c[2]; p0[NPP + 8] = c[3];
    }
  }
}

template <int MT_> struct MmaAcc { static constexpr int K = MT_ >= 2 ? 1 : 2; static constexpr int SETS = 3; };
#else
template <int MT_> struct MmaAcc { static constexpr int K = MT_ >= 4 ? 1 : (MT_ >= 2 ? 2 : 4); static constexpr int SETS = 2; };
#endif

// part_w[stream][NPP] = Wslice[MT_*16 rows, k-slice of this warp] * X[k-slice, streams]   (8 streams per n-tile)
// wa: A fragments (a0: row g, k tig; a1: row g+8, k tig; a2: row g, k tig+4; a3: row g+8, k tig+4), fp32.
template <int MT_, int KT_>
__device__ __forceinline__ void mma_contract(const float (&wa)[MT_][KT_][4], const float* xT, int SP, int kbase, int Kdim,
                                             int ntiles, float* part_w, int lane) {
  constexpr int NPP = MT_ * 16 + 4;
  constexpr int AK = MmaAcc<MT_>::K;
  constexpr int SETS = MmaAcc<MT_>::SETS;
  const int g = lane >> 2, tig = lane & 3;
  for (int nt = 0; nt < ntiles; ++nt) {
    float acc[MT_][AK][SETS][4];
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
      for (int a = 0; a < AK; ++a)
#pragma unroll
        for (int h = 0; h < SETS; ++h) { acc[mt][a][h][0] = acc[mt][a][h][1] = acc[mt][a][h][2] = acc[mt][a][h][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT_; ++kt) {
      {
        // no branch on the warp's real k-tile count: tiles past it carry zero weights and re-read the last valid rows, so
        // the loads of the whole unrolled loop can be hoisted ahead of the first MMA
        const float* xp = xT + (size_t)(min(kbase + kt * 8, Kdim - 8) + tig) * SP + nt * 8 + g;
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(xp[0], bh0, bl0);
        split_tf32(xp[4 * SP], bh1, bl1);
#pragma unroll
        for (int mt = 0; mt < MT_; ++mt) {
          uint32_t ah[4], al[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_tf32(wa[mt][kt][q], ah[q], al[q]);
          mma_tf32(acc[mt][kt % AK][1], al, bh0, bh1);
          mma_tf32(acc[mt][kt % AK][0], ah, bh0, bh1);
          mma_tf32(acc[mt][kt % AK][SETS - 1], ah, bl0, bl1);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt) {
      float c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float lo = 0.f, hi = 0.f;
#pragma unroll
        for (int a = 0; a < AK; ++a) {
          lo += acc[mt][a][1][q];
          if (SETS == 3) lo += acc[mt][a][2][q];
          hi += acc[mt][a][0][q];
        }
        c[q] = lo + hi;                                  // small terms first
      }
      // c0: (row g, stream 2tig)  c1: (g, 2tig+1)  c2: (g+8, 2tig)  c3: (g+8, 2tig+1); stored stream-major
      float* p0 = part_w + (size_t)(nt * 8 + 2 * tig) * NPP + mt * 16 + g;
      p0[0] = c[0]; p0[NPP] = c[1]; p0[8] = c[2]; p0[NPP + 8] = c[3];
    }
  }
}

// Same contraction with the A fragments read from shared memory in fragment order ws[kt][mt][q][lane] (conflict-free
// 32-word rows) instead of registers: the backward slice (K = 4C) needs 80+ registers per thread at cfg3, which together
// with the accumulators pushed the kernel into spills whose reloads sat on the recurrent chain.  kt runs over ktp tiles
// (a multiple of 4, zero-filled past the warp's real count).
template <int MT_, int KTP_>     // KTP_ > 0: compile-time tile count (fully unrolled, loads hoisted); 0: run-time ktp
__device__ __forceinline__ void mma_contract_ws(const float* ws, int ktp, const float* xT, int SP, int kbase, int Kdim,
                                                int ntiles, float* part_w, int lane) {
  constexpr int NPP = MT_ * 16 + 4;
  constexpr int AK = 4;
  const int g = lane >> 2, tig = lane & 3;
  for (int nt = 0; nt < ntiles; ++nt) {
    float acc[MT_][AK][2][4];
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
      for (int a = 0; a < AK; ++a)
#pragma unroll
        for (int h = 0; h < 2; ++h) { acc[mt][a][h][0] = acc[mt][a][h][1] = acc[mt][a][h][2] = acc[mt][a][h][3] = 0.f; }
    const int kend = KTP_ > 0 ? KTP_ : ktp;
#pragma unroll (KTP_ > 0 ? KTP_ / AK : 1)
    for (int kb = 0; kb < kend; kb += AK) {
#pragma unroll
      for (int a = 0; a < AK; ++a) {
        const int kt = kb + a;
        const float* xp = xT + (size_t)(min(kbase + kt * 8, Kdim - 8) + tig) * SP + nt * 8 + g;
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(xp[0], bh0, bl0);
        split_tf32(xp[4 * SP], bh1, bl1);
#pragma unroll
        for (int mt = 0; mt < MT_; ++mt) {
          const float* wp = ws + ((size_t)(kt * MT_ + mt) * 4) * 32 + lane;
          uint32_t ah[4], al[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_tf32(wp[q * 32], ah[q], al[q]);
          mma_tf32(acc[mt][a][1], al, bh0, bh1);
          mma_tf32(acc[mt][a][1], ah, bl0, bl1);
          mma_tf32(acc[mt][a][0], ah, bh0, bh1);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt) {
      float c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float lo = 0.f, hi = 0.f;
#pragma unroll
        for (int a = 0; a < AK; ++a) { lo += acc[mt][a][1][q]; hi += acc[mt][a][0][q]; }
        c[q] = lo + hi;
      }
      float* p0 = part_w + (size_t)(nt * 8 + 2 * tig) * NPP + mt * 16 + g;
      p0[0] = c[0]; p0[NPP] = c[1]; p0[8] = c[2]; p0[NPP + 8] = c[3];
    }
  }
}

// Fast form of the backward contraction (what cfg3 runs).  Against mma_contract_ws: the A fragments of a (k-tile, m-tile)
// are ONE 128-bit shared load per lane (layout [kt][mt][lane] float4) instead of four 32-bit loads; the three products of
// a k-tile go to three independent accumulator sets, k-interleaved four ways, so no MMA waits for the one issued before
// it; SP == 8 and an exact K split make the B addresses immediates (no per-tile integer arithmetic).  Measured before:
// 22 cycles per MMA per sub-partition against an issue floor of 8 (tools/micro/mma_rate.cu).  The tf32 residual of the
// weights is still taken on the fly: a second, pre-split copy in shared memory would double the A traffic
// (8 warps x 20 tiles x 1 KB = 164 KB per step) and make the step shared-memory-bound at ~1600 cycles.
// KTP_ k-tiles per warp (multiple of 4); xp0 = xT + (first k row of the warp + tig) * 8 + g.
template <int MT_, int KTP_>
__device__ __forceinline__ void mma_contract_fast(const float4* wfrag, const float* xp0, float* part_w, int lane) {
  constexpr int NPP = MT_ * 16 + 4;
  constexpr int AK = 4;
  static_assert(KTP_ % AK == 0, "k-tiles per warp must be a multiple of 4");
  const int g = lane >> 2, tig = lane & 3;
  float hh[MT_][AK][4], lh[MT_][AK][4], hl[MT_][AK][4];
#pragma unroll
  for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
    for (int a = 0; a < AK; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) { hh[mt][a][q] = 0.f; lh[mt][a][q] = 0.f; hl[mt][a][q] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < KTP_; ++kt) {
    const int a = kt % AK;
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(xp0[kt * 64], bh0, bl0);
    split_tf32(xp0[kt * 64 + 32], bh1, bl1);
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt) {
      const float4 w4 = wfrag[(kt * MT_ + mt) * 32 + lane];
      uint32_t ah[4], al[4];
      split_tf32(w4.x, ah[0], al[0]); split_tf32(w4.y, ah[1], al[1]); split_tf32(w4.z, ah[2], al[2]); split_tf32(w4.w, ah[3], al[3]);
      mma_tf32(lh[mt][a], al, bh0, bh1);
      mma_tf32(hl[mt][a], ah, bl0, bl1);
      mma_tf32(hh[mt][a], ah, bh0, bh1);
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT_; ++mt) {
    float c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int a = 0; a < AK; ++a) { lo += lh[mt][a][q] + hl[mt][a][q]; hi += hh[mt][a][q]; }
      c[q] = lo + hi;                                    // small terms first
    }
    float* p0 = part_w + (size_t)(2 * tig) * NPP + mt * 16 + g;
    p0[0] = c[0]; p0[NPP] = c[1]; p0[8] = c[2]; p0[NPP + 8] = c[3];
  }
}

struct MmaCta {
  int dir, pg, blk;         // chain = (dir, pg); blk = cell block inside the chain
  int sbeg, send;           // streams of this chain
};
__device__ __forceinline__ MmaCta mma_cta(const Launch& L) {
  MmaCta c;
  const int per_dir = L.nblk * L.pgroups;
  c.dir = blockIdx.x / per_dir;
  const int rem = blockIdx.x - c.dir * per_dir;
  c.pg = rem / L.nblk;
  c.blk = rem - c.pg * L.nblk;
  c.sbeg = c.pg * L.SGP;
  c.send = min(L.d[c.dir].S, c.sbeg + L.SGP);
  return c;
}

// ---------------------------------------------------------------- forward
template <int MT_, int KT_>
__global__ void __launch_bounds__(NT, 1) lstm_fwd_mma_kernel(Launch L) {
  extern __shared__ float smem[];
  const MmaCta cta = mma_cta(L);
  const DirDev& D = L.d[cta.dir];
  const int T = D.T, S = D.S, C = D.C, SX = D.SX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SP = L.SP, SG = L.SG;
  constexpr int NPP = MT_ * 16 + 4;
  const int K = C;
  const int c0 = cta.blk * D.cb, nc = max(0, min(D.cb, C - c0));
  if (nc == 0 || cta.sbeg >= S) return;                  // nobody waits on a CTA that owns no cells / no streams

  float* xT = smem;                                      // [K][SP]       m(t-1) of this chain's streams, stream-minor
  float* part = xT + (size_t)K * SP;                     // [NW][SG][NPP] per-warp partial gate sums
  float* cst = part + (size_t)NW * SG * NPP;             // [cb][SGP]     c(t-1)
  float* pst = cst + (size_t)D.cb * L.SGP;               // [3][cb]       peepholes

  // ---- one-time: weight slice -> A fragments (row n of the slice = cell n/4, gate n%4)
  const int ktiles = K >> 3, ktw = (ktiles + NW - 1) / NW;
  const int kt0 = warp * ktw, nkt = max(0, min(ktw, ktiles - kt0));
#if ASLP_FWD_FP16X3
  constexpr int KH = (KT_ + 1) / 2;                      // 16-wide k tiles per warp
  uint32_t wh[MT_][KH][4], wl[MT_][KH][4];
  {
    const int g = lane >> 2, tig = lane & 3;
    const int kend = (kt0 + nkt) * 8;                     // end of this warp's k range
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
      for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = mt * 16 + g + (q & 1) * 8, cl = n >> 2, gate = n & 3;
          const int k = kt0 * 8 + kh * 16 + 2 * tig + (q >> 1) * 8;
          const float* wrow = D.w_r + (size_t)(gate * C + c0 + cl) * D.ldwr;
          const float w0 = (cl < nc && k < kend && k < K) ? wrow[k] : 0.f;
          const float w1 = (cl < nc && k + 1 < kend && k + 1 < K) ? wrow[k + 1] : 0.f;
          if (fabsf(w0) > 32768.f || fabsf(w1) > 32768.f) __trap();     // outside the fp16 split's range: an error, never a silent inf
          split_h2(w0, w1, wh[mt][kh][q], wl[mt][kh][q]);
        }
  }
#else
  float wa[MT_][KT_][4];
  {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
      for (int kt = 0; kt < KT_; ++kt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = mt * 16 + g + (q & 1) * 8, cl = n >> 2, gate = n & 3;
          const int k = (kt0 + kt) * 8 + tig + (q >> 1) * 4;
          wa[mt][kt][q] = (kt < nkt && cl < nc) ? D.w_r[(size_t)(gate * C + c0 + cl) * D.ldwr + k] : 0.f;
        }
  }
#endif
  const int slot0 = D.reverse ? T + 1 : 0;
  for (int i = threadIdx.x; i < D.cb * L.SGP; i += NT) {
    const int cl = i / L.SGP, s = cta.sbeg + (i - cl * L.SGP);
    cst[i] = (cl < nc && s < cta.send) ? D.buf[((size_t)slot0 * S + s) * D.ldb + 4 * C + c0 + cl] : 0.f;
  }
  for (int i = threadIdx.x; i < 3 * D.cb; i += NT) {
    const int which = i / D.cb, cl = i - which * D.cb;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (cl < nc) ? p[c0 + cl] : 0.f;
  }
  for (int i = threadIdx.x; i < K * SP; i += NT) xT[i] = 0.f;
  __syncthreads();

  // ---- the (cell, stream-in-group) item this thread finishes every step
  const int s_local = threadIdx.x % SG, cl = threadIdx.x / SG;
  const bool owner = cl < nc;
  const int ngroups = (cta.send - cta.sbeg + SG - 1) / SG;         // sequential staging groups inside the chain
  const int nitems = T * ngroups;
  const int reverse = D.reverse, ldb = D.ldb;
  const float clip = D.clip;
  const int* seq_len = D.seq_len;
  float* const buf = D.buf;
  float* const xa = D.xa;
  const float pi = owner ? pst[cl] : 0.f, pf = owner ? pst[D.cb + cl] : 0.f, po = owner ? pst[2 * D.cb + cl] : 0.f;
  // step-invariant staging share when the chain has one staging group (the common case)
  const int sg4_0 = (min(cta.sbeg + SG, SX) - cta.sbeg) >> 2;
  constexpr int SB = (KT_ + 1) / 2;                               // items per thread when a group is 8 streams wide
  const StageDesc sd = stage_prepare<SB>(K, SX, SP, cta.sbeg, sg4_0);
  const bool hoisted = (ngroups == 1) && sd.fast;
  const size_t slot = (size_t)K * SX;
  auto item_row = [&](int it, int& t, int& s0) {
    if (ngroups == 1) { t = reverse ? T - it : 1 + it; s0 = cta.sbeg; return; }
    const int step = it / ngroups, grp = it - step * ngroups;
    t = reverse ? T - step : 1 + step;
    s0 = cta.sbeg + grp * SG;
  };
  // x*W_x^T + bias of a future item: pulled into L2 two items ahead (no register, no scoreboard), read into registers
  // right after the exchange wait of its own step, consumed after the contraction
  auto pre_ptr = [&](int it) -> const float* {
    int t, s0; item_row(it, t, s0);
    const int s = s0 + s_local;
    return (owner && it < nitems && s < cta.send) ? buf + ((size_t)t * S + s) * ldb + c0 + cl : nullptr;
  };
  auto warm = [&](int it) {
    const float* p = pre_ptr(it);
    if (p != nullptr) { prefetch_l2(p); prefetch_l2(p + C); prefetch_l2(p + 2 * C); prefetch_l2(p + 3 * C); }
  };
  if (ASLP_WARM_AHEAD > 0) { for (int w = 0; w < ASLP_WARM_AHEAD; ++w) warm(w); }
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  constexpr int PR = ASLP_PIPE_ROUNDS;                   // staging rounds in flight per step (recur.cuh)
  if (hoisted) {
    stage_issue<SB>(sd, xT, xa + (size_t)(reverse ? T + 1 : 0) * slot);
    if (PR > 1) { for (int r = 0; r < PR; ++r) cp_async_commit(); }          // the boundary slot is there: one real round
  }
  // warps none of whose threads finishes an item have nothing to do between the contraction and the next exchange wait:
  // they sleep towards the publish of the finishing warps before their staging rounds (warp-uniform)
  const bool warp_fin = __any_sync(0xffffffffu, owner && s_local < min(SG, cta.send - cta.sbeg));
  auto pipe_round = [&](int it, int t, int idle_ns) {
    if (PR > 1 && hoisted && it + 1 < nitems) {
      if (!warp_fin && idle_ns > 0) __nanosleep(idle_ns);
      stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
      cp_async_commit();
    }
  };

  for (int it = 0; it < nitems; ++it) {
    int t, s0; item_row(it, t, s0);
    const int tp = reverse ? t + 1 : t - 1;
    const int nstr = min(SG, cta.send - s0);             // streams of this staging group
    RECUR_TICK(k0);
    long long* dbg = (L.timing != nullptr && threadIdx.x == 0) ? &tacc[5] : nullptr;
    if (hoisted) {                                       // copies were issued at the end of the previous step
      if (PR > 1) stage_complete_pipe<SB, PR>(sd, xT, xa + (size_t)tp * slot, dbg);
      else stage_complete<SB>(sd, xT, xa + (size_t)tp * slot, dbg);
    }
    else stage_poll<8>(xT, SP, xa + (size_t)tp * slot, K, SX, s0, (min(s0 + SG, SX) - s0) >> 2, dbg);
    float pre[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const float* p = pre_ptr(it);
      if (p != nullptr) { pre[0] = p[0]; pre[1] = p[C]; pre[2] = p[2 * C]; pre[3] = p[3 * C]; }
    }
    RECUR_TICK(k1);
    __syncthreads();
    RECUR_TICK(k2);
#if ASLP_FWD_FP16X3
    mma_contract_h<MT_, KH>(wh, wl, xT, SP, kt0 * 8, K, (nstr + 7) >> 3, part + (size_t)warp * SG * NPP, lane);
#else
    mma_contract<MT_, KT_>(wa, xT, SP, kt0 * 8, K, (nstr + 7) >> 3, part + (size_t)warp * SG * NPP, lane);
#endif
    RECUR_TICK(k3);
    if (PR > 1 && hoisted) cp_async_wait_all();          // stale rounds of this step (identical values) are done before xT is released
    __syncthreads();
    RECUR_TICK(k4);
    const int s = s0 + s_local;
    const bool fin_item = owner && s_local < nstr;
    float yg = 0.f, yi = 0.f, yf = 0.f, yo = 0.f, yc = 0.f, yh = 0.f, ym = 0.f;
    if (fin_item) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const float4 p = *reinterpret_cast<const float4*>(part + ((size_t)w * SG + s_local) * NPP + cl * 4);
        sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
      }
      const int ci = cl * L.SGP + (s - cta.sbeg);
      const float cprev = cst[ci];
      yg = pre[0] + sum.x;
      yi = pre[1] + sum.y + cprev * pi;
      yf = pre[2] + sum.z + cprev * pf;
      yo = pre[3] + sum.w;
      yi = ref_sigmoid(yi); yf = ref_sigmoid(yf); yg = ref_tanh(yg);
      yc = yg * yi + cprev * yf;
      yc = fminf(fmaxf(yc, -clip), clip);
      yh = ref_tanh(yc);
      yo = ref_sigmoid(yo + yc * po);
      ym = yh * yo;
      if (seq_len != nullptr && t > seq_len[s]) { yg = yi = yf = yo = yc = yh = ym = 0.f; }
      st_pub(xa + ((size_t)t * C + c0 + cl) * SX + s, ym);       // publish m(t) first: it is what the other SMs wait for
      if (L.timing != nullptr && threadIdx.x == 0) tacc[7] += clock64() - k4;
      cst[ci] = yc;
    }
    // next step's exchange copies; xT is free (contraction done).  Pipelined form: first round right after the publish,
    // the others behind the bookkeeping stores and the prefetch arithmetic.  Single-round form (ASLP_PIPE_ROUNDS = 1):
    // after the bookkeeping stores (about one store-to-L2 latency after the publish, so the round usually finds the data);
    // ASLP_ISSUE_EARLY = 1 issues it right after the publish instead (A/B variant, tools/gpu_ab_fwd.sh).
    pipe_round(it, t, ASLP_PIPE_IDLE_NS0);
    if (PR == 1 && ASLP_ISSUE_EARLY && hoisted && it + 1 < nitems) stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
    if (fin_item) {
      float* o = buf + ((size_t)t * S + s) * ldb + c0 + cl;
      o[0] = yg; o[C] = yi; o[2 * C] = yf; o[3 * C] = yo; o[4 * C] = yc; o[5 * C] = yh; o[6 * C] = ym;
    }
    RECUR_TICK(e0);
    if (PR >= 3) pipe_round(it, t, ASLP_PIPE_IDLE_NS);
    if (PR == 1 && !ASLP_ISSUE_EARLY && hoisted && it + 1 < nitems) stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
    RECUR_TICK(e1);
    if (ASLP_WARM_AHEAD > 0) warm(it + ASLP_WARM_AHEAD);
    if (PR >= 2) pipe_round(it, t, ASLP_PIPE_IDLE_NS);
    if (L.timing != nullptr && threadIdx.x == 0) { tacc[8] += e0 - k4; tacc[9] += e1 - e0; tacc[10] += clock64() - e1; }
    if (L.timing != nullptr && threadIdx.x == 0) {
      const long long k5 = clock64();
      tacc[0] += k1 - k0; tacc[1] += k2 - k1; tacc[2] += k3 - k2; tacc[3] += k4 - k3; tacc[4] += k5 - k4;
    }
  }
  if (L.timing != nullptr && threadIdx.x == 0)
    for (int q = 0; q < 12; ++q) L.timing[(size_t)blockIdx.x * 12 + q] = tacc[q];
}

// ---------------------------------------------------------------- backward (R == 0 form)
// KT_ only sizes the staging share here (the weights live in shared memory).  PS_: fast contraction form
// (mma_contract_fast, float4 fragment layout); the planner turns it on when SP == 8 and K splits exactly over the warps.
template <int MT_, int KT_, bool PS_>
__global__ void __launch_bounds__(NT, 1) lstm_bwd_mma_kernel(Launch L) {
  extern __shared__ float smem[];
  const MmaCta cta = mma_cta(L);
  const DirDev& D = L.d[cta.dir];
  const int T = D.T, S = D.S, C = D.C, SX = D.SX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SP = L.SP, SG = L.SG;
  constexpr int NPP = MT_ * 16 + 4;
  const int K = 4 * C;
  const int c0 = cta.blk * D.cb, nc = max(0, min(D.cb, C - c0));
  if (nc == 0 || cta.sbeg >= S) return;

  float* xT = smem;                                      // [4C][SP]      dgifo(successor step) of this chain's streams
  float* part = xT + (size_t)K * SP;                     // [NW][SG][NPP]
  float* st = part + (size_t)NW * SG * NPP;              // [3][cb][SGP]  d_c, d_i, d_f of the successor step
  float* pst = st + (size_t)3 * D.cb * L.SGP;            // [3][cb]
  // ---- weight slice: column c0+n of W (d_m[c] = sum_q dgifo[q] W[q][c]); row n of the slice = own cell n.
  // Stored per warp in A-fragment order [kt][mt][q][lane] (a0: row g, k tig; a1: row g+8; a2: k tig+4; a3: both)
  const int ktiles = K >> 3, ktw = (ktiles + NW - 1) / NW;
  const int ktp = (ktw + 3) & ~3;
  const int kt0 = warp * ktw, nkt = max(0, min(ktw, ktiles - kt0));
  float* wsm = pst + ((3 * D.cb + 3) & ~3);              // [NW][ktp][MT_][4][32]  (PS_: [NW][ktp][MT_][32] float4)
  float* ws = wsm + (size_t)warp * ktp * MT_ * 128;
  {
    const int g = lane >> 2, tig = lane & 3;
    for (int kt = 0; kt < ktp; ++kt)
#pragma unroll
      for (int mt = 0; mt < MT_; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = mt * 16 + g + (q & 1) * 8;
          const int k = (kt0 + kt) * 8 + tig + (q >> 1) * 4;
          const float wv = (kt < nkt && n < nc) ? D.w_r[(size_t)k * D.ldwr + c0 + n] : 0.f;
          if (PS_) {
            ws[((size_t)(kt * MT_ + mt) * 32 + lane) * 4 + q] = wv;
          } else {
            ws[((size_t)(kt * MT_ + mt) * 4 + q) * 32 + lane] = wv;
          }
        }
  }
  for (int i = threadIdx.x; i < 3 * D.cb * L.SGP; i += NT) st[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * D.cb; i += NT) {
    const int which = i / D.cb, cl = i - which * D.cb;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (cl < nc) ? p[c0 + cl] : 0.f;
  }
  for (int i = threadIdx.x; i < K * SP; i += NT) xT[i] = 0.f;
  __syncthreads();

  const int s_local = threadIdx.x % SG, cl = threadIdx.x / SG;
  const bool owner = cl < nc;
  const int cc = c0 + (owner ? cl : 0);
  const int ngroups = (cta.send - cta.sbeg + SG - 1) / SG;
  const int nitems = T * ngroups;
  const int plane = D.cb * L.SGP;
  const int reverse = D.reverse, ldb = D.ldb, lddb = D.lddb;
  float* const buf = D.buf;
  float* const dbuf = D.dbuf;
  float* const xa = D.xa;
  const float pi = owner ? pst[cl] : 0.f, pf = owner ? pst[D.cb + cl] : 0.f, po = owner ? pst[2 * D.cb + cl] : 0.f;
  const int sg4_0 = (min(cta.sbeg + SG, SX) - cta.sbeg) >> 2;
  constexpr int SB = (KT_ + 1) / 2;
  const StageDesc sd = stage_prepare<SB>(K, SX, SP, cta.sbeg, sg4_0);
  const bool hoisted = (ngroups == 1) && sd.fast;
  const size_t slot = (size_t)K * SX;
  // backward visits time in the opposite order of the forward pass of this direction
  auto item_row = [&](int it, int& t, int& s0) {
    if (ngroups == 1) { t = reverse ? 1 + it : T - it; s0 = cta.sbeg; return; }
    const int step = it / ngroups, grp = it - step * ngroups;
    t = reverse ? 1 + step : T - step;
    s0 = cta.sbeg + grp * SG;
  };
  // forward activations of a future item: into L2 two items ahead, into registers after the exchange wait of their step
  auto warm = [&](int it) {
    if (!owner || it >= nitems) return;
    int t, s0; item_row(it, t, s0);
    const int s = s0 + s_local;
    if (s >= cta.send) return;
    const int tn = reverse ? t - 1 : t + 1, tp = reverse ? t + 1 : t - 1;
    const float* y = buf + ((size_t)t * S + s) * ldb + cc;
    prefetch_l2(y); prefetch_l2(y + C); prefetch_l2(y + 2 * C); prefetch_l2(y + 3 * C); prefetch_l2(y + 5 * C);
    prefetch_l2(buf + ((size_t)tp * S + s) * ldb + 4 * C + cc);
    prefetch_l2(buf + ((size_t)tn * S + s) * ldb + 2 * C + cc);
    prefetch_l2(dbuf + ((size_t)t * S + s) * lddb + 6 * C + cc);
  };
  if (ASLP_WARM_AHEAD > 0) { for (int w = 0; w < ASLP_WARM_AHEAD; ++w) warm(w); }
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (hoisted) stage_issue<SB>(sd, xT, xa + (size_t)(reverse ? 0 : T + 1) * slot);

  for (int it = 0; it < nitems; ++it) {
    int t, s0; item_row(it, t, s0);
    const int tn = reverse ? t - 1 : t + 1, tp = reverse ? t + 1 : t - 1;
    const int nstr = min(SG, cta.send - s0);
    const int s = s0 + s_local;
    RECUR_TICK(k0);
    long long* dbg = (L.timing != nullptr && threadIdx.x == 0) ? &tacc[5] : nullptr;
    if (hoisted) stage_complete<SB>(sd, xT, xa + (size_t)tn * slot, dbg);
    else stage_poll<20>(xT, SP, xa + (size_t)tn * slot, K, SX, s0, (min(s0 + SG, SX) - s0) >> 2, dbg);
    float yv[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // g, i, f, o, h at t ; c at the forward predecessor ; f at the successor
    float od = 0.f;                                      // out_diff share of d_m (preloaded in the m columns of dbuf)
    if (owner && s_local < nstr) {
      const float* y = buf + ((size_t)t * S + s) * ldb + cc;
      yv[0] = y[0]; yv[1] = y[C]; yv[2] = y[2 * C]; yv[3] = y[3 * C]; yv[4] = y[5 * C];
      yv[5] = buf[((size_t)tp * S + s) * ldb + 4 * C + cc];
      yv[6] = buf[((size_t)tn * S + s) * ldb + 2 * C + cc];
      od = dbuf[((size_t)t * S + s) * lddb + 6 * C + cc];
    }
    RECUR_TICK(k1);
    __syncthreads();
    RECUR_TICK(k2);
    constexpr int KTP = KT_ <= 32 ? ((KT_ + 3) & ~3) : 0;     // exact tile count known at compile time up to 32 tiles per warp
    if (PS_) {
      constexpr int KTPS = KTP > 0 ? KTP : 4;
      mma_contract_fast<MT_, KTPS>(reinterpret_cast<const float4*>(ws), xT + (size_t)(kt0 * 8 + (lane & 3)) * 8 + (lane >> 2),
                                   part + (size_t)warp * SG * NPP, lane);
    } else
    if (KTP > 0 && ktp == KTP) mma_contract_ws<MT_, KTP>(ws, ktp, xT, SP, kt0 * 8, K, (nstr + 7) >> 3, part + (size_t)warp * SG * NPP, lane);
    else mma_contract_ws<MT_, 0>(ws, ktp, xT, SP, kt0 * 8, K, (nstr + 7) >> 3, part + (size_t)warp * SG * NPP, lane);
    RECUR_TICK(k3);
    __syncthreads();
    RECUR_TICK(k4);
    if (owner && s_local < nstr) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) sum += part[((size_t)w * SG + s_local) * NPP + cl];
      const size_t row = (size_t)t * S + s;
      const int si = cl * L.SGP + (s - cta.sbeg);
      const float yg = yv[0], yi = yv[1], yf = yv[2], yo = yv[3], yh = yv[4], c_prev = yv[5], yf_next = yv[6];
      const float dm = sum + od;
      const float dc_n = st[si], di_n = st[plane + si], df_n = st[2 * plane + si];
      float dh = dm * yo;  dh = (1.0f - yh * yh) * dh;                 // DiffTanh(y_h, d_h)
      float dout = dm * yh;  dout = yo * (1.0f - yo) * dout;           // DiffSigmoid(y_o, d_o)
      const float dc = dh + dc_n * yf_next + di_n * pi + df_n * pf + dout * po;
      float df = dc * c_prev;  df = yf * (1.0f - yf) * df;
      float di = dc * yg;      di = yi * (1.0f - yi) * di;
      float dg = dc * yi;      dg = (1.0f - yg * yg) * dg;
      float* x = xa + ((size_t)t * K + cc) * SX + s;                   // publish dgifo(t) first
      st_pub(x, dg); st_pub(x + (size_t)C * SX, di); st_pub(x + (size_t)2 * C * SX, df); st_pub(x + (size_t)3 * C * SX, dout);
      if (L.timing != nullptr && threadIdx.x == 0) tacc[7] += clock64() - k4;
      if (ASLP_ISSUE_EARLY && hoisted && it + 1 < nitems) stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
      float* d = dbuf + row * lddb + cc;
      d[0] = dg; d[C] = di; d[2 * C] = df; d[3 * C] = dout; d[4 * C] = dc; d[5 * C] = dh; d[6 * C] = dm;
      st[si] = dc; st[plane + si] = di; st[2 * plane + si] = df;
    } else if (ASLP_ISSUE_EARLY && hoisted && it + 1 < nitems) {
      stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
    }
    RECUR_TICK(e0);
    if (!ASLP_ISSUE_EARLY && hoisted && it + 1 < nitems) stage_issue<SB>(sd, xT, xa + (size_t)t * slot);
    RECUR_TICK(e1);
    if (ASLP_WARM_AHEAD > 0) warm(it + ASLP_WARM_AHEAD);
    if (L.timing != nullptr && threadIdx.x == 0) { tacc[8] += e0 - k4; tacc[9] += e1 - e0; tacc[10] += clock64() - e1; }
    if (L.timing != nullptr && threadIdx.x == 0) {
      const long long k5 = clock64();
      tacc[0] += k1 - k0; tacc[1] += k2 - k1; tacc[2] += k3 - k2; tacc[3] += k4 - k3; tacc[4] += k5 - k4;
    }
  }
  if (L.timing != nullptr && threadIdx.x == 0)
    for (int q = 0; q < 12; ++q) L.timing[(size_t)blockIdx.x * 12 + q] = tacc[q];
}

// ---------------------------------------------------------------- planning + dispatch
struct MmaChoice { int mt, kt; };

inline int mma_rows(int cb, bool bwd) { return bwd ? cb : 4 * cb; }          // rows of the per-CTA weight slice
inline int mma_kdim(int C, bool bwd) { return bwd ? 4 * C : C; }
// fast contraction form applies: one staging group of 8 streams (SP == 8), K splits exactly into a multiple of 4 k-tiles per
// warp, and a kernel with that compile-time tile count exists
inline bool mma_presplit_shape(int C, int SG, int SGP) {
  const int kt = (4 * C) / 8;
  return SG == 8 && SGP == 8 && (4 * C) % (8 * NW) == 0 && (kt / NW) % 4 == 0 && (kt / NW == 8 || kt / NW == 20 || kt / NW == 32);
}
inline size_t mma_smem_floats(int C, int cb, int SG, int SGP, bool bwd, bool presplit = false) {
  const int SP = SG == 8 ? 8 : SG + 8;
  const size_t mt = (size_t)(mma_rows(cb, bwd) + 15) / 16;
  size_t fl = (size_t)mma_kdim(C, bwd) * SP + (size_t)NW * SG * (mt * 16 + 4) + (size_t)(bwd ? 3 : 1) * cb * SGP + 3 * (size_t)cb;
  if (bwd) {                                             // weight fragments in shared memory
    const size_t ktw = ((size_t)mma_kdim(C, bwd) / 8 + NW - 1) / NW, ktp = (ktw + 3) & ~(size_t)3;
    fl += 4 + (size_t)NW * ktp * mt * 128;
    (void)presplit;
  }
  return fl;
}
// register budget: MT*KT A fragments of 4 fp32 each
inline bool mma_fits(int C, int cb, bool bwd, MmaChoice* ch) {
  const int mt = (mma_rows(cb, bwd) + 15) / 16;
  const int kt = ((mma_kdim(C, bwd) / 8) + NW - 1) / NW;
  if (bwd) { if (!(mt <= 2 && kt <= 64)) return false; }   // weights in shared memory: only the capacity check of the planner limits K
  else     { if (!(mt <= 4 && kt <= 16 && mt * (kt <= 2 ? 2 : kt <= 5 ? 5 : kt <= 8 ? 8 : kt <= 12 ? 12 : 16) <= 30)) return false; }
  ch->mt = mt; ch->kt = kt;
  return true;
}

template <int MT_, int KT_> inline void* mma_fwd_ptr() { return (void*)lstm_fwd_mma_kernel<MT_, KT_>; }
template <int MT_, int KT_, bool PS_> inline void* mma_bwd_ptr() { return (void*)lstm_bwd_mma_kernel<MT_, KT_, PS_>; }

inline void* mma_pick_kernel(const MmaChoice& c, bool bwd, bool presplit = false) {
  if (bwd) {                                             // KT_ = staging share bound (items per thread = KT_/2 at 8 streams)
    if (presplit) {
      if (c.mt == 1) { if (c.kt == 8) return mma_bwd_ptr<1, 8, true>(); if (c.kt == 20) return mma_bwd_ptr<1, 20, true>(); if (c.kt == 32) return mma_bwd_ptr<1, 32, true>(); }
      else { if (c.kt == 8) return mma_bwd_ptr<2, 8, true>(); if (c.kt == 20) return mma_bwd_ptr<2, 20, true>(); if (c.kt == 32) return mma_bwd_ptr<2, 32, true>(); }
      return nullptr;
    }
    if (c.mt == 1) { if (c.kt <= 8) return mma_bwd_ptr<1, 8, false>(); if (c.kt <= 20) return mma_bwd_ptr<1, 20, false>(); if (c.kt <= 32) return mma_bwd_ptr<1, 32, false>(); return mma_bwd_ptr<1, 64, false>(); }
    if (c.kt <= 8) return mma_bwd_ptr<2, 8, false>(); if (c.kt <= 20) return mma_bwd_ptr<2, 20, false>(); if (c.kt <= 32) return mma_bwd_ptr<2, 32, false>(); return mma_bwd_ptr<2, 64, false>();
  }
#define ASLP_FWD_ROW(MTv)                                        \
  if (c.mt == MTv) {                                             \
    if (c.kt <= 2) return mma_fwd_ptr<MTv, 2>();                 \
    if (c.kt <= 5) return mma_fwd_ptr<MTv, 5>();                 \
    if (MTv * 8 <= 30 && c.kt <= 8) return mma_fwd_ptr<MTv, (MTv * 8 <= 30 ? 8 : 2)>();     \
    if (MTv * 12 <= 30 && c.kt <= 12) return mma_fwd_ptr<MTv, (MTv * 12 <= 30 ? 12 : 2)>();  \
    if (MTv * 16 <= 30 && c.kt <= 16) return mma_fwd_ptr<MTv, (MTv * 16 <= 30 ? 16 : 2)>();  \
  }
  ASLP_FWD_ROW(1) ASLP_FWD_ROW(2) ASLP_FWD_ROW(3) ASLP_FWD_ROW(4)
#undef ASLP_FWD_ROW
  return nullptr;
}
