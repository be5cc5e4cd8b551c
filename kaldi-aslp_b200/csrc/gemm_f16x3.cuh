// gemm_f16x3.cuh -- the fp32-grade chunk GEMM as THREE kind::f16 tensor-core passes over operands that were split ONCE, ahead of the
// main loop, into fp16 hi / lo planes in HBM (included by gemm.cu inside its anonymous namespace).
//
// Why: the 3xTF32 kernel splits both operand tiles in shared memory at every pipeline stage (4 splitter warps: 32 KB read + 32 KB
// written per k-block) and its three TF32 passes read 96 KB of tiles per k-block; it is bound by shared-memory traffic at ~1630
// cycles per k-block of 32 (187 TFLOP/s useful), while the single-pass TF32 kernel runs at the L2 -> SM fill rate (~800 cycles,
// 440 TFLOP/s) on the same shapes (profiles/r02_gemm_ab.jsonl).  With the split done by a separate HBM-bound pass
//   x * 2^e = h1 + h2 (+ O(2^-22)),  h1 = fp16(x 2^e), h2 = fp16(x 2^e - h1),  2^e = the power of two that brings the largest magnitude
//                                                                             of the operand's row (along K) into [2^14, 2^15)
// the main loop has no splitter, TMA brings the four planes (same bytes per k as the fp32 tiles), the three passes
// lo*hi + hi*lo + hi*hi are kind::f16 MMAs (K = 16 per instruction at the rate TF32 does K = 8, half the tile bytes read per pass),
// and the epilogue multiplies by 2^-(e_row + e_col), exact.  Products of fp16 values are exact in the fp32 accumulator, the dropped
// lo*lo term is O(2^-22): the same class of error as 3xTF32 (tests/test_gpu_gemm.py holds both to 2e-5 of max |C|).
// Every operand becomes K-major in its plane (the transposing split handles the MN-major cases), so ONE kernel serves all four
// transpose combinations.  Scaling per row of op(A) and per column of op(B) (i.e. along K) keeps >= 22 significant bits for every
// element within 2^-18 of its row's maximum and degrades gracefully below (absolute error <= 2^-40 of the row maximum).
#pragma once
#include <cuda_fp16.h>

constexpr int HK = 64;                         // halves per 128-byte swizzled row = K per pipeline stage
constexpr int HTILE = BM * HK * 2;             // 16 KB: one 128 x 64 fp16 tile
constexpr int HSTAGES = 3;
constexpr int HSTAGE_BYTES = 4 * HTILE;        // A_hi, A_lo, B_hi, B_lo

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// 2^e with mx * 2^e in [2^14, 2^15); inv = 2^-e.  Zero / denormal / non-finite rows are left unscaled.
__device__ __forceinline__ float split_scale(float mx, float& inv) {
  const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu);
  if (e == 0 || e == 255) { inv = 1.0f; return 1.0f; }
  int se = 127 + 14 - (e - 127);
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  inv = __uint_as_float((unsigned)(254 - se) << 23);
  return __uint_as_float((unsigned)se << 23);
}
__device__ __forceinline__ void split_store4(__half* hi, __half* lo, float4 x, float s) {
  const float a0 = x.x * s, a1 = x.y * s, a2 = x.z * s, a3 = x.w * s;
  const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1), h2 = __float2half_rn(a2), h3 = __float2half_rn(a3);
  const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
  const __half l2 = __float2half_rn(a2 - __half2float(h2)), l3 = __float2half_rn(a3 - __half2float(h3));
  __half2 hp0 = __halves2half2(h0, h1), hp1 = __halves2half2(h2, h3), lp0 = __halves2half2(l0, l1), lp1 = __halves2half2(l2, l3);
  *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<uint32_t*>(&hp0), *reinterpret_cast<uint32_t*>(&hp1));
  *reinterpret_cast<uint2*>(lo) = make_uint2(*reinterpret_cast<uint32_t*>(&lp0), *reinterpret_cast<uint32_t*>(&lp1));
}

// K-major source: X [rows][cols] (ld floats, 16-byte aligned rows) -> hi / lo [rows][ldp] (ldp % 8 == 0), inv[rows].  Warp per row;
// the row is read twice (the second pass comes from L1 / L2): 4 B read + 4 B written per element from HBM.
__global__ void __launch_bounds__(256) presplit_rows_kernel(const float* __restrict__ X, int ld, int rows, int cols, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, int ldp, float* __restrict__ inv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c4 = cols >> 2;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float* x = X + (size_t)r * ld;
    float mx = 0.f;
    for (int c = lane; c < c4; c += 32) {
      const float4 v = *reinterpret_cast<const float4*>(x + 4 * c);
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int c = 4 * c4 + lane; c < cols; c += 32) mx = fmaxf(mx, fabsf(x[c]));
    mx = warp_max(mx);
    float iv;
    const float s = split_scale(mx, iv);
    if (lane == 0) inv[r] = iv;
    __half* h = hi + (size_t)r * ldp;
    __half* l = lo + (size_t)r * ldp;
    for (int c = lane; c < c4; c += 32) split_store4(h + 4 * c, l + 4 * c, *reinterpret_cast<const float4*>(x + 4 * c), s);
    for (int c = 4 * c4 + lane; c < cols; c += 32) {
      const float a = x[c] * s;
      const __half hh = __float2half_rn(a);
      h[c] = hh;
      l[c] = __float2half_rn(a - __half2float(hh));
    }
  }
}

// MN-major source: X [krows][mcols] (the operand's K runs down the rows) -> planes [mcols][ldp] with K contiguous.
// pass 1: column maxima (as float bit patterns: non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) presplit_colmax_kernel(const float* __restrict__ X, int ld, int krows, int mcols, unsigned* __restrict__ cmax) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  float mx = 0.f;
  if (c < mcols)
    for (int r = blockIdx.y * 8 + (threadIdx.x >> 5); r < krows; r += gridDim.y * 8) mx = fmaxf(mx, fabsf(X[(size_t)r * ld + c]));
  __shared__ float sh[8][32];
  sh[threadIdx.x >> 5][threadIdx.x & 31] = mx;
  __syncthreads();
  if (threadIdx.x < 32 && c < mcols) {
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, sh[i][threadIdx.x]);
    atomicMax(cmax + c, __float_as_uint(mx));
  }
}
// pass 2: 64 (k) x 32 (m) tiles through shared memory: coalesced 128-byte reads along m, 128-byte writes along k
__global__ void __launch_bounds__(256) presplit_transpose_kernel(const float* __restrict__ X, int ld, int krows, int mcols, const unsigned* __restrict__ cmax,
                                                                 __half* __restrict__ hi, __half* __restrict__ lo, int ldp, float* __restrict__ inv) {
  __shared__ float tile[64][33];
  const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + ty + 8 * i, m = m0 + tx;
    tile[ty + 8 * i][tx] = (k < krows && m < mcols) ? X[(size_t)k * ld + m] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ml = ty + 8 * i, m = m0 + ml;
    if (m >= mcols) continue;
    float iv;
    const float s = split_scale(__uint_as_float(cmax[m]), iv);
    if (blockIdx.y == 0 && tx == 0) inv[m] = iv;
    const int k = k0 + 2 * tx;
    if (k < krows) {                              // ldp is even and >= krows rounded up to 8: k + 1 stays inside the row's pitch
      const float a0 = tile[2 * tx][ml] * s, a1 = tile[2 * tx + 1][ml] * s;
      const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
      *reinterpret_cast<__half2*>(hi + (size_t)m * ldp + k) = __halves2half2(h0, h1);
      *reinterpret_cast<__half2*>(lo + (size_t)m * ldp + k) = __halves2half2(__float2half_rn(a0 - __half2float(h0)), __float2half_rn(a1 - __half2float(h1)));
    }
  }
}

// ------------------------------------------------------------------ the kernel: C = alpha * (A B^T) * inv_a[m] * inv_b[n] + beta C (+ bias, clip)
//   warp 0 : TMA producer, four 16 KB tiles per stage      warp 1 : TMEM allocator + MMA issuer      warps 2..5 : epilogue
__global__ void __launch_bounds__(192, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBh,
                  const __grid_constant__ CUtensorMap tmBl, EpiParams p, const float* __restrict__ inv_a, const float* __restrict__ inv_b) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + HSTAGES * HSTAGE_BYTES;
  auto full_bar  = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (HSTAGES + s); };
  auto tmem_full_bar  = [&](int a) { return bar_base + 8u * (2 * HSTAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * HSTAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + HSTAGES * HSTAGE_BYTES + 8 * (2 * HSTAGES + 4));
  float* stage_base = reinterpret_cast<float*>(smem_gen + HSTAGES * HSTAGE_BYTES + BAR_REGION);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + HK - 1) / HK;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int num_items = tiles_mn * p.splits;
  auto item_coords = [&](int item, int& z, int& m0, int& n0, int& kb_begin, int& kb_end) {
    z = item / tiles_mn;
    const int t = item - z * tiles_mn;
    const int tm = t / p.tiles_n;
    m0 = tm * BM; n0 = (t - tm * p.tiles_n) * BN;
    kb_begin = z * p.kb_per_split;
    kb_end = min(num_kb, kb_begin + p.kb_per_split);
  };

  if (threadIdx.x == 64) {      // descriptor fetch under the barrier / TMEM set-up instead of in front of the first load
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < HSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int z, m0, n0, kb_begin, kb_end;
      item_coords(item, z, m0, n0, kb_begin, kb_end);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = smem_base + s * HSTAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), HSTAGE_BYTES);
          tma_load_2d(sa, &tmAh, full_bar(s), kb * HK, m0);
          tma_load_2d(sa + HTILE, &tmAl, full_bar(s), kb * HK, m0);
          tma_load_2d(sa + 2 * HTILE, &tmBh, full_bar(s), kb * HK, n0);
          tma_load_2d(sa + 3 * HTILE, &tmBl, full_bar(s), kb * HK, n0);
        }
        __syncwarp();
        if (++s == HSTAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int s = 0; uint32_t ph = 0;
    int acc_idx = 0; uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int z, m0, n0, kb_begin, kb_end;
      item_coords(item, z, m0, n0, kb_begin, kb_end);
      mbar_wait(tmem_empty_bar(acc_idx), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc_idx * BN);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * HSTAGE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HK / 16; ++k) {                 // one MMA = 16 halves = 32 bytes along the swizzled row
            const uint64_t ah = make_desc(sa + k * 32, 16u, 1024u, 2u), al = make_desc(sa + HTILE + k * 32, 16u, 1024u, 2u);
            const uint64_t bh = make_desc(sa + 2 * HTILE + k * 32, 16u, 1024u, 2u), bl = make_desc(sa + 3 * HTILE + k * 32, 16u, 1024u, 2u);
            const uint32_t acc = (kb > kb_begin || k > 0) ? 1u : 0u;
            tc_mma_f16(tmem_d, al, bh, idesc, acc);          // small terms first
            tc_mma_f16(tmem_d, ah, bl, idesc, 1u);
            tc_mma_f16(tmem_d, ah, bh, idesc, 1u);
          }
          tc_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == HSTAGES) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) tc_commit(tmem_full_bar(acc_idx));
      __syncwarp();
      if (++acc_idx == 2) { acc_idx = 0; acc_ph ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    int acc_idx = 0; uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int z, m0, n0, kb_begin, kb_end;
      item_coords(item, z, m0, n0, kb_begin, kb_end);
      mbar_wait(tmem_full_bar(acc_idx), acc_ph);
      tc_fence_after();
      const bool has_work = kb_end > kb_begin;
      float* stg = stage_base + (size_t)q * 32 * STG_PITCH;
      const int sub_r = lane >> 3, sub_c = (lane & 7) << 2;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_idx * BN + c * 32), v);
        const int nb = n0 + c * 32;
        if (nb < p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          __syncwarp();
          const int col = nb + sub_c;
          const int nvalid = p.N - col;
          float cs[4] = {0.f, 0.f, 0.f, 0.f};                // 2^-e of the four columns of this lane
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) if (jj < nvalid) cs[jj] = inv_b[col + jj];
          float4 oldc[8];                                    // beta operands of the chunk, all in flight together (see gemm.cu)
          const bool pre_beta = p.partial == nullptr && p.beta != 0.f && nvalid >= 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = m0 + q * 32 + it * 4 + sub_r;
            oldc[it] = (pre_beta && row < p.M) ? *reinterpret_cast<const float4*>(p.C + (size_t)row * p.ldc + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + sub_r;
            const int row = m0 + q * 32 + rl;
            if (row < p.M && nvalid > 0) {
              float4 o = *reinterpret_cast<const float4*>(stg + rl * STG_PITCH + sub_c);
              const float rs = inv_a[row];                   // powers of two: the un-scaling is exact
              o.x *= rs * cs[0]; o.y *= rs * cs[1]; o.z *= rs * cs[2]; o.w *= rs * cs[3];
              if (p.partial != nullptr) {
                if (!has_work) o = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(p.partial + ((size_t)z * p.M + row) * p.ldp + col) = o;
              } else {
                float* dst = p.C + (size_t)row * p.ldc + col;
                o.x *= p.alpha; o.y *= p.alpha; o.z *= p.alpha; o.w *= p.alpha;
                if (nvalid >= 4) {
                  if (p.beta != 0.f) {
                    const float4 old = oldc[it];
                    o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                  }
                  if (p.bias != nullptr) {
                    const float4 b = *reinterpret_cast<const float4*>(p.bias + col);
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                  }
                  if (p.clip > 0.f) {
                    o.x = fminf(fmaxf(o.x, -p.clip), p.clip); o.y = fminf(fmaxf(o.y, -p.clip), p.clip);
                    o.z = fminf(fmaxf(o.z, -p.clip), p.clip); o.w = fminf(fmaxf(o.w, -p.clip), p.clip);
                  }
                  *reinterpret_cast<float4*>(dst) = o;
                } else {
                  const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int jj = 0; jj < 3; ++jj) {
                    if (jj < nvalid) {
                      float x = ov[jj];
                      if (p.beta != 0.f) x += p.beta * dst[jj];
                      if (p.bias != nullptr) x += p.bias[col + jj];
                      if (p.clip > 0.f) x = fminf(fmaxf(x, -p.clip), p.clip);
                      dst[jj] = x;
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(acc_idx));
      if (++acc_idx == 2) { acc_idx = 0; acc_ph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem_base) : "memory");
  }
}
