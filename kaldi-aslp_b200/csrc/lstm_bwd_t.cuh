// lstm_bwd_t.cuh -- "transposed" backward recurrence (included by lstm.cu after lstm_mma.cuh; R == 0 / folded form).
//
// lstm_bwd_mma_kernel makes every CTA of a chain gather ALL 4C rows of dgifo(t+1) (40 KB at cfg3) and contract them with
// its own 16 weight columns: 80 CTAs x 40 KB = 3.2 MB cross L2 per step, ~500 cycles at the L2-slice cap on top of the
// round trip, and the gather needs 10 copy instructions per thread.  Here the contraction is turned around: a CTA
// contracts its OWN 64 dgifo rows (4 gates x 16 cells, already in its shared memory) against the matching 64 ROWS of W'
// for ALL C columns,
//     z_j[c] = sum_{q in rows of CTA j} dgifo[q] W'[q][c]            (M = C: one 16-row MMA tile per consumer CTA, K = 64)
// and publishes the C partial sums straight from the MMA accumulators; the owner of cell block b gathers the 20 partials
// z_j[16 cells of b] (one contiguous 10 KB block, 2.5 copies per thread) and adds them in a fixed order.  Per step the
// chain moves 0.8 MB read + 0.8 MB written instead of 3.2 + 0.16, the K split across warps and its shared-memory partial
// reduction disappear (a warp owns whole m-tiles), and the B operand is 2 KB instead of 40.
// Every partial block has ONE producer and ONE consumer, so the exchange is a ring of TR time slots that the consumer
// re-arms with the sentinel after reading (a fence orders the re-arm before the consumer's next publish; a producer
// overwrites a slot only TR-1 steps later, after it has consumed a publish of that consumer).  Same data-is-its-own-flag
// protocol (NaN sentinel), same cell derivative chain and buffer layout as lstm_bwd_mma_kernel.
#pragma once

constexpr int TR = 4;                    // ring slots of the partial-sum exchange
#ifndef ASLP_BWD_T_DELAY_NS
#define ASLP_BWD_T_DELAY_NS 0
#endif
#ifndef ASLP_PIPE_BWD_GAP_NS
#define ASLP_PIPE_BWD_GAP_NS 0   // extra spacing before the last gather round of a step (pipelined polling)
#endif
#ifndef ASLP_BWD_T_FP16
#define ASLP_BWD_T_FP16 1        // 1: fp16 two-term split with per-stream power-of-two scaling of dgifo (8-stream form); 0: 3xTF32
#endif

__device__ __forceinline__ void st_pub2(float* p, float a, float b) {
  asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}

// shared memory: wT [nblk][8][32] float4 | Bsm [64][8] | Psm [nblk][16][8] | st [3][16][8] | pst [3][16]
inline size_t bwd_t_smem_floats(int nblk) { return (size_t)nblk * 1024 + 512 + (size_t)nblk * 128 + 384 + 48 + 32; }   // + column maxima [4][8]
// exchange floats per direction: [TR][pgroups][consumer][producer][16][8]
inline size_t bwd_t_exchange_floats(int nblk, int pgroups) { return (size_t)TR * pgroups * nblk * nblk * 128; }

static __global__ void xch_t_init_kernel(float* x, size_t per_slot, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    x[i] = i < per_slot ? 0.f : __uint_as_float(SENTINEL);       // slot 0 = the zero derivative boundary, the rest armed
}

// z[c][s] = sum_k W'[own row k][c] * dgifo[k][s] for the NM m-tiles (16 columns c each) warp `warp` owns: warp, warp+8, ...
// NM is a template parameter so that the loop has no branch: with a run-time "is this tile real" test inside, every
// A-fragment load sat behind its branch right in front of the MMAs that use it (40 cycles per MMA instead of ~10).
template <int NM>
__device__ __forceinline__ void bwd_t_contract(const float4* wT, const float* Bsm, float* out, int nblk, int warp, int lane) {
  const int g = lane >> 2, tig = lane & 3;
  float hh[NM][4], lh[NM][4], hl[NM][4];
#pragma unroll
  for (int mi = 0; mi < NM; ++mi)
#pragma unroll
    for (int q = 0; q < 4; ++q) { hh[mi][q] = 0.f; lh[mi][q] = 0.f; hl[mi][q] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < 8; ++kt) {
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(Bsm[(kt * 8 + tig) * 8 + g], bh0, bl0);
    split_tf32(Bsm[(kt * 8 + tig + 4) * 8 + g], bh1, bl1);
#pragma unroll
    for (int mi = 0; mi < NM; ++mi) {
      const float4 w4 = wT[((warp + 8 * mi) * 8 + kt) * 32 + lane];
      uint32_t ah[4], al[4];
      split_tf32(w4.x, ah[0], al[0]); split_tf32(w4.y, ah[1], al[1]); split_tf32(w4.z, ah[2], al[2]); split_tf32(w4.w, ah[3], al[3]);
      mma_tf32(lh[mi], al, bh0, bh1);
      mma_tf32(hl[mi], ah, bl0, bl1);
      mma_tf32(hh[mi], ah, bh0, bh1);
    }
  }
#pragma unroll
  for (int mi = 0; mi < NM; ++mi) {
    float* dst = out + (size_t)(warp + 8 * mi) * nblk * 128;           // block of consumer CTA warp + 8 mi
    // c0: (cell g, stream 2tig)  c1: (g, 2tig+1)  c2: (g+8, 2tig)  c3: (g+8, 2tig+1); small terms first
    st_pub2(dst + g * 8 + 2 * tig, (lh[mi][0] + hl[mi][0]) + hh[mi][0], (lh[mi][1] + hl[mi][1]) + hh[mi][1]);
    st_pub2(dst + (g + 8) * 8 + 2 * tig, (lh[mi][2] + hl[mi][2]) + hh[mi][2], (lh[mi][3] + hl[mi][3]) + hh[mi][3]);
  }
}

// fp16 form of the contraction above.  dgifo has no lower bound, so every stream column of the B operand is first scaled
// by a power of two that brings its largest magnitude (over the CTA's 64 rows; maxima left in cmx[4][8] by the finishing
// warps) to [2^14, 2^15): exact, no fp16 overflow, and an element keeps >= 22 significant bits unless it is more than 2^24
// below its column's maximum, where its contribution to the partial sum is below fp32 resolution of the dominant terms
// anyway.  The accumulators are scaled back per stream column (exact) before they are published.  Half the MMAs of 3xTF32.
// A fragments of W'[own 64 rows][all C columns] for m16n8k16, pre-split into fp16 halves: hi [nblk][4][32] uint4, then lo
// (same bytes as the tf32 layout); a0: row g, k 2tig..+1; a1: row g+8; a2: row g, k 2tig+8..+9; a3: row g+8;
// row = cell column c of W' (m-tile mt covers cells mt*16 .. +15), k = gate*16 + own cell
__device__ __forceinline__ void bwd_t_fill_h(const DirDev& D, uint4* wTh, uint4* wTl, int nblk, int c0) {
  const int C = D.C;
  for (int i = threadIdx.x; i < nblk * 128; i += NT) {
    const int mt = i >> 7, kh = (i >> 5) & 3, ln = i & 31, gg = ln >> 2, tt = ln & 3;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = mt * 16 + gg + (q & 1) * 8;
      float w[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k = kh * 16 + 2 * tt + (q >> 1) * 8 + u, gate = k >> 4, cl = k & 15;
        w[u] = (c < C && c0 + cl < C) ? D.w_r[(size_t)(gate * C + c0 + cl) * D.ldwr + c] : 0.f;
        if (fabsf(w[u]) > 32768.f) __trap();             // outside the fp16 split's range: an error, never a silent inf
      }
      split_h2(w[0], w[1], hi[q], lo[q]);
    }
    wTh[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    wTl[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
__device__ __forceinline__ float pow2_scale(float mx, float& inv) {
  const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu);
  if (e == 0 || e == 255) { inv = 1.0f; return 1.0f; }     // zero / denormal column (or inf: nothing to save)
  int se = 127 + 14 - (e - 127);
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  inv = __uint_as_float((unsigned)(254 - se) << 23);
  return __uint_as_float((unsigned)se << 23);
}
template <int NM>
__device__ __forceinline__ void bwd_t_contract_h(const uint4* wh, const uint4* wl, const float* Bsm, const float* cmx, float* out, int nblk,
                                                 int warp, int lane) {
  const int g = lane >> 2, tig = lane & 3;
  auto colmax = [&](int s) { return fmaxf(fmaxf(cmx[s], cmx[8 + s]), fmaxf(cmx[16 + s], cmx[24 + s])); };
  float inv_g, inv0, inv1;
  const float sc_g = pow2_scale(colmax(g), inv_g);
  pow2_scale(colmax(2 * tig), inv0);
  pow2_scale(colmax(2 * tig + 1), inv1);
  float hh[NM][4], lh[NM][4], hl[NM][4];
#pragma unroll
  for (int mi = 0; mi < NM; ++mi)
#pragma unroll
    for (int q = 0; q < 4; ++q) { hh[mi][q] = 0.f; lh[mi][q] = 0.f; hl[mi][q] = 0.f; }
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const float* col = Bsm + (kh * 16 + 2 * tig) * 8 + g;
    uint32_t bh0, bl0, bh1, bl1;
    split_h2(col[0] * sc_g, col[8] * sc_g, bh0, bl0);            // k = 2tig, 2tig+1
    split_h2(col[64] * sc_g, col[72] * sc_g, bh1, bl1);          // k = 2tig+8, 2tig+9
#pragma unroll
    for (int mi = 0; mi < NM; ++mi) {
      const uint4 h4 = wh[((warp + 8 * mi) * 4 + kh) * 32 + lane], l4 = wl[((warp + 8 * mi) * 4 + kh) * 32 + lane];
      const uint32_t ah[4] = {h4.x, h4.y, h4.z, h4.w}, al[4] = {l4.x, l4.y, l4.z, l4.w};
      mma_f16(lh[mi], al, bh0, bh1);
      mma_f16(hl[mi], ah, bl0, bl1);
      mma_f16(hh[mi], ah, bh0, bh1);
    }
  }
#pragma unroll
  for (int mi = 0; mi < NM; ++mi) {
    float* dst = out + (size_t)(warp + 8 * mi) * nblk * 128;
    st_pub2(dst + g * 8 + 2 * tig, ((lh[mi][0] + hl[mi][0]) + hh[mi][0]) * inv0, ((lh[mi][1] + hl[mi][1]) + hh[mi][1]) * inv1);
    st_pub2(dst + (g + 8) * 8 + 2 * tig, ((lh[mi][2] + hl[mi][2]) + hh[mi][2]) * inv0, ((lh[mi][3] + hl[mi][3]) + hh[mi][3]) * inv1);
  }
  (void)inv_g;
}

__global__ void __launch_bounds__(NT, 1) lstm_bwd_t_kernel(Launch L) {
  extern __shared__ __align__(16) float smem[];
  const MmaCta cta = mma_cta(L);
  const DirDev& D = L.d[cta.dir];
  const int T = D.T, S = D.S, C = D.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = L.nblk;                               // == C / 16: m-tile index == consumer CTA index
  const int c0 = cta.blk * 16;
  if (cta.sbeg >= S) return;                             // a chain without streams: all of its CTAs leave together

#if !ASLP_BWD_T_FP16
  float4* wT = reinterpret_cast<float4*>(smem);          // A fragments of W'[own 64 rows][all C columns], transposed
#endif
  float* Bsm = smem + (size_t)nblk * 1024;               // dgifo of the step just finished: [gate*16 + cell][stream]
  float* Psm = Bsm + 512;                                // gathered partial sums [producer][cell][stream]
  float* st = Psm + (size_t)nblk * 128;                  // d_c, d_i, d_f of the successor step [3][cell][stream]
  float* pst = st + 384;                                 // peepholes [3][cell]
  float* cmx = pst + 48;                                 // per-warp column maxima of |dgifo| [4][8] (fp16 form)

  // ---- one-time: A fragments (a0: row g, k tig; a1: row g+8, k tig; a2: row g, k tig+4; a3: row g+8, k tig+4);
  // row = cell column c of W' (m-tile mt covers cells mt*16 .. +15), k = gate*16 + own cell
#if ASLP_BWD_T_FP16
  uint4* wTh = reinterpret_cast<uint4*>(smem);
  uint4* wTl = wTh + (size_t)nblk * 128;
  bwd_t_fill_h(D, wTh, wTl, nblk, c0);
  for (int i = threadIdx.x; i < 32; i += NT) cmx[i] = 0.f;
#else
  for (int i = threadIdx.x; i < nblk * 256; i += NT) {
    const int mt = i >> 8, kt = (i >> 5) & 7, ln = i & 31, gg = ln >> 2, tt = ln & 3;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = mt * 16 + gg + (q & 1) * 8;
      const int k = kt * 8 + tt + (q >> 1) * 4, gate = k >> 4, cl = k & 15;
      v[q] = (c < C && c0 + cl < C) ? D.w_r[(size_t)(gate * C + c0 + cl) * D.ldwr + c] : 0.f;
    }
    wT[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
#endif
  for (int i = threadIdx.x; i < 384; i += NT) st[i] = 0.f;
  for (int i = threadIdx.x; i < 512; i += NT) Bsm[i] = 0.f;
  for (int i = threadIdx.x; i < 48; i += NT) {
    const int which = i >> 4, cl = i & 15;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (c0 + cl < C) ? p[c0 + cl] : 0.f;
  }
  __syncthreads();

  // exchange addressing: block(slot, consumer, producer) of this chain
  float* const X = D.xa;
  const size_t slot_stride = (size_t)L.pgroups * nblk * nblk * 128;
  const size_t chain_off = (size_t)cta.pg * nblk * nblk * 128;
  auto own_block = [&](int slot) { return X + (size_t)slot * slot_stride + chain_off + (size_t)cta.blk * nblk * 128; };   // [producer][16][8]
  const int items = nblk * 32;                           // float4 items of the own block
  constexpr int PR = ASLP_PIPE_ROUNDS;                   // gather rounds in flight per step (pipelined polling, recur.cuh)
  auto issue_gather = [&](int slot) {
    const float* src = own_block(slot);
    for (int i = threadIdx.x; i < items; i += NT) cp_async16(Psm + i * 4, src + i * 4);
    if (PR > 1) cp_async_commit();
  };
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};        // debug phase clocks (aslp_lstm_debug_timing), thread 0
  auto gather_clean = [&]() {
    unsigned bad = 0;
    for (int i = threadIdx.x; i < items; i += NT) bad |= sentinel_in(*reinterpret_cast<const float4*>(Psm + i * 4));
    return bad == 0;
  };
  auto complete_gather = [&](int slot) {
    const float* src = own_block(slot);
    unsigned rounds = 0;
    if (PR > 1) {
      // the PR rounds were issued during the previous step: leave on the first one that shows no sentinel
      bool ok = false;
      if (PR >= 3) { cp_async_wait_group<2>(); if (L.timing != nullptr && threadIdx.x == 0) tacc[5] += 1; ok = gather_clean(); }
      if (!ok) { cp_async_wait_group<1>(); if (L.timing != nullptr && threadIdx.x == 0) tacc[5] += 1; ok = gather_clean(); }
      if (!ok) { cp_async_wait_group<0>(); if (L.timing != nullptr && threadIdx.x == 0) tacc[5] += 1; ok = gather_clean(); }
      while (!ok) {
        for (int i = threadIdx.x; i < items; i += NT) cp_async16(Psm + i * 4, src + i * 4);
        cp_async_wait_all();
        if (L.timing != nullptr && threadIdx.x == 0) tacc[5] += 1;
        ok = gather_clean();
        if (++rounds > POLL_LIMIT) __trap();
      }
      return;
    }
    for (;;) {
      cp_async_wait_all();
      if (L.timing != nullptr && threadIdx.x == 0) tacc[5] += 1;
      bool again = false;
      for (int i = threadIdx.x; i < items; i += NT) {
        const float4 v = *reinterpret_cast<const float4*>(Psm + i * 4);
        if (sentinel_in(v)) { cp_async16(Psm + i * 4, src + i * 4); again = true; }
      }
      if (!again) break;
      if (++rounds > POLL_LIMIT) __trap();
    }
  };

  // ---- the (cell, stream) item a finishing thread owns
  const bool fin = threadIdx.x < 128;
  const int cl = (threadIdx.x >> 3) & 15, s_local = threadIdx.x & 7;
  const int s = cta.sbeg + s_local;
  const bool live = fin && s < cta.send && c0 + cl < C;
  const int cc = c0 + cl;
  const int reverse = D.reverse, ldb = D.ldb, lddb = D.lddb;
  float* const buf = D.buf;
  float* const dbuf = D.dbuf;
  const float pi = fin ? pst[cl] : 0.f, pf = fin ? pst[16 + cl] : 0.f, po = fin ? pst[32 + cl] : 0.f;
  auto row_t = [&](int it) { return reverse ? 1 + it : T - it; };      // backward visits time against the forward order
  auto warm = [&](int it) {
    if (!live || it >= T) return;
    const int t = row_t(it), tn = reverse ? t - 1 : t + 1, tp = reverse ? t + 1 : t - 1;
    const float* y = buf + ((size_t)t * S + s) * ldb + cc;
    prefetch_l2(y); prefetch_l2(y + C); prefetch_l2(y + 2 * C); prefetch_l2(y + 3 * C); prefetch_l2(y + 5 * C);
    prefetch_l2(buf + ((size_t)tp * S + s) * ldb + 4 * C + cc);
    prefetch_l2(buf + ((size_t)tn * S + s) * ldb + 2 * C + cc);
    prefetch_l2(dbuf + ((size_t)t * S + s) * lddb + 6 * C + cc);
  };
  warm(0); warm(1);
  issue_gather(0);
  if (PR > 1) { for (int r = 1; r < PR; ++r) cp_async_commit(); }     // slot 0 is the boundary: one real round, PR groups

  for (int it = 0; it < T; ++it) {
    const int slot = it % TR;
    const int t = row_t(it), tn = reverse ? t - 1 : t + 1, tp = reverse ? t + 1 : t - 1;
    RECUR_TICK(k0);
    complete_gather(slot);
    RECUR_TICK(k1);
    float yv[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // g, i, f, o, h at t ; c at the forward predecessor ; f at the successor
    float od = 0.f;                                      // out_diff share of d_m (preloaded in the m columns of dbuf)
    if (live) {
      const float* y = buf + ((size_t)t * S + s) * ldb + cc;
      yv[0] = y[0]; yv[1] = y[C]; yv[2] = y[2 * C]; yv[3] = y[3 * C]; yv[4] = y[5 * C];
      yv[5] = buf[((size_t)tp * S + s) * ldb + 4 * C + cc];
      yv[6] = buf[((size_t)tn * S + s) * ldb + 2 * C + cc];
      od = dbuf[((size_t)t * S + s) * lddb + 6 * C + cc];
    }
    __syncthreads();                                     // every thread's share of the gather is in Psm
    RECUR_TICK(k2);
    float dg = 0.f, di = 0.f, df = 0.f, dout = 0.f, dc = 0.f, dh = 0.f, dm = 0.f;
    if (fin) {
      if (live) {
        // fixed order (deterministic): four interleaved partial chains, so the nblk shared-memory loads are in flight
        // together instead of one load-add round trip per producer
        const float* pp = Psm + cl * 8 + s_local;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int p = 0;
        for (; p + 4 <= nblk; p += 4) { s0 += pp[p * 128]; s1 += pp[(p + 1) * 128]; s2 += pp[(p + 2) * 128]; s3 += pp[(p + 3) * 128]; }
        for (; p < nblk; ++p) s0 += pp[p * 128];
        const float sum = (s0 + s1) + (s2 + s3);
        const int si = cl * 8 + s_local;
        const float yg = yv[0], yi = yv[1], yf = yv[2], yo = yv[3], yh = yv[4], c_prev = yv[5], yf_next = yv[6];
        dm = sum + od;
        const float dc_n = st[si], di_n = st[128 + si], df_n = st[256 + si];
        dh = dm * yo;  dh = (1.0f - yh * yh) * dh;                       // DiffTanh(y_h, d_h)
        dout = dm * yh;  dout = yo * (1.0f - yo) * dout;                 // DiffSigmoid(y_o, d_o)
        dc = dh + dc_n * yf_next + di_n * pi + df_n * pf + dout * po;
        df = dc * c_prev;  df = yf * (1.0f - yf) * df;
        di = dc * yg;      di = yi * (1.0f - yi) * di;
        dg = dc * yi;      dg = (1.0f - yg * yg) * dg;
        st[si] = dc; st[128 + si] = di; st[256 + si] = df;
      }
      const int bi = cl * 8 + s_local;                   // B operand of this step's contraction (zeros for padding streams)
      Bsm[bi] = dg; Bsm[128 + bi] = di; Bsm[256 + bi] = df; Bsm[384 + bi] = dout;
#if ASLP_BWD_T_FP16
      // largest |dgifo| of every stream column over this warp's four cells (lanes with equal lane & 7)
      float m4 = fmaxf(fmaxf(fabsf(dg), fabsf(di)), fmaxf(fabsf(df), fabsf(dout)));
      m4 = fmaxf(m4, ::__shfl_xor_sync(0xffffffffu, m4, 8));      // (global overload: <cuda_fp16.h> sits inside this namespace and its __half shuffles would hide the float one)
      m4 = fmaxf(m4, ::__shfl_xor_sync(0xffffffffu, m4, 16));
      if (lane < 8) cmx[warp * 8 + lane] = m4;
#endif
    } else {
      __threadfence();                                   // the re-arm stores of the previous step (below) precede this step's publish
    }
    RECUR_TICK(k3);
    if (PR > 1) cp_async_wait_all();                     // this step's stale gather rounds are done before the slot is re-armed / Psm reused
    __syncthreads();                                     // Bsm complete, re-arm fenced
    RECUR_TICK(k4);
    if (it + 1 < T) {
      // ---- contraction of the own 64 dgifo rows against all C columns, published straight from the accumulators
      const int nslot = (it + 1) % TR;
      float* out = X + (size_t)nslot * slot_stride + chain_off + (size_t)cta.blk * 128;      // + consumer * nblk * 128
      const int nm = (nblk - warp + 7) >> 3;             // m-tiles of this warp (warp-uniform)
#if ASLP_BWD_T_FP16
      if (nm >= 3) bwd_t_contract_h<3>(wTh, wTl, Bsm, cmx, out, nblk, warp, lane);
      else if (nm == 2) bwd_t_contract_h<2>(wTh, wTl, Bsm, cmx, out, nblk, warp, lane);
      else if (nm == 1) bwd_t_contract_h<1>(wTh, wTl, Bsm, cmx, out, nblk, warp, lane);
#else
      if (nm >= 3) bwd_t_contract<3>(wT, Bsm, out, nblk, warp, lane);
      else if (nm == 2) bwd_t_contract<2>(wT, Bsm, out, nblk, warp, lane);
      else if (nm == 1) bwd_t_contract<1>(wT, Bsm, out, nblk, warp, lane);
#endif
    }
    RECUR_TICK(k5);
    // ---- off the chain: re-arm the slot consumed at the top of this step, start the next gather, and only then write the
    // bookkeeping of this step (they overlap the gather's round trip).  The re-arm and the prefetch address arithmetic put
    // about one store-to-L2 latency between the publish and the gather, so its first round usually finds the data.
    const bool more = it + 1 < T;
    if (PR > 1 && more) issue_gather((it + 1) % TR);     // round 0 right behind the publish (Psm is free: its readers passed the barrier)
    if (!fin) {
      float* blk = own_block(slot);
      const float4 sent = make_float4(__uint_as_float(SENTINEL), __uint_as_float(SENTINEL), __uint_as_float(SENTINEL), __uint_as_float(SENTINEL));
      for (int i = threadIdx.x - 128; i < items; i += NT - 128) *reinterpret_cast<float4*>(blk + i * 4) = sent;
    }
    if (more) {
      warm(it + 2);
#if ASLP_BWD_T_DELAY_NS > 0
      __nanosleep(ASLP_BWD_T_DELAY_NS);
#endif
      if (PR == 1 || PR >= 3) issue_gather((it + 1) % TR);   // single-round form: Psm is free, its readers passed the barrier above
    }
    if (live) {
      float* d = dbuf + ((size_t)t * S + s) * lddb + cc;
      d[0] = dg; d[C] = di; d[2 * C] = df; d[3 * C] = dout; d[4 * C] = dc; d[5 * C] = dh; d[6 * C] = dm;
    }
    if (PR >= 2 && more) {
#if ASLP_PIPE_BWD_GAP_NS > 0
      __nanosleep(ASLP_PIPE_BWD_GAP_NS);
#endif
      issue_gather((it + 1) % TR);
    }
    if (L.timing != nullptr && threadIdx.x == 0) {
      const long long k6 = clock64();
      tacc[0] += k1 - k0; tacc[1] += k2 - k1; tacc[2] += k3 - k2; tacc[3] += k4 - k3; tacc[4] += k5 - k4; tacc[6] += k6 - k5;
    }
  }
  if (L.timing != nullptr && threadIdx.x == 0)
    for (int q = 0; q < 12; ++q) L.timing[(size_t)blockIdx.x * 12 + q] = tacc[q];
}

// ---------------------------------------------------------------- wide form: 8*NS streams per chain (NS = 2, 3, 4)
// Same algorithm for minibatches with many streams (BASELINE cfg2: 100 streams, 512 cells): a chain of 8 streams per CTA
// set would need more CTAs than the chip has, so a chain carries NS n-tiles of 8 streams.  The A fragment of an
// (m-tile, k-tile) is loaded once and used for NS x 3 MMAs; every thread finishes 16 * 8NS / 256 (cell, stream) items; the
// re-arm + fence is done by all threads after the finish (the 4 idle warps of the 8-stream form do not exist here).
inline size_t bwd_tn_smem_floats(int nblk, int ns) { return (size_t)nblk * 1024 + 64 * 40 + (size_t)nblk * 128 * ns + 384 * ns + 48 + 64; }   // + column maxima [2][32]
// Bsm row pitch: the rows one fragment load touches (tig, or 2 tig for the k16 shape) x 8 streams hit 32 different banks
template <int NS> struct BwdTnPitch { static constexpr int value = ASLP_BWD_T_FP16 ? 8 * NS + 4 : ((NS == 2 || NS == 4) ? 8 * NS + 8 : 8 * NS); };

template <int NM, int NS>
__device__ __forceinline__ void bwd_tn_contract(const float4* wT, const float* Bsm, float* out, int nblk, int warp, int lane) {
  constexpr int SW = 8 * NS;                             // streams per chain = row length of an exchange block
  constexpr int BP = BwdTnPitch<NS>::value;
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll 1
  for (int mi = 0; mi < NM; ++mi) {
    const int mt = warp + 8 * mi;
    float hh[NS][4], lh[NS][4], hl[NS][4];
#pragma unroll
    for (int nt = 0; nt < NS; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) { hh[nt][q] = 0.f; lh[nt][q] = 0.f; hl[nt][q] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      const float4 w4 = wT[(mt * 8 + kt) * 32 + lane];
      uint32_t ah[4], al[4];
      split_tf32(w4.x, ah[0], al[0]); split_tf32(w4.y, ah[1], al[1]); split_tf32(w4.z, ah[2], al[2]); split_tf32(w4.w, ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < NS; ++nt) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(Bsm[(kt * 8 + tig) * BP + nt * 8 + g], bh0, bl0);
        split_tf32(Bsm[(kt * 8 + tig + 4) * BP + nt * 8 + g], bh1, bl1);
        mma_tf32(lh[nt], al, bh0, bh1);
        mma_tf32(hl[nt], ah, bl0, bl1);
        mma_tf32(hh[nt], ah, bh0, bh1);
      }
    }
    float* dst = out + (size_t)mt * nblk * 16 * SW;      // block of consumer CTA mt: [producer][16][SW]
#pragma unroll
    for (int nt = 0; nt < NS; ++nt) {
      st_pub2(dst + g * SW + nt * 8 + 2 * tig, (lh[nt][0] + hl[nt][0]) + hh[nt][0], (lh[nt][1] + hl[nt][1]) + hh[nt][1]);
      st_pub2(dst + (g + 8) * SW + nt * 8 + 2 * tig, (lh[nt][2] + hl[nt][2]) + hh[nt][2], (lh[nt][3] + hl[nt][3]) + hh[nt][3]);
    }
  }
}

// fp16 form (see bwd_t_contract_h): the B fragments of all NS n-tiles are scaled and split once per step and warp, then
// every m-tile costs 2 fragment loads and 3 NS MMAs per 16 k.  cmx: bit patterns of the per-stream maxima of |dgifo|.
template <int NM, int NS>
__device__ __forceinline__ void bwd_tn_contract_h(const uint4* wh, const uint4* wl, const float* Bsm, const unsigned* cmx, float* out, int nblk,
                                                  int warp, int lane) {
  constexpr int SW = 8 * NS;
  constexpr int BP = BwdTnPitch<NS>::value;
  const int g = lane >> 2, tig = lane & 3;
  uint32_t bh[NS][4][2], bl[NS][4][2];
  float inv[NS][2];
#pragma unroll
  for (int nt = 0; nt < NS; ++nt) {
    float unused;
    const float sc = pow2_scale(__uint_as_float(cmx[nt * 8 + g]), unused);
    pow2_scale(__uint_as_float(cmx[nt * 8 + 2 * tig]), inv[nt][0]);
    pow2_scale(__uint_as_float(cmx[nt * 8 + 2 * tig + 1]), inv[nt][1]);
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const float* col = Bsm + (kh * 16 + 2 * tig) * BP + nt * 8 + g;
      split_h2(col[0] * sc, col[BP] * sc, bh[nt][kh][0], bl[nt][kh][0]);
      split_h2(col[8 * BP] * sc, col[9 * BP] * sc, bh[nt][kh][1], bl[nt][kh][1]);
    }
  }
#pragma unroll 1
  for (int mi = 0; mi < NM; ++mi) {
    const int mt = warp + 8 * mi;
    float hh[NS][4], lh[NS][4], hl[NS][4];
#pragma unroll
    for (int nt = 0; nt < NS; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) { hh[nt][q] = 0.f; lh[nt][q] = 0.f; hl[nt][q] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const uint4 h4 = wh[(mt * 4 + kh) * 32 + lane], l4 = wl[(mt * 4 + kh) * 32 + lane];
      const uint32_t ah[4] = {h4.x, h4.y, h4.z, h4.w}, al[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int nt = 0; nt < NS; ++nt) {
        mma_f16(lh[nt], al, bh[nt][kh][0], bh[nt][kh][1]);
        mma_f16(hl[nt], ah, bl[nt][kh][0], bl[nt][kh][1]);
        mma_f16(hh[nt], ah, bh[nt][kh][0], bh[nt][kh][1]);
      }
    }
    float* dst = out + (size_t)mt * nblk * 16 * SW;      // block of consumer CTA mt: [producer][16][SW]
#pragma unroll
    for (int nt = 0; nt < NS; ++nt) {
      st_pub2(dst + g * SW + nt * 8 + 2 * tig, ((lh[nt][0] + hl[nt][0]) + hh[nt][0]) * inv[nt][0], ((lh[nt][1] + hl[nt][1]) + hh[nt][1]) * inv[nt][1]);
      st_pub2(dst + (g + 8) * SW + nt * 8 + 2 * tig, ((lh[nt][2] + hl[nt][2]) + hh[nt][2]) * inv[nt][0], ((lh[nt][3] + hl[nt][3]) + hh[nt][3]) * inv[nt][1]);
    }
  }
}

template <int NS>
__global__ void __launch_bounds__(NT, 1) lstm_bwd_tn_kernel(Launch L) {
  constexpr int SW = 8 * NS;                             // streams per chain
  constexpr int BP = BwdTnPitch<NS>::value;
  constexpr int ITEMS = 16 * SW;                         // (cell, stream) items of a CTA
  constexpr int IPT = (ITEMS + NT - 1) / NT;             // items per thread (1 or 2)
  extern __shared__ __align__(16) float smem[];
  const MmaCta cta = mma_cta(L);
  const DirDev& D = L.d[cta.dir];
  const int T = D.T, S = D.S, C = D.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = L.nblk;
  const int c0 = cta.blk * 16;
  if (cta.sbeg >= S) return;

#if !ASLP_BWD_T_FP16
  float4* wT = reinterpret_cast<float4*>(smem);          // [nblk][8][32] float4
#endif
  float* Bsm = smem + (size_t)nblk * 1024;               // [64][BP]
  float* Psm = Bsm + 64 * 40;                            // [nblk][16][SW]
  float* st = Psm + (size_t)nblk * 16 * SW;              // [3][16][SW]
  float* pst = st + 3 * ITEMS;                           // [3][16]
  unsigned* cmx = reinterpret_cast<unsigned*>(pst + 48); // per-stream maxima of |dgifo| (bit patterns), double-buffered [2][32] (fp16 form)

#if ASLP_BWD_T_FP16
  uint4* wTh = reinterpret_cast<uint4*>(smem);
  uint4* wTl = wTh + (size_t)nblk * 128;
  bwd_t_fill_h(D, wTh, wTl, nblk, c0);
  for (int i = threadIdx.x; i < 64; i += NT) cmx[i] = 0u;
#else
  for (int i = threadIdx.x; i < nblk * 256; i += NT) {
    const int mt = i >> 8, kt = (i >> 5) & 7, ln = i & 31, gg = ln >> 2, tt = ln & 3;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = mt * 16 + gg + (q & 1) * 8;
      const int k = kt * 8 + tt + (q >> 1) * 4, gate = k >> 4, cl = k & 15;
      v[q] = (c < C && c0 + cl < C) ? D.w_r[(size_t)(gate * C + c0 + cl) * D.ldwr + c] : 0.f;
    }
    wT[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
#endif
  for (int i = threadIdx.x; i < 3 * ITEMS; i += NT) st[i] = 0.f;
  for (int i = threadIdx.x; i < 64 * BP; i += NT) Bsm[i] = 0.f;
  for (int i = threadIdx.x; i < 48; i += NT) {
    const int which = i >> 4, cl = i & 15;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (c0 + cl < C) ? p[c0 + cl] : 0.f;
  }
  __syncthreads();

  float* const X = D.xa;
  const size_t blk_floats = (size_t)16 * SW;             // one (consumer, producer) block
  const size_t slot_stride = (size_t)L.pgroups * nblk * nblk * blk_floats;
  const size_t chain_off = (size_t)cta.pg * nblk * nblk * blk_floats;
  auto own_block = [&](int slot) { return X + (size_t)slot * slot_stride + chain_off + (size_t)cta.blk * nblk * blk_floats; };
  const int items4 = nblk * 4 * SW;                      // float4 items of the own block
  auto issue_gather = [&](int slot) {
    const float* src = own_block(slot);
    for (int i = threadIdx.x; i < items4; i += NT) cp_async16(Psm + i * 4, src + i * 4);
  };
  auto complete_gather = [&](int slot) {
    const float* src = own_block(slot);
    unsigned rounds = 0;
    for (;;) {
      cp_async_wait_all();
      bool again = false;
      for (int i = threadIdx.x; i < items4; i += NT) {
        const float4 v = *reinterpret_cast<const float4*>(Psm + i * 4);
        if (sentinel_in(v)) { cp_async16(Psm + i * 4, src + i * 4); again = true; }
      }
      if (!again) break;
      if (++rounds > POLL_LIMIT) __trap();
    }
  };

  const int reverse = D.reverse, ldb = D.ldb, lddb = D.lddb;
  float* const buf = D.buf;
  float* const dbuf = D.dbuf;
  auto row_t = [&](int it) { return reverse ? 1 + it : T - it; };
  issue_gather(0);

  for (int it = 0; it < T; ++it) {
    const int slot = it % TR;
    const int t = row_t(it), tn = reverse ? t - 1 : t + 1, tp = reverse ? t + 1 : t - 1;
    complete_gather(slot);
    float yv[IPT][7], od[IPT];
    bool live[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const int item = threadIdx.x + r * NT, cl = item / SW, sl = item - cl * SW, s = cta.sbeg + sl;
      live[r] = item < ITEMS && s < cta.send && c0 + cl < C;
#pragma unroll
      for (int q = 0; q < 7; ++q) yv[r][q] = 0.f;
      od[r] = 0.f;
      if (live[r]) {
        const int cc = c0 + cl;
        const float* y = buf + ((size_t)t * S + s) * ldb + cc;
        yv[r][0] = y[0]; yv[r][1] = y[C]; yv[r][2] = y[2 * C]; yv[r][3] = y[3 * C]; yv[r][4] = y[5 * C];
        yv[r][5] = buf[((size_t)tp * S + s) * ldb + 4 * C + cc];
        yv[r][6] = buf[((size_t)tn * S + s) * ldb + 2 * C + cc];
        od[r] = dbuf[((size_t)t * S + s) * lddb + 6 * C + cc];
      }
    }
    __syncthreads();                                     // every thread's share of the gather is in Psm
    float dv[IPT][7];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const int item = threadIdx.x + r * NT;
#pragma unroll
      for (int q = 0; q < 7; ++q) dv[r][q] = 0.f;
      if (item < ITEMS) {
        const int cl = item / SW;
        if (live[r]) {
          // fixed order (deterministic), four interleaved chains: the nblk loads are in flight together
          const float* pp = Psm + item;
          float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
          int p = 0;
          for (; p + 4 <= nblk; p += 4) { q0 += pp[(size_t)p * ITEMS]; q1 += pp[(size_t)(p + 1) * ITEMS]; q2 += pp[(size_t)(p + 2) * ITEMS]; q3 += pp[(size_t)(p + 3) * ITEMS]; }
          for (; p < nblk; ++p) q0 += pp[(size_t)p * ITEMS];
          const float sum = (q0 + q1) + (q2 + q3);
          const float pi = pst[cl], pf = pst[16 + cl], po = pst[32 + cl];
          const float yg = yv[r][0], yi = yv[r][1], yf = yv[r][2], yo = yv[r][3], yh = yv[r][4], c_prev = yv[r][5], yf_next = yv[r][6];
          const float dm = sum + od[r];
          const float dc_n = st[item], di_n = st[ITEMS + item], df_n = st[2 * ITEMS + item];
          float dh = dm * yo;  dh = (1.0f - yh * yh) * dh;
          float dout = dm * yh;  dout = yo * (1.0f - yo) * dout;
          const float dc = dh + dc_n * yf_next + di_n * pi + df_n * pf + dout * po;
          float df = dc * c_prev;  df = yf * (1.0f - yf) * df;
          float di = dc * yg;      di = yi * (1.0f - yi) * di;
          float dg = dc * yi;      dg = (1.0f - yg * yg) * dg;
          st[item] = dc; st[ITEMS + item] = di; st[2 * ITEMS + item] = df;
          dv[r][0] = dg; dv[r][1] = di; dv[r][2] = df; dv[r][3] = dout; dv[r][4] = dc; dv[r][5] = dh; dv[r][6] = dm;
        }
        const int sl = item - cl * SW;                   // B operand rows: gate * 16 + cell
        Bsm[(0 * 16 + cl) * BP + sl] = dv[r][0]; Bsm[(1 * 16 + cl) * BP + sl] = dv[r][1];
        Bsm[(2 * 16 + cl) * BP + sl] = dv[r][2]; Bsm[(3 * 16 + cl) * BP + sl] = dv[r][3];
#if ASLP_BWD_T_FP16
        const float m4 = fmaxf(fmaxf(fabsf(dv[r][0]), fabsf(dv[r][1])), fmaxf(fabsf(dv[r][2]), fabsf(dv[r][3])));
        if (m4 > 0.f) atomicMax(&cmx[(it & 1) * 32 + sl], __float_as_uint(m4));      // non-negative floats order like their bits
#endif
      }
    }
#if ASLP_BWD_T_FP16
    if (threadIdx.x < 32) cmx[((it + 1) & 1) * 32 + threadIdx.x] = 0u;   // next step's maxima; its last readers passed the barrier above
#endif
    {                                                    // re-arm the slot just consumed; visible before the next publish
      float* blk = own_block(slot);
      const float4 sent = make_float4(__uint_as_float(SENTINEL), __uint_as_float(SENTINEL), __uint_as_float(SENTINEL), __uint_as_float(SENTINEL));
      for (int i = threadIdx.x; i < items4; i += NT) *reinterpret_cast<float4*>(blk + i * 4) = sent;
      __threadfence();
    }
    __syncthreads();                                     // Bsm complete, re-arm fenced
    if (it + 1 < T) {
      const int nslot = (it + 1) % TR;
      float* out = X + (size_t)nslot * slot_stride + chain_off + (size_t)cta.blk * blk_floats;   // + consumer * nblk * blk_floats
      const int nm = (nblk - warp + 7) >> 3;             // m-tiles of this warp (warp-uniform)
#if ASLP_BWD_T_FP16
      const unsigned* cm = cmx + (it & 1) * 32;
      if (nm >= 4) bwd_tn_contract_h<4, NS>(wTh, wTl, Bsm, cm, out, nblk, warp, lane);
      else if (nm == 3) bwd_tn_contract_h<3, NS>(wTh, wTl, Bsm, cm, out, nblk, warp, lane);
      else if (nm == 2) bwd_tn_contract_h<2, NS>(wTh, wTl, Bsm, cm, out, nblk, warp, lane);
      else if (nm == 1) bwd_tn_contract_h<1, NS>(wTh, wTl, Bsm, cm, out, nblk, warp, lane);
#else
      if (nm >= 4) bwd_tn_contract<4, NS>(wT, Bsm, out, nblk, warp, lane);
      else if (nm == 3) bwd_tn_contract<3, NS>(wT, Bsm, out, nblk, warp, lane);
      else if (nm == 2) bwd_tn_contract<2, NS>(wT, Bsm, out, nblk, warp, lane);
      else if (nm == 1) bwd_tn_contract<1, NS>(wT, Bsm, out, nblk, warp, lane);
#endif
      issue_gather(nslot);
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      if (live[r]) {
        const int item = threadIdx.x + r * NT, cl = item / SW, sl = item - cl * SW;
        float* d = dbuf + ((size_t)t * S + cta.sbeg + sl) * lddb + c0 + cl;
        d[0] = dv[r][0]; d[C] = dv[r][1]; d[2 * C] = dv[r][2]; d[3 * C] = dv[r][3]; d[4 * C] = dv[r][4]; d[5 * C] = dv[r][5]; d[6 * C] = dv[r][6];
      }
    }
  }
}
