// Persistent LSTM-family recurrence for sm_100a: ONE cooperative launch walks all T time steps of
// every direction, instead of the reference's per-step 2 small GEMMs + 13 (fwd) / 15 (bwd)
// pointwise launches, each followed by a device sync
// (src/aslp-nnet/nnet-blstm-projected-streams-lc.h:552-717 fwd, :729-960 bwd;
//  nnet-recurrent-component.cc:235-480; nnet-lstm-projected-streams.h:313-617).
//
// Design (B200-first):
//  * the cells of a direction are partitioned over CTAs; each CTA keeps ITS slice of W_gifo_r and
//    W_r_m resident in shared memory for the whole sequence (weights are read from HBM once);
//  * cross-CTA exchange of m(t) / r(t) (and dgifo(t) / d_r(t) backward) goes through L2 in a
//    stream-minor exchange buffer [T+2][dim][S]; there is NO grid barrier: the buffer is pre-filled
//    with a NaN sentinel and consumers poll the DATA ITSELF (every fp32 word is its own flag), so a
//    CTA proceeds the moment its inputs land;
//  * a work unit is (4 rows) x (16 streams): lanes split the contraction dimension, a butterfly
//    shuffle-reduce leaves each lane pair with the 4 row sums of one stream, and the whole cell
//    update (peepholes, sigmoid/tanh, clamp, output gate) happens in registers right there;
//  * activations keep the reference buffer layout [(T+2)S, 7C+R] = [g i f o c h m r] so the chunk
//    wgrad GEMMs (lc.h:981-1000) read strided views of it.
#include "common.cuh"
#include "scratch.cuh"
#include "recur.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>

namespace {

using namespace recur;

struct DirDev {
  int T, S, C, R, Rr;        // Rr = recurrent input dim (R, or C when R == 0)
  int reverse;
  float* buf; int ldb;
  float* dbuf; int lddb;
  const float* w_r; int ldwr;
  const float* w_rm; int ldwrm;
  const float* peep_i; const float* peep_f; const float* peep_o;
  const int* seq_len;
  float clip;
  int SX;                    // stream stride of the exchange buffers (S rounded up to 4)
  float* xa;                 // fwd: m exchange  [T+2][C][SX]   | bwd: dgifo exchange [T+2][4C][SX]
  float* xb;                 // fwd: r exchange  [T+2][R][SX]   | bwd: d_r  exchange [T+2][R][SX]
  int cb, rb;                // cells / r-rows owned per CTA
};
struct Launch {
  DirDev d[2];
  int ndirs, nblk;           // CTAs per direction
  int SG;                    // streams staged per group (multiple of 16)
  int SP;                    // staging stride in floats (SG + 4: conflict-free 128-bit reads)
  int pgroups, SGP;          // tensor-core form: parallel stream groups per direction and streams per group
  long long* timing;         // debug: per-CTA clock64 totals [nCTA][12] = {wait for CTA, own poll, wait for CTA's polls, units}; or NULL
};
#define RECUR_TICK(var) const long long var = (L.timing != nullptr && threadIdx.x == 0) ? clock64() : 0

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(NT, 1) lstm_fwd_kernel(Launch L) {
  extern __shared__ float smem[];
  const int dir = blockIdx.x / L.nblk, blk = blockIdx.x % L.nblk;
  const DirDev& D = L.d[dir];
  const int T = D.T, S = D.S, C = D.C, R = D.R, Rr = D.Rr, SX = D.SX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SP = L.SP, SG = L.SG;

  const int c0 = blk * D.cb, nc = max(0, min(D.cb, C - c0));          // own cells
  const int j0 = blk * D.rb, nr = (R > 0) ? max(0, min(D.rb, R - j0)) : 0;   // own projection rows
  const int ncp = (D.cb + 3) & ~3, nrp = (D.rb + 3) & ~3;              // padded to whole units

  // shared memory carve-up
  float* wA = smem;                                 // [cb*4][Rr]   row = cell_local*4 + gate
  float* wB = wA + (size_t)D.cb * 4 * Rr;           // [nrp][C]
  float* xT = wB + (size_t)nrp * C;                 // [max(Rr, C)][SP]  staging (A then B share it)
  float* cst = xT + (size_t)max(Rr, C) * SP;        // [cb][SXs] c(t-1) state, SXs = stream count padded
  const int SXs = SX;
  float* pst = cst + (size_t)D.cb * SXs;            // peepholes [3][cb]

  // one-time: weights, peepholes, initial cell state
  for (int i = threadIdx.x; i < D.cb * 4 * Rr; i += NT) {
    const int row = i / Rr, k = i - row * Rr;
    const int cl = row >> 2, g = row & 3;
    wA[i] = (cl < nc) ? D.w_r[(size_t)(g * C + c0 + cl) * D.ldwr + k] : 0.f;
  }
  for (int i = threadIdx.x; i < nrp * C; i += NT) {
    const int row = i / C, k = i - row * C;
    wB[i] = (row < nr) ? D.w_rm[(size_t)(j0 + row) * D.ldwrm + k] : 0.f;
  }
  const int slot0 = D.reverse ? T + 1 : 0;
  for (int i = threadIdx.x; i < D.cb * SXs; i += NT) {
    const int cl = i / SXs, s = i - cl * SXs;
    cst[i] = (cl < nc && s < S) ? D.buf[((size_t)slot0 * S + s) * D.ldb + 4 * C + c0 + cl] : 0.f;
  }
  for (int i = threadIdx.x; i < 3 * D.cb; i += NT) {
    const int which = i / D.cb, cl = i - which * D.cb;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (cl < nc) ? p[c0 + cl] : 0.f;
  }
  __syncthreads();

  const int ngroups = (S + SG - 1) / SG;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float* xrec = (R > 0) ? D.xb : D.xa;            // what feeds the gates: r (projected) or m
  const int my_s_in_chunk = lane_stream(lane);

  for (int step = 0; step < T; ++step) {
    const int t = D.reverse ? T - step : 1 + step;
    const int tp = D.reverse ? t + 1 : t - 1;
    for (int grp = 0; grp < ngroups; ++grp) {
      const int s0 = grp * SG;
      const int sg = min(SG, SX - s0);                   // staged streams (multiple of 4)
      const int nchunks = (min(SG, S - s0) + 15) / 16;
      // ---------------- phase A: gates + cell update for own cells
      RECUR_TICK(k0);
      __syncthreads();                                   // previous readers of xT are done
      RECUR_TICK(k1);
      stage_poll<8>(xT, SP, xrec + (size_t)tp * Rr * SX, Rr, SX, s0, sg >> 2);
      RECUR_TICK(k2);
      __syncthreads();
      RECUR_TICK(k3);
      for (int u = warp; u < nc * nchunks; u += NW) {
        const int cl = u / nchunks, ch = u - cl * nchunks;
        const int s = s0 + ch * 16 + my_s_in_chunk;
        const bool active = ((lane & 1) == 0) && (s < S);
        const size_t row = (size_t)t * S + s;
        float pre[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[g] = D.buf[row * D.ldb + g * C + c0 + cl];   // x*W_x^T + bias (issued early)
        }
        float acc[4][16], sum[4];
        RECUR_TICK(u0);
        unit_dot<4>(acc, wA + (size_t)cl * 4 * Rr, Rr, Rr, xT, SP, ch * 16, lane);
        RECUR_TICK(u1);
        unit_reduce<4>(acc, sum, lane);
        RECUR_TICK(u2);
        if (L.timing != nullptr && threadIdx.x == 0) { tacc[4] += u1 - u0; tacc[5] += u2 - u1; tacc[6] += u0 - k3; }
        if (active) {
          const float cprev = cst[cl * SXs + s];
          float yg = pre[0] + sum[0];
          float yi = pre[1] + sum[1] + cprev * pst[cl];
          float yf = pre[2] + sum[2] + cprev * pst[D.cb + cl];
          float yo = pre[3] + sum[3];
          yi = ref_sigmoid(yi); yf = ref_sigmoid(yf); yg = ref_tanh(yg);
          float yc = yg * yi + cprev * yf;
          yc = fminf(fmaxf(yc, -D.clip), D.clip);
          float yh = ref_tanh(yc);
          yo = ref_sigmoid(yo + yc * pst[2 * D.cb + cl]);
          float ym = yh * yo;
          if (D.seq_len != nullptr && t > D.seq_len[s]) { yg = yi = yf = yo = yc = yh = ym = 0.f; }
          float* o = D.buf + row * D.ldb + c0 + cl;
          o[0] = yg; o[C] = yi; o[2 * C] = yf; o[3 * C] = yo; o[4 * C] = yc; o[5 * C] = yh; o[6 * C] = ym;
          cst[cl * SXs + s] = yc;
          st_pub(D.xa + ((size_t)t * C + c0 + cl) * SX + s, ym);   // publish m(t): the store is the flag
        }
      }
      if (L.timing != nullptr && threadIdx.x == 0) {
        const long long k4 = clock64();
        tacc[0] += k1 - k0; tacc[1] += k2 - k1; tacc[2] += k3 - k2; tacc[3] += k4 - k3;
      }
      // ---------------- phase B: r(t) = m(t) W_rm^T for own projection rows
      if (R > 0) {
        __syncthreads();
        stage_poll<8>(xT, SP, D.xa + (size_t)t * C * SX, C, SX, s0, sg >> 2);
        __syncthreads();
        const int nru = (nr + 3) >> 2;
        for (int u = warp; u < nru * nchunks; u += NW) {
          const int ru = u / nchunks, ch = u - ru * nchunks;
          const int s = s0 + ch * 16 + my_s_in_chunk;
          float acc[4][16], sum[4];
          unit_dot<4>(acc, wB + (size_t)ru * 4 * C, C, C, xT, SP, ch * 16, lane);
          unit_reduce<4>(acc, sum, lane);
          if (((lane & 1) == 0) && s < S) {
            const size_t row = (size_t)t * S + s;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = ru * 4 + q;
              if (j < nr) {
                D.buf[row * D.ldb + 7 * C + j0 + j] = sum[q];
                st_pub(D.xb + ((size_t)t * R + j0 + j) * SX + s, sum[q]);   // publish r(t)
              }
            }
          }
        }
      }
    }
  }
  if (L.timing != nullptr && threadIdx.x == 0)
    for (int q = 0; q < 8; ++q) L.timing[(size_t)blockIdx.x * 12 + q] = tacc[q];
}

// ---------------------------------------------------------------- backward (BPTT, reference "version 1")
__global__ void __launch_bounds__(NT, 1) lstm_bwd_kernel(Launch L) {
  extern __shared__ float smem[];
  const int dir = blockIdx.x / L.nblk, blk = blockIdx.x % L.nblk;
  const DirDev& D = L.d[dir];
  const int T = D.T, S = D.S, C = D.C, R = D.R, SX = D.SX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SP = L.SP, SG = L.SG;
  const int G4 = 4 * C;

  const int c0 = blk * D.cb, nc = max(0, min(D.cb, C - c0));
  const int j0 = blk * D.rb, nr = (R > 0) ? max(0, min(D.rb, R - j0)) : 0;
  const int ncp = (D.cb + 3) & ~3, nrp = (D.rb + 3) & ~3;

  // R > 0 : w1[nrp][4C] = columns j of W_r (d_r = dgifo(tn) W_r), w2[ncp][R] = columns c of W_rm (d_m = d_r W_rm)
  // R == 0: w1[ncp][4C] = columns c of W_r (d_m = out_diff + dgifo(tn) W_r)
  const int n1 = (R > 0) ? nrp : ncp;
  float* w1 = smem;
  float* w2 = w1 + (size_t)n1 * G4;
  float* xT = w2 + ((R > 0) ? (size_t)ncp * R : 0);
  float* st = xT + (size_t)max(G4, R) * SP;          // state [3][cb][SX]: d_c(tn), d_i(tn), d_f(tn)
  float* pst = st + (size_t)3 * D.cb * SX;           // peepholes [3][cb]

  for (int i = threadIdx.x; i < n1 * G4; i += NT) {
    const int row = i / G4, q = i - row * G4;
    float v = 0.f;
    if (R > 0) { if (row < nr) v = D.w_r[(size_t)q * D.ldwr + j0 + row]; }
    else       { if (row < nc) v = D.w_r[(size_t)q * D.ldwr + c0 + row]; }
    w1[i] = v;
  }
  if (R > 0) {
    for (int i = threadIdx.x; i < ncp * R; i += NT) {
      const int row = i / R, j = i - row * R;
      w2[i] = (row < nc) ? D.w_rm[(size_t)j * D.ldwrm + c0 + row] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < 3 * D.cb * SX; i += NT) st[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * D.cb; i += NT) {
    const int which = i / D.cb, cl = i - which * D.cb;
    const float* p = which == 0 ? D.peep_i : (which == 1 ? D.peep_f : D.peep_o);
    pst[i] = (cl < nc) ? p[c0 + cl] : 0.f;
  }
  __syncthreads();

  const int ngroups = (S + SG - 1) / SG;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int my_s_in_chunk = lane_stream(lane);
  const int ocol = (R > 0) ? 7 * C : 6 * C;          // where out_diff was preloaded in dbuf (r or m columns)

  for (int step = 0; step < T; ++step) {
    // backward visits time in the opposite order of the forward pass of this direction
    const int t = D.reverse ? 1 + step : T - step;
    const int tn = D.reverse ? t - 1 : t + 1;        // the forward successor (already back-propagated)
    const int tp = D.reverse ? t + 1 : t - 1;        // the forward predecessor
    for (int grp = 0; grp < ngroups; ++grp) {
      const int s0 = grp * SG;
      const int sg = min(SG, SX - s0);
      const int nchunks = (min(SG, S - s0) + 15) / 16;
      // ---------------- stage dgifo(tn)  [4C][streams]
      RECUR_TICK(k0);
      __syncthreads();
      RECUR_TICK(k1);
      stage_poll<20>(xT, SP, D.xa + (size_t)tn * G4 * SX, G4, SX, s0, sg >> 2);
      RECUR_TICK(k2);
      __syncthreads();
      RECUR_TICK(k3);
      if (R > 0) {
        // phase B1: d_r(t) for own projection rows
        const int nru = (nr + 3) >> 2;
        for (int u = warp; u < nru * nchunks; u += NW) {
          const int ru = u / nchunks, ch = u - ru * nchunks;
          const int s = s0 + ch * 16 + my_s_in_chunk;
          const bool active = ((lane & 1) == 0) && s < S;
          const size_t row = (size_t)t * S + s;
          float od[4] = {0.f, 0.f, 0.f, 0.f};
          if (active) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (ru * 4 + q < nr) od[q] = D.dbuf[row * D.lddb + 7 * C + j0 + ru * 4 + q];
          }
          float acc[4][16], sum[4];
          unit_dot<4>(acc, w1 + (size_t)ru * 4 * G4, G4, G4, xT, SP, ch * 16, lane);
          unit_reduce<4>(acc, sum, lane);
          if (active) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = ru * 4 + q;
              if (j < nr) {
                const float dr = od[q] + sum[q];
                D.dbuf[row * D.lddb + 7 * C + j0 + j] = dr;
                st_pub(D.xb + ((size_t)t * R + j0 + j) * SX + s, dr);      // publish d_r(t)
              }
            }
          }
        }
        __syncthreads();
        stage_poll<8>(xT, SP, D.xb + (size_t)t * R * SX, R, SX, s0, sg >> 2);
        __syncthreads();
      }
      // ---------------- d_m(t) for own cells, then the cell derivative chain
      const int ncu = (nc + 3) >> 2;
      for (int u = warp; u < ncu * nchunks; u += NW) {
        const int cu = u / nchunks, ch = u - cu * nchunks;
        const int s = s0 + ch * 16 + my_s_in_chunk;
        const bool active = ((lane & 1) == 0) && s < S;
        const size_t row = (size_t)t * S + s;
        // issue the activation loads before the contraction so their latency hides behind it
        float yv[4][7];   // g, i, f, o, h at t ; c(tp) ; f(tn)
        float od[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cl = cu * 4 + q;
          const bool ok = active && cl < nc;
          const int cc = c0 + (ok ? cl : 0);
          const float* y = D.buf + (ok ? row : 0) * D.ldb + cc;
          yv[q][0] = ok ? y[0] : 0.f;       yv[q][1] = ok ? y[C] : 0.f;
          yv[q][2] = ok ? y[2 * C] : 0.f;   yv[q][3] = ok ? y[3 * C] : 0.f;
          yv[q][4] = ok ? y[5 * C] : 0.f;
          yv[q][5] = ok ? D.buf[((size_t)tp * S + s) * D.ldb + 4 * C + cc] : 0.f;
          yv[q][6] = ok ? D.buf[((size_t)tn * S + s) * D.ldb + 2 * C + cc] : 0.f;
          od[q] = (ok && R == 0) ? D.dbuf[row * D.lddb + 6 * C + cc] : 0.f;   // out_diff preloaded in the m columns
        }
        float acc[4][16], sum[4];
        if (R > 0) unit_dot<4>(acc, w2 + (size_t)cu * 4 * R, R, R, xT, SP, ch * 16, lane);
        else       unit_dot<4>(acc, w1 + (size_t)cu * 4 * G4, G4, G4, xT, SP, ch * 16, lane);
        unit_reduce<4>(acc, sum, lane);
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int cl = cu * 4 + q;
            if (cl < nc) {
              const int cc = c0 + cl;
              const float yg = yv[q][0], yi = yv[q][1], yf = yv[q][2], yo = yv[q][3], yh = yv[q][4];
              const float c_prev = yv[q][5], yf_next = yv[q][6];
              float* d = D.dbuf + row * D.lddb + cc;
              const float dm = sum[q] + od[q];
              const float dc_n = st[(0 * D.cb + cl) * SX + s], di_n = st[(1 * D.cb + cl) * SX + s], df_n = st[(2 * D.cb + cl) * SX + s];
              float dh = dm * yo;  dh = (1.0f - yh * yh) * dh;                 // DiffTanh(y_h, d_h)
              float dout = dm * yh;  dout = yo * (1.0f - yo) * dout;           // DiffSigmoid(y_o, d_o)
              const float dc = dh + dc_n * yf_next + di_n * pst[cl] + df_n * pst[D.cb + cl] + dout * pst[2 * D.cb + cl];
              float df = dc * c_prev;  df = yf * (1.0f - yf) * df;
              float di = dc * yg;      di = yi * (1.0f - yi) * di;
              float dg = dc * yi;      dg = (1.0f - yg * yg) * dg;
              d[0] = dg; d[C] = di; d[2 * C] = df; d[3 * C] = dout; d[4 * C] = dc; d[5 * C] = dh; d[6 * C] = dm;
              st[(0 * D.cb + cl) * SX + s] = dc; st[(1 * D.cb + cl) * SX + s] = di; st[(2 * D.cb + cl) * SX + s] = df;
              float* x = D.xa + ((size_t)t * G4 + cc) * SX + s;                // publish dgifo(t)
              st_pub(x, dg); st_pub(x + (size_t)C * SX, di); st_pub(x + (size_t)2 * C * SX, df); st_pub(x + (size_t)3 * C * SX, dout);
            }
          }
        }
      }
      if (L.timing != nullptr && threadIdx.x == 0) {
        const long long k4 = clock64();
        tacc[0] += k1 - k0; tacc[1] += k2 - k1; tacc[2] += k3 - k2; tacc[3] += k4 - k3;
      }
    }
  }
  if (L.timing != nullptr && threadIdx.x == 0)
    for (int q = 0; q < 8; ++q) L.timing[(size_t)blockIdx.x * 12 + q] = tacc[q];
}

#include "lstm_mma.cuh"
#include "lstm_bwd_t.cuh"

// ---------------------------------------------------------------- host side
struct Plan { Launch L; size_t smem; size_t ws_bytes; bool mma; MmaChoice mc; bool presplit; bool transposed; int tn_ns; };

size_t ws_per_dir(int T, int S, int C, int R, bool bwd) {
  const size_t SX = (size_t)(S + 3) / 4 * 4;
  const size_t da = bwd ? 4 * (size_t)C : (size_t)C;
  size_t bytes = ((size_t)(T + 2) * da * SX + (size_t)(T + 2) * R * SX) * sizeof(float);
  // transposed backward form: ring of TR slots x [chains][C/16][C/16][16][streams per chain]; chains x streams <= 2S + 8
  if (bwd && R == 0 && C % 16 == 0) {
    const size_t ring = (size_t)4 * (C / 16) * (C / 16) * 16 * (2 * (size_t)S + 8) * sizeof(float);
    if (ring > bytes) bytes = ring;
  }
  return bytes;
}

void fill_dir(DirDev& D, const aslp_lstm_dir_t& a) {
  D.T = a.T; D.S = a.S; D.C = a.C; D.R = a.R; D.Rr = a.R > 0 ? a.R : a.C; D.reverse = a.reverse;
  D.buf = a.buf; D.ldb = a.ldb; D.dbuf = a.dbuf; D.lddb = a.lddb;
  D.w_r = a.w_r; D.ldwr = a.ldwr; D.w_rm = a.w_rm; D.ldwrm = a.ldwrm;
  D.peep_i = a.peep_i; D.peep_f = a.peep_f; D.peep_o = a.peep_o; D.seq_len = a.seq_len_dev; D.clip = a.cell_clip;
  D.SX = (a.S + 3) / 4 * 4;
}

// tensor-core form: pick the largest number of parallel stream groups whose per-CTA weight slice still fits the
// register budget (more groups = fewer CTAs per chain = more cells per CTA, but proportionally less exchange traffic)
int make_plan_mma(const aslp_lstm_dir_t* dirs, int ndirs, bool bwd, void* ws, size_t ws_bytes, Plan* P) {
  Launch& L = P->L;
  P->mma = false; P->presplit = false; P->transposed = false; P->tn_ns = 0;
  L.ndirs = ndirs;
  const int C = dirs[0].C, S = dirs[0].S;
  for (int i = 0; i < ndirs; ++i)
    if (dirs[i].R != 0 || dirs[i].C != C || dirs[i].S != S || C % 8 != 0) return ASLP_STATUS_INVALID_VALUE;
  const int sms = aslp_num_sms();
  for (int pg = (S + 7) / 8; pg >= 1; --pg) {
    int SGP = (S + pg - 1) / pg;
    SGP = SGP <= 8 ? 8 : (SGP + 15) / 16 * 16;
    if ((pg - 1) * SGP >= S) continue;                    // would leave a group without streams
    int nblk = sms / (ndirs * pg);
    if (nblk < 1) continue;
    if (nblk > C) nblk = C;
    int cb = (C + nblk - 1) / nblk;
    // fill the last 16-row MMA tile (fwd: 4 rows per cell, bwd: 1): same work per CTA, fewer CTAs in the exchange
    cb = bwd ? (cb + 15) / 16 * 16 : (cb + 3) / 4 * 4;
    if (cb > C) cb = C;
    nblk = (C + cb - 1) / cb;
    MmaChoice mc;
    if (!mma_fits(C, cb, bwd, &mc)) continue;
    int SG = SGP;
    while (SG > 16 && (cb * SG > NT || mma_smem_floats(C, cb, SG, SGP, bwd) * sizeof(float) > 220 * 1024)) SG -= 16;
    if (cb * SG > NT || mma_smem_floats(C, cb, SG, SGP, bwd) * sizeof(float) > 220 * 1024) continue;
    L.nblk = nblk; L.pgroups = pg; L.SGP = SGP; L.SG = SG; L.SP = SG == 8 ? 8 : SG + 8;
    size_t need_ws = 0;
    char* wsp = (char*)ws;
    for (int i = 0; i < ndirs; ++i) {
      DirDev& D = L.d[i];
      fill_dir(D, dirs[i]);
      D.cb = cb; D.rb = 0;
      const size_t da = bwd ? 4 * (size_t)C : (size_t)C;
      D.xa = (float*)(wsp + need_ws);
      need_ws += (size_t)(D.T + 2) * da * D.SX * sizeof(float);
      D.xb = nullptr;
    }
    if (ws == nullptr || ws_bytes < need_ws) { aslp_set_last_error_msg("LSTM workspace too small", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
    // backward: fast contraction form (float4 fragment layout, immediate B addresses) when the shape allows
    static const bool no_presplit = std::getenv("ASLP_LSTM_NO_PRESPLIT") != nullptr;
    P->presplit = bwd && !no_presplit && mma_presplit_shape(C, SG, SGP) && mma_pick_kernel(mc, true, true) != nullptr &&
                  mma_smem_floats(C, cb, SG, SGP, bwd, true) * sizeof(float) <= 224 * 1024;
    P->smem = mma_smem_floats(C, cb, SG, SGP, bwd, P->presplit) * sizeof(float);
    P->ws_bytes = need_ws;
    P->mc = mc;
    P->mma = true;
    return 0;
  }
  return ASLP_STATUS_INVALID_VALUE;
}

// transposed backward form (lstm_bwd_t.cuh): 16 cells per CTA (one MMA tile per consumer), nblk = C / 16 CTAs per chain, the
// largest number of stream groups (chains) per direction that still fits the chip; a chain carries 8, 16, 24 or 32 streams
int make_plan_bwd_t(const aslp_lstm_dir_t* dirs, int ndirs, void* ws, size_t ws_bytes, Plan* P) {
  Launch& L = P->L;
  P->mma = false; P->presplit = false; P->transposed = false; P->tn_ns = 0;
  const char* tenv = std::getenv("ASLP_LSTM_BWD_T");
  if (tenv != nullptr && tenv[0] == '0') return ASLP_STATUS_INVALID_VALUE;
  const int C = dirs[0].C, S = dirs[0].S, T = dirs[0].T;
  for (int i = 0; i < ndirs; ++i)
    if (dirs[i].R != 0 || dirs[i].C != C || dirs[i].S != S || dirs[i].T != T) return ASLP_STATUS_INVALID_VALUE;
  if (C % 16 != 0 || C / 16 > 32) return ASLP_STATUS_INVALID_VALUE;
  const int nblk = C / 16, sms = aslp_num_sms();
  const size_t per_dir = ws_per_dir(T, S, C, 0, true);
  for (int pg = (S + 7) / 8; pg >= 1; --pg) {
    const int SW = (((S + pg - 1) / pg) + 7) / 8 * 8;
    if (SW > 32) break;                                  // fewer groups only make the chains wider
    if ((pg - 1) * SW >= S) continue;                    // would leave a chain without streams
    if (ndirs * pg * nblk > sms) continue;
    const int ns = SW / 8;
    if (ns == 1 && nblk > 24) continue;                  // the 8-stream kernel has at most three m-tiles per warp
    const size_t smem = (ns == 1 ? bwd_t_smem_floats(nblk) : bwd_tn_smem_floats(nblk, ns)) * sizeof(float);
    if (smem > 224 * 1024) continue;
    const size_t ring = (size_t)TR * pg * nblk * nblk * 16 * SW * sizeof(float);
    if (ring > per_dir) continue;
    if (ws == nullptr || ws_bytes < per_dir * ndirs) { aslp_set_last_error_msg("LSTM workspace too small", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
    L.ndirs = ndirs; L.nblk = nblk; L.pgroups = pg; L.SGP = SW; L.SG = SW; L.SP = SW;
    for (int i = 0; i < ndirs; ++i) {
      DirDev& D = L.d[i];
      fill_dir(D, dirs[i]);
      D.cb = 16; D.rb = 0;
      D.xa = (float*)((char*)ws + per_dir * i);
      D.xb = nullptr;
    }
    P->smem = smem; P->ws_bytes = per_dir * ndirs; P->transposed = true; P->tn_ns = ns; P->mma = true;
    P->mc.mt = 1; P->mc.kt = 8;
    return 0;
  }
  return ASLP_STATUS_INVALID_VALUE;
}

int make_plan(const aslp_lstm_dir_t* dirs, int ndirs, bool bwd, void* ws, size_t ws_bytes, Plan* P) {
  Launch& L = P->L;
  P->mma = false; P->presplit = false; P->transposed = false; P->tn_ns = 0;
  L.pgroups = 1; L.SGP = 0;
  L.ndirs = ndirs;
  const int sms = aslp_num_sms();
  int nblk = sms / ndirs;
  int maxS = 0;
  for (int i = 0; i < ndirs; ++i) {
    // fewer CTAs than SM share if there are not enough cells
    if (dirs[i].C < nblk) nblk = dirs[i].C;
    if (dirs[i].S > maxS) maxS = dirs[i].S;
  }
  if (nblk < 1) nblk = 1;
  L.nblk = nblk;
  size_t need_ws = 0, max_smem = 0;
  // stream group: the largest multiple of 16 streams whose staging buffer still fits shared memory
  int SG = ((maxS + 15) / 16) * 16;
  char* wsp = (char*)ws;
  for (;; SG -= 16) {
    max_smem = 0; need_ws = 0;
    for (int i = 0; i < ndirs; ++i) {
      const aslp_lstm_dir_t& a = dirs[i];
      DirDev& D = L.d[i];
      D.T = a.T; D.S = a.S; D.C = a.C; D.R = a.R; D.Rr = a.R > 0 ? a.R : a.C; D.reverse = a.reverse;
      D.buf = a.buf; D.ldb = a.ldb; D.dbuf = a.dbuf; D.lddb = a.lddb;
      D.w_r = a.w_r; D.ldwr = a.ldwr; D.w_rm = a.w_rm; D.ldwrm = a.ldwrm;
      D.peep_i = a.peep_i; D.peep_f = a.peep_f; D.peep_o = a.peep_o; D.seq_len = a.seq_len_dev; D.clip = a.cell_clip;
      D.SX = (a.S + 3) / 4 * 4;
      D.cb = (a.C + nblk - 1) / nblk;
      D.rb = a.R > 0 ? (a.R + nblk - 1) / nblk : 0;
      const size_t da = bwd ? 4 * (size_t)a.C : (size_t)a.C;
      D.xa = (float*)(wsp + need_ws);
      need_ws += (size_t)(a.T + 2) * da * D.SX * sizeof(float);
      D.xb = (float*)(wsp + need_ws);
      need_ws += (size_t)(a.T + 2) * a.R * D.SX * sizeof(float);
      const int SP = SG + 4;
      const size_t ncp = (D.cb + 3) & ~3, nrp = (D.rb + 3) & ~3;
      size_t fl;
      if (!bwd) fl = (size_t)D.cb * 4 * D.Rr + nrp * a.C + (size_t)max(D.Rr, a.C) * SP + (size_t)D.cb * D.SX + 3 * D.cb;
      else fl = (a.R > 0 ? nrp : ncp) * 4 * (size_t)a.C + (a.R > 0 ? ncp * a.R : 0) + (size_t)max(4 * a.C, a.R) * SP + (size_t)3 * D.cb * D.SX + 3 * D.cb;
      if (fl * sizeof(float) > max_smem) max_smem = fl * sizeof(float);
    }
    if (max_smem <= 220 * 1024 || SG <= 16) break;
  }
  L.SG = SG; L.SP = SG + 4;
  P->smem = max_smem;
  P->ws_bytes = need_ws;
  if (max_smem > 227 * 1024) { aslp_set_last_error_msg("LSTM slice does not fit shared memory (C/R too large for this SM count)", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  if (ws == nullptr || ws_bytes < need_ws) { aslp_set_last_error_msg("LSTM workspace too small", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  return 0;
}

int init_exchange(cudaStream_t st, const Plan& P, bool bwd) {
  for (int i = 0; i < P.L.ndirs; ++i) {
    const DirDev& D = P.L.d[i];
    const int blocks = aslp_num_sms() * 8;
    if (!bwd) {
      const int slot0 = D.reverse ? D.T + 1 : 0;
      // m exchange: boundary m from buf col 6C ; r exchange: boundary r from col 7C
      xch_init_kernel<<<blocks, 256, 0, st>>>(D.xa, D.C, D.T, D.S, D.SX, slot0, D.buf, D.ldb, 6 * D.C);
      ASLP_CHECK_LAUNCH();
      if (D.R > 0) {
        xch_init_kernel<<<blocks, 256, 0, st>>>(D.xb, D.R, D.T, D.S, D.SX, slot0, D.buf, D.ldb, 7 * D.C);
        ASLP_CHECK_LAUNCH();
      }
    } else {
      const int slotn = D.reverse ? 0 : D.T + 1;       // derivative boundary is zero
      xch_init_kernel<<<blocks, 256, 0, st>>>(D.xa, 4 * D.C, D.T, D.S, D.SX, slotn, nullptr, 0, 0);
      ASLP_CHECK_LAUNCH();
      if (D.R > 0) {
        xch_init_kernel<<<blocks, 256, 0, st>>>(D.xb, D.R, D.T, D.S, D.SX, slotn, nullptr, 0, 0);
        ASLP_CHECK_LAUNCH();
      }
    }
  }
  return 0;
}

// optional per-launch device timing of the persistent kernels (bench.py's roofline object): CUDA events on the
// launching stream around the cooperative launch only; resolved lazily in aslp_lstm_profile_read().
struct ProfRec { cudaEvent_t a, b; bool bwd; };
bool g_prof_on = false;
long long* g_timing = nullptr;     // device buffer for the per-CTA phase clocks (debug hook), [grid][8]
std::vector<ProfRec> g_prof;

int run(aslp_stream_t s, const aslp_lstm_dir_t* dirs, int ndirs, void* ws, size_t ws_bytes, bool bwd) {
  cudaStream_t st = (cudaStream_t)s;
  ASLP_REQUIRE(ndirs == 1 || ndirs == 2);
  for (int i = 0; i < ndirs; ++i) {
    ASLP_REQUIRE(dirs[i].T > 0 && dirs[i].S > 0 && dirs[i].C > 0 && dirs[i].R >= 0);
    ASLP_REQUIRE(dirs[i].buf != nullptr && dirs[i].w_r != nullptr);
    ASLP_REQUIRE(dirs[i].R == 0 || dirs[i].w_rm != nullptr);
    ASLP_REQUIRE(!bwd || dirs[i].dbuf != nullptr);
  }
  Plan P;
  // ASLP_LSTM_KERNEL=simt forces the SIMT contraction (tests run both forms); default: tensor-core form when it applies
  const char* force = std::getenv("ASLP_LSTM_KERNEL");
  const bool want_mma = !(force != nullptr && std::strcmp(force, "simt") == 0);
  int rc = (want_mma && bwd) ? make_plan_bwd_t(dirs, ndirs, ws, ws_bytes, &P) : ASLP_STATUS_INVALID_VALUE;
  if (rc != 0) rc = want_mma ? make_plan_mma(dirs, ndirs, bwd, ws, ws_bytes, &P) : ASLP_STATUS_INVALID_VALUE;
  if (rc != 0) rc = make_plan(dirs, ndirs, bwd, ws, ws_bytes, &P);
  if (rc != 0) return rc;
  if (force != nullptr && std::strcmp(force, "mma") == 0 && !P.mma) {
    aslp_set_last_error_msg("ASLP_LSTM_KERNEL=mma but the tensor-core recurrence does not apply to this shape", __FILE__, __LINE__);
    return ASLP_STATUS_INVALID_VALUE;
  }
  if (P.transposed) {
    const size_t per_slot = (size_t)P.L.pgroups * P.L.nblk * P.L.nblk * 16 * P.L.SGP;
    for (int i = 0; i < P.L.ndirs; ++i) {
      xch_t_init_kernel<<<aslp_num_sms(), 256, 0, st>>>(P.L.d[i].xa, per_slot, per_slot * TR);
      ASLP_CHECK_LAUNCH();
    }
  } else {
    rc = init_exchange(st, P, bwd);
    if (rc != 0) return rc;
  }
  void* kfn = P.transposed ? (P.tn_ns == 1 ? (void*)lstm_bwd_t_kernel : P.tn_ns == 2 ? (void*)lstm_bwd_tn_kernel<2>
                                              : P.tn_ns == 3 ? (void*)lstm_bwd_tn_kernel<3> : (void*)lstm_bwd_tn_kernel<4>)
                           : (P.mma ? mma_pick_kernel(P.mc, bwd, P.presplit) : (bwd ? (void*)lstm_bwd_kernel : (void*)lstm_fwd_kernel));
  if (kfn == nullptr) { aslp_set_last_error_msg("no tensor-core recurrence kernel for this shape", __FILE__, __LINE__); return ASLP_STATUS_UNKNOWN_ERROR; }
  ASLP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  P.L.timing = g_timing;
  void* args[] = {(void*)&P.L};
  // cooperative launch: guarantees all CTAs are co-resident (the polling exchange needs that)
  ProfRec rec;
  if (g_prof_on) { cudaEventCreate(&rec.a); cudaEventCreate(&rec.b); rec.bwd = bwd; cudaEventRecord(rec.a, st); }
  ASLP_CUDA(cudaLaunchCooperativeKernel(kfn, dim3(P.L.nblk * P.L.pgroups * ndirs), dim3(NT), args, P.smem, st));
  ASLP_COUNT_LAUNCH();
  if (g_prof_on) { cudaEventRecord(rec.b, st); g_prof.push_back(rec); }
  return 0;
}

}  // namespace

extern "C" {

size_t aslp_lstm_workspace_bytes(int T, int S, int C, int R, int ndirs, int backward) {
  return (size_t)ndirs * ws_per_dir(T, S, C, R, backward != 0);
}
int aslp_lstm_profile(int enable) {
  g_prof_on = enable != 0;
  for (ProfRec& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  return 0;
}
int aslp_lstm_debug_timing(long long* dev_buf) { g_timing = dev_buf; return 0; }
int aslp_lstm_profile_read(double* fwd_ms, int* fwd_launches, double* bwd_ms, int* bwd_launches) {
  double f = 0, b = 0; int nf = 0, nb = 0;
  for (ProfRec& r : g_prof) {
    ASLP_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    ASLP_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.bwd) { b += ms; ++nb; } else { f += ms; ++nf; }
  }
  if (fwd_ms) *fwd_ms = f; if (fwd_launches) *fwd_launches = nf; if (bwd_ms) *bwd_ms = b; if (bwd_launches) *bwd_launches = nb;
  return 0;
}
int aslp_lstm_seq_fwd(aslp_stream_t s, const aslp_lstm_dir_t* dirs, int ndirs, void* workspace, size_t workspace_bytes) {
  return run(s, dirs, ndirs, workspace, workspace_bytes, false);
}
int aslp_lstm_seq_bwd(aslp_stream_t s, const aslp_lstm_dir_t* dirs, int ndirs, void* workspace, size_t workspace_bytes) {
  return run(s, dirs, ndirs, workspace, workspace_bytes, true);
}

}  // extern "C"
