// Persistent GruStreams recurrence (src/aslp-nnet/nnet-gru-streams.h:238-441) on the same machinery as lstm.cu:
// hidden units partitioned over CTAs, W_zr_h / W_m_g slices resident in shared memory for all T steps,
// sentinel-polled exchange of h(t) and g(t) (backward: dz|dr(t) and dm(t)) through L2, no grid barrier.
//   fwd : zr = sigmoid(pre_zr + h(t-1) W_zr_h^T) ; g = r .* h(t-1) ; m = tanh(pre_m + g W_m_g^T) ; h = h(t-1) - h(t-1).*z + z.*m
//   bwd : dh = od + dzr(t+1) W_zr_h + dh(t+1) - dh(t+1).*z(t+1) + dg(t+1).*r(t+1) ; dm = (1-m^2)(dh.*z) ; dg = dm W_m_g ;
//         dr = r(1-r)(dg.*h(t-1)) ; dz = z(1-z)(dh.*m - dh.*h(t-1))
#include "common.cuh"
#include "recur.cuh"

namespace {
using namespace recur;

struct GruDev {
  int T, S, H, SX;
  float* buf; int ldb;
  float* dbuf; int lddb;
  const float* w_zr_h; int ldwzr;
  const float* w_m_g; int ldwmg;
  float* xa;      // fwd: h exchange [T+2][H][SX]    | bwd: dz|dr exchange [T+2][2H][SX]
  float* xb;      // fwd: g exchange [T+2][H][SX]    | bwd: dm exchange    [T+2][H][SX]
  int hb;         // hidden units per CTA
  int SG, SP;
};

__global__ void __launch_bounds__(NT, 1) gru_fwd_kernel(GruDev D) {
  extern __shared__ float smem[];
  const int T = D.T, S = D.S, H = D.H, SX = D.SX, SP = D.SP, SG = D.SG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * D.hb, nh = max(0, min(D.hb, H - j0));
  const int hbp = (D.hb + 3) & ~3;
  float* w1 = smem;                               // [hbp*2][H] row = ul*2 + {z, r}
  float* w2 = w1 + (size_t)hbp * 2 * H;           // [hbp][H]
  float* xT = w2 + (size_t)hbp * H;               // [H][SP]
  float* hst = xT + (size_t)H * SP;               // [hb][SX] h(t-1) of own units
  float* zst = hst + (size_t)D.hb * SX;           // [hb][SX] z(t) of own units
  for (int i = threadIdx.x; i < hbp * 2 * H; i += NT) {
    const int row = i / H, k = i - row * H;
    const int ul = row >> 1, g = row & 1;
    w1[i] = (ul < nh) ? D.w_zr_h[(size_t)(g * H + j0 + ul) * D.ldwzr + k] : 0.f;
  }
  for (int i = threadIdx.x; i < hbp * H; i += NT) {
    const int row = i / H, k = i - row * H;
    w2[i] = (row < nh) ? D.w_m_g[(size_t)(j0 + row) * D.ldwmg + k] : 0.f;
  }
  for (int i = threadIdx.x; i < D.hb * SX; i += NT) {
    const int ul = i / SX, s = i - ul * SX;
    hst[i] = (ul < nh && s < S) ? D.buf[(size_t)s * D.ldb + 4 * H + j0 + ul] : 0.f;     // row block 0 = carried state
  }
  __syncthreads();
  const int ngroups = (S + SG - 1) / SG;
  const int my_s = lane_stream(lane);
  for (int t = 1; t <= T; ++t) {
    for (int grp = 0; grp < ngroups; ++grp) {
      const int s0 = grp * SG, sg = min(SG, SX - s0), nchunks = (min(SG, S - s0) + 15) / 16;
      __syncthreads();
      stage_poll<8>(xT, SP, D.xa + (size_t)(t - 1) * H * SX, H, SX, s0, sg >> 2);
      __syncthreads();
      const int npair = (nh + 1) >> 1;
      for (int u = warp; u < npair * nchunks; u += NW) {
        const int p = u / nchunks, ch = u - p * nchunks;
        const int s = s0 + ch * 16 + my_s;
        const bool active = ((lane & 1) == 0) && s < S;
        const size_t row = (size_t)t * S + s;
        float pre[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
#pragma unroll
          for (int q = 0; q < 2; ++q) if (2 * p + q < nh) { pre[2 * q] = D.buf[row * D.ldb + j0 + 2 * p + q]; pre[2 * q + 1] = D.buf[row * D.ldb + H + j0 + 2 * p + q]; }
        }
        float acc[4][16], sum[4];
        unit_dot<4>(acc, w1 + (size_t)p * 4 * H, H, H, xT, SP, ch * 16, lane);
        unit_reduce<4>(acc, sum, lane);
        if (active) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int ul = 2 * p + q;
            if (ul < nh) {
              const float z = ref_sigmoid(pre[2 * q] + sum[2 * q]), r = ref_sigmoid(pre[2 * q + 1] + sum[2 * q + 1]);
              const float g = r * hst[ul * SX + s];
              float* o = D.buf + row * D.ldb + j0 + ul;
              o[0] = z; o[H] = r; o[3 * H] = g;
              zst[ul * SX + s] = z;
              st_pub(D.xb + ((size_t)t * H + j0 + ul) * SX + s, g);
            }
          }
        }
      }
      __syncthreads();
      stage_poll<8>(xT, SP, D.xb + (size_t)t * H * SX, H, SX, s0, sg >> 2);
      __syncthreads();
      const int nquad = (nh + 3) >> 2;
      for (int u = warp; u < nquad * nchunks; u += NW) {
        const int qd = u / nchunks, ch = u - qd * nchunks;
        const int s = s0 + ch * 16 + my_s;
        const bool active = ((lane & 1) == 0) && s < S;
        const size_t row = (size_t)t * S + s;
        float pre[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (4 * qd + q < nh) pre[q] = D.buf[row * D.ldb + 2 * H + j0 + 4 * qd + q];
        }
        float acc[4][16], sum[4];
        unit_dot<4>(acc, w2 + (size_t)qd * 4 * H, H, H, xT, SP, ch * 16, lane);
        unit_reduce<4>(acc, sum, lane);
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ul = 4 * qd + q;
            if (ul < nh) {
              const float m = ref_tanh(pre[q] + sum[q]);
              const float hp = hst[ul * SX + s], z = zst[ul * SX + s];
              float h = hp;                 // y_h = h(t-1); y_h -= h(t-1).*z; y_h += z.*m   (same operation order as the reference)
              h = h - hp * z;
              h = h + z * m;
              float* o = D.buf + row * D.ldb + j0 + ul;
              o[2 * H] = m; o[4 * H] = h;
              hst[ul * SX + s] = h;
              st_pub(D.xa + ((size_t)t * H + j0 + ul) * SX + s, h);
            }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 1) gru_bwd_kernel(GruDev D) {
  extern __shared__ float smem[];
  const int T = D.T, S = D.S, H = D.H, SX = D.SX, SP = D.SP, SG = D.SG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * D.hb, nh = max(0, min(D.hb, H - j0));
  const int hbp = (D.hb + 3) & ~3;
  float* w1 = smem;                               // [hbp][2H]  column k of W_zr_h
  float* w2 = w1 + (size_t)hbp * 2 * H;           // [hbp][H]   column k of W_m_g
  float* xT = w2 + (size_t)hbp * H;               // [2H][SP]
  float* st = xT + (size_t)2 * H * SP;            // [2][hb][SX]: dh(t+1), dg(t+1)
  for (int i = threadIdx.x; i < hbp * 2 * H; i += NT) {
    const int row = i / (2 * H), q = i - row * 2 * H;
    w1[i] = (row < nh) ? D.w_zr_h[(size_t)q * D.ldwzr + j0 + row] : 0.f;
  }
  for (int i = threadIdx.x; i < hbp * H; i += NT) {
    const int row = i / H, j = i - row * H;
    w2[i] = (row < nh) ? D.w_m_g[(size_t)j * D.ldwmg + j0 + row] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * D.hb * SX; i += NT) st[i] = 0.f;
  __syncthreads();
  const int ngroups = (S + SG - 1) / SG;
  const int my_s = lane_stream(lane);
  const int nquad = (nh + 3) >> 2;
  for (int t = T; t >= 1; --t) {
    for (int grp = 0; grp < ngroups; ++grp) {
      const int s0 = grp * SG, sg = min(SG, SX - s0), nchunks = (min(SG, S - s0) + 15) / 16;
      __syncthreads();
      stage_poll<16>(xT, SP, D.xa + (size_t)(t + 1) * 2 * H * SX, 2 * H, SX, s0, sg >> 2);
      __syncthreads();
      for (int u = warp; u < nquad * nchunks; u += NW) {
        const int qd = u / nchunks, ch = u - qd * nchunks;
        const int s = s0 + ch * 16 + my_s;
        const bool active = ((lane & 1) == 0) && s < S;
        const size_t row = (size_t)t * S + s, rown = (size_t)(t + 1) * S + s, rowp = (size_t)(t - 1) * S + s;
        float v[4][6];    // od, z, m, h(t-1), z(t+1), r(t+1)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ul = 4 * qd + q;
          const bool okk = active && ul < nh;
          const int k = j0 + (okk ? ul : 0);
          v[q][0] = okk ? D.dbuf[row * D.lddb + 4 * H + k] : 0.f;
          v[q][1] = okk ? D.buf[row * D.ldb + k] : 0.f;
          v[q][2] = okk ? D.buf[row * D.ldb + 2 * H + k] : 0.f;
          v[q][3] = okk ? D.buf[rowp * D.ldb + 4 * H + k] : 0.f;
          v[q][4] = okk ? D.buf[rown * D.ldb + k] : 0.f;
          v[q][5] = okk ? D.buf[rown * D.ldb + H + k] : 0.f;
        }
        float acc[4][16], sum[4];
        unit_dot<4>(acc, w1 + (size_t)qd * 4 * 2 * H, 2 * H, 2 * H, xT, SP, ch * 16, lane);
        unit_reduce<4>(acc, sum, lane);
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ul = 4 * qd + q;
            if (ul < nh) {
              const float dh_n = st[ul * SX + s], dg_n = st[(D.hb + ul) * SX + s];
              float dh = v[q][0] + sum[q];
              dh = dh + dh_n;
              dh = dh - dh_n * v[q][4];
              dh = dh + dg_n * v[q][5];
              const float z = v[q][1], m = v[q][2], hp = v[q][3];
              float dm = dh * z;  dm = (1.0f - m * m) * dm;
              float dz = dh * m;  dz = dz - dh * hp;  dz = z * (1.0f - z) * dz;
              float* d = D.dbuf + row * D.lddb + j0 + ul;
              d[0] = dz; d[2 * H] = dm; d[4 * H] = dh;
              st[ul * SX + s] = dh;
              st_pub(D.xb + ((size_t)t * H + j0 + ul) * SX + s, dm);
              st_pub(D.xa + ((size_t)t * 2 * H + j0 + ul) * SX + s, dz);
            }
          }
        }
      }
      __syncthreads();
      stage_poll<8>(xT, SP, D.xb + (size_t)t * H * SX, H, SX, s0, sg >> 2);
      __syncthreads();
      for (int u = warp; u < nquad * nchunks; u += NW) {
        const int qd = u / nchunks, ch = u - qd * nchunks;
        const int s = s0 + ch * 16 + my_s;
        const bool active = ((lane & 1) == 0) && s < S;
        const size_t row = (size_t)t * S + s, rowp = (size_t)(t - 1) * S + s;
        float rv[4], hp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ul = 4 * qd + q;
          const bool okk = active && ul < nh;
          rv[q] = okk ? D.buf[row * D.ldb + H + j0 + ul] : 0.f;
          hp[q] = okk ? D.buf[rowp * D.ldb + 4 * H + j0 + ul] : 0.f;
        }
        float acc[4][16], sum[4];
        unit_dot<4>(acc, w2 + (size_t)qd * 4 * H, H, H, xT, SP, ch * 16, lane);
        unit_reduce<4>(acc, sum, lane);
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ul = 4 * qd + q;
            if (ul < nh) {
              const float dg = sum[q];
              float dr = dg * hp[q];  dr = rv[q] * (1.0f - rv[q]) * dr;
              float* d = D.dbuf + row * D.lddb + j0 + ul;
              d[H] = dr; d[3 * H] = dg;
              st[(D.hb + ul) * SX + s] = dg;
              st_pub(D.xa + ((size_t)t * 2 * H + H + j0 + ul) * SX + s, dr);
            }
          }
        }
      }
    }
  }
}

size_t gru_ws(int T, int S, int H, bool bwd) {
  const size_t SX = (size_t)(S + 3) / 4 * 4;
  return (size_t)(T + 2) * ((bwd ? 2 : 1) * (size_t)H + H) * SX * sizeof(float);
}

int run_gru(aslp_stream_t s, const aslp_gru_t* g, void* ws, size_t ws_bytes, bool bwd) {
  cudaStream_t st = (cudaStream_t)s;
  ASLP_REQUIRE(g != nullptr && g->T > 0 && g->S > 0 && g->H > 0 && g->buf != nullptr && g->w_zr_h != nullptr && g->w_m_g != nullptr);
  ASLP_REQUIRE(!bwd || g->dbuf != nullptr);
  const size_t need = gru_ws(g->T, g->S, g->H, bwd);
  if (ws == nullptr || ws_bytes < need) { aslp_set_last_error_msg("GRU workspace too small", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  GruDev D;
  D.T = g->T; D.S = g->S; D.H = g->H; D.SX = (g->S + 3) / 4 * 4;
  D.buf = g->buf; D.ldb = g->ldb; D.dbuf = g->dbuf; D.lddb = g->lddb;
  D.w_zr_h = g->w_zr_h; D.ldwzr = g->ldwzr; D.w_m_g = g->w_m_g; D.ldwmg = g->ldwmg;
  const size_t da = (bwd ? 2 : 1) * (size_t)g->H;
  D.xa = (float*)ws;
  D.xb = D.xa + (size_t)(g->T + 2) * da * D.SX;
  int nblk = aslp_num_sms();
  if (g->H < nblk) nblk = g->H;
  D.hb = (g->H + nblk - 1) / nblk;
  nblk = (g->H + D.hb - 1) / D.hb;
  const size_t hbp = (D.hb + 3) & ~3;
  size_t smem = 0;
  int SG = ((g->S + 15) / 16) * 16;
  for (;; SG -= 16) {
    const size_t SP = SG + 4;
    const size_t fl = hbp * 2 * g->H + hbp * g->H + (bwd ? 2 : 1) * (size_t)g->H * SP + 2 * (size_t)D.hb * D.SX;
    smem = fl * sizeof(float);
    if (smem <= 220 * 1024 || SG <= 16) break;
  }
  if (smem > 227 * 1024) { aslp_set_last_error_msg("GRU slice does not fit shared memory", __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; }
  D.SG = SG; D.SP = SG + 4;
  const int blocks = aslp_num_sms() * 8;
  if (!bwd) {
    xch_init_kernel<<<blocks, 256, 0, st>>>(D.xa, g->H, g->T, g->S, D.SX, 0, g->buf, g->ldb, 4 * g->H);   // h boundary = carried state
    ASLP_CHECK_LAUNCH();
    xch_init_kernel<<<blocks, 256, 0, st>>>(D.xb, g->H, g->T, g->S, D.SX, -1, nullptr, 0, 0);
    ASLP_CHECK_LAUNCH();
  } else {
    xch_init_kernel<<<blocks, 256, 0, st>>>(D.xa, 2 * g->H, g->T, g->S, D.SX, g->T + 1, nullptr, 0, 0);   // dzr(T+1) = 0
    ASLP_CHECK_LAUNCH();
    xch_init_kernel<<<blocks, 256, 0, st>>>(D.xb, g->H, g->T, g->S, D.SX, -1, nullptr, 0, 0);
    ASLP_CHECK_LAUNCH();
  }
  void* kfn = bwd ? (void*)gru_bwd_kernel : (void*)gru_fwd_kernel;
  ASLP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&D};
  ASLP_CUDA(cudaLaunchCooperativeKernel(kfn, dim3(nblk), dim3(NT), args, smem, st));
  ASLP_COUNT_LAUNCH();
  return 0;
}

}  // namespace

extern "C" {
size_t aslp_gru_workspace_bytes(int T, int S, int H, int backward) { return gru_ws(T, S, H, backward != 0); }
int aslp_gru_seq_fwd(aslp_stream_t s, const aslp_gru_t* g, void* workspace, size_t workspace_bytes) { return run_gru(s, g, workspace, workspace_bytes, false); }
int aslp_gru_seq_bwd(aslp_stream_t s, const aslp_gru_t* g, void* workspace, size_t workspace_bytes) { return run_gru(s, g, workspace, workspace_bytes, true); }
}
