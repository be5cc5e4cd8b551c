// BatchNormalization (src/aslp-nnet/nnet-batch-normalization.h:139-284): the reference spends ~12
// elementwise launches forward and ~25 backward plus two fp32->fp64 matrix conversions; here the
// forward is ONE column-statistics pass (sum x, sum x^2 in double; the variance about the fp32 mean follows exactly from
// them) + 1 normalise pass, and the backward 1 statistics pass + 1 elementwise pass; x-hat is recomputed from x, mean and
// inv_std instead of being stored and re-read.  All 128-bit, deterministic (two-phase column reductions, no atomics).
// HBM passes over the [rows, cols] matrix: forward 2 reads + 1 write (12 B/elem, 8 algorithmic), backward 4 reads + 1
// write (20 B/elem, 16 algorithmic; the second read of a pass pair hits L2 when the minibatch fits it).
#include "common.cuh"
#include "scratch.cuh"

namespace {

constexpr int MAXV = 4;      // column sums produced per pass

__device__ __forceinline__ float4 ld4z(const float* p, int nv) {
  if (nv == 4) return *reinterpret_cast<const float4*>(p);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nv > 0) v.x = p[0];
  if (nv > 1) v.y = p[1];
  if (nv > 2) v.z = p[2];
  return v;
}
__device__ __forceinline__ void st4z(float* p, float4 v, int nv) {
  if (nv == 4) { *reinterpret_cast<float4*>(p) = v; return; }
  if (nv > 0) p[0] = v.x;
  if (nv > 1) p[1] = v.y;
  if (nv > 2) p[2] = v.z;
}

// MODE 0 (fwd): v0 = sum (double)x, v1 = sum (double)fl(x*x) (the reference's running sum), v2 = sum (double)x*(double)x
// MODE 2 (bwd): v0 = sum xhat*dy, v1 = sum dy, v2 = sum (x-mean)*dy, v3 = sum (x-mean);  xhat = (x-mean)*inv_std (a = inv_std)
// block = 256 threads = 32 column quads x 8 row lanes ; partial[chunk][v][colpad] in double
template <int MODE>
__global__ void bn_partial_kernel(double* partial, const float* x, int ldx, const float* a, int lda, const float* b, int ldb,
                                  const float* mean, int rows, int cols, int rows_per_chunk) {
  __shared__ double red[8][32][4];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cq) * 4;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  const int colpad = ((cols + 3) >> 2) << 2;
  constexpr int NV = (MODE == 0) ? 3 : 4;
  double acc[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[v][j] = 0.0;
  if (c < cols) {
    const int nv = min(4, cols - c);
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, iv[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE != 0) {
      const float4 m4 = ld4z(mean + c, nv); mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
      const float4 i4 = ld4z(a + c, nv); iv[0] = i4.x; iv[1] = i4.y; iv[2] = i4.z; iv[3] = i4.w;
    }
    // four rows per iteration, all loads issued before the first use
    constexpr int U = 4;
    for (int r = r0 + rl; r < r1; r += 8 * U) {
      float4 x4[U], d4[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + 8 * u;
        const bool live = rr < r1;
        x4[u] = live ? ld4z(x + (size_t)rr * ldx + c, nv) : make_float4(0.f, 0.f, 0.f, 0.f);
        d4[u] = (MODE == 2 && live) ? ld4z(b + (size_t)rr * ldb + c, nv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + 8 * u < r1) {
          const float xv[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w};
          const float dv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double xd = (double)xv[j];
              acc[0][j] += xd; acc[1][j] += (double)(xv[j] * xv[j]); acc[2][j] += xd * xd;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float xm = xv[j] - mu[j];
              const float hv = xm * iv[j];
              acc[0][j] += (double)(hv * dv[j]); acc[1][j] += (double)dv[j]; acc[2][j] += (double)(xm * dv[j]); acc[3][j] += (double)xm;
            }
          }
        }
      }
    }
  }
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rl][cq][j] = acc[v][j];
    __syncthreads();
    if (rl == 0 && c < cols) {
      double* dst = partial + ((size_t)blockIdx.y * MAXV + v) * colpad + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k][cq][j];
        dst[j] = s;
      }
    }
  }
}

// Finalize kernels: block = 32 columns x 8 chunk lanes.  Every lane sums each 8th chunk partial (independent loads), the 8
// lanes of a column meet in shared memory in a fixed order (deterministic).  The one-thread-per-column form was a chain of
// chunks x NV dependent L2 loads: 38-44 us for 74 chunks, longer than the statistics pass itself.
template <int NVAL>
__device__ __forceinline__ void sum_chunks_block(const double* partial, int chunks, int colpad, int c, bool col_ok, double (&out)[NVAL]) {
  __shared__ double red[NVAL][8][33];
  const int cl = threadIdx.x & 31, kl = threadIdx.x >> 5;
  double s[NVAL];
#pragma unroll
  for (int v = 0; v < NVAL; ++v) s[v] = 0.0;
  if (col_ok)
    for (int k = kl; k < chunks; k += 8)
#pragma unroll
      for (int v = 0; v < NVAL; ++v) s[v] += partial[((size_t)k * MAXV + v) * colpad + c];
#pragma unroll
  for (int v = 0; v < NVAL; ++v) red[v][kl][cl] = s[v];
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NVAL; ++v) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[v][j][cl];
    out[v] = t;
  }
}

__global__ void __launch_bounds__(256) bn_fwd_fin_kernel(const double* partial, int chunks, int rows, int cols, float var_floor, float* mean,
                                                         float* inv_std, double* acc_mean, double* acc_var) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int colpad = ((cols + 3) >> 2) << 2;
  double sv[3];
  sum_chunks_block<3>(partial, chunks, colpad, c, c < cols, sv);
  if (c >= cols || (threadIdx.x >> 5) != 0) return;
  const double s1 = sv[0], s2f = sv[1], s2 = sv[2];
  const float mu = (float)(s1 * (double)(1.0f / rows));                                     // AddRowSumMat(1/N, in)
  mean[c] = mu;
  if (acc_mean != nullptr) acc_mean[c] += s1;
  if (acc_var != nullptr) acc_var[c] += s2f;
  // (1/N) sum (x - mu)^2 about the ROUNDED fp32 mean, as the reference's second pass computes it, from the exact sums
  const double m = (double)mu;
  double ssq = s2 - 2.0 * m * s1 + (double)rows * m * m;
  if (ssq < 0.0) ssq = 0.0;
  float var = (float)(ssq * (double)(1.0f / rows));
  var += var_floor;                         // Add(var_floor); ApplyPow(0.5); InvertElements()
  inv_std[c] = 1.0f / sqrtf(var);
}

__global__ void bn_normalize_kernel(float* out, int ldo, float* xhat, int ldh, const float* in, int ldi, int rows, int cols,
                                    const float* scale, const float* shift, const float* mean, const float* inv_std) {
  const int n4 = (cols + 3) >> 2;
  const long long total = (long long)rows * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i - (long long)r * n4) << 2, nv = min(4, cols - c);
    const float4 x = ld4z(in + (size_t)r * ldi + c, nv), m = ld4z(mean + c, nv), iv = ld4z(inv_std + c, nv);
    const float4 sc = ld4z(scale + c, nv), sh = ld4z(shift + c, nv);
    float4 h, o;
    h.x = (x.x - m.x) * iv.x; h.y = (x.y - m.y) * iv.y; h.z = (x.z - m.z) * iv.z; h.w = (x.w - m.w) * iv.w;
    o.x = h.x * sc.x + sh.x; o.y = h.y * sc.y + sh.y; o.z = h.z * sc.z + sh.z; o.w = h.w * sc.w + sh.w;
    if (xhat != nullptr) st4z(xhat + (size_t)r * ldh + c, h, nv);
    st4z(out + (size_t)r * ldo + c, o, nv);
  }
}

// bwd finalize: dscale/dshift (momentum) and the per-column dvar, dmean used by the elementwise pass
__global__ void __launch_bounds__(256) bn_bwd_fin_kernel(const double* partial, int chunks, int rows, int cols, const float* scale,
                                                         const float* inv_std, float momentum, float* dscale, float* dshift,
                                                         float* dvar_out, float* dmean_out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int colpad = ((cols + 3) >> 2) << 2;
  double sv[4];
  sum_chunks_block<4>(partial, chunks, colpad, c, c < cols, sv);
  if (c >= cols || (threadIdx.x >> 5) != 0) return;
  const float s_hd = (float)sv[0], s_d = (float)sv[1], s_xd = (float)sv[2], s_x = (float)sv[3];
  dscale[c] = momentum * dscale[c] + s_hd;
  dshift[c] = momentum * dshift[c] + s_d;
  const float iv = inv_std[c], sc = scale[c];
  const float dvar = -0.5f * iv * iv * iv * sc * s_xd;               // sum (x-mean) * (dy*scale) * (-0.5 ivar^3)
  const float dmean = -(sc * iv) * s_d - (2.0f / rows) * dvar * s_x; // sum(-g*ivar) - sum (x-mean)*(2/N)*dvar
  dvar_out[c] = dvar; dmean_out[c] = dmean;
}
__global__ void bn_bwd_apply_kernel(float* din, int ldd, const float* in, int ldi, const float* dy, int ldo, int rows, int cols,
                                    const float* scale, const float* mean, const float* inv_std, const float* dvar, const float* dmean) {
  const int n4 = (cols + 3) >> 2;
  const long long total = (long long)rows * n4;
  const float two_n = 2.0f / rows, inv_n = 1.0f / rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i - (long long)r * n4) << 2, nv = min(4, cols - c);
    const float4 x = ld4z(in + (size_t)r * ldi + c, nv), d = ld4z(dy + (size_t)r * ldo + c, nv);
    const float4 sc = ld4z(scale + c, nv), m = ld4z(mean + c, nv), iv = ld4z(inv_std + c, nv), dv = ld4z(dvar + c, nv), dm = ld4z(dmean + c, nv);
    float4 o;
    o.x = d.x * sc.x * iv.x + (x.x - m.x) * two_n * dv.x + inv_n * dm.x;
    o.y = d.y * sc.y * iv.y + (x.y - m.y) * two_n * dv.y + inv_n * dm.y;
    o.z = d.z * sc.z * iv.z + (x.z - m.z) * two_n * dv.z + inv_n * dm.z;
    o.w = d.w * sc.w * iv.w + (x.w - m.w) * two_n * dv.w + inv_n * dm.w;
    st4z(din + (size_t)r * ldd + c, o, nv);
  }
}

struct RedPlan { int col_blocks, chunks, rows_per_chunk; size_t bytes; };
// one full wave of the statistics kernel: chunks = resident blocks (occupancy query: the kernels are register-limited to
// 2-4 blocks per SM) / column blocks
template <typename KernelT>
RedPlan plan_reduce(KernelT kernel, int rows, int cols) {
  RedPlan p;
  p.col_blocks = aslp_div_up(cols, 128);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  int chunks = (aslp_num_sms() * per_sm) / p.col_blocks;
  if (chunks > aslp_div_up(rows, 32)) chunks = aslp_div_up(rows, 32);
  if (chunks < 1) chunks = 1;
  p.rows_per_chunk = aslp_div_up(rows, chunks);
  p.chunks = aslp_div_up(rows, p.rows_per_chunk);
  p.bytes = (size_t)p.chunks * MAXV * ((cols + 3) / 4 * 4) * sizeof(double);
  return p;
}
inline int ew_grid(long long rows, long long cols) {
  long long b = (rows * ((cols + 3) / 4) + 255) / 256;
  const long long cap = (long long)aslp_num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

int aslp_bn_fwd_train(aslp_stream_t s, float* out, int ldo, float* xhat, int ldx, const float* in, int ldi, int rows, int cols,
                      const float* scale, const float* shift, float var_floor, float* mean, float* inv_std, double* acc_mean,
                      double* acc_var) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldo % 4 == 0 && ldi % 4 == 0 && (xhat == nullptr || ldx % 4 == 0));
  cudaStream_t st = (cudaStream_t)s;
  const RedPlan p = plan_reduce(bn_partial_kernel<0>, rows, cols);
  double* partial = (double*)aslp_scratch(st, p.bytes);
  if (partial == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  dim3 grid(p.col_blocks, p.chunks);
  bn_partial_kernel<0><<<grid, 256, 0, st>>>(partial, in, ldi, nullptr, 0, nullptr, 0, nullptr, rows, cols, p.rows_per_chunk);
  ASLP_CHECK_LAUNCH();
  bn_fwd_fin_kernel<<<aslp_div_up(cols, 32), 256, 0, st>>>(partial, p.chunks, rows, cols, var_floor, mean, inv_std, acc_mean, acc_var);
  ASLP_CHECK_LAUNCH();
  bn_normalize_kernel<<<ew_grid(rows, cols), 256, 0, st>>>(out, ldo, xhat, ldx, in, ldi, rows, cols, scale, shift, mean, inv_std);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_bn_fwd_eval(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int cols, const float* scale,
                     const float* shift, const float* mean, const float* inv_std) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldo % 4 == 0 && ldi % 4 == 0);
  bn_normalize_kernel<<<ew_grid(rows, cols), 256, 0, (cudaStream_t)s>>>(out, ldo, nullptr, 0, in, ldi, rows, cols, scale, shift, mean, inv_std);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_bn_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* xhat, int ldx, const float* out_diff,
                int ldo, int rows, int cols, const float* scale, const float* mean, const float* inv_std, float momentum,
                float* dscale, float* dshift) {
  if (rows == 0 || cols == 0) return 0;
  ASLP_REQUIRE(ldi % 4 == 0 && ldo % 4 == 0 && (in_diff == nullptr || ldd % 4 == 0));
  (void)xhat; (void)ldx;
  cudaStream_t st = (cudaStream_t)s;
  const RedPlan p = plan_reduce(bn_partial_kernel<2>, rows, cols);
  const size_t colpad = (size_t)(cols + 3) / 4 * 4;
  double* partial = (double*)aslp_scratch(st, p.bytes + 2 * colpad * sizeof(float));
  if (partial == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  float* dvar = (float*)((char*)partial + p.bytes);
  float* dmean = dvar + colpad;
  dim3 grid(p.col_blocks, p.chunks);
  bn_partial_kernel<2><<<grid, 256, 0, st>>>(partial, in, ldi, inv_std, 0, out_diff, ldo, mean, rows, cols, p.rows_per_chunk);   // xhat recomputed, not read
  ASLP_CHECK_LAUNCH();
  bn_bwd_fin_kernel<<<aslp_div_up(cols, 32), 256, 0, st>>>(partial, p.chunks, rows, cols, scale, inv_std, momentum, dscale, dshift, dvar, dmean);
  ASLP_CHECK_LAUNCH();
  if (in_diff != nullptr) {
    bn_bwd_apply_kernel<<<ew_grid(rows, cols), 256, 0, st>>>(in_diff, ldd, in, ldi, out_diff, ldo, rows, cols, scale, mean, inv_std, dvar, dmean);
    ASLP_CHECK_LAUNCH();
  }
  return 0;
}

}  // extern "C"
