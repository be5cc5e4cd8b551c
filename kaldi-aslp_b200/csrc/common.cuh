// Shared device/host helpers for libaslp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/aslp_b200.h"

// ---- launch accounting (bench.py reports gpu_launches from this counter) ----
extern unsigned long long g_aslp_launches;
#define ASLP_COUNT_LAUNCH() (++g_aslp_launches)

#define ASLP_CHECK_LAUNCH()                                   \
  do {                                                        \
    ASLP_COUNT_LAUNCH();                                      \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) { aslp_set_last_error(e__, __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; } \
  } while (0)

#define ASLP_CUDA(call)                                       \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) { aslp_set_last_error(e__, __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; } \
  } while (0)

#define ASLP_REQUIRE(cond)                                    \
  do { if (!(cond)) { aslp_set_last_error_msg("invalid argument: " #cond, __FILE__, __LINE__); return ASLP_STATUS_INVALID_VALUE; } } while (0)

void aslp_set_last_error(cudaError_t e, const char* file, int line);
void aslp_set_last_error_msg(const char* msg, const char* file, int line);

static inline int aslp_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
int aslp_num_sms();

// ---- device math with the reference's formulas (matrix/kaldi-vector.cc:885-936) ----
// Branch-free forms of the reference's overflow-safe split (x > 0: 1/(1+exp(-x)), else exp(x)/(exp(x)+1)); both
// branches share e = exp(-|x|) and the denominator, so selecting the numerator gives the SAME value without divergence.
// exp and the division use the SFU forms (ex2.approx on x*log2(e), rcp.approx): ~2 ulp each, i.e. <= 3e-7 on outputs in
// [0,1] -- far inside the 1e-4 parity bound -- and a third of the instructions of expf + IEEE division, which is what the
// step time of the recurrent chains and the throughput of the Sigmoid/Tanh kernels are made of.
// -DASLP_PRECISE_MATH restores expf and IEEE division.
#ifdef ASLP_PRECISE_MATH
__device__ __forceinline__ float aslp_exp(float x) { return expf(x); }
__device__ __forceinline__ float aslp_div(float a, float b) { return a / b; }
__device__ __forceinline__ float aslp_log1p_of_exp_neg(float d) { return log1pf(expf(-d)); }
#else
__device__ __forceinline__ float aslp_exp(float x) { return __expf(x); }
__device__ __forceinline__ float aslp_div(float a, float b) { return __fdividef(a, b); }
// log(1 + exp(-d)), d >= 0: the argument of the log lies in (1, 2], where lg2.approx has an absolute error < 4e-7
__device__ __forceinline__ float aslp_log1p_of_exp_neg(float d) { return __logf(1.0f + __expf(-d)); }
#endif
__device__ __forceinline__ float ref_sigmoid(float x) {
  const float e = aslp_exp(-fabsf(x));
  return aslp_div(x > 0.0f ? 1.0f : e, 1.0f + e);
}
// x > 0: -1 + 2/(1+exp(-x)^2), else 1 - 2/(1+exp(x)^2): the two are exact negations of each other at equal |x|
__device__ __forceinline__ float ref_tanh(float x) {
  const float ie = aslp_exp(-fabsf(x));
  const float v = -1.0f + aslp_div(2.0f, 1.0f + ie * ie);
  return x > 0.0f ? v : -v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses (read-once / write-once data: bypass L1 allocation)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
