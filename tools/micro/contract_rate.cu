// Stand-alone timing of the backward-recurrence contraction loop (same instruction mix as mma_contract_fast<1, 20> in
// csrc/lstm_mma.cuh: per k-tile one LDS.128 of the A fragment, two LDS of B, the tf32 splits, three HMMA.1688 into
// independent accumulators), 8 warps per CTA, one CTA per SM, inputs in shared memory.  Tells whether the ~2500 cycles the
// kernel spends per step in its contraction are inherent in the loop or come from around it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o contract_rate contract_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int KTP, int MODE>     // MODE 0: full loop; 1: no splits of A (hi only, 1 MMA); 2: loads + splits only (no MMA)
__global__ void __launch_bounds__(256, 1) probe(long long* out, int iters) {
  extern __shared__ __align__(16) float sm[];
  float* xT = sm;                        // [K = 8 warps * KTP * 8][8]
  float4* w = reinterpret_cast<float4*>(sm + 8 * KTP * 8 * 8);   // [8 warps][KTP][32] float4
  float* part = reinterpret_cast<float*>(w + 8 * KTP * 32);     // [8 warps][8][20]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < 8 * KTP * 64; i += 256) xT[i] = 1e-3f * (i % 97);
  for (int i = threadIdx.x; i < 8 * KTP * 32; i += 256) w[i] = make_float4(1e-3f * (i % 13), 2e-3f, 3e-3f * (i % 7), 1e-3f);
  __syncthreads();
  const float4* wf = w + warp * KTP * 32;
  const float* xp0 = xT + (size_t)(warp * KTP * 8 + tig) * 8 + g;
  float* pw = part + warp * 8 * 20;
  long long total = 0;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    float hh[4][4], lh[4][4], hl[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) { hh[a][q] = 0.f; lh[a][q] = 0.f; hl[a][q] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KTP; ++kt) {
      const int a = kt % 4;
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(xp0[kt * 64], bh0, bl0);
      split_tf32(xp0[kt * 64 + 32], bh1, bl1);
      const float4 w4 = wf[kt * 32 + lane];
      uint32_t ah[4], al[4];
      split_tf32(w4.x, ah[0], al[0]); split_tf32(w4.y, ah[1], al[1]); split_tf32(w4.z, ah[2], al[2]); split_tf32(w4.w, ah[3], al[3]);
      if (MODE == 0) { mma_tf32(lh[a], al, bh0, bh1); mma_tf32(hl[a], ah, bl0, bl1); mma_tf32(hh[a], ah, bh0, bh1); }
      else if (MODE == 1) { mma_tf32(hh[a], ah, bh0, bh1); }
      else { hh[a][0] += __uint_as_float(al[0] ^ bl0); hh[a][1] += __uint_as_float(al[1] ^ bl1); hh[a][2] += __uint_as_float(al[2] ^ bh0); hh[a][3] += __uint_as_float(al[3] ^ bh1); }
    }
    float c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) { lo += lh[a][q] + hl[a][q]; hi += hh[a][q]; }
      c[q] = lo + hi;
    }
    float* p0 = pw + (2 * tig) * 20 + g;
    p0[0] = c[0]; p0[20] = c[1]; p0[8] = c[2]; p0[28] = c[3];
    total += clock64() - t0;
    xT[(it * 37 + threadIdx.x) % (8 * KTP * 64)] += c[0] * 1e-9f;      // keep the loop body live across iterations
  }
  if (lane == 0) out[blockIdx.x * 8 + warp] = total;
}

template <int KTP, int MODE>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8 * sizeof(long long));
  const int iters = 2000;
  const size_t smem = (size_t)(8 * KTP * 64 + 8 * KTP * 128 + 8 * 8 * 20) * sizeof(float);
  cudaFuncSetAttribute(probe<KTP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<KTP, MODE><<<148, 256, smem>>>(d, iters);
  cudaDeviceSynchronize();
  static long long h[148 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148 * 8; ++i) cyc += h[i]; cyc /= 148 * 8;
  printf("{\"loop\": \"%s\", \"k_tiles_per_warp\": %d, \"cycles_per_step\": %.0f, \"cycles_per_k_tile\": %.1f, \"smem_bytes\": %zu, \"err\": \"%s\"}\n",
         name, KTP, cyc / iters, cyc / iters / KTP, smem, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<20, 0>("3xTF32 full");
  run<20, 1>("1 MMA per tile (loads + splits kept)");
  run<20, 2>("loads + splits only");
  run<8, 0>("3xTF32 full");
  return 0;
}
