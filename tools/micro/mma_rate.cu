// Issue-rate probe of the warp-level (legacy) MMA shapes on sm_100a: cycles per instruction per SM sub-partition with
// W warps resident per sub-partition and 4 independent accumulator chains per warp.  Development aid for the recurrence
// kernels (which contraction form to use); build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int KIND>
__global__ void probe(long long* out, int iters, float seed) {
  float c[4][4];
  for (int a = 0; a < 4; ++a) for (int q = 0; q < 4; ++q) c[a][q] = seed * (a + q);
  unsigned a0 = __float_as_uint(seed), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3]) : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int a = 0; a < 4; ++a) for (int q = 0; q < 4; ++q) s += c[a][q];
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = (long long)s; }
}

template <int KIND>
void run(const char* name, int warps_per_cta) {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  const int iters = 4096;
  probe<KIND><<<148, warps_per_cta * 32>>>(d, iters, 1e-3f);
  probe<KIND><<<148, warps_per_cta * 32>>>(d, iters, 1e-3f);
  cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[2 * i]; cyc /= 148;
  const double per_warp = cyc / (iters * 4.0);
  const double per_smsp = per_warp / (warps_per_cta / 4.0 < 1 ? 1 : warps_per_cta / 4.0);
  printf("{\"mma\": \"%s\", \"warps_per_cta\": %d, \"cycles_per_mma_per_warp\": %.2f, \"cycles_per_mma_per_smsp\": %.2f, \"err\": \"%s\"}\n",
         name, warps_per_cta, per_warp, per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("m16n8k8.tf32", w);
    run<1>("m16n8k16.f16", w);
    run<2>("m16n8k16.bf16", w);
    run<3>("m16n8k8.f16", w);
  }
  return 0;
}
