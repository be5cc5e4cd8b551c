// How fast can the recurrence's per-step all-gather through L2 be?  (development aid, not product code)
//   P1  two CTAs on different SMs bounce one word: st.relaxed.gpu -> poll with ld.relaxed.gpu / ld.volatile; cycles per hop.
//   P2  the real pattern: NC CTAs of a chain, each publishes its BYTES-byte slice of step t and gathers all NC slices (data is its
//       own flag: a NaN sentinel in every word) -- by (a) every thread spinning on its own 16-byte word(s) with ld.relaxed.gpu.v4 and
//       writing them to shared memory, (b) cp.async.cg into shared memory + inspect + re-issue (what the product kernels do).
//       Steps are data-dependent (a CTA publishes step t+1 only after it holds all of step t).  Cycles per step = the hand-off floor.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_exchange l2_exchange.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t SENT = 0x7fc0dead;     // a quiet NaN nobody computes

__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_relaxed4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed4(uint4* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- P1: ping-pong of one word between CTA 0 and CTA `peer`
__global__ void pingpong_kernel(uint32_t* flag, int rounds, int peer, long long* out) {
  if (blockIdx.x != 0 && blockIdx.x != peer) return;
  if (threadIdx.x != 0) return;
  const bool first = blockIdx.x == 0;
  const long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    if (first) {
      st_relaxed(flag, 2 * r - 1);
      while (ld_relaxed(flag) != (uint32_t)(2 * r)) {}
    } else {
      while (ld_relaxed(flag) != (uint32_t)(2 * r - 1)) {}
      st_relaxed(flag, 2 * r);
    }
  }
  if (first) out[0] = clock64() - t0;
}

// ---- P2: chain all-gather.  xch [steps % RING][NC][WORDS] uint32; WORDS = BYTES / 4 per CTA per step.
template <int MODE>
__global__ void __launch_bounds__(256, 1) gather_kernel(uint32_t* xch, int NC, int WORDS, int steps, long long* out, unsigned* check) {
  extern __shared__ uint32_t sm[];                 // [NC * WORDS]
  constexpr int RING = 4;
  const int chain = blockIdx.x / NC, me = blockIdx.x % NC;
  uint32_t* base = xch + (size_t)chain * RING * NC * WORDS;
  const int total4 = NC * WORDS / 4;               // 16-byte items to gather per step
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < steps; ++t) {
    uint32_t* slot = base + (size_t)(t % RING) * NC * WORDS;
    // publish own slice (value depends on what was gathered: keeps the chain honest)
    for (int w = threadIdx.x; w < WORDS; w += blockDim.x) st_relaxed(slot + me * WORDS + w, (uint32_t)(t * 131 + me + (acc & 1)));
    // re-arm the slot two steps ahead (its last readers finished at step t - 2: every CTA has published t - 1 since)
    {
      uint32_t* rearm = base + (size_t)((t + 2) % RING) * NC * WORDS;
      for (int w = threadIdx.x; w < WORDS; w += blockDim.x) st_relaxed(rearm + me * WORDS + w, SENT);
    }
    if (MODE == 0) {
      for (int i = threadIdx.x; i < total4; i += blockDim.x) {
        uint4 v;
        do { v = ld_relaxed4(reinterpret_cast<const uint4*>(slot) + i); } while (v.x == SENT || v.y == SENT || v.z == SENT || v.w == SENT);
        reinterpret_cast<uint4*>(sm)[i] = v;
      }
    } else {
      // cp.async.cg 16 B per item, wait, inspect, re-issue the ones that still show the sentinel
      bool pending = true;
      while (pending) {
        for (int i = threadIdx.x; i < total4; i += blockDim.x) {
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<uint4*>(sm) + i);
          const uint4 cur = reinterpret_cast<uint4*>(sm)[i];
          if (t == 0 || cur.x == SENT || cur.y == SENT || cur.z == SENT || cur.w == SENT || true)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(reinterpret_cast<const uint4*>(slot) + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        bool bad = false;
        for (int i = threadIdx.x; i < total4; i += blockDim.x) {
          const uint4 v = reinterpret_cast<uint4*>(sm)[i];
          bad |= (v.x == SENT || v.y == SENT || v.z == SENT || v.w == SENT);
        }
        pending = __syncthreads_or(bad);
      }
    }
    __syncthreads();
    acc += sm[(threadIdx.x * 7) % (NC * WORDS)];
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; check[blockIdx.x] = acc; }
}

template <int MODE>
void run_gather(int NC, int chains, int BYTES, int steps) {
  const int WORDS = BYTES / 4, grid = NC * chains;
  uint32_t* xch; long long* dout; unsigned* dchk;
  const size_t n = (size_t)chains * 4 * NC * WORDS;
  CK(cudaMalloc(&xch, n * 4)); CK(cudaMalloc(&dout, grid * 8)); CK(cudaMalloc(&dchk, grid * 4));
  std::vector<uint32_t> h(n, SENT);
  CK(cudaMemcpy(xch, h.data(), n * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)NC * WORDS * 4;
  CK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  void* args[] = {&xch, &NC, (void*)&WORDS, &steps, &dout, &dchk};
  CK(cudaLaunchCooperativeKernel((void*)gather_kernel<MODE>, dim3(grid), dim3(256), args, smem, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("P2 mode %d NC %d: %s\n", MODE, NC, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> ho(grid);
  CK(cudaMemcpy(ho.data(), dout, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : ho) mx = v > mx ? v : mx;
  printf("P2 all-gather through L2: %s, %2d CTAs/chain x %d chains, %4d B per CTA per step (%5.1f KB gathered): %.0f cycles per step\n",
         MODE == 0 ? "ld.relaxed.gpu.v4 spin" : "cp.async.cg + inspect ", NC, chains, BYTES, NC * BYTES / 1024.0, (double)mx / steps);
  cudaFree(xch); cudaFree(dout); cudaFree(dchk);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sms %d\n", prop.name, prop.multiProcessorCount);
  uint32_t* flag; long long* dout;
  CK(cudaMalloc(&flag, 256)); CK(cudaMalloc(&dout, 64));
  for (int peer : {1, 2, 37, 74, 100, 147}) {
    CK(cudaMemset(flag, 0, 256));
    const int rounds = 2000;
    void* args[] = {&flag, (void*)&rounds, &peer, &dout};
    CK(cudaLaunchCooperativeKernel((void*)pingpong_kernel, dim3(148), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dout, 8, cudaMemcpyDeviceToHost));
    printf("P1 ping-pong CTA 0 <-> CTA %3d: %.0f cycles per round trip (%.0f per hop)\n", peer, (double)c / rounds, (double)c / rounds / 2);
  }
  for (int chains : {1, 4}) {
    for (int NC : {10, 20, 27}) {
      const int BYTES = 320 * 8 * 4 / NC / 16 * 16;     // the chain's 320 cells x 8 streams x 4 B shared among its CTAs
      run_gather<0>(NC, chains, BYTES, 2000);
      run_gather<1>(NC, chains, BYTES, 2000);
    }
  }
  run_gather<0>(20, 4, 2048, 2000);                      // backward: 20 partial blocks of 16 cells x 8 streams... per CTA 20 x 512 B in
  run_gather<1>(20, 4, 2048, 2000);
  return 0;
}
