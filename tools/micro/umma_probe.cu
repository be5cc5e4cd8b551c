// Design probe for the tcgen05 / cluster form of the LSTM recurrence (development aid, not product code).
//   T1  numerics of tcgen05.mma kind::f16 with the A operand resident in TMEM (written with tcgen05.st 32x32b) and B in
//       shared memory in the un-swizzled K-major canonical layout; both LBO/SBO assignments are tried.
//   T2  time of one recurrence step's MMA chain (NMMA x [128 x 16 x 16], A in TMEM) + commit + mbarrier wake-up + tcgen05.ld.
//   T3  cluster all-gather through distributed shared memory: every CTA of a cluster of CS pushes BYTES to each peer with
//       st.async (16 B, mbarrier complete_tx) per round, rounds are data-dependent (a CTA sends round r+1 only after all of
//       round r has landed) -- cycles per round = the hand-off cost of a recurrence step.  Also the bulk-copy variant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) { if (++spins > (1u << 22)) __trap(); }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// ---------------------------------------------------------------- T1 + T2
// A [128][K] fp16 row-major (global), B [16][K] fp16 row-major (global); D [128][16] fp32.  K multiple of 16, <= 320.
// variant 0: cores contiguous along K: LBO = 128, SBO = (K/8)*128 ; variant 1: the two fields swapped
__global__ void __launch_bounds__(160, 1) ts_mma_kernel(const __half* A, const __half* B, float* D, int K, int variant, int reps, int nmma,
                                                       long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B -> smem canonical K-major no-swizzle: core (ng, kc) = 8 rows x 16 B at ng*(K/8)*128 + kc*128, row r at +16 r
  const int kcores = K / 8;
  for (int i = threadIdx.x; i < 16 * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const uint32_t off = (uint32_t)(n / 8) * kcores * 128 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__half*>(smem + off) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tm_a = tmem + 32, tm_d = tmem;        // D: columns 0..15, A: columns 32 .. 32 + K/2
  if (warp < 4) {
    // row = 32 warp + lane; K/2 packed columns, 8 at a time
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const __half2 h = __halves2half2(A[(size_t)row * K + 2 * (c0 + j)], A[(size_t)row * K + 2 * (c0 + j) + 1]);
        v[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st8(tm_a + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 acc, F16 x F16, K-major both, N=16, M=128
  const uint32_t sb = smem_u32(smem);
  const uint32_t lbo = variant == 0 ? 128u : (uint32_t)kcores * 128u;
  const uint32_t sbo = variant == 0 ? (uint32_t)kcores * 128u : 128u;
  uint32_t ph = 0;
  long long t_issue = 0, t_total = 0;
  for (int r = 0; r < reps; ++r) {
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 128) {
      t0 = clock64();
      // numerics pass (nmma == 0): the real K loop; timing pass: nmma MMAs cycling through the k-steps
      const int steps = nmma > 0 ? nmma : K / 16;
      for (int j = 0; j < steps; ++j) {
        const int kk = j % (K / 16);
        mma_f16_ts(tm_d, tm_a + kk * 8, make_desc(sb + kk * 256, lbo, sbo), idesc, j > 0 ? 1u : 0u);
      }
      tc_commit(smem_u32(&bars[0]));
      t1 = clock64();
    }
    if (warp < 4) {
      mbar_wait(smem_u32(&bars[0]), ph);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tm_d + ((uint32_t)(warp * 32) << 16), v);
      if (r == reps - 1)
        for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
      tc_fence_before();
    }
    ph ^= 1u;
    __syncthreads();
    if (threadIdx.x == 128) { const long long t2 = clock64(); t_issue += t1 - t0; t_total += t2 - t0; }
  }
  if (threadIdx.x == 128 && cycles != nullptr) { cycles[0] = t_issue; cycles[1] = t_total; }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory"); }
}


// ---------------------------------------------------------------- T4: issue rate of tcgen05.mma by shape / operand source / accumulator rotation
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// SS: A from shared memory; M 64|128; N multiple of 8 (M=64) / 16 (M=128); NACC independent accumulators used round-robin.
// Everything about an MMA is a compile-time constant so that the issue loop is MMA after MMA (8 per trip).
template <int M, int N, int SS, int NACC>
__global__ void __launch_bounds__(160, 1) rate_kernel(int nmma, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t sb = smem_u32(smem);             // B: 32 KB region, A (ss): next 64 KB
  const uint32_t sa = sb + 32 * 1024;
  constexpr int KS = 8;                           // k-steps cycled through (K = 128 image)
  const uint64_t bd0 = make_desc(sb, 128, KS * 2 * 128), ad0 = make_desc(sa, 128, KS * 2 * 128);
  uint32_t ph = 0;
  long long t_issue = 0, t_total = 0;
  for (int r = 0; r < reps; ++r) {
    long long t0 = 0, t1 = 0;
    if (warp == 4) {                               // warp-uniform: the whole warp walks the loop, one elected lane issues
      t0 = clock64();
      for (int j0 = 0; j0 < nmma; j0 += 8) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t d = tmem + (uint32_t)((kk % NACC) * N);
            const uint32_t acc = (j0 > 0 || kk >= NACC) ? 1u : 0u;
            if (SS) mma_f16_ss(d, ad0 + (uint64_t)(kk * 16), bd0 + (uint64_t)(kk * 16), idesc, acc);
            else    mma_f16_ts(d, tmem + 256 + kk * 8, bd0 + (uint64_t)(kk * 16), idesc, acc);
          }
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(smem_u32(&bars[0]));
      __syncwarp();
      t1 = clock64();
    }
    if (warp < 4) { mbar_wait(smem_u32(&bars[0]), ph); tc_fence_after(); tc_fence_before(); }
    ph ^= 1u;
    __syncthreads();
    if (threadIdx.x == 128) { const long long t2 = clock64(); t_issue += t1 - t0; t_total += t2 - t0; }
  }
  if (threadIdx.x == 128) { cycles[0] = t_issue; cycles[1] = t_total; }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory"); }
}
template <int M, int N, int SS, int NACC>
void run_rate(long long* dcyc) {
  CK(cudaFuncSetAttribute(rate_kernel<M, N, SS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int nmma : {8, 24, 64}) {
    const int reps = 100;
    rate_kernel<M, N, SS, NACC><<<1, 160, 97 * 1024>>>(nmma, reps, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T4 M=%d N=%d ss=%d nacc=%d: CUDA error %s\n", M, N, SS, NACC, cudaGetErrorString(e)); exit(1); }
    long long cy[2]; CK(cudaMemcpy(cy, dcyc, 16, cudaMemcpyDeviceToHost));
    printf("T4 M=%3d N=%3d A=%s accumulators=%d chain=%2d: issue %.0f cycles (%.1f per MMA), issue->all done %.0f cycles (%.1f per MMA)\n", M, N, SS ? "smem" : "tmem",
           NACC, nmma, (double)cy[0] / reps, (double)cy[0] / reps / nmma, (double)cy[1] / reps, (double)cy[1] / reps / nmma);
  }
}

// ---------------------------------------------------------------- T3: cluster all-gather
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async16(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               :: "r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode 0: st.async 16 B by 128 threads; mode 1: cp.async.bulk shared::cta -> shared::cluster, one thread per destination
// BYTES per destination per round (multiple of 16).  Buffers: recv[2][CS][BYTES], stage[BYTES].
__global__ void __launch_bounds__(128, 1) allgather_kernel(int CS, int BYTES, int rounds, int mode, long long* out, unsigned* check) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  const uint32_t rank = cluster_rank();
  uint8_t* recv = smem;                          // [2][CS][BYTES]
  uint8_t* stage = smem + (size_t)2 * CS * BYTES;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(smem_u32(&bars[0]), (uint32_t)CS * BYTES);
    mbar_expect_tx(smem_u32(&bars[1]), (uint32_t)CS * BYTES);
  }
  cluster_sync();
  const int vecs = BYTES / 16;                   // 16-byte stores per destination
  uint32_t ph[2] = {0, 0};
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const int b = r & 1;
    const uint32_t val = (uint32_t)r * 1000u + rank;
    if (mode == 0) {
      // thread -> (vector index, destination subset)
      for (int i = threadIdx.x; i < vecs * CS; i += 128) {
        const int d = i / vecs, v = i - d * vecs;
        const uint32_t la = smem_u32(recv + ((size_t)b * CS + rank) * BYTES + v * 16);
        st_async16(mapa(la, d), mapa(smem_u32(&bars[b]), d), val, val + 1, val + 2, (uint32_t)v);
      }
    } else {
      for (int i = threadIdx.x; i < BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(stage)[i] = val;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x < CS) {
        const int d = threadIdx.x;
        const uint32_t la = smem_u32(recv + ((size_t)b * CS + rank) * BYTES);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(mapa(la, d)), "r"(smem_u32(stage)), "r"((uint32_t)BYTES), "r"(mapa(smem_u32(&bars[b]), d)) : "memory");
      }
    }
    // wait for all CS contributions of this round
    mbar_wait(smem_u32(&bars[b]), ph[b]);
    ph[b] ^= 1u;
    // consume: every thread reads one word per source (keeps the data dependence honest)
    for (int s = 0; s < CS; ++s) acc += *reinterpret_cast<volatile uint32_t*>(recv + ((size_t)b * CS + s) * BYTES + (threadIdx.x * 16) % BYTES);
    __syncthreads();                             // everyone has read buffer b of round r
    if (threadIdx.x == 0) mbar_expect_tx(smem_u32(&bars[b]), (uint32_t)CS * BYTES);   // re-arm for round r+2 before we send r+1
    if (mode == 1) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
  }
  const long long t1 = clock64();
  // expected: sum over rounds, sources of (r*1000 + s) for threads whose offset hits word 0 of a vector
  if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; check[blockIdx.x] = acc; }
  cluster_sync();
}

static float h2f(__half h) { return __half2float(h); }

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sms %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  // ------------------------------------------------ T1
  for (int K : {64, 320}) {
    std::vector<__half> hA(128 * K), hB(16 * K);
    srand(1);
    for (auto& x : hA) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& x : hB) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    __half *dA, *dB; float* dD; long long* dcyc;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 16 * 4)); CK(cudaMalloc(&dcyc, 16));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    const int smem = 16 * K * 2 + 1024;
    CK(cudaFuncSetAttribute(ts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int variant = 0; variant < 2; ++variant) {
      CK(cudaMemset(dD, 0, 128 * 16 * 4));
      ts_mma_kernel<<<1, 160, smem>>>(dA, dB, dD, K, variant, 1, 0, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("T1 K=%d variant=%d: CUDA error %s\n", K, variant, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * 16);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)h2f(hA[m * K + k]) * h2f(hB[n * K + k]);
        maxerr = fmax(maxerr, fabs(ref - hD[m * 16 + n])); maxref = fmax(maxref, fabs(ref));
      }
      printf("T1 ts_mma K=%d variant=%d (%s): max_abs_err %.3e (max |ref| %.3f) %s\n", K, variant,
             variant == 0 ? "LBO=K-adjacent core, SBO=N-group" : "swapped", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
    }
    // ---------------------------------------------- T2 (use the K=320 buffers)
    if (K == 320) {
      for (int nmma : {1, 20, 60, 72, 120}) {
        const int reps = 200;
        ts_mma_kernel<<<1, 160, smem>>>(dA, dB, dD, K, 0, reps, nmma, dcyc);
        CK(cudaDeviceSynchronize());
        long long c[2]; CK(cudaMemcpy(c, dcyc, 16, cudaMemcpyDeviceToHost));
        printf("T2 chain of %3d MMAs [128x16x16, A in TMEM]: issue %.0f cycles, issue->commit->wake->tcgen05.ld->barrier %.0f cycles per round\n",
               nmma, (double)c[0] / reps, (double)c[1] / reps);
      }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dcyc);
  }

  // ------------------------------------------------ T4
  {
    long long* dcyc; CK(cudaMalloc(&dcyc, 16));
    run_rate<128, 16, 0, 1>(dcyc); run_rate<128, 16, 0, 4>(dcyc); run_rate<128, 16, 1, 1>(dcyc); run_rate<128, 16, 1, 4>(dcyc);
    run_rate<128, 32, 0, 1>(dcyc); run_rate<128, 64, 0, 1>(dcyc); run_rate<128, 128, 0, 1>(dcyc); run_rate<128, 256, 0, 1>(dcyc); run_rate<128, 256, 1, 1>(dcyc);
    run_rate<64, 8, 0, 1>(dcyc); run_rate<64, 8, 0, 4>(dcyc); run_rate<64, 16, 0, 1>(dcyc); run_rate<64, 8, 1, 1>(dcyc);
    cudaFree(dcyc);
  }
  if (argc > 1 && atoi(argv[1]) == 4) return 0;
  // ------------------------------------------------ T3
  CK(cudaFuncSetAttribute(allgather_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(allgather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int CS : {2, 4, 8, 10, 16}) {
    for (int nclusters : {1, 4}) {
      for (int mode = 0; mode < 2; ++mode) {
        for (int BYTES : {16, 512, 1024, 2048}) {
          if (mode == 1 && BYTES == 16) continue;
          const int rounds = 500;
          const int grid = CS * nclusters;
          long long* dout; unsigned* dchk;
          CK(cudaMalloc(&dout, grid * 8)); CK(cudaMalloc(&dchk, grid * 4));
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128);
          cfg.dynamicSmemBytes = (size_t)2 * CS * BYTES + BYTES + 1024;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          int maxc = -1;
          cudaError_t eo = cudaOccupancyMaxActiveClusters(&maxc, allgather_kernel, &cfg);
          if (eo != cudaSuccess) { printf("T3 CS=%d: occupancy query failed: %s\n", CS, cudaGetErrorString(eo)); cudaGetLastError(); cudaFree(dout); cudaFree(dchk); continue; }
          if (BYTES == 16 && mode == 0 && nclusters == 1) printf("T3 CS=%d: max active clusters %d\n", CS, maxc);
          if (maxc < nclusters) { cudaFree(dout); cudaFree(dchk); continue; }
          cudaError_t el = cudaLaunchKernelEx(&cfg, allgather_kernel, CS, BYTES, rounds, mode, dout, dchk);
          if (el != cudaSuccess) { printf("T3 CS=%d launch failed: %s\n", CS, cudaGetErrorString(el)); cudaGetLastError(); cudaFree(dout); cudaFree(dchk); continue; }
          cudaError_t es = cudaDeviceSynchronize();
          if (es != cudaSuccess) { printf("T3 CS=%d mode=%d BYTES=%d: run failed: %s\n", CS, mode, BYTES, cudaGetErrorString(es)); return 1; }
          std::vector<long long> ho(grid); std::vector<unsigned> hc(grid);
          CK(cudaMemcpy(ho.data(), dout, grid * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc.data(), dchk, grid * 4, cudaMemcpyDeviceToHost));
          long long mx = 0; for (auto v : ho) mx = v > mx ? v : mx;
          // thread 0 reads word 0 of each source's block: value r*1000 + s
          unsigned expect = 0; for (int r = 0; r < rounds; ++r) for (int s = 0; s < CS; ++s) expect += (unsigned)r * 1000u + s;
          printf("T3 allgather CS=%2d clusters=%d mode=%s bytes/dest=%4d: %.0f cycles/round (%.1f B/cycle in per CTA) data %s\n", CS, nclusters,
                 mode == 0 ? "st.async" : "bulk    ", BYTES, (double)mx / rounds, (double)(CS - 1) * BYTES / ((double)mx / rounds),
                 hc[0] == expect ? "ok" : "WRONG");
          cudaFree(dout); cudaFree(dchk);
        }
      }
    }
  }
  return 0;
}
