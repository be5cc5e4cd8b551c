// Dense contractions for the aslp-nnet path: CuMatrixBase::AddMatMat
// (src/aslp-cudamatrix/cu-matrix.cc:1027-1062 -> cublasSgemm, cublas-wrappers.h:28-38) replaced by a
// hand-written sm_100a kernel: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ->
// tcgen05.mma kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue
// (alpha/beta/bias/clip fused).  fp32 operands stay fp32 in HBM; the tensor core reads them as
// TF32.  The default precision is a 3-pass split (a = hi + lo; hi*hi + lo*hi + hi*lo) that
// restores fp32-grade products so the reference's 1e-4 parity bound holds.
//
// Persistent: one CTA per SM walks the (split, tile) work items; TWO TMEM accumulators (2 x 128 columns) so that the
// epilogue of one tile (TMEM -> registers -> HBM) overlaps the main loop of the next.
//
//   warp 0      : TMA producer (one elected lane), runs ahead across tile boundaries
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM lane quarter = warp_id % 4)
//   warps 6..9  : hi/lo splitter (3xTF32 only)
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "scratch.cuh"

// persistent-grid cap of the calling thread (aslp_gemm_set_cta_limit): work issued on a side stream leaves SMs to a concurrently
// running persistent kernel instead of taking one CTA per SM
static thread_local int t_cta_limit = 0;

namespace {

constexpr int BM = 128, BN = 128, BK = 32;           // BK floats = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB per operand tile
constexpr uint32_t SPIN_LIMIT = 1u << 18;
constexpr int STG_PITCH = 36;                        // epilogue staging tile pitch (floats)
constexpr int BAR_REGION = 256;                      // mbarriers + the TMEM slot, after the stage buffers
constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;    // one 32 x 32 chunk per epilogue warp
#ifndef ASLP_SPLIT_WARPS
#define ASLP_SPLIT_WARPS 4                           // hi/lo splitter warps of the 3xTF32 kernel
#endif
#ifndef ASLP_SPLIT_UNROLL
#define ASLP_SPLIT_UNROLL 4                          // float4 loads a splitter thread keeps in flight
#endif
constexpr int SPLIT_THREADS = 32 * ASLP_SPLIT_WARPS;
constexpr int SPLIT_UNROLL = ASLP_SPLIT_UNROLL;
constexpr int NT3 = 192 + SPLIT_THREADS;             // threads of the 3-pass kernel

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();   // turn a protocol bug into an error, not a hang
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// one lane of a CONVERGED warp; unlike `if (lane == 0)` the compiler knows the branch is taken by a single thread and emits the
// uniform-datapath instructions (UTCHMMA, UTMALDG) back to back instead of wrapping each in an ELECT / BRA.U.ANY loop -- measured
// 19 vs 46 cycles per tcgen05.mma issued (profiles/r02_umma_dsmem_probe.txt)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B)
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B -- the ONLY legal shared-memory layout for
// MN-major 32-bit (tf32) operands: 32-byte chunks swizzled inside the 128 B row, atom = 4 k-rows x 128 B
// (cutlass/gemm/collective/builders/sm100_common.inl:92; TMA side: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// extended epilogue of aslp_gemm_ex (include/aslp_b200.h): activation of the result, multiplication by the derivative of
// an activation at its output y, and the SGD apply W -= lr * C -- each otherwise a launch of its own after the product
struct EpiExt { int act; const float* dy; int ldy; int dkind; float* w; int ldw; float lr; };
__device__ __forceinline__ float epi_act(int kind, float x) {        // kind - 1 = ASLP_ACT_*
  if (kind == 1 + ASLP_ACT_SIGMOID) return ref_sigmoid(x);
  if (kind == 1 + ASLP_ACT_TANH) return ref_tanh(x);
  return fmaxf(x, 0.f);
}
__device__ __forceinline__ float epi_dact(int kind, float y, float e) {   // as act_bwd_kernel (pointwise.cu)
  if (kind == ASLP_ACT_SIGMOID) return y * (1.0f - y) * e;
  if (kind == ASLP_ACT_TANH) return (1.0f - y * y) * e;
  return y > 0.f ? e : 0.f;
}

#ifdef ASLP_GEMM_DEBUG_TIMES
__device__ unsigned long long g_dbg_t[148 * 8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DBG_T(slot) do { if ((threadIdx.x & 31) == 0 && threadIdx.x == 64) g_dbg_t[blockIdx.x * 8 + (slot)] = gtime(); } while (0)
#else
#define DBG_T(slot) do {} while (0)
#endif
constexpr int TB_PITCH = 132;                             // floats per row of a CTA's raw tile in shared memory (cluster reduction)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
struct EpiParams {
  float* C; int ldc;
  int M, N, K;
  float alpha, beta;
  const float* bias;
  float clip;
  float* partial;      // split-K: [splits][M][ldp] raw accumulators, else NULL
  int ldp;
  int kb_per_split;
  int tiles_m, tiles_n, splits;   // persistent scheduling: work item = (split z, tile row, tile column), z slowest
  // split-K reduced INSIDE the launch: the `splits` CTAs of a tile form a thread-block cluster (blockIdx = tile * splits + z,
  // cluster rank = z); each leaves its raw accumulator tile in its OWN shared memory (the operand stages are free by then),
  // the cluster meets, and CTA z reduces rows [z, z+1) * 128 / splits of the tile over all peers through distributed shared
  // memory in ascending z (the order and rounding of splitk_reduce_kernel) with the full epilogue.  No partial tile goes to
  // global memory, no second launch.  (The first form of this -- partials in global memory, arrival counters, 128 threads
  // per SM pulling 73 KB out of L2 -- took 8-19 us for the reduction alone: profiles/r02_gemm_in_launch_reduce.txt.)
  int cluster;         // non-zero: launched with cluster dimension = splits (<= 8), one work item per CTA
  EpiExt ext;
};

// ------------------------------------------------------------------ the kernel
template <bool A_MN, bool B_MN, int PASSES, bool INL>     // INL: with the in-launch split-K reduction (its own instantiation:
__global__ void __launch_bounds__(PASSES == 1 ? 192 : NT3, 1)   // the default kernel stays at its register count and code size)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, EpiParams p) {
  constexpr int STAGES = (PASSES == 1) ? 6 : 3;
  constexpr int STAGE_BYTES = (PASSES == 1) ? 2 * TILE_BYTES : 4 * TILE_BYTES;
  DBG_T(5);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B wants 1024 B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // barriers live after the stage buffers
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar  = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  // two TMEM accumulators (2 x 128 columns): the epilogue of one tile overlaps the main loop of the next
  auto tmem_full_bar  = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));
  float* stage_base = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + BAR_REGION);   // [4 warps][32][STG_PITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int num_items = tiles_mn * p.splits;
  // item -> (z, m0, n0, k-block range); consecutive items walk the tile columns of one row, so CTAs running at the same
  // time share the A rows in L2
  auto item_coords = [&](int item, int& z, int& m0, int& n0, int& kb_begin, int& kb_end) {
    z = (INL && p.cluster) ? item % p.splits : item / tiles_mn;
    const int t = (INL && p.cluster) ? item / p.splits : item - z * tiles_mn;
    const int tm = t / p.tiles_n;
    m0 = tm * BM; n0 = (t - tm * p.tiles_n) * BN;
    kb_begin = z * p.kb_per_split;
    kb_end = min(num_kb, kb_begin + p.kb_per_split);
  };

  if (threadIdx.x == 64) {      // descriptor fetch under the barrier / TMEM set-up instead of in front of the first load
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), SPLIT_THREADS);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: 2 x 128 fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  DBG_T(0);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (the warp stays converged; one elected lane issues) =====================
    {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int z, m0, n0, kb_begin, kb_end;
        item_coords(item, z, m0, n0, kb_begin, kb_end);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint32_t sb = sa + TILE_BYTES;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), 2 * TILE_BYTES);
            if (!A_MN) {
              tma_load_2d(sa, &tmA, full_bar(s), kb * BK, m0);              // box {32 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * 4096, &tmA, full_bar(s), m0 + 32 * j, kb * BK);  // box {32 m, 32 k}
            }
            if (!B_MN) {
              tma_load_2d(sb, &tmB, full_bar(s), kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d(sb + j * 4096, &tmB, full_bar(s), n0 + 32 * j, kb * BK);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the warp stays converged; one elected lane issues) =====================
    {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c=F32, a=b=TF32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); one MMA (K=8) = 32 B along the row
      // MN-major: 4 boxes of [32 k][128 B]; LBO = box stride 4096 B; swizzle atom = 4 k-rows -> SBO = 512 B;
      //           one MMA consumes 8 k-rows = 1024 B
      const uint32_t a_lbo = A_MN ? 4096u : 16u, b_lbo = B_MN ? 4096u : 16u;
      const uint32_t a_sbo = A_MN ? 512u : 1024u, b_sbo = B_MN ? 512u : 1024u;
      const uint32_t a_lay = A_MN ? 1u : 2u, b_lay = B_MN ? 1u : 2u;
      const uint32_t a_step = A_MN ? 1024u : 32u, b_step = B_MN ? 1024u : 32u;
      int s = 0; uint32_t ph = 0;
      int acc_idx = 0; uint32_t acc_ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int z, m0, n0, kb_begin, kb_end;
        item_coords(item, z, m0, n0, kb_begin, kb_end);
        mbar_wait(tmem_empty_bar(acc_idx), acc_ph ^ 1u);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc_idx * BN);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(PASSES == 1 ? full_bar(s) : split_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint32_t sb = sa + TILE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t da = make_desc(sa + k * a_step, a_lbo, a_sbo, a_lay);
              const uint64_t db = make_desc(sb + k * b_step, b_lbo, b_sbo, b_lay);
              const uint32_t acc = (kb > kb_begin || k > 0) ? 1u : 0u;
              if (PASSES == 1) {
                tc_mma_tf32(tmem_d, da, db, idesc, acc);
              } else {
                const uint64_t da_lo = make_desc(sa + 2 * TILE_BYTES + k * a_step, a_lbo, a_sbo, a_lay);
                const uint64_t db_lo = make_desc(sb + 2 * TILE_BYTES + k * b_step, b_lbo, b_sbo, b_lay);
                tc_mma_tf32(tmem_d, da_lo, db, idesc, acc);     // small terms first
                tc_mma_tf32(tmem_d, da, db_lo, idesc, 1u);
                tc_mma_tf32(tmem_d, da, db, idesc, 1u);
              }
            }
            tc_commit(empty_bar(s));            // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc_commit(tmem_full_bar(acc_idx));    // accumulator complete (fires at once for an empty k range)
        __syncwarp();
        if (++acc_idx == 2) { acc_idx = 0; acc_ph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may read
    int acc_idx = 0; uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int z, m0, n0, kb_begin, kb_end;
      item_coords(item, z, m0, n0, kb_begin, kb_end);
      mbar_wait(tmem_full_bar(acc_idx), acc_ph);
      DBG_T(4);
      tc_fence_after();
      const bool has_work = kb_end > kb_begin;
      // Coalesced epilogue.  tcgen05.ld hands lane r the 32 columns of accumulator row r; storing that straight to HBM makes
      // every store instruction touch 32 different rows (32 half-filled sectors, 32 LSU passes) and was what bounded the
      // K <= 1280 shapes of this path.  The chunk goes through a padded per-warp staging tile instead (pitch 36 floats:
      // conflict-free for 128-bit writes by row and 128-bit reads by row quarter), and is written out with 8 lanes covering
      // 128 contiguous bytes of one row, 4 rows per instruction; beta reads of the old C are coalesced the same way.
      float* stg = stage_base + (size_t)q * 32 * STG_PITCH;
      const int sub_r = lane >> 3, sub_c = (lane & 7) << 2;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_idx * BN + c * 32), v);
        const int nb = n0 + c * 32;
        if (INL && p.cluster) {                 // raw accumulator chunk -> this CTA's tile buffer (row = TMEM lane, pitch 132: conflict-free)
          float* tb = reinterpret_cast<float*>(smem_gen) + (size_t)(q * 32 + lane) * TB_PITCH + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tb + j) = has_work ? make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (nb < p.N) {                  // warp-uniform
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          __syncwarp();
          const int col = nb + sub_c;
          const int nvalid = p.N - col;         // > 0: columns col .. col+3 that exist
          // beta: the old C values of the whole chunk are fetched first, all eight loads in flight together.  Loaded one by
          // one inside the store loop below they serialise behind the stores (the compiler cannot prove that `dst` of one row
          // is not the next row's source): 32 dependent HBM round trips per tile, which made the momentum weight-gradient
          // product of a 256-frame minibatch take 30 us instead of 14.
          float4 oldc[8];
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
              if (p.partial != nullptr) {
                // ldp is padded to a multiple of 4, so a float4 never crosses the row end
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
          __syncwarp();                         // the staging tile is rewritten by the next chunk
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld inside tc_ld32): hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(acc_idx));
      if (++acc_idx == 2) { acc_idx = 0; acc_ph ^= 1u; }
    }
  } else {
    // ===================== hi/lo splitter (3xTF32) =====================
    if (PASSES == 3) {
      const int t = threadIdx.x - 192;            // 0..127
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int z, m0, n0, kb_begin, kb_end;
        item_coords(item, z, m0, n0, kb_begin, kb_end);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(s), ph);
          float4* hi4 = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES);                   // A then B, 32 KB
          float4* lo4 = reinterpret_cast<float4*>(smem_gen + s * STAGE_BYTES + 2 * TILE_BYTES);  // A_lo then B_lo
#pragma unroll (SPLIT_UNROLL)
          for (int i = 0; i < (2 * TILE_BYTES / 16) / SPLIT_THREADS; ++i) {
            const int idx = t + i * SPLIT_THREADS;
            // the tensor core reads a TF32 operand as the top 19 bits of the fp32 word, so the raw tile already IS the
            // "hi" operand; only the residual has to be written (one third less shared-memory traffic per stage)
            const float4 x = hi4[idx];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            lo4[idx] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
          mbar_arrive(split_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  }

  if constexpr (INL) if (p.cluster) {
    // ---- in-launch reduction over the cluster, by ALL warps of the CTA.  (Left to the four epilogue warps -- one warp per
    // scheduler, ~2000 dependent instructions each -- this phase took 7-16 us whatever the partials were read from; the work
    // is instruction latency, not bytes: profiles/r02_gemm_in_launch_reduce.txt.)
    int z, m0, n0, kb_begin, kb_end;
    item_coords(blockIdx.x, z, m0, n0, kb_begin, kb_end);
    DBG_T(1);
    cluster_sync_all();                                   // every split's raw tile is in its CTA's shared memory
    DBG_T(2);
    const int rp = (BM + p.splits - 1) / p.splits;        // CTA z reduces rows [z * rp, (z + 1) * rp) of the tile
    for (int it = threadIdx.x; it < rp * 32; it += blockDim.x) {
      const int rr = it >> 5, c4 = (it & 31) << 2;        // a thread: 4 columns of one row, over all splits in ascending z
      const int rt = z * rp + rr, row = m0 + rt, col = n0 + c4;
      if (rt >= BM || row >= p.M || col >= p.N) continue;
      const uint32_t laddr = smem_base + (uint32_t)(rt * TB_PITCH + c4) * 4u;
      float4 v[8];
#pragma unroll
      for (int zz = 0; zz < 8; ++zz) v[zz] = zz < p.splits ? ld_dsmem_v4(mapa_u32(laddr, (uint32_t)zz)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float ad[4], ay[4], aw[4], bia[4];                  // the epilogue's operands, in flight with the partials
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool okj = col + j < p.N;
        ad[j] = (okj && p.beta != 0.f) ? p.C[(size_t)row * p.ldc + col + j] : 0.f;
        ay[j] = (okj && p.ext.dy != nullptr) ? p.ext.dy[(size_t)row * p.ext.ldy + col + j] : 0.f;
        aw[j] = (okj && p.ext.w != nullptr) ? p.ext.w[(size_t)row * p.ext.ldw + col + j] : 0.f;
        bia[j] = (okj && p.bias != nullptr) ? p.bias[col + j] : 0.f;
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int zz = 0; zz < 8; ++zz)
        if (zz < p.splits) { acc.x += v[zz].x; acc.y += v[zz].y; acc.z += v[zz].z; acc.w += v[zz].w; }
      const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (col + j < p.N) {
          float vv = p.alpha * o[j];
          if (p.beta != 0.f) vv += p.beta * ad[j];
          if (p.bias != nullptr) vv += bia[j];
          if (p.clip > 0.f) vv = fminf(fmaxf(vv, -p.clip), p.clip);
          if (p.ext.act != 0) vv = epi_act(p.ext.act, vv);
          if (p.ext.dy != nullptr) vv = epi_dact(p.ext.dkind, ay[j], vv);
          p.C[(size_t)row * p.ldc + col + j] = vv;
          if (p.ext.w != nullptr) p.ext.w[(size_t)row * p.ext.ldw + col + j] = aw[j] + (-p.ext.lr) * vv;
        }
      }
    }
    DBG_T(3);
    cluster_sync_all();                                   // nobody leaves (and frees its shared memory) while a peer still reads its tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem_base) : "memory");
  }
}

// split-K second phase: C = alpha * sum_z partial[z] + beta*C + bias, clip, then the extended epilogue
__global__ void splitk_reduce_kernel(float* C, int ldc, const float* partial, int ldp, int splits, int M, int N,
                                     float alpha, float beta, const float* bias, float clip, EpiExt x) {
  const int n4 = (N + 3) / 4;
  const long long total = (long long)M * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i % n4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)z * M + r) * ldp + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float o[4] = {acc.x, acc.y, acc.z, acc.w};
    float* dst = C + (size_t)r * ldc + c;
    for (int j = 0; j < 4 && c + j < N; ++j) {
      float v = alpha * o[j];
      if (beta != 0.f) v += beta * dst[j];
      if (bias != nullptr) v += bias[c + j];
      if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
      if (x.act != 0) v = epi_act(x.act, v);
      if (x.dy != nullptr) v = epi_dact(x.dkind, x.dy[(size_t)r * x.ldy + c + j], v);
      dst[j] = v;
      if (x.w != nullptr) { float* wp = x.w + (size_t)r * x.ldw + c + j; *wp = *wp + (-x.lr) * v; }
    }
  }
}

#include "gemm_f16x3.cuh"

// ------------------------------------------------------------------ CUDA-core fp32 GEMM (odd shapes / exact fp32)
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_fp32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, float beta, float* __restrict__ C,
                                                        int ldc, const float* __restrict__ bias, float clip) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int mm, kk;
      if (TA) { mm = i % TM; kk = i / TM; } else { kk = i % TK; mm = i / TK; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) v = TA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
      As[kk][mm] = v;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int nn, kk;
      if (TB) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < K) v = TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
      if (bias != nullptr) v += bias[n];
      if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// operand stored row-major [outer_extent rows][inner_extent cols], row stride ld floats
bool make_tmap(CUtensorMap* tm, const float* base, int inner_extent, int outer_extent, int ld, int box_inner, int box_outer, bool mn_major) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner_extent, (cuuint64_t)outer_extent};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <bool A_MN, bool B_MN, int PASSES, bool INL>
int launch_tc(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p_in, int splits) {
  constexpr int STAGES = (PASSES == 1) ? 6 : 3;
  constexpr int STAGE_BYTES = (PASSES == 1) ? 2 * TILE_BYTES : 4 * TILE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + BAR_REGION + STG_BYTES;
  static_assert(8 * (3 * STAGES + 4) + 16 <= BAR_REGION, "barrier region too small");
  static bool attr_set = false;
  if (!attr_set) {
    ASLP_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<A_MN, B_MN, PASSES, INL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  EpiParams p = p_in;
  p.tiles_m = aslp_div_up(p.M, BM); p.tiles_n = aslp_div_up(p.N, BN); p.splits = splits;
  // persistent: one CTA per SM walks the (split, tile) work items with a stride of the grid size
  const long long items = (long long)p.tiles_m * p.tiles_n * splits;
  const int sm_cap = (t_cta_limit > 0 && t_cta_limit < aslp_num_sms()) ? t_cta_limit : aslp_num_sms();
  const int grid = (int)(items < sm_cap ? items : sm_cap);
  if (INL && p.cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)items, 1, 1);            // one work item per CTA; blockIdx = tile * splits + z
    cfg.blockDim = dim3(PASSES == 1 ? 192 : NT3, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)splits; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ASLP_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<A_MN, B_MN, PASSES, INL>, ta, tb, p));
    ASLP_COUNT_LAUNCH();
    return 0;
  }
  gemm_tf32_kernel<A_MN, B_MN, PASSES, INL><<<grid, PASSES == 1 ? 192 : NT3, SMEM, st>>>(ta, tb, p);
  ASLP_CHECK_LAUNCH();
  return 0;
}

// fp16 plane stored [rows][ldp] with K (extent k) contiguous; box = {64 halves, 128 rows}, 128-byte swizzle
bool make_tmap_h(CUtensorMap* tm, const __half* base, int k_extent, int rows, int ldp) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldp * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)HK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// splits op(X) [R rows of the product, K] into K-major fp16 planes; `mn_major`: X is stored [K][R]
int presplit_operand(cudaStream_t st, const float* X, int ld, int R, int K, bool mn_major, __half* hi, __half* lo, int ldp, float* inv, unsigned* cmax) {
  if (!mn_major) {
    int blocks = aslp_div_up(R, 8);
    if (blocks > aslp_num_sms() * 16) blocks = aslp_num_sms() * 16;
    presplit_rows_kernel<<<blocks, 256, 0, st>>>(X, ld, R, K, hi, lo, ldp, inv);
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  ASLP_CUDA(cudaMemsetAsync(cmax, 0, sizeof(unsigned) * R, st));
  int gy = aslp_div_up(K, 8 * 16);
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  presplit_colmax_kernel<<<dim3(aslp_div_up(R, 32), gy), 256, 0, st>>>(X, ld, K, R, cmax);
  ASLP_CHECK_LAUNCH();
  presplit_transpose_kernel<<<dim3(aslp_div_up(R, 32), aslp_div_up(K, 64)), 256, 0, st>>>(X, ld, K, R, cmax, hi, lo, ldp, inv);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int pick_splits_h(int M, int N, int K);

// the whole fp16-split product: two operand splits + the three-pass kernel (+ split-K reduce)
int gemm_f16x3(cudaStream_t st, bool a_mn, bool b_mn, const EpiParams& p_in, const float* A, int lda, const float* B, int ldb, void* workspace,
               size_t workspace_bytes) {
  const int M = p_in.M, N = p_in.N, K = p_in.K;
  const int ldp = (K + 7) / 8 * 8;
  const size_t plane_a = (size_t)M * ldp, plane_b = (size_t)N * ldp;
  const size_t bytes = 2 * (plane_a + plane_b) * sizeof(__half) + ((size_t)M + N) * 2 * sizeof(float) + 64;
  char* scr = (char*)aslp_scratch(st, bytes);
  if (scr == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  __half* a_hi = (__half*)scr; __half* a_lo = a_hi + plane_a;
  __half* b_hi = a_lo + plane_a; __half* b_lo = b_hi + plane_b;
  float* inv_a = (float*)(((uintptr_t)(b_lo + plane_b) + 15) & ~(uintptr_t)15);
  float* inv_b = inv_a + M;
  unsigned* cmax = (unsigned*)(inv_b + N);                // max(M, N) words are enough; M + N are reserved
  int rc = presplit_operand(st, A, lda, M, K, a_mn, a_hi, a_lo, ldp, inv_a, cmax);
  if (rc != 0) return rc;
  rc = presplit_operand(st, B, ldb, N, K, b_mn, b_hi, b_lo, ldp, inv_b, cmax);
  if (rc != 0) return rc;
  CUtensorMap tah, tal, tbh, tbl;
  if (!(make_tmap_h(&tah, a_hi, K, M, ldp) && make_tmap_h(&tal, a_lo, K, M, ldp) && make_tmap_h(&tbh, b_hi, K, N, ldp) && make_tmap_h(&tbl, b_lo, K, N, ldp))) {
    aslp_set_last_error_msg("cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    return ASLP_STATUS_EXECUTION_FAILED;
  }
  constexpr int SMEM = HSTAGES * HSTAGE_BYTES + 1024 + BAR_REGION + STG_BYTES;
  static_assert(8 * (2 * HSTAGES + 4) + 16 <= BAR_REGION, "barrier region too small");
  static bool attr_set = false;
  if (!attr_set) {
    ASLP_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  EpiParams p = p_in;
  int splits = pick_splits_h(M, N, K);
  const size_t ldpart = ((size_t)N + 3) / 4 * 4;
  if (splits > 1 && (workspace == nullptr || workspace_bytes < (size_t)splits * M * ldpart * sizeof(float))) splits = 1;
  const int num_kb = aslp_div_up(K, HK);
  p.partial = splits > 1 ? (float*)workspace : nullptr;
  p.ldp = (int)ldpart;
  p.kb_per_split = aslp_div_up(num_kb, splits);
  p.tiles_m = aslp_div_up(M, BM); p.tiles_n = aslp_div_up(N, BN); p.splits = splits;
  const long long items = (long long)p.tiles_m * p.tiles_n * splits;
  const int sm_cap = (t_cta_limit > 0 && t_cta_limit < aslp_num_sms()) ? t_cta_limit : aslp_num_sms();
  const int grid = (int)(items < sm_cap ? items : sm_cap);
  gemm_f16x3_kernel<<<grid, 192, SMEM, st>>>(tah, tal, tbh, tbl, p, inv_a, inv_b);
  ASLP_CHECK_LAUNCH();
  if (splits > 1) {
    const long long total = (long long)M * ((N + 3) / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(p.C, p.ldc, (const float*)workspace, (int)ldpart, splits, M, N, p.alpha, p.beta, p.bias, p.clip, EpiExt{0, nullptr, 0, 0, nullptr, 0, 0.f});
    ASLP_CHECK_LAUNCH();
  }
  return 0;
}

// Split-K policy.  (a) fill the chip: one work item per SM when the output has few tiles (wgrad shapes, small minibatches);
// (b) bound the length of one TMEM accumulation chain to 32 k-blocks (K = 1024): the tensor core's fp32
// accumulate truncates, so its error grows with the chain length, while the split partials are summed with
// IEEE round-to-nearest adds in the reduce kernel.  Large square GEMMs (many tiles) are left unsplit.
int pick_splits(int M, int N, int K) {
  const int tiles = aslp_div_up(M, BM) * aslp_div_up(N, BN);
  const int num_kb = aslp_div_up(K, BK);
  const int sms = aslp_num_sms();
  // (c) short K with few tiles (the products of a 256-frame minibatch: 16 .. 64 tiles, 8 .. 14 k-blocks): still split, down to
  // two k-blocks per item -- unsplit they leave 84 .. 132 SMs idle and a tile's whole k chain on one SM
  if (tiles >= sms || num_kb < 4) return 1;
  // one wave of work items (floor(sms / tiles) splits), not two: the k-loops of these products are a handful of blocks either
  // way, while every extra split is another M x N partial for the reduce pass to read -- cfg1 0.304 -> 0.278 ms per minibatch,
  // cfg4 1.018 -> 0.976 ms (ASLP_GEMM_SPLIT_WAVES=2 restores the old count)
  static const int waves = getenv("ASLP_GEMM_SPLIT_WAVES") ? atoi(getenv("ASLP_GEMM_SPLIT_WAVES")) : 1;
  int splits = waves >= 2 ? aslp_div_up(waves * sms, tiles) : (sms / tiles > 0 ? sms / tiles : 1);
  const int by_chain = aslp_div_up(num_kb, 32);
  if (by_chain > splits) splits = by_chain;
  const int by_kb = num_kb >= 16 ? num_kb / 4 : num_kb / 2;
  if (splits > by_kb) splits = by_kb;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  const int per = aslp_div_up(num_kb, splits);     // make every split non-empty
  return aslp_div_up(num_kb, per);
}

// k-blocks are 64 wide here; the split count never exceeds pick_splits' (the workspace the caller sized with aslp_gemm_workspace_bytes)
int pick_splits_h(int M, int N, int K) {
  int splits = pick_splits(M, N, K);
  const int num_kb = aslp_div_up(K, HK);
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  const int per = aslp_div_up(num_kb, splits);
  return aslp_div_up(num_kb, per);
}

// How many clusters of `cs` CTAs of the 3-pass kernel the device runs at once (each CTA owns a whole SM's shared memory and a
// cluster must sit inside one GPC: 16 clusters of 8 do NOT fit a B200, and a second wave doubles the launch).  Queried once.
int cluster_fit(int cs) {
  static int fit[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
  if (cs < 2 || cs > 8) return 0;
  if (fit[cs] >= 0) return fit[cs];
  constexpr int SMEM3 = 3 * 4 * TILE_BYTES + 1024 + BAR_REGION + STG_BYTES;
  auto kern = gemm_tf32_kernel<false, false, 3, true>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3) != cudaSuccess) { cudaGetLastError(); return fit[cs] = 0; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cs * aslp_num_sms()), 1, 1);
  cfg.blockDim = dim3(NT3, 1, 1);
  cfg.dynamicSmemBytes = SMEM3;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return fit[cs] = n;
}

}  // namespace

extern "C" {
int aslp_gemm_cluster_fit(int cluster_size) { return cluster_fit(cluster_size); }
#ifdef ASLP_GEMM_DEBUG_TIMES
int aslp_gemm_debug_times(unsigned long long* host) { return (int)cudaMemcpyFromSymbol(host, g_dbg_t, sizeof(unsigned long long) * 148 * 8); }
#endif

int aslp_gemm_set_cta_limit(int max_ctas) { t_cta_limit = max_ctas; return 0; }

size_t aslp_gemm_workspace_bytes(int M, int N, int K) {
  const int splits = pick_splits(M, N, K);
  if (splits <= 1) return 0;
  const size_t ldp = ((size_t)N + 3) / 4 * 4;
  return (size_t)splits * M * ldp * sizeof(float);
}

// `ext` != NULL: fold the extended epilogue into the split-K reduce when the product takes that path (*ext_done = true)
static int gemm_impl(aslp_stream_t s, int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int lda,
                     const float* B, int ldb, float beta, float* C, int ldc, const float* bias, float clip, int precision,
                     void* workspace, size_t workspace_bytes, const EpiExt* ext, bool* ext_done, bool reduce_in_launch, int max_splits) {
  cudaStream_t st = (cudaStream_t)s;
  if (ext_done != nullptr) *ext_done = false;
  ASLP_REQUIRE(M >= 0 && N >= 0 && K >= 0);
  if (M == 0 || N == 0) return 0;
  ASLP_REQUIRE(A != nullptr && B != nullptr && C != nullptr);
  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                       (lda % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && (bias == nullptr || (uintptr_t)bias % 16 == 0);
  const bool tiny = (long long)M * N * K < (1ll << 18) || K == 0;
  if (precision == ASLP_GEMM_FP32 || !aligned || tiny) {
    dim3 grid(aslp_div_up(N, 64), aslp_div_up(M, 64));
    if (!trans_a && !trans_b) gemm_fp32_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip);
    else if (!trans_a && trans_b) gemm_fp32_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip);
    else if (trans_a && !trans_b) gemm_fp32_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip);
    else gemm_fp32_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip);
    ASLP_CHECK_LAUNCH();
    return 0;
  }
  // op(A)[M,K]: !trans_a -> stored [M][K] (K-major) ; trans_a -> stored [K][M] (M-major)
  // op(B)[K,N]:  trans_b -> stored [N][K] (K-major) ; !trans_b -> stored [K][N] (N-major)
  bool a_mn = trans_a != 0, b_mn = trans_b == 0;
  static const bool via_transpose = getenv("ASLP_GEMM_MN_TRANSPOSE") != nullptr;   // bring-up switch: K-major only
  if (via_transpose && (a_mn || b_mn)) {
    const size_t lda_t = ((size_t)K + 3) / 4 * 4, ldb_t = lda_t;
    float* scr = (float*)aslp_scratch(st, ((a_mn ? (size_t)M * lda_t : 0) + (b_mn ? (size_t)N * ldb_t : 0)) * sizeof(float));
    if (scr == nullptr) { aslp_set_last_error_msg("scratch allocation failed", __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
    if (a_mn) { int rc = aslp_transpose(s, scr, (int)lda_t, A, lda, K, M); if (rc) return rc; A = scr; lda = (int)lda_t; scr += (size_t)M * lda_t; a_mn = false; }
    if (b_mn) { int rc = aslp_transpose(s, scr, (int)ldb_t, B, ldb, K, N); if (rc) return rc; B = scr; ldb = (int)ldb_t; b_mn = false; }
  }
  // default fp32-grade mode: operands split once into fp16 hi / lo planes, three kind::f16 passes (gemm_f16x3.cuh).  The two split
  // launches only pay off on chunk-sized products; ASLP_GEMM_SPLIT=tf32 keeps the in-loop 3xTF32 split for every shape (A/B runs).
  static const char* split_env = getenv("ASLP_GEMM_SPLIT");
  static const bool force_tf32_split = split_env != nullptr && split_env[0] == 't';
  // Measured (profiles/r02_gemm_ab.jsonl, times include the split launches): 16000x1280x640 NT 0.148 -> 0.107 ms, 16000x640x1280 NN
  // 0.151 -> 0.123 ms, 4096^3 0.644 -> 0.389 ms; but 16000x320x320 0.033 -> 0.043 ms (the split of A costs more than it saves) and
  // the TN weight gradient 1280x640x16000 0.157 -> 0.221 ms (two large TRANSPOSING splits: a column-maximum pass plus a tile
  // transpose each).  So: chunk-sized products only, and only when every MN-major operand is small (a weight matrix).
  static const double f16_min_work = getenv("ASLP_GEMM_F16_MIN_WORK") ? atof(getenv("ASLP_GEMM_F16_MIN_WORK")) : 3.0e9;
  const bool cheap_split = (!a_mn || (long long)M * K <= (1ll << 22)) && (!b_mn || (long long)N * K <= (1ll << 22));
  if (precision == ASLP_GEMM_F16X3 ||
      (precision == ASLP_GEMM_3XTF32 && !force_tf32_split && cheap_split && (double)M * N * K >= f16_min_work && K >= 2 * HK)) {
    EpiParams pe;
    pe.C = C; pe.ldc = ldc; pe.M = M; pe.N = N; pe.K = K; pe.alpha = alpha; pe.beta = beta; pe.bias = bias; pe.clip = clip;
    pe.partial = nullptr; pe.ldp = 0; pe.kb_per_split = 0; pe.tiles_m = pe.tiles_n = pe.splits = 0;
    pe.cluster = 0; pe.ext = EpiExt{0, nullptr, 0, 0, nullptr, 0, 0.f};
    return gemm_f16x3(st, a_mn, b_mn, pe, A, lda, B, ldb, workspace, workspace_bytes);
  }
  CUtensorMap ta, tb;
  bool ok = a_mn ? make_tmap(&ta, A, M, K, lda, 32, BK, true) : make_tmap(&ta, A, K, M, lda, BK, BM, false);
  ok = ok && (b_mn ? make_tmap(&tb, B, N, K, ldb, 32, BK, true) : make_tmap(&tb, B, K, N, ldb, BK, BN, false));
  if (!ok) { aslp_set_last_error_msg("cuTensorMapEncodeTiled failed", __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; }

  int splits = pick_splits(M, N, K);
  // aslp_gemm_ex (max_splits > 0): the split count is one whose clusters all run at once -- tiles <= cluster_fit(splits) -- whether
  // or not this call reduces inside the launch, so that the two forms add the same partial sums in the same order
  bool cluster_ok = false;
  if (max_splits > 0 && splits > 1 && precision != ASLP_GEMM_TF32) {
    const int tiles = aslp_div_up(M, BM) * aslp_div_up(N, BN);
    int cs = splits < max_splits ? splits : max_splits;
    while (cs > 1 && tiles > cluster_fit(cs)) --cs;
    if (cs > 1) { splits = cs; cluster_ok = true; }
  }
  const size_t ldp = ((size_t)N + 3) / 4 * 4;
  if (splits > 1 && (workspace == nullptr || workspace_bytes < (size_t)splits * M * ldp * sizeof(float))) splits = 1;   // caller gave no room: single pass
  const int num_kb = aslp_div_up(K, BK);
  EpiParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.bias = bias; p.clip = clip;
  p.partial = splits > 1 ? (float*)workspace : nullptr;
  p.ldp = (int)ldp;
  p.kb_per_split = aslp_div_up(num_kb, splits);
  p.cluster = 0; p.ext = EpiExt{0, nullptr, 0, 0, nullptr, 0, 0.f};
  // In-launch reduction: when the caller asked for it (aslp_gemm_ex, reduce_in_launch) and the splits of a tile fit one
  // thread-block cluster (portable size: 8).  Clusters are co-scheduled by the hardware, so nothing else about the launch matters.
  const bool in_launch = cluster_ok && reduce_in_launch && (t_cta_limit <= 0 || t_cta_limit >= aslp_num_sms());
  if (in_launch) {
    p.cluster = 1;
    p.partial = nullptr;
    if (ext != nullptr) p.ext = *ext;
  }
  const bool one_pass = precision == ASLP_GEMM_TF32;      // ASLP_GEMM_3XTF32 below the fp16-split threshold: the in-loop split
  int rc;
#define ASLP_DISPATCH(AM, BMN)                                             \
  rc = one_pass ? launch_tc<AM, BMN, 1, false>(st, ta, tb, p, splits)       \
     : in_launch ? launch_tc<AM, BMN, 3, true>(st, ta, tb, p, splits)      \
                 : launch_tc<AM, BMN, 3, false>(st, ta, tb, p, splits)
  if (!a_mn && !b_mn) { ASLP_DISPATCH(false, false); }
  else if (!a_mn && b_mn) { ASLP_DISPATCH(false, true); }
  else if (a_mn && !b_mn) { ASLP_DISPATCH(true, false); }
  else { ASLP_DISPATCH(true, true); }
#undef ASLP_DISPATCH
  if (rc != 0) return rc;
  if (in_launch) { if (ext != nullptr && ext_done != nullptr) *ext_done = true; return 0; }
  if (splits > 1) {
    const long long total = (long long)M * ((N + 3) / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > aslp_num_sms() * 8) blocks = aslp_num_sms() * 8;
    const EpiExt none{0, nullptr, 0, 0, nullptr, 0, 0.f};
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(C, ldc, (const float*)workspace, (int)ldp, splits, M, N, alpha, beta, bias, clip, ext != nullptr ? *ext : none);
    ASLP_CHECK_LAUNCH();
    if (ext != nullptr && ext_done != nullptr) *ext_done = true;
  }
  return 0;
}

int aslp_gemm(aslp_stream_t s, int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias, float clip, int precision,
              void* workspace, size_t workspace_bytes) {
  return gemm_impl(s, trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip, precision, workspace, workspace_bytes, nullptr, nullptr, false, 0);
}

int aslp_gemm_ex(aslp_stream_t s, int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int lda,
                 const float* B, int ldb, float beta, float* C, int ldc, const float* bias, float clip, int precision,
                 void* workspace, size_t workspace_bytes, const aslp_gemm_epilogue_t* epi) {
  if (epi == nullptr) return aslp_gemm(s, trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip, precision, workspace, workspace_bytes);
  ASLP_REQUIRE(epi->act >= 0 && epi->act <= 3);
  ASLP_REQUIRE(epi->dact_y == nullptr || (epi->dact_kind >= 0 && epi->dact_kind <= 2));
  if (M == 0 || N == 0) return 0;
  EpiExt x{epi->act, epi->dact_y, epi->dact_ldy, epi->dact_kind, epi->update_w, epi->update_ldw, epi->update_lr};
  bool done = false;
  static const bool allow_in_launch = getenv("ASLP_GEMM_REDUCE_IN_LAUNCH") == nullptr || getenv("ASLP_GEMM_REDUCE_IN_LAUNCH")[0] != '0';
  int rc = gemm_impl(s, trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, clip, precision, workspace, workspace_bytes, &x, &done,
                     epi->reduce_in_launch != 0 && allow_in_launch, 8);
  if (rc != 0 || done) return rc;
  // the product did not go through the split-K reduce (large or odd shapes): the same steps as launches of their own
  if (x.act != 0) { rc = aslp_act_fwd(s, x.act - 1, C, ldc, C, ldc, M, N); if (rc != 0) return rc; }
  if (x.dy != nullptr) { rc = aslp_act_bwd(s, x.dkind, C, ldc, x.dy, x.ldy, C, ldc, M, N); if (rc != 0) return rc; }
  if (x.w != nullptr) { rc = aslp_axpby(s, x.w, x.ldw, C, ldc, M, N, -x.lr, 1.0f); if (rc != 0) return rc; }
  return 0;
}

}  // extern "C"
