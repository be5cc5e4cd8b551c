// aslp-parallel averaging on ONE packed fp32 arena + an NCCL communicator over NVLink/NVSwitch.
// Replaces MpiNode (src/aslp-parallel/mpi-node.h:18-101) and the per-tensor
// device -> host -> MPI_Allreduce -> host -> device staging of BspWorker / BmufWorker / SodWorker
// (bsp-worker.cc:33-58, bmuf-worker.cc:37-68, sod-worker.cc:37-61): one ncclAllReduce per sync on
// the whole arena, with the pre-scale / BMUF filter / SOD optimizer as single fused passes.
#include "common.cuh"
#include <nccl.h>
#include <string.h>

struct aslp_comm {
  ncclComm_t comm;
  int rank, nranks;
};

namespace {

inline int blocks_for(size_t n) {
  size_t b = (n / 4 + 255) / 256;
  const size_t cap = (size_t)aslp_num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

__global__ void scale_kernel(float* dst, const float* w, size_t n, float f) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = w[i] * f;
}
__global__ void diff_kernel(float* g, const float* a, const float* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) g[i] = a[i] - b[i];
}
// bmuf-worker.cc:55-66
__global__ void bmuf_kernel(float* w, float* w_prev, float* delta_prev, const float* g, size_t n, float mom, float lr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float delta = mom * delta_prev[i] + (1.0f - mom) * lr * g[i];
    const float nw = w_prev[i] + delta;
    w[i] = nw; w_prev[i] = nw; delta_prev[i] = delta;
  }
}
// optimizer.h:21-170 ; g is the summed (w_prev - w), the optimizers update w in place
__global__ void sod_kernel(int opt, float* w, const float* g, float* s1, float* s2, size_t n, float lr, float p1, float p2, float eps, int step) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float wi = w[i];
    switch (opt) {   // formulas of src/aslp-parallel/optimizer.h (floors at eps = 1e-8 as ApplyFloor does there)
      case ASLP_OPT_SGD: wi -= lr * gi; break;                                                      // :36-47
      case ASLP_OPT_MOMENTUM: { const float v = p1 * s1[i] + lr * gi; s1[i] = v; wi -= v; } break;   // :48-61
      case ASLP_OPT_ADAGRAD: { const float a = s1[i] + gi * gi; s1[i] = a; wi -= lr * gi * (1.0f / sqrtf(fmaxf(a, eps))); } break;   // :62-81
      case ASLP_OPT_RMSPROP: { const float a = 0.9f * s1[i] + 0.1f * gi * gi; s1[i] = a; wi -= lr * gi * (1.0f / sqrtf(fmaxf(a, eps))); } break;  // :82-101
      case ASLP_OPT_ADADELTA: {                                                                     // :102-129
        const float a = p1 * s1[i] + (1.0f - p1) * gi * gi;
        const float upd = (1.0f / sqrtf(fmaxf(a, eps))) * sqrtf(fmaxf(s2[i], eps)) * gi;
        s1[i] = a; wi -= upd; s2[i] = p1 * s2[i] + (1.0f - p1) * upd * upd;
      } break;
      case ASLP_OPT_ADAM: {                                                                         // :130-159
        const float m = p1 * s1[i] + (1.0f - p1) * gi;
        const float v = p2 * s2[i] + (1.0f - p2) * gi * gi;
        s1[i] = m; s2[i] = v;
        const float c1 = 1.0f / (1.0f - powf(p1, (float)step)), c2 = 1.0f / (1.0f - powf(p2, (float)step));
        wi -= lr * c1 * m * (1.0f / sqrtf(fmaxf(v * c2, eps)));
      } break;
    }
    w[i] = wi;
  }
}

// ---- multi-tensor forms: the model's parameter tensors stay where the components own them (GetGpuParams views);
// a device table {ptr, arena offset, n} lets ONE launch gather / scatter all of them against a packed arena.
// blockIdx.y = tensor, blockIdx.x strides over its elements.
enum { MT_PACK_SCALE = 0, MT_PACK_DIFF = 1, MT_UNPACK = 2, MT_BMUF = 3, MT_SOD = 4, MT_PACK_WEIGHTED = 5 };
struct MtArgs {
  const aslp_tensor_ref_t* table;
  float* arena;            // PACK_*: destination; UNPACK: source; BMUF/SOD: the all-reduced g
  float* w_prev;           // arena-shaped
  float* s1; float* s2;    // BMUF: s1 = delta_prev ; SOD: optimizer states
  float a, b, p1, p2, eps; // PACK_SCALE: a = factor; PACK_DIFF: a = sign; BMUF: a = momentum, b = learn rate; SOD: a = lr
  int opt, step;
  const int* count;        // PACK_WEIGHTED: the job's frame total, on the device (a = this rank's frames)
};
template <int MODE>
__global__ void multi_tensor_kernel(MtArgs A) {
  const aslp_tensor_ref_t t = A.table[blockIdx.y];
  // bsp-worker.cc:44: float factor = float(num_worker_samples) / num_all_samples -- the same fp32 division, taken on the device
  const float weight = MODE == MT_PACK_WEIGHTED ? A.a / (float)(*A.count) : 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < t.n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = t.offset + i;
    if (MODE == MT_PACK_SCALE) A.arena[j] = t.ptr[i] * A.a;
    else if (MODE == MT_PACK_WEIGHTED) A.arena[j] = t.ptr[i] * weight;
    else if (MODE == MT_PACK_DIFF) A.arena[j] = A.a * (t.ptr[i] - A.w_prev[j]);
    else if (MODE == MT_UNPACK) t.ptr[i] = A.arena[j];
    else if (MODE == MT_BMUF) {
      const float delta = A.a * A.s1[j] + (1.0f - A.a) * A.b * A.arena[j];
      const float nw = A.w_prev[j] + delta;
      t.ptr[i] = nw; A.w_prev[j] = nw; A.s1[j] = delta;
    }
  }
}
// SOD: apply the optimizer to the scattered weights with arena-shaped g / states, then w_prev = w (sod-worker.cc:55-59)
__global__ void multi_tensor_sod_kernel(MtArgs A) {
  const aslp_tensor_ref_t t = A.table[blockIdx.y];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < t.n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = t.offset + i;
    const float gi = A.arena[j], lr = A.a, p1 = A.p1, p2 = A.p2, eps = A.eps;
    float wi = t.ptr[i];
    switch (A.opt) {
      case ASLP_OPT_SGD: wi -= lr * gi; break;
      case ASLP_OPT_MOMENTUM: { const float v = p1 * A.s1[j] + lr * gi; A.s1[j] = v; wi -= v; } break;
      case ASLP_OPT_ADAGRAD: { const float a = A.s1[j] + gi * gi; A.s1[j] = a; wi -= lr * gi * (1.0f / sqrtf(fmaxf(a, eps))); } break;
      case ASLP_OPT_RMSPROP: { const float a = 0.9f * A.s1[j] + 0.1f * gi * gi; A.s1[j] = a; wi -= lr * gi * (1.0f / sqrtf(fmaxf(a, eps))); } break;
      case ASLP_OPT_ADADELTA: {
        const float a = p1 * A.s1[j] + (1.0f - p1) * gi * gi;
        const float upd = (1.0f / sqrtf(fmaxf(a, eps))) * sqrtf(fmaxf(A.s2[j], eps)) * gi;
        A.s1[j] = a; wi -= upd; A.s2[j] = p1 * A.s2[j] + (1.0f - p1) * upd * upd;
      } break;
      case ASLP_OPT_ADAM: {
        const float m = p1 * A.s1[j] + (1.0f - p1) * gi, v = p2 * A.s2[j] + (1.0f - p2) * gi * gi;
        A.s1[j] = m; A.s2[j] = v;
        const float c1 = 1.0f / (1.0f - powf(p1, (float)A.step)), c2 = 1.0f / (1.0f - powf(p2, (float)A.step));
        wi -= lr * c1 * m * (1.0f / sqrtf(fmaxf(v * c2, eps)));
      } break;
    }
    t.ptr[i] = wi;
    A.w_prev[j] = wi;
  }
}
template <int MODE>
int launch_mt(aslp_stream_t s, const MtArgs& A, int ntensors) {
  if (ntensors == 0) return 0;
  dim3 grid(aslp_num_sms() * 2 / (ntensors < 8 ? 1 : 4) + 1, ntensors);
  multi_tensor_kernel<MODE><<<grid, 256, 0, (cudaStream_t)s>>>(A);
  ASLP_CHECK_LAUNCH();
  return 0;
}

#define ASLP_NCCL(call)                                                                              \
  do {                                                                                               \
    ncclResult_t r__ = (call);                                                                       \
    if (r__ != ncclSuccess) { aslp_set_last_error_msg(ncclGetErrorString(r__), __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; } \
  } while (0)

}  // namespace

extern "C" {

int aslp_sync_scale(aslp_stream_t s, float* dst, const float* w, size_t n, float factor) {
  if (n == 0) return 0;
  scale_kernel<<<blocks_for(n * 4), 256, 0, (cudaStream_t)s>>>(dst, w, n, factor);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_sync_diff(aslp_stream_t s, float* g, const float* a, const float* b, size_t n) {
  if (n == 0) return 0;
  diff_kernel<<<blocks_for(n * 4), 256, 0, (cudaStream_t)s>>>(g, a, b, n);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_sync_bmuf_apply(aslp_stream_t s, float* w, float* w_prev, float* delta_prev, const float* g_sum, size_t n, float momentum, float learn_rate) {
  if (n == 0) return 0;
  bmuf_kernel<<<blocks_for(n * 4), 256, 0, (cudaStream_t)s>>>(w, w_prev, delta_prev, g_sum, n, momentum, learn_rate);
  ASLP_CHECK_LAUNCH();
  return 0;
}
int aslp_sync_sod_apply(aslp_stream_t s, int opt, float* w, const float* g, float* state1, float* state2, size_t n, float lr, float p1,
                        float p2, float eps, int step) {
  if (n == 0) return 0;
  ASLP_REQUIRE(opt >= ASLP_OPT_SGD && opt <= ASLP_OPT_ADAM);
  sod_kernel<<<blocks_for(n * 4), 256, 0, (cudaStream_t)s>>>(opt, w, g, state1, state2, n, lr, p1, p2, eps, step);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_sync_pack(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, float factor) {
  MtArgs A = {}; A.table = table_dev; A.arena = arena; A.a = factor;
  return launch_mt<MT_PACK_SCALE>(s, A, ntensors);
}
int aslp_sync_pack_weighted(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, int frames, const int* frames_all_dev) {
  ASLP_REQUIRE(frames_all_dev != nullptr && frames >= 0);
  MtArgs A = {}; A.table = table_dev; A.arena = arena; A.a = (float)frames; A.count = frames_all_dev;
  return launch_mt<MT_PACK_WEIGHTED>(s, A, ntensors);
}
int aslp_sync_pack_diff(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, const float* w_prev_arena, float sign) {
  MtArgs A = {}; A.table = table_dev; A.arena = arena; A.w_prev = const_cast<float*>(w_prev_arena); A.a = sign;
  return launch_mt<MT_PACK_DIFF>(s, A, ntensors);
}
int aslp_sync_unpack(aslp_stream_t s, const float* arena, const aslp_tensor_ref_t* table_dev, int ntensors) {
  MtArgs A = {}; A.table = table_dev; A.arena = const_cast<float*>(arena);
  return launch_mt<MT_UNPACK>(s, A, ntensors);
}
int aslp_sync_bmuf_apply_packed(aslp_stream_t s, const aslp_tensor_ref_t* table_dev, int ntensors, float* w_prev_arena, float* delta_prev_arena,
                                const float* g_sum_arena, float momentum, float learn_rate) {
  MtArgs A = {}; A.table = table_dev; A.arena = const_cast<float*>(g_sum_arena); A.w_prev = w_prev_arena; A.s1 = delta_prev_arena;
  A.a = momentum; A.b = learn_rate;
  return launch_mt<MT_BMUF>(s, A, ntensors);
}
int aslp_sync_sod_apply_packed(aslp_stream_t s, int opt, const aslp_tensor_ref_t* table_dev, int ntensors, const float* g_sum_arena, float* state1,
                               float* state2, float* w_prev_arena, float lr, float p1, float p2, float eps, int step) {
  ASLP_REQUIRE(opt >= ASLP_OPT_SGD && opt <= ASLP_OPT_ADAM);
  if (ntensors == 0) return 0;
  MtArgs A = {}; A.table = table_dev; A.arena = const_cast<float*>(g_sum_arena); A.w_prev = w_prev_arena; A.s1 = state1; A.s2 = state2;
  A.a = lr; A.p1 = p1; A.p2 = p2; A.eps = eps; A.opt = opt; A.step = step;
  dim3 grid(aslp_num_sms() / 2 + 1, ntensors);
  multi_tensor_sod_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(A);
  ASLP_CHECK_LAUNCH();
  return 0;
}

int aslp_comm_unique_id(char id_out[128]) {
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ASLP_NCCL(ncclGetUniqueId(&id));
  memcpy(id_out, &id, 128);
  return 0;
}
int aslp_comm_init(aslp_comm_t* c, const char id[128], int nranks, int rank) {
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  aslp_comm* cc = new aslp_comm();
  cc->rank = rank; cc->nranks = nranks;
  ncclResult_t r = ncclCommInitRank(&cc->comm, nranks, uid, rank);
  if (r != ncclSuccess) { delete cc; aslp_set_last_error_msg(ncclGetErrorString(r), __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; }
  *c = cc;
  return 0;
}
int aslp_comm_destroy(aslp_comm_t c) {
  if (c == nullptr) return 0;
  ncclCommDestroy(c->comm);
  delete c;
  return 0;
}
int aslp_comm_rank(aslp_comm_t c, int* rank, int* nranks) {
  ASLP_REQUIRE(c != nullptr);
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return 0;
}
int aslp_comm_allreduce_sum_f32(aslp_comm_t c, aslp_stream_t s, float* buf, size_t n) {
  ASLP_REQUIRE(c != nullptr);
  ASLP_NCCL(ncclAllReduce(buf, buf, n, ncclFloat32, ncclSum, c->comm, (cudaStream_t)s));
  ASLP_COUNT_LAUNCH();
  return 0;
}
int aslp_comm_allreduce_sum_f64(aslp_comm_t c, aslp_stream_t s, double* buf, size_t n) {
  ASLP_REQUIRE(c != nullptr);
  ASLP_NCCL(ncclAllReduce(buf, buf, n, ncclFloat64, ncclSum, c->comm, (cudaStream_t)s));
  ASLP_COUNT_LAUNCH();
  return 0;
}
int aslp_comm_allreduce_sum_i32(aslp_comm_t c, aslp_stream_t s, int* buf, size_t n) {
  ASLP_REQUIRE(c != nullptr);
  ASLP_NCCL(ncclAllReduce(buf, buf, n, ncclInt32, ncclSum, c->comm, (cudaStream_t)s));
  ASLP_COUNT_LAUNCH();
  return 0;
}
int aslp_comm_barrier(aslp_comm_t c, aslp_stream_t s) {
  ASLP_REQUIRE(c != nullptr);
  static thread_local int* dummy = nullptr;
  if (dummy == nullptr) { ASLP_CUDA(cudaMalloc(&dummy, sizeof(int))); ASLP_CUDA(cudaMemset(dummy, 0, sizeof(int))); }
  ASLP_NCCL(ncclAllReduce(dummy, dummy, 1, ncclInt32, ncclSum, c->comm, (cudaStream_t)s));
  ASLP_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return 0;
}

// point-to-point exchange of a packed arena with one peer (the async server modes: MPI_Send / MPI_Recv / MPI_Sendrecv of
// easgd-worker.cc:49-56, asgd-worker.cc:47-58 -- one message for the whole model instead of one per tensor)
int aslp_comm_send_f32(aslp_comm_t c, aslp_stream_t s, const float* buf, size_t n, int peer) {
  ASLP_REQUIRE(c != nullptr && peer >= 0 && peer < c->nranks && peer != c->rank);
  ASLP_NCCL(ncclSend(buf, n, ncclFloat32, peer, c->comm, (cudaStream_t)s));
  ASLP_COUNT_LAUNCH();
  return 0;
}
int aslp_comm_recv_f32(aslp_comm_t c, aslp_stream_t s, float* buf, size_t n, int peer) {
  ASLP_REQUIRE(c != nullptr && peer >= 0 && peer < c->nranks && peer != c->rank);
  ASLP_NCCL(ncclRecv(buf, n, ncclFloat32, peer, c->comm, (cudaStream_t)s));
  ASLP_COUNT_LAUNCH();
  return 0;
}
int aslp_comm_sendrecv_f32(aslp_comm_t c, aslp_stream_t s, const float* sendbuf, float* recvbuf, size_t n, int peer) {
  ASLP_REQUIRE(c != nullptr && peer >= 0 && peer < c->nranks && peer != c->rank && sendbuf != recvbuf);
  ASLP_NCCL(ncclGroupStart());
  ncclResult_t r1 = ncclSend(sendbuf, n, ncclFloat32, peer, c->comm, (cudaStream_t)s);
  ncclResult_t r2 = ncclRecv(recvbuf, n, ncclFloat32, peer, c->comm, (cudaStream_t)s);
  ASLP_NCCL(ncclGroupEnd());
  ASLP_NCCL(r1); ASLP_NCCL(r2);
  ASLP_COUNT_LAUNCH();
  return 0;
}

}  // extern "C"
