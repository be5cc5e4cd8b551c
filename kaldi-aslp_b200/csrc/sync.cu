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

}  // extern "C"
