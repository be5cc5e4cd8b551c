// Runtime plumbing of libaslp_b200.so: device selection, memory, streams, error text.
// Replaces CuDevice (src/aslp-cudamatrix/cu-device.h:40-170) for the hot path.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

unsigned long long g_aslp_launches = 0;
// bumped whenever this library or its caller (through aslp_free) releases device memory: a captured step graph holds raw
// pointers, so it is only replayed while no allocation it may have seen has gone away (aslp_alloc_epoch)
static unsigned long long g_alloc_epoch = 0;
static thread_local char g_err[512] = "";

void aslp_set_last_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
}
void aslp_set_last_error_msg(const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s at %s:%d", msg, file, line);
}

int aslp_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" {

const char* aslp_last_error(void) { return g_err; }
unsigned long long aslp_launch_count(void) { return g_aslp_launches; }

int aslp_device_count(int* n) { ASLP_CUDA(cudaGetDeviceCount(n)); return 0; }
int aslp_set_device(int dev) { ASLP_CUDA(cudaSetDevice(dev)); return 0; }
int aslp_get_device(int* dev) { ASLP_CUDA(cudaGetDevice(dev)); return 0; }
int aslp_malloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) { aslp_set_last_error(e, __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  return 0;
}
int aslp_free(void* p) { ++g_alloc_epoch; ASLP_CUDA(cudaFree(p)); return 0; }
unsigned long long aslp_alloc_epoch(void) { return g_alloc_epoch; }
int aslp_count_launches(unsigned long long n) { g_aslp_launches += n; return 0; }

// ---- static-shape step replay: stream capture -> executable graph (host/matrix.h CuStepGraph) ----
int aslp_graph_begin(aslp_stream_t s) {
  ASLP_CUDA(cudaStreamBeginCapture((cudaStream_t)s, cudaStreamCaptureModeRelaxed));
  return 0;
}
int aslp_graph_end(aslp_stream_t s, void** exec, int* kernel_nodes) {
  ASLP_REQUIRE(exec != nullptr);
  *exec = nullptr;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)s, &g);
  if (e != cudaSuccess || g == nullptr) {
    cudaGetLastError();                                   // a failed capture leaves a sticky-looking error behind: clear it
    aslp_set_last_error(e == cudaSuccess ? cudaErrorUnknown : e, __FILE__, __LINE__);
    if (g != nullptr) cudaGraphDestroy(g);
    return ASLP_STATUS_EXECUTION_FAILED;
  }
  if (kernel_nodes != nullptr) {
    size_t n = 0;
    *kernel_nodes = 0;
    if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n > 0) {
      cudaGraphNode_t* nodes = new cudaGraphNode_t[n];
      if (cudaGraphGetNodes(g, nodes, &n) == cudaSuccess)
        for (size_t i = 0; i < n; ++i) {
          cudaGraphNodeType t;
          if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++*kernel_nodes;
        }
      delete[] nodes;
    }
  }
  cudaGraphExec_t x = nullptr;
  e = cudaGraphInstantiate(&x, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { cudaGetLastError(); aslp_set_last_error(e, __FILE__, __LINE__); return ASLP_STATUS_EXECUTION_FAILED; }
  *exec = (void*)x;
  return 0;
}
int aslp_graph_launch(void* exec, aslp_stream_t s) {
  ASLP_REQUIRE(exec != nullptr);
  ASLP_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)s));
  return 0;
}
int aslp_graph_destroy(void* exec) {
  if (exec != nullptr) ASLP_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
  return 0;
}
int aslp_malloc_host(void** p, size_t bytes) {
  cudaError_t e = cudaMallocHost(p, bytes ? bytes : 16);
  if (e != cudaSuccess) { aslp_set_last_error(e, __FILE__, __LINE__); return ASLP_STATUS_MEMOPS_FAILED; }
  return 0;
}
int aslp_free_host(void* p) { ASLP_CUDA(cudaFreeHost(p)); return 0; }
int aslp_memset(aslp_stream_t s, void* d, int v, size_t n) { ASLP_CUDA(cudaMemsetAsync(d, v, n, (cudaStream_t)s)); return 0; }
int aslp_memset2d(aslp_stream_t s, void* d, size_t pitch, int v, size_t w, size_t h) {
  if (w == 0 || h == 0) return 0;
  ASLP_CUDA(cudaMemset2DAsync(d, pitch, v, w, h, (cudaStream_t)s)); return 0; }
int aslp_memcpy_h2d(aslp_stream_t s, void* d, const void* h, size_t n) { ASLP_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s)); return 0; }
int aslp_memcpy_d2h(aslp_stream_t s, void* h, const void* d, size_t n) { ASLP_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s)); return 0; }
int aslp_memcpy_d2d(aslp_stream_t s, void* d, const void* src, size_t n) { ASLP_CUDA(cudaMemcpyAsync(d, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)s)); return 0; }
int aslp_memcpy2d_h2d(aslp_stream_t s, void* d, size_t dp, const void* h, size_t sp, size_t w, size_t ht) {
  if (w == 0 || ht == 0) return 0;
  // dense rows on both sides: one linear copy (a pitched copy of thousands of short rows is several times slower)
  if (dp == w && sp == w) { ASLP_CUDA(cudaMemcpyAsync(d, h, w * ht, cudaMemcpyHostToDevice, (cudaStream_t)s)); return 0; }
  ASLP_CUDA(cudaMemcpy2DAsync(d, dp, h, sp, w, ht, cudaMemcpyHostToDevice, (cudaStream_t)s)); return 0; }
int aslp_memcpy2d_d2h(aslp_stream_t s, void* h, size_t dp, const void* d, size_t sp, size_t w, size_t ht) {
  if (w == 0 || ht == 0) return 0;
  // dense rows on both sides: one linear copy (a pitched copy of thousands of short rows is several times slower)
  if (dp == w && sp == w) { ASLP_CUDA(cudaMemcpyAsync(h, d, w * ht, cudaMemcpyDeviceToHost, (cudaStream_t)s)); return 0; }
  ASLP_CUDA(cudaMemcpy2DAsync(h, dp, d, sp, w, ht, cudaMemcpyDeviceToHost, (cudaStream_t)s)); return 0; }
int aslp_memcpy2d_d2d(aslp_stream_t s, void* d, size_t dp, const void* src, size_t sp, size_t w, size_t ht) {
  if (w == 0 || ht == 0) return 0;
  // dense rows on both sides: one linear copy (a pitched copy of thousands of short rows is several times slower)
  if (dp == w && sp == w) { ASLP_CUDA(cudaMemcpyAsync(d, src, w * ht, cudaMemcpyDeviceToDevice, (cudaStream_t)s)); return 0; }
  ASLP_CUDA(cudaMemcpy2DAsync(d, dp, src, sp, w, ht, cudaMemcpyDeviceToDevice, (cudaStream_t)s)); return 0; }
int aslp_stream_create(aslp_stream_t* s) { cudaStream_t st; ASLP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); *s = (aslp_stream_t)st; return 0; }
int aslp_stream_destroy(aslp_stream_t s) { ASLP_CUDA(cudaStreamDestroy((cudaStream_t)s)); return 0; }
int aslp_stream_sync(aslp_stream_t s) { ASLP_CUDA(cudaStreamSynchronize((cudaStream_t)s)); return 0; }
int aslp_device_sync(void) { ASLP_CUDA(cudaDeviceSynchronize()); return 0; }
int aslp_event_record(aslp_stream_t s, void** event) {
  if (*event == nullptr) { cudaEvent_t e; ASLP_CUDA(cudaEventCreate(&e)); *event = (void*)e; }
  ASLP_CUDA(cudaEventRecord((cudaEvent_t)*event, (cudaStream_t)s));
  return 0;
}
int aslp_stream_wait_event(aslp_stream_t s, void* event) {
  ASLP_REQUIRE(event != nullptr);
  ASLP_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)event, 0));
  return 0;
}
int aslp_event_sync(void* event) {
  ASLP_REQUIRE(event != nullptr);
  ASLP_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return 0;
}
int aslp_event_destroy(void* event) {
  if (event != nullptr) ASLP_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int aslp_event_elapsed_ms(void* a, void* b, float* ms) {
  ASLP_CUDA(cudaEventSynchronize((cudaEvent_t)b));
  ASLP_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}

}  // extern "C"

// ---- per-(device, stream) scratch arena (see scratch.cuh) ----
#include "scratch.cuh"
#include <map>
#include <mutex>
#include <utility>
namespace {
struct ScratchBuf { void* ptr; size_t bytes; };
std::map<std::pair<int, cudaStream_t>, ScratchBuf> g_scratch;
std::mutex g_scratch_mu;
}
void* aslp_scratch(cudaStream_t stream, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  ScratchBuf& b = g_scratch[std::make_pair(dev, stream)];
  const size_t want = bytes + ASLP_SCRATCH_RESERVED;
  if (b.ptr == nullptr || b.bytes < want) {
    if (b.ptr != nullptr) { cudaStreamSynchronize(stream); cudaFree(b.ptr); b.ptr = nullptr; ++g_alloc_epoch; }
    size_t cap = want < (8u << 20) ? (8u << 20) : want * 2;
    if (cudaMalloc(&b.ptr, cap) != cudaSuccess) { b.ptr = nullptr; b.bytes = 0; return nullptr; }
    cudaMemsetAsync(b.ptr, 0, ASLP_SCRATCH_RESERVED, stream);
    b.bytes = cap;
  }
  return static_cast<char*>(b.ptr) + ASLP_SCRATCH_RESERVED;
}

// a helper thread's stream is going away (CuThreadDetach): drop its scratch arena
extern "C" int aslp_scratch_release(aslp_stream_t s) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = g_scratch.find(std::make_pair(dev, (cudaStream_t)s));
  if (it != g_scratch.end()) { if (it->second.ptr != nullptr) { cudaFree(it->second.ptr); ++g_alloc_epoch; } g_scratch.erase(it); }
  return 0;
}
