/* ctc.h -- the warp-ctc C API, kept symbol-for-symbol so that
 * src/aslp-nnet/warp-ctc.cc:119 (EvalGpu) links against libaslp_b200.so unchanged.
 * Interface replaced: src/warp-ctc/include/ctc.h:16-122 (status enum, ctcComputeInfo,
 * compute_ctc_loss, get_workspace_size).  Only the GPU location is served here:
 * info.loc == CTC_CPU returns CTC_STATUS_EXECUTION_FAILED (there is no CPU path).
 *
 * Semantics follow CpuCTC (src/warp-ctc/include/detail/cpu_ctc.h:158-428):
 *   - activations are UNNORMALISED, laid out (t, n, k) contiguous with row stride ==
 *     alphabet_size (stream-interleaved rows t*minibatch + n); softmax is applied inside;
 *   - gradients (same layout, must be zeroed by the caller as in the reference) receive
 *     softmax - posterior for t < input_lengths[n]; utterances with L + repeats > T get
 *     cost 0 and no gradient (cpu_ctc.h:193-195); blank label = 0;
 *   - activations / gradients / workspace are DEVICE pointers, labels / lengths / costs
 *     are HOST pointers (as for warp-ctc's CTC_GPU path);
 *   - costs[n] = -log p(labels_n | activations_n).
 */
#ifndef ASLP_B200_CTC_H_
#define ASLP_B200_CTC_H_
#ifdef __cplusplus
#include <cstddef>
extern "C" {
#else
#include <stddef.h>
#endif

typedef struct CUstream_st* CUstream;

typedef enum {
  CTC_STATUS_SUCCESS = 0,
  CTC_STATUS_MEMOPS_FAILED = 1,
  CTC_STATUS_INVALID_VALUE = 2,
  CTC_STATUS_EXECUTION_FAILED = 3,
  CTC_STATUS_UNKNOWN_ERROR = 4
} ctcStatus_t;

const char* ctcGetStatusString(ctcStatus_t status);

typedef enum { CTC_CPU = 0, CTC_GPU = 1 } ctcComputeLocation;

struct ctcComputeInfo {
  ctcComputeLocation loc;
  union {
    unsigned int num_threads;
    CUstream stream;
  };
};

ctcStatus_t compute_ctc_loss(const float* const activations, float* gradients,
                             const int* const flat_labels, const int* const label_lengths,
                             const int* const input_lengths, int alphabet_size, int minibatch,
                             float* costs, void* workspace, struct ctcComputeInfo info);

ctcStatus_t get_workspace_size(const int* const label_lengths, const int* const input_lengths,
                               int alphabet_size, int minibatch, struct ctcComputeInfo info,
                               size_t* size_bytes);

#ifdef __cplusplus
}
#endif
#endif
