// cu-workspace.h -- one growable device workspace shared by the sequential ops of the process-wide stream
// (split-K partials, LSTM/GRU exchange buffers, CTC alphas), so the hot path never allocates per minibatch
// (the reference's WarpCtc::EvalGpu cudaMalloc/cudaFrees three buffers per minibatch, warp-ctc.cc:74-95,173-182).
#ifndef ASLP_HOST_CU_WORKSPACE_H_
#define ASLP_HOST_CU_WORKSPACE_H_
#include "matrix.h"

namespace kaldi {
// device pointer to at least `bytes`; contents are scratch: valid until the next CuWorkspace() call
void* CuWorkspace(size_t bytes);
// GEMM precision used by the components: ASLP_GEMM_3XTF32 (default, fp32-grade) unless the environment variable
// ASLP_GEMM_PRECISION is "tf32" (single-pass TF32, looser bound) or "fp32" (CUDA cores)
int GemmPrecision();
void SetGemmPrecision(int p);
}  // namespace kaldi
#endif
