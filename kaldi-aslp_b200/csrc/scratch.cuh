// Library-internal scratch: one lazily allocated device buffer per (device, stream) for reduction
// partials and grid-barrier counters.  Allocated once at first use, never per call, so the hot
// path does not cudaMalloc (the reference's WarpCtc::EvalGpu mallocs per minibatch,
// src/aslp-nnet/warp-ctc.cc:74-95 -- that is what this avoids).
#pragma once
#include <stddef.h>
#include <cuda_runtime.h>

// returns a device pointer to >= bytes of scratch private to `stream`; NULL on failure.
// The first 4 KB of every scratch is reserved and kept ZERO between kernels (barrier counters).
void* aslp_scratch(cudaStream_t stream, size_t bytes);
constexpr size_t ASLP_SCRATCH_RESERVED = 4096;
