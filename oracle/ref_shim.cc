// ref_shim.cc -- TEST INFRASTRUCTURE ONLY (part of the oracle build recipe, see oracle/Makefile).
// The reference defines SequenceDataReader::Done() with the `inline` keyword inside data-reader.cc
// (src/aslp-nnet/data-reader.cc:196-198) although its trainer mains call it from another translation unit.  Current
// g++ emits no out-of-line body for it, so aslp-nnet-train-lstm-streams would jump through an unresolved symbol.
// This file supplies that one out-of-line body (what older compilers emitted); the reference sources stay untouched.
#include "aslp-nnet/data-reader.h"

namespace kaldi {
namespace aslp_nnet {
bool SequenceDataReader::Done() { return (read_done_ && feature_reader_->Done()); }
}  // namespace aslp_nnet
}  // namespace kaldi
