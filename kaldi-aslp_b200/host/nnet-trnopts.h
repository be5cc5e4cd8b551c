// nnet-trnopts.h -- NnetTrainOptions with the reference's option names and defaults
// (src/aslp-nnet/nnet-trnopts.h:29-47: --learn-rate 0.008 --momentum 0 --l2-penalty 0 --l1-penalty 0).
#ifndef ASLP_HOST_NNET_TRNOPTS_H_
#define ASLP_HOST_NNET_TRNOPTS_H_
#include "base.h"
#include "parse-options.h"

namespace kaldi {
namespace aslp_nnet {

struct NnetTrainOptions {
  BaseFloat learn_rate, momentum, l2_penalty, l1_penalty;
  NnetTrainOptions() : learn_rate(0.008f), momentum(0.0f), l2_penalty(0.0f), l1_penalty(0.0f) {}
  void Register(OptionsItf* opts) {
    opts->Register("learn-rate", &learn_rate, "Learning rate");
    opts->Register("momentum", &momentum, "Momentum");
    opts->Register("l2-penalty", &l2_penalty, "L2 penalty (weight decay)");
    opts->Register("l1-penalty", &l1_penalty, "L1 penalty (promote sparsity)");
  }
  friend std::ostream& operator<<(std::ostream& os, const NnetTrainOptions& o) {
    return os << "NnetTrainOptions : learn_rate" << o.learn_rate << ", momentum" << o.momentum
              << ", l2_penalty" << o.l2_penalty << ", l1_penalty" << o.l1_penalty;
  }
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
