// nnet-train-step.h -- one cross-entropy training step, Propagate + Xent::Eval + Backpropagate (with the updates inside,
// src/aslp-nnet/nnet-nnet.cc:126-129), as the trainers run it minibatch after minibatch
// (src/aslp-nnetbin/aslp-nnet-train-frame.cc:110-124, aslp-nnet-train-perutt.cc:170-205).
//
// At the BASELINE cfg1 minibatch (256 frames through a 5-layer DNN) the device work of a step is ~60 kernels of 2-10 us:
// the step time is what the HOST needs to enqueue them (measured: 0.64 of 0.66 ms per step).  When the net consists of
// components without per-step host state and the minibatch shape repeats, the step is recorded once from the compute stream
// and replayed as one graph launch (CuStepGraph, host/matrix.h).  What changes from step to step reaches the recording
// through fixed device buffers filled OUTSIDE of it: the features are copied into `in_`, targets and frame weights into
// Xent's staging buffer (one asynchronous upload each); Xent's host bookkeeping runs every step.  The arithmetic is the
// same kernels in the same order, so a replayed step is bit-identical to the enqueued one (tests/test_gpu_step_graph.py).
#ifndef ASLP_HOST_NNET_TRAIN_STEP_H_
#define ASLP_HOST_NNET_TRAIN_STEP_H_
#include <sstream>
#include "cu-workspace.h"
#include "nnet-loss.h"
#include "nnet-nnet.h"

namespace kaldi {
namespace aslp_nnet {

class XentTrainStep {
 public:
  // Same effect as  nnet->Propagate(feats, &out); xent->Eval(frame_weights, out, post, &diff); nnet->Backpropagate(diff, NULL);
  // The net's output of the step stays available through Output().
  void Run(Nnet* nnet, Xent* xent, const CuMatrixBase<BaseFloat>& feats, const VectorBase<BaseFloat>& frame_weights, const Posterior& post) {
    const int32 rows = feats.NumRows(), cols = feats.NumCols();
    double nf = -1.0;
    const bool replayable = CuStepGraph::Enabled() && rows > 0 && nnet->StepReplayable();
    if (replayable) nf = xent->StageSparse(frame_weights, post, nnet->OutputDim());
    if (nf < 0.0) {                                      // dense targets, or a net with per-step host state: the plain sequence
      nnet->Propagate(feats, &out_);
      xent->Eval(frame_weights, out_, post, &diff_);
      nnet->Backpropagate(diff_, NULL);
      return;
    }
    if (in_.NumRows() != rows || in_.NumCols() != cols) in_.Resize(rows, cols, kUndefined);
    in_.CopyFromMat(feats);
    const NnetTrainOptions& o = nnet->GetTrainOptions();
    std::ostringstream key;
    key << nnet << ' ' << xent << ' ' << rows << ' ' << cols << ' ' << o.learn_rate << ' ' << o.momentum << ' ' << o.l2_penalty << ' ' << o.l1_penalty
        << ' ' << in_.Data() << ' ' << nnet->NumComponents() << ' ' << GemmPrecision();
    if (graph_.Begin(key.str())) {
      Enqueue(nnet, xent);
      if (!graph_.End()) Enqueue(nnet, xent);
    }
    xent->Progress(nf);
  }
  const CuMatrix<BaseFloat>& Output() const { return out_; }
  long long Replays() const { return graph_.Replays(); }
 private:
  void Enqueue(Nnet* nnet, Xent* xent) {
    nnet->Propagate(in_, &out_);
    xent->LaunchSparse(out_, &diff_);
    nnet->Backpropagate(diff_, NULL);
  }
  CuStepGraph graph_;
  CuMatrix<BaseFloat> in_, out_, diff_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
