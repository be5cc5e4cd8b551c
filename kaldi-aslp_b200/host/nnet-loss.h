// nnet-loss.h -- Xent (src/aslp-nnet/nnet-loss.{h,cc}:63-200) and WarpCtc (src/aslp-nnet/warp-ctc.{h,cc}) with the
// reference's method names, report line formats and host-side bookkeeping (bit-exact counters, the 6-sigma loss
// guard state machine, greedy token error rate), over the fused Xent kernel and the CTC kernels of libaslp_b200.
#ifndef ASLP_HOST_NNET_LOSS_H_
#define ASLP_HOST_NNET_LOSS_H_
#include "matrix.h"

namespace kaldi {

// hmm/posterior.h: per frame a list of (pdf-id, weight)
typedef std::vector<std::vector<std::pair<int32, BaseFloat>>> Posterior;

namespace aslp_nnet {

// common interface of the frame-level objectives (nnet-loss.h:33-71); the trainer mains hold a LossItf*
class LossItf {
 public:
  LossItf() {}
  virtual ~LossItf() {}
  virtual void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& target,
                    CuMatrix<BaseFloat>* diff) = 0;
  virtual void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& target,
                    CuMatrix<BaseFloat>* diff) = 0;
  // without frame weights (all ones)
  virtual void Eval(const CuMatrixBase<BaseFloat>& net_out, const Posterior& target, CuMatrix<BaseFloat>* diff) {
    tmp_frame_weights_.Resize(static_cast<int32>(target.size()), kUndefined);
    tmp_frame_weights_.Set(1.0f);
    Eval(tmp_frame_weights_, net_out, target, diff);
  }
  virtual std::string Report() = 0;
  virtual BaseFloat AvgLoss() = 0;
 protected:
  Vector<BaseFloat> tmp_frame_weights_;
};

void PosteriorToMatrix(const Posterior& post, int32 num_cols, CuMatrix<BaseFloat>* mat);      // nnet-utils.h:318-338

class Xent : public LossItf {
 public:
  Xent();
  ~Xent();
  using LossItf::Eval;
  // dense targets (soft labels)
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& targets, CuMatrix<BaseFloat>* diff);
  // posterior targets: one fused sparse pass when every frame has at most one pdf, dense otherwise
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& post, CuMatrix<BaseFloat>* diff);
  std::string Report();
  BaseFloat AvgLoss();
  double Frames() { Fetch(); return frames_; }
  double Correct() { Fetch(); return correct_; }
  // The sparse Eval in three parts, so that the device part can sit inside a recorded step (CuStepGraph) while the host
  // parts run every step: StageSparse packs target index, target weight and frame weight of every frame into one
  // page-locked slot and uploads them with ONE asynchronous copy into a fixed device buffer (returns the weighted frame
  // count, or -1 when some frame has more than one target -- the caller then takes the dense Eval); LaunchSparse is the
  // kernel alone; Progress is the host-side bookkeeping.
  double StageSparse(const VectorBase<BaseFloat>& frame_weights, const Posterior& post, int32 num_pdf);
  void LaunchSparse(const CuMatrixBase<BaseFloat>& net_out, CuMatrix<BaseFloat>* diff);
  void Progress(double num_frames);
 private:
  void Fetch();                 // device accumulators -> host (synchronises)
  double* stats_dev_;           // [ce, entropy, likelihood, correct, frames] accumulated on the device
  double frames_, correct_, loss_, entropy_, likelyhood_;
  double frames_progress_, base_[5];
  CuMatrix<BaseFloat> tgt_mat_;
  CuVector<BaseFloat> frame_w_dev_;
  static const int kStageSlots = 8;
  float* stage_host_[kStageSlots];          // page-locked [3][rows_pad]: index bits, target weight, frame weight
  void* stage_event_[kStageSlots];
  size_t stage_cap_[kStageSlots];
  CuVector<BaseFloat> stage_dev_;           // the same three rows on the device (fixed address while the minibatch size repeats)
  int32 stage_rows_, stage_pad_;
  unsigned stage_next_;
};

// mean square error (nnet-loss.h:133-171, nnet-loss.cc:205-290): diff = w (y - t); one fused pass, the loss stays on the device
// until a report asks for it
class Mse : public LossItf {
 public:
  Mse();
  ~Mse();
  using LossItf::Eval;
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& target, CuMatrix<BaseFloat>* diff);
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& target, CuMatrix<BaseFloat>* diff);
  std::string Report();
  BaseFloat AvgLoss();
 private:
  double Loss();                // device accumulator -> host (synchronises)
  double* loss_dev_;
  double frames_, frames_progress_, loss_base_;
  int32 num_tgt_;
  CuVector<BaseFloat> frame_weights_;
  CuMatrix<BaseFloat> tgt_mat_;
};

// 'multitask,<type1>,<dim1>,<weight1>,...,<typeN>,<dimN>,<weightN>' over column ranges of the output (nnet-loss.h:173-222)
class MultiTaskLoss : public LossItf {
 public:
  MultiTaskLoss() {}
  ~MultiTaskLoss();
  using LossItf::Eval;
  void InitFromString(const std::string& s);
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& target, CuMatrix<BaseFloat>* diff) {
    KALDI_ERR << "This is not supposed to be called!";
  }
  void Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& target, CuMatrix<BaseFloat>* diff);
  std::string Report();
  BaseFloat AvgLoss();
 private:
  std::vector<LossItf*> loss_vec_;
  std::vector<int32> loss_dim_, loss_dim_offset_;
  std::vector<BaseFloat> loss_weights_;
  CuMatrix<BaseFloat> tgt_mat_;
};

class WarpCtc {
 public:
  WarpCtc();
  // CTC training over multiple sequences; net_out rows are stream-interleaved (t * num_seq + s). diff receives
  // d(loss)/d(activation) clipped to [-1, 1] after the average-loss guard (warp-ctc.cc:33-286)
  void Eval(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out,
            const std::vector<std::vector<int32>>& labels, CuMatrix<BaseFloat>* diff);
  void ErrorRate(const std::vector<int>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out, std::vector<std::vector<int>>& label);
  void SetReportStep(int32 report_step) { report_step_ = report_step; }
  std::string Report();
  float NumErrorTokens() const { return error_num_; }
  int32 NumRefTokens() const { return ref_num_; }
  void SetUseGpu(bool use_gpu);       // only true is served: there is no CPU path
  const std::vector<float>& LastCosts() const { return costs_; }
  int32 NumRejected() const { return rejected_num_; }      // utterances whose diff the loss guard zeroed so far (warp-ctc.cc:318-330)
 private:
  void StatAndAverageLossCheck(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt,
                               const std::vector<float>& pzx_host, CuMatrix<BaseFloat>* diff);
  int32 frames_, sequences_num_, ref_num_;
  float error_num_;
  int32 frames_progress_, ref_num_progress_;
  float error_num_progress_;
  int32 sequences_progress_;
  double obj_progress_;
  int32 report_step_;
  double obj_;
  double loss_sum_, loss_square_sum_, loss_sum_bak_, loss_square_sum_bak_;
  int32 normal_num_, stat_period_;
  std::vector<float> costs_;
  int32 rejected_num_ = 0;
  CuArrayInt maxid_;
};

// Eesen-style CTC on the softmax OUTPUTS (probabilities), errors back-propagated through the softmax
// (src/aslp-nnet/ctc-loss.{h,cc}).  The reference computes it only on the GPU (2T+1 kernel launches per minibatch,
// cu-kernels.cu:3276-3534); here one launch of aslp_ctc_eesen.  Statistics, the 6-sigma loss guard
// (CTC_GRAD_CHECK == AVG_LOSS_CHECK, ctc-loss.h:36; window 100 utterances) and the report lines as in the reference.
class Ctc {
 public:
  Ctc();
  void EvalParallel(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out,
                    std::vector<std::vector<int32>>& label, CuMatrix<BaseFloat>* diff);
  // single sequence (ctc-loss.cc:33-113): the multi-sequence path with one stream
  void Eval(const CuMatrixBase<BaseFloat>& net_out, const std::vector<int32>& label, CuMatrix<BaseFloat>* diff);
  void ErrorRateMSeq(const std::vector<int>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out, std::vector<std::vector<int>>& label);
  void SetReportStep(int32 report_step) { report_step_ = report_step; }
  std::string Report();
  float NumErrorTokens() const { return error_num_; }
  int32 NumRefTokens() const { return ref_num_; }
  const std::vector<float>& LastObj() const { return pzx_; }      // -log p(z|x) per sequence of the last call
 private:
  void StatAndAverageLossCheck(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt,
                               const std::vector<float>& pzx_host, CuMatrix<BaseFloat>* diff);
  int32 frames_, sequences_num_, ref_num_;
  float error_num_;
  int32 frames_progress_, ref_num_progress_;
  float error_num_progress_;
  int32 sequences_progress_;
  double obj_progress_;
  int32 report_step_;
  double obj_;
  double loss_sum_, loss_square_sum_, loss_sum_bak_, loss_square_sum_bak_;
  int32 normal_num_, stat_period_;
  std::vector<float> pzx_;
  CuArrayInt labels_dev_, seq_len_dev_;
  CuVector<BaseFloat> pzx_dev_;
};

int32 LevenshteinEditDistance(const std::vector<int32>& ref, const std::vector<int32>& hyp, int32* ins, int32* del, int32* sub);

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
