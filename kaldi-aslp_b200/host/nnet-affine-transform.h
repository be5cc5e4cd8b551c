// nnet-affine-transform.h -- AffineTransform and LinearTransform over the tcgen05 GEMM.
// Reference: src/aslp-nnet/nnet-affine-transform.h:186-245, nnet-linear-transform.h:126-158.
//   fwd   : out = in W^T + bias            one GEMM, bias fused in the epilogue
//   bwd   : in_diff = out_diff W           one GEMM
//   update: W_corr = mmt W_corr + diff^T in (momentum as the GEMM's beta), bias_corr column sum,
//           optional L2 / L1 / max-norm, W -= lr W_corr
#ifndef ASLP_HOST_NNET_AFFINE_TRANSFORM_H_
#define ASLP_HOST_NNET_AFFINE_TRANSFORM_H_
#include <stdlib.h>
#include "cu-workspace.h"
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class AffineTransform : public UpdatableComponent {
 public:
  AffineTransform(int32 dim_in, int32 dim_out, bool has_bias = true)
      : UpdatableComponent(dim_in, dim_out), has_bias_(has_bias), linearity_(dim_out, dim_in), bias_(has_bias ? dim_out : 0),
        linearity_corr_(dim_out, dim_in), bias_corr_(has_bias ? dim_out : 0), learn_rate_coef_(1.0f), bias_learn_rate_coef_(1.0f), max_norm_(0.0f) {}
  Component* Copy() const { return new AffineTransform(*this); }
  ComponentType GetType() const { return has_bias_ ? kAffineTransform : kLinearTransform; }

  void InitData(std::istream& is) {
    float bias_mean = -2.0f, bias_range = 2.0f, param_stddev = 0.1f, norm_init_scale = 1.0f;
    bool gauss_init = true;
    std::string token;
    while (!is.eof()) {      // same keys as nnet-affine-transform.h:70-84 (<NormInit> switches to Glorot-uniform)
      ReadToken(is, false, &token);
      if (token == "<NormInit>") { ReadBasicType(is, false, &norm_init_scale); gauss_init = false; }
      else if (token == "<ParamStddev>") ReadBasicType(is, false, &param_stddev);
      else if (token == "<BiasMean>") ReadBasicType(is, false, &bias_mean);
      else if (token == "<BiasRange>") ReadBasicType(is, false, &bias_range);
      else if (token == "<LearnRateCoef>") ReadBasicType(is, false, &learn_rate_coef_);
      else if (token == "<BiasLearnRateCoef>") ReadBasicType(is, false, &bias_learn_rate_coef_);
      else if (token == "<MaxNorm>") ReadBasicType(is, false, &max_norm_);
      else KALDI_ERR << "Unknown token " << token << ", a typo in config? (ParamStddev|BiasMean|BiasRange|LearnRateCoef|BiasLearnRateCoef)";
      is >> std::ws;
    }
    if (!gauss_init) {
      const float scale = norm_init_scale * sqrt(6.0 / (output_dim_ + input_dim_));
      InitMatParam(&linearity_, scale);
      if (has_bias_) InitVecParam(&bias_, scale);
    } else {
      Matrix<BaseFloat> mat(output_dim_, input_dim_);
      for (int32 r = 0; r < output_dim_; r++)
        for (int32 c = 0; c < input_dim_; c++) mat(r, c) = param_stddev * RandGauss();
      linearity_ = mat;
      if (has_bias_) {
        Vector<BaseFloat> vec(output_dim_);
        for (int32 i = 0; i < output_dim_; i++) vec(i) = bias_mean + (RandUniform() - 0.5) * bias_range;
        bias_ = vec;
      }
    }
  }

  void ReadData(std::istream& is, bool binary) {
    if ('<' == Peek(is, binary)) {
      ExpectToken(is, binary, "<LearnRateCoef>"); ReadBasicType(is, binary, &learn_rate_coef_);
      if (has_bias_) { ExpectToken(is, binary, "<BiasLearnRateCoef>"); ReadBasicType(is, binary, &bias_learn_rate_coef_); }
    }
    if (has_bias_ && '<' == Peek(is, binary)) { ExpectToken(is, binary, "<MaxNorm>"); ReadBasicType(is, binary, &max_norm_); }
    if (has_bias_ && '<' == Peek(is, binary)) { float tmp; ExpectToken(is, binary, "<ClipGradient>"); ReadBasicType(is, binary, &tmp); }
    linearity_.Read(is, binary);
    if (has_bias_) bias_.Read(is, binary);
    KALDI_ASSERT(linearity_.NumRows() == output_dim_ && linearity_.NumCols() == input_dim_);
    KALDI_ASSERT(!has_bias_ || bias_.Dim() == output_dim_);
    linearity_corr_.Resize(output_dim_, input_dim_, kSetZero);
    if (has_bias_) bias_corr_.Resize(output_dim_, kSetZero);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<LearnRateCoef>"); WriteBasicType(os, binary, learn_rate_coef_);
    if (has_bias_) {
      WriteToken(os, binary, "<BiasLearnRateCoef>"); WriteBasicType(os, binary, bias_learn_rate_coef_);
      WriteToken(os, binary, "<MaxNorm>"); WriteBasicType(os, binary, max_norm_);
    }
    linearity_.Write(os, binary);
    if (has_bias_) bias_.Write(os, binary);
  }

  int32 NumParams() const { return linearity_.NumRows() * linearity_.NumCols() + bias_.Dim(); }
  void GetParams(Vector<BaseFloat>* wei_copy) const {
    wei_copy->Resize(NumParams());
    CopyRowsToVec(linearity_, wei_copy->Data());
    if (has_bias_) { Vector<float> b; bias_.CopyToVec(&b); for (int32 i = 0; i < b.Dim(); ++i) (*wei_copy)(linearity_.NumRows() * linearity_.NumCols() + i) = b(i); }
  }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
    params->clear();
    params->push_back(std::make_pair(linearity_.Data(), linearity_.NumRows() * linearity_.Stride()));
    if (has_bias_) params->push_back(std::make_pair(bias_.Data(), bias_.Dim()));
  }
  std::string Info() const { return std::string("\n  linearity") + MomentStatistics(linearity_) + (has_bias_ ? "\n  bias" + MomentStatistics(bias_) : ""); }
  std::string InfoGradient() const {
    return std::string("\n  linearity_grad") + MomentStatistics(linearity_corr_) + ", lr-coef " + ToString(learn_rate_coef_) + ", max-norm " + ToString(max_norm_) +
           (has_bias_ ? "\n  bias_grad" + MomentStatistics(bias_corr_) + ", lr-coef " + ToString(bias_learn_rate_coef_) : "");
  }

  // (split-K workspace also for the forward / backward products: at a 256-frame minibatch the output has 16 tiles, i.e.
  // 16 of 148 SMs busy for the whole K chain -- 29 us per 256 x 1024 x 1024 product measured -- unless K is split)
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    const size_t wsb = aslp_gemm_workspace_bytes(in.NumRows(), output_dim_, input_dim_);
    aslp_gemm_epilogue_t epi = {};
    epi.reduce_in_launch = ReduceInLaunch();
    ASLP_OK(aslp_gemm_ex(CuStream(), 0, 1, in.NumRows(), output_dim_, input_dim_, 1.0f, in.Data(), in.Stride(), linearity_.Data(), linearity_.Stride(),
                         0.0f, out->Data(), out->Stride(), has_bias_ ? bias_.Data() : nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb, &epi));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    const size_t wsb = aslp_gemm_workspace_bytes(out_diff.NumRows(), input_dim_, output_dim_);
    aslp_gemm_epilogue_t epi = {};
    epi.reduce_in_launch = ReduceInLaunch();
    ASLP_OK(aslp_gemm_ex(CuStream(), 0, 0, out_diff.NumRows(), input_dim_, output_dim_, 1.0f, out_diff.Data(), out_diff.Stride(), linearity_.Data(),
                         linearity_.Stride(), 0.0f, in_diff->Data(), in_diff->Stride(), nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb, &epi));
  }
  // Forward with the following Sigmoid / Tanh / ReLU applied in the product's epilogue (Nnet::Propagate pairs the two
  // components when the product takes the split-K reduce pass; the pre-activation is then never written)
  void PropagateFused(const CuMatrixBase<BaseFloat>& in, int act_kind, CuMatrixBase<BaseFloat>* act_out) {
    const size_t wsb = aslp_gemm_workspace_bytes(in.NumRows(), output_dim_, input_dim_);
    aslp_gemm_epilogue_t epi = {};
    epi.act = 1 + act_kind; epi.reduce_in_launch = ReduceInLaunch();
    ASLP_OK(aslp_gemm_ex(CuStream(), 0, 1, in.NumRows(), output_dim_, input_dim_, 1.0f, in.Data(), in.Stride(), linearity_.Data(), linearity_.Stride(),
                         0.0f, act_out->Data(), act_out->Stride(), has_bias_ ? bias_.Data() : nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb, &epi));
  }
  // Backward with the derivative of the PRECEDING activation (output y) applied in the epilogue: what this component and the
  // activation in front of it would write in two passes, d(loss)/d(input of the activation), in one
  void BackpropagateFused(const CuMatrixBase<BaseFloat>& out_diff, const CuMatrixBase<BaseFloat>& act_y, int act_kind, CuMatrixBase<BaseFloat>* act_in_diff) {
    const size_t wsb = aslp_gemm_workspace_bytes(out_diff.NumRows(), input_dim_, output_dim_);
    aslp_gemm_epilogue_t epi = {};
    epi.dact_y = act_y.Data(); epi.dact_ldy = act_y.Stride(); epi.dact_kind = act_kind; epi.reduce_in_launch = ReduceInLaunch();
    ASLP_OK(aslp_gemm_ex(CuStream(), 0, 0, out_diff.NumRows(), input_dim_, output_dim_, 1.0f, out_diff.Data(), out_diff.Stride(), linearity_.Data(),
                         linearity_.Stride(), 0.0f, act_in_diff->Data(), act_in_diff->Stride(), nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb, &epi));
  }
  // Split-K reduced inside the launch (the splits of a tile as a thread-block cluster, aslp_gemm_epilogue_t::reduce_in_launch).
  // OFF unless ASLP_GEMM_REDUCE_IN_LAUNCH=1: at par with the separate reduce pass (cfg1 0.276 ms either way, cfg4 0.956 vs
  // 0.976 ms per minibatch; profiles/r02_gemm_in_launch_reduce.txt).
  static int ReduceInLaunch() {
    static const bool on = getenv("ASLP_GEMM_REDUCE_IN_LAUNCH") != nullptr && getenv("ASLP_GEMM_REDUCE_IN_LAUNCH")[0] == '1';
    return on && CuOnComputeStream() ? 1 : 0;
  }
  // true when a product of this layer at `rows` frames goes through the split-K reduce pass, i.e. when folding a neighbouring
  // pointwise step into it costs nothing (few-tile shapes: the launch-bound minibatches)
  bool SmallBatchShape(int32 rows, bool backward) const {
    const int32 n = backward ? input_dim_ : output_dim_, k = backward ? output_dim_ : input_dim_;
    return static_cast<long long>(rows) * n * k >= (1ll << 18) && rows % 4 == 0 && aslp_gemm_workspace_bytes(rows, n, k) > 0;
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    UpdateLinearity(input, diff);
    UpdateBias(diff);
  }
  // The two halves of Update touch disjoint state (W, W_corr / bias, bias_corr), so Nnet::Backpropagate may run the weight
  // half on the side stream, under the backward products of the layers below (nnet-nnet.cc).
  void UpdateLinearity(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    aslp_stream_t st = CuStream();
    const BaseFloat lr = opts_.learn_rate * learn_rate_coef_;
    const BaseFloat mmt = opts_.momentum, l2 = opts_.l2_penalty, l1 = opts_.l1_penalty;
    const int32 num_frames = input.NumRows();
    const size_t wsb = aslp_gemm_workspace_bytes(output_dim_, input_dim_, num_frames);
    // without weight decay the SGD apply W -= lr * W_corr rides in the weight-gradient product's epilogue (the decay terms
    // below act on W between the two, so with them the apply stays a pass of its own)
    const bool fused_apply = (l2 == 0.0 && l1 == 0.0);
    aslp_gemm_epilogue_t epi = {};
    if (fused_apply) { epi.update_w = linearity_.Data(); epi.update_ldw = linearity_.Stride(); epi.update_lr = lr; }
    epi.reduce_in_launch = ReduceInLaunch();
    ASLP_OK(aslp_gemm_ex(st, 1, 0, output_dim_, input_dim_, num_frames, 1.0f, diff.Data(), diff.Stride(), input.Data(), input.Stride(), mmt,
                         linearity_corr_.Data(), linearity_corr_.Stride(), nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb, &epi));
    if (!fused_apply) {
      // (LinearTransform regularises with the bare learn rate, nnet-linear-transform.h:140-155)
      const BaseFloat lr_reg = has_bias_ ? lr : opts_.learn_rate;
      if (l2 != 0.0) linearity_.Scale(1.0f - lr_reg * l2 * num_frames);                // W += (-lr*l2*N) * W
      if (l1 != 0.0) ASLP_OK(aslp_regularize_l1(st, linearity_.Data(), linearity_.Stride(), linearity_corr_.Data(), linearity_corr_.Stride(),
                                                output_dim_, input_dim_, lr_reg * l1 * num_frames, lr_reg));
      linearity_.AddMat(-lr, linearity_corr_);
    }
    if (max_norm_ > 0.0) ASLP_OK(aslp_max_norm_rows(st, linearity_.Data(), linearity_.Stride(), output_dim_, input_dim_, max_norm_));
  }
  void UpdateBias(const CuMatrixBase<BaseFloat>& diff) {
    // bias_corr = mmt * bias_corr + column sums of diff; bias -= lr_bias * bias_corr: one launch
    if (has_bias_) ASLP_OK(aslp_bias_grad_update(CuStream(), bias_.Data(), bias_corr_.Data(), diff.Data(), diff.Stride(), diff.NumRows(), output_dim_,
                                                 opts_.momentum, opts_.learn_rate * bias_learn_rate_coef_));
  }

  const CuVector<BaseFloat>& GetBias() const { return bias_; }
  void SetBias(const CuVector<BaseFloat>& bias) { KALDI_ASSERT(bias.Dim() == bias_.Dim()); bias_ = bias; }
  const CuMatrix<BaseFloat>& GetLinearity() const { return linearity_; }
  void SetLinearity(const CuMatrixBase<BaseFloat>& l) { KALDI_ASSERT(l.NumRows() == linearity_.NumRows() && l.NumCols() == linearity_.NumCols()); linearity_.CopyFromMat(l); }
  const CuVector<BaseFloat>& GetBiasCorr() const { return bias_corr_; }
  const CuMatrix<BaseFloat>& GetLinearityCorr() const { return linearity_corr_; }

 protected:
  bool has_bias_;
  CuMatrix<BaseFloat> linearity_;
  CuVector<BaseFloat> bias_;
  CuMatrix<BaseFloat> linearity_corr_;
  CuVector<BaseFloat> bias_corr_;
  BaseFloat learn_rate_coef_, bias_learn_rate_coef_, max_norm_;
};

class LinearTransform : public AffineTransform {
 public:
  LinearTransform(int32 dim_in, int32 dim_out) : AffineTransform(dim_in, dim_out, false) {}
  Component* Copy() const { return new LinearTransform(*this); }
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
