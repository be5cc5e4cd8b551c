// nnet-misc-components.h -- BatchNormalization, CompactFsmn, RowConvolution over their fused kernels.
// Reference: src/aslp-nnet/nnet-batch-normalization.h:139-284, nnet-cfsmn-component.h:169-264,
// nnet-row-convolution.{h,cc}.
#ifndef ASLP_HOST_NNET_MISC_COMPONENTS_H_
#define ASLP_HOST_NNET_MISC_COMPONENTS_H_
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class BatchNormalization : public UpdatableComponent {
 public:
  BatchNormalization(int32 dim_in, int32 dim_out)
      : UpdatableComponent(dim_in, dim_out), var_floor_(0.0000001f), num_acc_frames_(0), acc_cleaned_(false) {}
  Component* Copy() const { return new BatchNormalization(*this); }
  ComponentType GetType() const { return kBatchNormalization; }

  void InitData(std::istream& is) {
    num_acc_frames_ = 0;
    scale_.Resize(output_dim_); scale_.Set(1.0f);
    shift_.Resize(output_dim_, kSetZero);
    KALDI_ASSERT(output_dim_ > 0 && input_dim_ > 0);
    acc_means_.Resize(output_dim_, kSetZero);
    acc_vars_.Resize(output_dim_, kSetZero);
    AllocWork();
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<NumAccFrames>");
    ReadBasicType(is, binary, &num_acc_frames_);
    acc_means_.Read(is, binary);
    acc_vars_.Read(is, binary);
    shift_.Read(is, binary);
    scale_.Read(is, binary);
    KALDI_ASSERT(acc_means_.Dim() == acc_vars_.Dim() && acc_means_.Dim() == shift_.Dim() && acc_means_.Dim() == scale_.Dim());
    AllocWork();
    if (num_acc_frames_ <= 0.0) return;
    // global statistics for the eval path (nnet-batch-normalization.h:76-93)
    const float var_floor = 1e-10f;
    Vector<double> acc_mean, acc_var;
    acc_means_.CopyToVec(&acc_mean);
    acc_vars_.CopyToVec(&acc_var);
    Vector<BaseFloat> mean_h(acc_mean.Dim()), ivar_h(acc_mean.Dim());
    for (int32 d = 0; d < acc_mean.Dim(); d++) {
      const BaseFloat mean = acc_mean(d) / num_acc_frames_;
      BaseFloat var = acc_var(d) / num_acc_frames_ - mean * mean;
      if (var <= var_floor) { KALDI_WARN << "Very small variance " << var << " flooring to " << var_floor; var = var_floor; }
      mean_h(d) = mean;
      ivar_h(d) = 1.0 / sqrt(var + var_floor_);
    }
    mean_vec_ = mean_h;
    var_vec_ = ivar_h;
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<NumAccFrames>");
    WriteBasicType(os, binary, num_acc_frames_);
    // quirk kept for byte compatibility: the reference's CuVector<Real>::Write goes through a Vector<BaseFloat>
    // (src/aslp-cudamatrix/cu-vector.cc:851-855), so the fp64 running sums land on disk as "FV"; Read promotes them back
    WriteAsFloat(acc_means_, os, binary);
    WriteAsFloat(acc_vars_, os, binary);
    shift_.Write(os, binary);
    scale_.Write(os, binary);
  }
  int32 NumParams() const { return shift_.Dim() + scale_.Dim(); }
  void GetParams(Vector<BaseFloat>* w) const {
    w->Resize(NumParams());
    Vector<float> a, b;
    shift_.CopyToVec(&a); scale_.CopyToVec(&b);
    for (int32 i = 0; i < a.Dim(); ++i) (*w)(i) = a(i);
    for (int32 i = 0; i < b.Dim(); ++i) (*w)(a.Dim() + i) = b(i);
  }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* p) {
    p->clear();
    p->push_back(std::make_pair(shift_.Data(), shift_.Dim()));
    p->push_back(std::make_pair(scale_.Data(), scale_.Dim()));
  }
  // device fp64 running sums + host frame counter for the end-of-epoch allreduce (mpi-node.h:76-91)
  double* GetAccStats(std::vector<std::pair<double*, int>>* params) {
    params->clear();
    params->push_back(std::make_pair(acc_means_.Data(), acc_means_.Dim()));
    params->push_back(std::make_pair(acc_vars_.Data(), acc_vars_.Dim()));
    return &num_acc_frames_;
  }
  std::string Info() const { return std::string("\n  batch_normaliztion"); }
  void CleanAccs() { acc_means_.SetZero(); acc_vars_.SetZero(); num_acc_frames_ = 0; }

  void FeedforwardFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    aslp_stream_t st = CuStream();
    if (num_acc_frames_ <= 0) {       // local statistics, nothing accumulated
      ASLP_OK(aslp_bn_fwd_train(st, out->Data(), out->Stride(), nullptr, 0, in.Data(), in.Stride(), in.NumRows(), output_dim_,
                                scale_.Data(), shift_.Data(), var_floor_, mean_vec_.Data(), var_vec_.Data(), nullptr, nullptr));
    } else {
      ASLP_OK(aslp_bn_fwd_eval(st, out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), output_dim_, scale_.Data(), shift_.Data(),
                               mean_vec_.Data(), var_vec_.Data()));
    }
  }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    if (!acc_cleaned_) { acc_cleaned_ = true; CleanAccs(); }      // running sums restart with the first training propagate (:178-181)
    // no x-hat buffer: the backward kernels recompute it from in, mean_vec_ and var_vec_ (= 1/std)
    ASLP_OK(aslp_bn_fwd_train(CuStream(), out->Data(), out->Stride(), nullptr, 0, in.Data(), in.Stride(), in.NumRows(), output_dim_,
                              scale_.Data(), shift_.Data(), var_floor_, mean_vec_.Data(), var_vec_.Data(), acc_means_.Data(), acc_vars_.Data()));
    num_acc_frames_ += in.NumRows();
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_bn_bwd(CuStream(), in_diff ? in_diff->Data() : nullptr, in_diff ? in_diff->Stride() : 0, in.Data(), in.Stride(), nullptr,
                        0, out_diff.Data(), out_diff.Stride(), in.NumRows(), output_dim_, scale_.Data(), mean_vec_.Data(), var_vec_.Data(),
                        opts_.momentum, dscale_.Data(), dshift_.Data()));
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {       // plain -lr * d (no lr coefficient, :279-283)
    const BaseFloat lr = opts_.learn_rate;
    const int32 ld = (output_dim_ + 3) / 4 * 4;
    ASLP_OK(aslp_axpby(CuStream(), scale_.Data(), ld, dscale_.Data(), ld, 1, output_dim_, -lr, 1.0f));
    ASLP_OK(aslp_axpby(CuStream(), shift_.Data(), ld, dshift_.Data(), ld, 1, output_dim_, -lr, 1.0f));
  }

 private:
  static void WriteAsFloat(const CuVector<double>& v, std::ostream& os, bool binary) {
    Vector<double> d;
    v.CopyToVec(&d);
    Vector<float> f(d.Dim());
    for (int32 i = 0; i < d.Dim(); ++i) f(i) = static_cast<float>(d(i));
    f.Write(os, binary);
  }
  void AllocWork() {
    mean_vec_.Resize(output_dim_, kSetZero);
    var_vec_.Resize(output_dim_); var_vec_.Set(1.0f);
    dscale_.Resize(output_dim_, kSetZero);
    dshift_.Resize(output_dim_, kSetZero);
  }
  CuVector<BaseFloat> mean_vec_, var_vec_, scale_, dscale_, shift_, dshift_;
  BaseFloat var_floor_;
  CuVector<double> acc_means_, acc_vars_;
  double num_acc_frames_;
  bool acc_cleaned_;
};

class CompactFsmn : public UpdatableComponent {
 public:
  CompactFsmn(int32 dim_in, int32 dim_out)
      : UpdatableComponent(dim_in, dim_out), max_frames_(3000), learn_rate_coef_(1.0f), past_context_(0), future_context_(0), clip_gradient_(0.0f) {}
  Component* Copy() const { return new CompactFsmn(*this); }
  ComponentType GetType() const { return kCompactFsmn; }

  void InitData(std::istream& is) {
    int32 past_context = 30, future_context = 30;
    float vec_coef_mean = 0.0f, vec_coef_range = 1.0f;
    ProtoOptions po("(PastContext|FutureContext|VecCoefMean|VecCoefRange|LearnRateCoef)");
    po.Int("<PastContext>", &past_context); po.Int("<FutureContext>", &future_context);
    po.Float("<LearnRateCoef>", &learn_rate_coef_); po.Float("<VecCoefMean>", &vec_coef_mean); po.Float("<VecCoefRange>", &vec_coef_range);
    po.Float("<ClipGradient>", &clip_gradient_);
    po.Parse(is);
    const int32 num_row = past_context + future_context + 1, num_col = input_dim_;
    const float param_scale = 0.5 * sqrt(6.0 / (num_col + num_row));
    // host Matrix::SetRandUniform with a fresh RandomState, then (u - 0.5) * 2 * scale  (nnet-cfsmn-component.h:52-56,86-89)
    Matrix<BaseFloat> m(num_row, num_col);
    RandomState rs;
    for (int32 r = 0; r < num_row; ++r)
      for (int32 c = 0; c < num_col; ++c) { float u = RandUniform(&rs); u += -0.5f; u *= 2 * param_scale; m(r, c) = u; }
    vec_coef_ = m;
    vec_coef_corr_.Resize(num_row, num_col, kSetZero);
    past_context_ = past_context; future_context_ = future_context;
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<PastContext>"); ReadBasicType(is, binary, &past_context_);
    ExpectToken(is, binary, "<FutureContext>"); ReadBasicType(is, binary, &future_context_);
    ExpectToken(is, binary, "<LearnRateCoef>"); ReadBasicType(is, binary, &learn_rate_coef_);
    vec_coef_.Read(is, binary);
    vec_coef_corr_.Resize(vec_coef_.NumRows(), vec_coef_.NumCols(), kSetZero);
    KALDI_ASSERT(vec_coef_.NumCols() == input_dim_);
    KALDI_ASSERT(vec_coef_.NumRows() == past_context_ + future_context_ + 1);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<PastContext>"); WriteBasicType(os, binary, past_context_);
    WriteToken(os, binary, "<FutureContext>"); WriteBasicType(os, binary, future_context_);
    WriteToken(os, binary, "<LearnRateCoef>"); WriteBasicType(os, binary, learn_rate_coef_);
    vec_coef_.Write(os, binary);
  }
  int32 NumParams() const { return vec_coef_.NumRows() * vec_coef_.NumCols(); }
  void GetParams(Vector<BaseFloat>* w) const { w->Resize(NumParams()); CopyRowsToVec(vec_coef_, w->Data()); }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* p) {
    p->clear();
    p->push_back(std::make_pair(vec_coef_.Data(), vec_coef_.NumRows() * vec_coef_.Stride()));
  }
  std::string Info() const { return std::string("\n  vec_coef") + MomentStatistics(vec_coef_); }
  std::string InfoGradient() const { return std::string("\n  vec_coef_grad") + MomentStatistics(vec_coef_corr_) + ", lr-coef " + ToString(learn_rate_coef_); }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    KALDI_ASSERT(in.NumRows() <= max_frames_);             // the reference's hard limit (:38,174)
    KALDI_ASSERT(in.NumCols() == vec_coef_.NumCols());
    ASLP_OK(aslp_fsmn_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), in.NumCols(), vec_coef_.Data(),
                          vec_coef_.Stride(), past_context_, future_context_));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    KALDI_ASSERT(in.NumRows() <= max_frames_);
    aslp_stream_t st = CuStream();
    ASLP_OK(aslp_fsmn_coef_grad(st, vec_coef_corr_.Data(), vec_coef_corr_.Stride(), in.Data(), in.Stride(), out_diff.Data(), out_diff.Stride(),
                                in.NumRows(), in.NumCols(), past_context_, future_context_, clip_gradient_));
    ASLP_OK(aslp_fsmn_bwd(st, in_diff->Data(), in_diff->Stride(), out_diff.Data(), out_diff.Stride(), in.NumRows(), in.NumCols(), vec_coef_.Data(),
                          vec_coef_.Stride(), past_context_, future_context_));
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    vec_coef_.AddMat(-opts_.learn_rate * learn_rate_coef_, vec_coef_corr_);
  }
 private:
  CuMatrix<BaseFloat> vec_coef_, vec_coef_corr_;
  int32 max_frames_;
  BaseFloat learn_rate_coef_;
  int32 past_context_, future_context_;
  BaseFloat clip_gradient_;
};

class RowConvolution : public UpdatableComponent {
 public:
  RowConvolution(int32 dim_in, int32 dim_out) : UpdatableComponent(dim_in, dim_out), future_ctx_(0) {}
  Component* Copy() const { return new RowConvolution(*this); }
  ComponentType GetType() const { return kRowConvolution; }
  void InitData(std::istream& is) {
    ProtoOptions po("(FutureContext)"); po.Int("<FutureContext>", &future_ctx_); po.Parse(is);
    KALDI_ASSERT(future_ctx_ > 0);
    Matrix<BaseFloat> mat(input_dim_, future_ctx_ + 1);
    for (int r = 0; r < mat.NumRows(); r++)
      for (int c = 0; c < mat.NumCols(); c++) mat(r, c) = 1.0 * RandGauss();
    w_ = mat;
    AllocWork();
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<FutureContext>"); ReadBasicType(is, binary, &future_ctx_);
    w_.Read(is, binary);
    AllocWork();
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<FutureContext>"); WriteBasicType(os, binary, future_ctx_);
    w_.Write(os, binary);
  }
  int32 NumParams() const { return w_.NumRows() * w_.NumCols(); }
  void GetParams(Vector<BaseFloat>* w) const { w->Resize(NumParams()); CopyRowsToVec(w_, w->Data()); }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* p) { p->clear(); p->push_back(std::make_pair(w_.Data(), w_.NumRows() * w_.Stride())); }
  std::string Info() const { return std::string("  ") + "\n  w_ " + MomentStatistics(w_); }
  std::string InfoGradient() const { return std::string("  ") + "\n w_diff_ " + MomentStatistics(w_diff_) + "\n w_corr_ " + MomentStatistics(w_corr_); }
  void SetSeqLengths(const std::vector<int32>& sequence_lengths) { sequence_lengths_ = sequence_lengths; seq_len_dev_ = sequence_lengths; }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    const int32 S = static_cast<int32>(sequence_lengths_.size());
    KALDI_ASSERT(S > 0 && in.NumRows() % S == 0);
    ASLP_OK(aslp_rowconv_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows() / S, S, input_dim_, w_.Data(), w_.Stride(),
                             future_ctx_, seq_len_dev_.Data()));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    const int32 S = static_cast<int32>(sequence_lengths_.size());
    ASLP_OK(aslp_rowconv_bwd(CuStream(), in_diff->Data(), in_diff->Stride(), w_diff_.Data(), w_diff_.Stride(), in.Data(), in.Stride(), out_diff.Data(),
                             out_diff.Stride(), in.NumRows() / S, S, input_dim_, w_.Data(), w_.Stride(), future_ctx_, seq_len_dev_.Data()));
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {     // w_corr = mmt*w_corr + w_diff ; w -= lr*w_corr (.cc:161-169)
    ASLP_OK(aslp_axpby(CuStream(), w_corr_.Data(), w_corr_.Stride(), w_diff_.Data(), w_diff_.Stride(), w_corr_.NumRows(), w_corr_.NumCols(), 1.0f, opts_.momentum));
    w_.AddMat(-opts_.learn_rate, w_corr_);
  }
 private:
  void AllocWork() { w_diff_.Resize(input_dim_, future_ctx_ + 1, kSetZero); w_corr_.Resize(input_dim_, future_ctx_ + 1, kSetZero); }
  int32 future_ctx_;
  CuMatrix<BaseFloat> w_, w_diff_, w_corr_;
  std::vector<int32> sequence_lengths_;
  CuArrayInt seq_len_dev_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
