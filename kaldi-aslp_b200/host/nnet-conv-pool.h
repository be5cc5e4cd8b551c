// nnet-conv-pool.h -- ConvolutionalComponent and MaxPoolingComponent (the CNN front end of the CTC recipes,
// aslp_scripts/aslp_nnet/run_ctc_cnn_1dnn_2blstm.sh:67-68) over aslp_gemm and the four kernels of csrc/conv_pool.cu.
// Reference: src/aslp-nnet/nnet-convolutional-component.h (proto keys :94-119, file format :167-221, Propagate :263-307,
// Backpropagate :377-400, Update :403-447: NO momentum, NO l1/l2, gradients summed over patch positions, optional max-norm)
// and src/aslp-nnet/nnet-max-pooling-component.h (:50-98, :100-156).
//   fwd   : patches = gather(in)  [frames*P, filter_dim];  out viewed as [frames*P, num_filters] = patches F^T + bias   ONE GEMM
//   bwd   : patch_diffs = out_diff F  (ONE GEMM);  in_diff = gather-sum of the patch positions that read each input column
//   update: F_grad = out_diff^T patches (ONE GEMM, split-K), b_grad = column sums; F -= lr c F_grad; b -= lr c_b b_grad
// The single-GEMM view needs the [frames, P*num_filters] matrices to be dense with aligned rows (num_filters a multiple of 4:
// every recipe's shape); otherwise the same calls run once per patch position on column ranges, as the reference does.
#ifndef ASLP_HOST_NNET_CONV_POOL_H_
#define ASLP_HOST_NNET_CONV_POOL_H_
#include "cu-workspace.h"
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class ConvolutionalComponent : public UpdatableComponent {
 public:
  ConvolutionalComponent(int32 dim_in, int32 dim_out)
      : UpdatableComponent(dim_in, dim_out), patch_dim_(0), patch_step_(0), patch_stride_(0), learn_rate_coef_(1.0f), bias_learn_rate_coef_(1.0f),
        max_norm_(0.0f) {}
  Component* Copy() const { return new ConvolutionalComponent(*this); }
  ComponentType GetType() const { return kConvolutionalComponent; }

  void InitData(std::istream& is) {
    float bias_mean = -2.0f, bias_range = 2.0f, param_stddev = 0.1f, norm_init_scale = 1.0f;
    bool gauss_init = true;
    std::string token;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<NormInit>") { ReadBasicType(is, false, &norm_init_scale); gauss_init = false; }
      else if (token == "<ParamStddev>") ReadBasicType(is, false, &param_stddev);
      else if (token == "<BiasMean>") ReadBasicType(is, false, &bias_mean);
      else if (token == "<BiasRange>") ReadBasicType(is, false, &bias_range);
      else if (token == "<PatchDim>") ReadBasicType(is, false, &patch_dim_);
      else if (token == "<PatchStep>") ReadBasicType(is, false, &patch_step_);
      else if (token == "<PatchStride>") ReadBasicType(is, false, &patch_stride_);
      else if (token == "<LearnRateCoef>") ReadBasicType(is, false, &learn_rate_coef_);
      else if (token == "<BiasLearnRateCoef>") ReadBasicType(is, false, &bias_learn_rate_coef_);
      else if (token == "<MaxNorm>") ReadBasicType(is, false, &max_norm_);
      else KALDI_ERR << "Unknown token " << token << ", a typo in config? (ParamStddev|BiasMean|BiasRange|PatchDim|PatchStep|PatchStride)";
      is >> std::ws;
    }
    const Geometry g = CheckGeometry();
    KALDI_LOG << "num_splice " << g.num_splice;
    KALDI_LOG << "num_patches " << g.num_patches;
    KALDI_LOG << "filter_dim " << g.filter_dim;
    KALDI_LOG << "num_filters " << g.num_filters;
    filters_.Resize(g.num_filters, g.filter_dim);
    bias_.Resize(g.num_filters);
    if (!gauss_init) {
      const float scale = norm_init_scale * sqrt(6.0 / (g.num_filters + g.filter_dim));
      InitMatParam(&filters_, scale);
      InitVecParam(&bias_, scale);
    } else {                                   // same draw order as the reference: all filter rows, then the biases
      Matrix<BaseFloat> mat(g.num_filters, g.filter_dim);
      for (int32 r = 0; r < g.num_filters; r++)
        for (int32 c = 0; c < g.filter_dim; c++) mat(r, c) = param_stddev * RandGauss();
      filters_ = mat;
      Vector<BaseFloat> vec(g.num_filters);
      for (int32 i = 0; i < g.num_filters; i++) vec(i) = bias_mean + (RandUniform() - 0.5) * bias_range;
      bias_ = vec;
    }
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<PatchDim>"); ReadBasicType(is, binary, &patch_dim_);
    ExpectToken(is, binary, "<PatchStep>"); ReadBasicType(is, binary, &patch_step_);
    ExpectToken(is, binary, "<PatchStride>"); ReadBasicType(is, binary, &patch_stride_);
    ExpectToken(is, binary, "<LearnRateCoef>"); ReadBasicType(is, binary, &learn_rate_coef_);
    ExpectToken(is, binary, "<BiasLearnRateCoef>"); ReadBasicType(is, binary, &bias_learn_rate_coef_);
    ExpectToken(is, binary, "<MaxNorm>"); ReadBasicType(is, binary, &max_norm_);
    ExpectToken(is, binary, "<Filters>"); filters_.Read(is, binary);
    ExpectToken(is, binary, "<Bias>"); bias_.Read(is, binary);
    const Geometry g = CheckGeometry();
    KALDI_ASSERT(g.num_filters == filters_.NumRows() && g.num_filters == bias_.Dim() && g.filter_dim == filters_.NumCols());
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<PatchDim>"); WriteBasicType(os, binary, patch_dim_);
    WriteToken(os, binary, "<PatchStep>"); WriteBasicType(os, binary, patch_step_);
    WriteToken(os, binary, "<PatchStride>"); WriteBasicType(os, binary, patch_stride_);
    WriteToken(os, binary, "<LearnRateCoef>"); WriteBasicType(os, binary, learn_rate_coef_);
    WriteToken(os, binary, "<BiasLearnRateCoef>"); WriteBasicType(os, binary, bias_learn_rate_coef_);
    WriteToken(os, binary, "<MaxNorm>"); WriteBasicType(os, binary, max_norm_);
    WriteToken(os, binary, "<Filters>"); filters_.Write(os, binary);
    WriteToken(os, binary, "<Bias>"); bias_.Write(os, binary);
  }
  int32 NumParams() const { return filters_.NumRows() * filters_.NumCols() + bias_.Dim(); }
  void GetParams(Vector<BaseFloat>* wei_copy) const {
    wei_copy->Resize(NumParams());
    CopyRowsToVec(filters_, wei_copy->Data());
    Vector<float> b;
    bias_.CopyToVec(&b);
    for (int32 i = 0; i < b.Dim(); ++i) (*wei_copy)(filters_.NumRows() * filters_.NumCols() + i) = b(i);
  }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
    params->clear();
    params->push_back(std::make_pair(filters_.Data(), filters_.NumRows() * filters_.Stride()));
    params->push_back(std::make_pair(bias_.Data(), bias_.Dim()));
  }
  std::string Info() const { return std::string("\n  filters") + MomentStatistics(filters_) + "\n  bias" + MomentStatistics(bias_); }
  std::string InfoGradient() const {
    return std::string("\n  filters_grad") + MomentStatistics(filters_grad_) + ", lr-coef " + ToString(learn_rate_coef_) + ", max-norm " +
           ToString(max_norm_) + "\n  bias_grad" + MomentStatistics(bias_grad_) + ", lr-coef " + ToString(bias_learn_rate_coef_);
  }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    const Geometry g = CheckGeometry();
    const int32 rows = in.NumRows();
    aslp_stream_t st = CuStream();
    patches_.Resize(rows * g.num_patches, g.filter_dim, kUndefined);
    ASLP_OK(aslp_conv_gather_patches(st, patches_.Data(), patches_.Stride(), in.Data(), in.Stride(), rows, g.num_patches, g.num_splice, patch_dim_,
                                     patch_step_, patch_stride_));
    if (Dense(g, out->Stride())) {
      ASLP_OK(aslp_gemm(st, 0, 1, rows * g.num_patches, g.num_filters, g.filter_dim, 1.0f, patches_.Data(), patches_.Stride(), filters_.Data(),
                        filters_.Stride(), 0.0f, out->Data(), g.num_filters, bias_.Data(), 0.0f, GemmPrecision(), nullptr, 0));
    } else {
      for (int32 p = 0; p < g.num_patches; p++)
        ASLP_OK(aslp_gemm(st, 0, 1, rows, g.num_filters, g.filter_dim, 1.0f, patches_.Data() + static_cast<size_t>(p) * patches_.Stride(),
                          g.num_patches * patches_.Stride(), filters_.Data(), filters_.Stride(), 0.0f, out->Data() + p * g.num_filters, out->Stride(),
                          bias_.Data(), 0.0f, GemmPrecision(), nullptr, 0));
    }
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    const Geometry g = CheckGeometry();
    const int32 rows = out_diff.NumRows();
    aslp_stream_t st = CuStream();
    patch_diffs_.Resize(rows * g.num_patches, g.filter_dim, kUndefined);
    if (Dense(g, out_diff.Stride())) {
      ASLP_OK(aslp_gemm(st, 0, 0, rows * g.num_patches, g.filter_dim, g.num_filters, 1.0f, out_diff.Data(), g.num_filters, filters_.Data(),
                        filters_.Stride(), 0.0f, patch_diffs_.Data(), patch_diffs_.Stride(), nullptr, 0.0f, GemmPrecision(), nullptr, 0));
    } else {
      for (int32 p = 0; p < g.num_patches; p++)
        ASLP_OK(aslp_gemm(st, 0, 0, rows, g.filter_dim, g.num_filters, 1.0f, out_diff.Data() + p * g.num_filters, out_diff.Stride(), filters_.Data(),
                          filters_.Stride(), 0.0f, patch_diffs_.Data() + static_cast<size_t>(p) * patch_diffs_.Stride(),
                          g.num_patches * patch_diffs_.Stride(), nullptr, 0.0f, GemmPrecision(), nullptr, 0));
    }
    ASLP_OK(aslp_conv_scatter_patch_diffs(st, in_diff->Data(), in_diff->Stride(), patch_diffs_.Data(), patch_diffs_.Stride(), rows, g.num_patches,
                                          g.num_splice, patch_dim_, patch_step_, patch_stride_));
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    const Geometry g = CheckGeometry();
    const int32 rows = diff.NumRows();
    KALDI_ASSERT(patches_.NumRows() == rows * g.num_patches);           // the patches of the Propagate this diff belongs to
    aslp_stream_t st = CuStream();
    const BaseFloat lr = opts_.learn_rate;
    filters_grad_.Resize(g.num_filters, g.filter_dim, kUndefined);      // no momentum: the gradient is rebuilt every time (:410-411)
    bias_grad_.Resize(g.num_filters, kUndefined);
    if (Dense(g, diff.Stride())) {
      const int32 k = rows * g.num_patches;
      const size_t wsb = aslp_gemm_workspace_bytes(g.num_filters, g.filter_dim, k);
      ASLP_OK(aslp_gemm(st, 1, 0, g.num_filters, g.filter_dim, k, 1.0f, diff.Data(), g.num_filters, patches_.Data(), patches_.Stride(), 0.0f,
                        filters_grad_.Data(), filters_grad_.Stride(), nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb));
      ASLP_OK(aslp_col_sum(st, bias_grad_.Data(), diff.Data(), g.num_filters, k, g.num_filters, 1.0f, 0.0f, 0.0f));
    } else {
      const size_t wsb = aslp_gemm_workspace_bytes(g.num_filters, g.filter_dim, rows);
      for (int32 p = 0; p < g.num_patches; p++) {
        const float beta = p == 0 ? 0.0f : 1.0f;
        ASLP_OK(aslp_gemm(st, 1, 0, g.num_filters, g.filter_dim, rows, 1.0f, diff.Data() + p * g.num_filters, diff.Stride(),
                          patches_.Data() + static_cast<size_t>(p) * patches_.Stride(), g.num_patches * patches_.Stride(), beta, filters_grad_.Data(),
                          filters_grad_.Stride(), nullptr, 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb));
        // the column range of a patch position does not start on a 16-byte boundary here: reduce a dense copy of it
        diff_patch_.Resize(rows, g.num_filters, kUndefined);
        ASLP_OK(aslp_memcpy2d_d2d(st, diff_patch_.Data(), sizeof(float) * diff_patch_.Stride(), diff.Data() + p * g.num_filters,
                                  sizeof(float) * diff.Stride(), sizeof(float) * g.num_filters, rows));
        ASLP_OK(aslp_col_sum(st, bias_grad_.Data(), diff_patch_.Data(), diff_patch_.Stride(), rows, g.num_filters, 1.0f, beta, 0.0f));
      }
    }
    filters_.AddMat(-lr * learn_rate_coef_, filters_grad_);
    const int32 ldb = (g.num_filters + 3) / 4 * 4;
    ASLP_OK(aslp_axpby(st, bias_.Data(), ldb, bias_grad_.Data(), ldb, 1, g.num_filters, -lr * bias_learn_rate_coef_, 1.0f));
    if (max_norm_ > 0.0) ASLP_OK(aslp_max_norm_rows(st, filters_.Data(), filters_.Stride(), g.num_filters, g.filter_dim, max_norm_));
  }

 private:
  struct Geometry { int32 num_splice, num_patches, filter_dim, num_filters; };
  Geometry CheckGeometry() const {             // the reference's sanity checks (:124-138)
    KALDI_ASSERT(patch_dim_ > 0 && patch_step_ > 0 && patch_stride_ > 0);
    KALDI_ASSERT(input_dim_ % patch_stride_ == 0);
    KALDI_ASSERT((patch_stride_ - patch_dim_) % patch_step_ == 0);
    Geometry g;
    g.num_splice = input_dim_ / patch_stride_;
    g.num_patches = 1 + (patch_stride_ - patch_dim_) / patch_step_;
    g.filter_dim = g.num_splice * patch_dim_;
    KALDI_ASSERT(output_dim_ % g.num_patches == 0);
    g.num_filters = output_dim_ / g.num_patches;
    return g;
  }
  // the [frames, P*num_filters] matrix can be addressed as a row-major [frames*P, num_filters] one with 16-byte aligned rows
  static bool Dense(const Geometry& g, int32 stride) { return stride == g.num_patches * g.num_filters && g.num_filters % 4 == 0; }
  int32 patch_dim_, patch_step_, patch_stride_;
  CuMatrix<BaseFloat> filters_;                           // row = vectorised rectangular filter [num_filters, filter_dim]
  CuVector<BaseFloat> bias_;
  CuMatrix<BaseFloat> filters_grad_;
  CuVector<BaseFloat> bias_grad_;
  BaseFloat learn_rate_coef_, bias_learn_rate_coef_, max_norm_;
  CuMatrix<BaseFloat> patches_;                           // [frames * num_patches, filter_dim] of the last Propagate
  CuMatrix<BaseFloat> patch_diffs_;
  CuMatrix<BaseFloat> diff_patch_;                        // dense copy of one patch position's derivative columns (odd shapes only)
};

class MaxPoolingComponent : public Component {
 public:
  MaxPoolingComponent(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out), pool_size_(0), pool_step_(0), pool_stride_(0) {}
  Component* Copy() const { return new MaxPoolingComponent(*this); }
  ComponentType GetType() const { return kMaxPoolingComponent; }
  void InitData(std::istream& is) {
    std::string token;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<PoolSize>") ReadBasicType(is, false, &pool_size_);
      else if (token == "<PoolStep>") ReadBasicType(is, false, &pool_step_);
      else if (token == "<PoolStride>") ReadBasicType(is, false, &pool_stride_);
      else KALDI_ERR << "Unknown token " << token << ", a typo in config? (PoolSize|PoolStep|PoolStride)";
      is >> std::ws;
    }
    KALDI_ASSERT(pool_size_ != 0 && pool_step_ != 0 && pool_stride_ != 0);
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<PoolSize>"); ReadBasicType(is, binary, &pool_size_);
    ExpectToken(is, binary, "<PoolStep>"); ReadBasicType(is, binary, &pool_step_);
    ExpectToken(is, binary, "<PoolStride>"); ReadBasicType(is, binary, &pool_stride_);
    int32 num_patches, num_pools;
    PoolGeometry(&num_patches, &num_pools);
    KALDI_ASSERT(output_dim_ == num_pools * pool_stride_);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<PoolSize>"); WriteBasicType(os, binary, pool_size_);
    WriteToken(os, binary, "<PoolStep>"); WriteBasicType(os, binary, pool_step_);
    WriteToken(os, binary, "<PoolStride>"); WriteBasicType(os, binary, pool_stride_);
  }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    int32 num_patches, num_pools;
    PoolGeometry(&num_patches, &num_pools);
    KALDI_ASSERT(output_dim_ == num_pools * pool_stride_);
    ASLP_OK(aslp_maxpool_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), num_pools, pool_size_, pool_step_, pool_stride_));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    int32 num_patches, num_pools;
    PoolGeometry(&num_patches, &num_pools);
    ASLP_OK(aslp_maxpool_bwd(CuStream(), in_diff->Data(), in_diff->Stride(), in.Data(), in.Stride(), out.Data(), out.Stride(), out_diff.Data(),
                             out_diff.Stride(), in.NumRows(), num_patches, num_pools, pool_size_, pool_step_, pool_stride_));
  }

 private:
  void PoolGeometry(int32* num_patches, int32* num_pools) const {
    KALDI_ASSERT(pool_size_ > 0 && pool_step_ > 0 && pool_stride_ > 0);
    KALDI_ASSERT(input_dim_ % pool_stride_ == 0);
    *num_patches = input_dim_ / pool_stride_;
    KALDI_ASSERT((*num_patches - pool_size_) % pool_step_ == 0);
    *num_pools = 1 + (*num_patches - pool_size_) / pool_step_;
  }
  int32 pool_size_, pool_step_, pool_stride_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
