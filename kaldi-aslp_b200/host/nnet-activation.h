// nnet-activation.h -- Softmax, Sigmoid, Tanh, ReLU (src/aslp-nnet/nnet-activation.h:35-60,153-199,276-298)
// and the copy layers InputLayer / OutputLayer / ScaleLayer (src/aslp-nnet/nnet-io.h:20-102),
// Splice / AddShift / Rescale (src/aslp-nnet/nnet-various.h).  Each pass is one vectorised HBM-bound kernel.
#ifndef ASLP_HOST_NNET_ACTIVATION_H_
#define ASLP_HOST_NNET_ACTIVATION_H_
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class Softmax : public Component {
 public:
  Softmax(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new Softmax(*this); }
  ComponentType GetType() const { return kSoftmax; }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_softmax_rows(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), in.NumCols()));
  }
  // the loss already delivers y - t: the backward pass is a copy (nnet-activation.h:51-59)
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    in_diff->CopyFromMat(out_diff);
  }
};

template <int KIND, Component::ComponentType TYPE>
class PointwiseActivation : public Component {
 public:
  PointwiseActivation(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new PointwiseActivation(*this); }
  ComponentType GetType() const { return TYPE; }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_act_fwd(CuStream(), KIND, out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), in.NumCols()));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    // sigmoid / tanh use y; the ReLU mask Heaviside(x) is taken from y as well (y = max(x, 0) is positive exactly where x is), so
    // that no activation needs its input again -- Nnet may have folded the forward pass into the producing Affine's epilogue
    const CuMatrixBase<BaseFloat>& ref = out;
    ASLP_OK(aslp_act_bwd(CuStream(), KIND, in_diff->Data(), in_diff->Stride(), ref.Data(), ref.Stride(), out_diff.Data(), out_diff.Stride(),
                         out_diff.NumRows(), out_diff.NumCols()));
  }
};
typedef PointwiseActivation<ASLP_ACT_SIGMOID, Component::kSigmoid> Sigmoid;
typedef PointwiseActivation<ASLP_ACT_TANH, Component::kTanh> Tanh;
typedef PointwiseActivation<ASLP_ACT_RELU, Component::kReLU> ReLU;

// ---- copy layers (nnet-io.h) ----
template <Component::ComponentType TYPE>
class CopyLayer : public Component {
 public:
  CopyLayer(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new CopyLayer(*this); }
  ComponentType GetType() const { return TYPE; }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) { out->CopyFromMat(in); }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) { in_diff->CopyFromMat(out_diff); }
};
typedef CopyLayer<Component::kInputLayer> InputLayer;
typedef CopyLayer<Component::kOutputLayer> OutputLayer;

class ScaleLayer : public Component {
 public:
  ScaleLayer(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out), scale_(1.0f) {}
  Component* Copy() const { return new ScaleLayer(*this); }
  ComponentType GetType() const { return kScaleLayer; }
  void InitData(std::istream& is) { ProtoOptions po("(Scale)"); po.Float("<Scale>", &scale_); po.Parse(is); }
  void ReadData(std::istream& is, bool binary) { ExpectToken(is, binary, "<Scale>"); ReadBasicType(is, binary, &scale_); }
  void WriteData(std::ostream& os, bool binary) const { WriteToken(os, binary, "<Scale>"); WriteBasicType(os, binary, scale_); }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_axpby(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), in.NumCols(), scale_, 0.0f));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_axpby(CuStream(), in_diff->Data(), in_diff->Stride(), out_diff.Data(), out_diff.Stride(), out_diff.NumRows(), out_diff.NumCols(), scale_, 0.0f));
  }
 private:
  float scale_;
};

// ---- Splice (nnet-various.h:43-178) ----
class Splice : public Component {
 public:
  Splice(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new Splice(*this); }
  ComponentType GetType() const { return kSplice; }
  void InitData(std::istream& is) {
    std::vector<int32> frame_offsets;
    std::string token;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<ReadVector>") {
        ReadIntegerVector(is, false, &frame_offsets);
      } else if (token == "<BuildVector>") {          // e.g. <BuildVector> -5:5 </BuildVector>  or  -5:1:5  or single values
        while (!is.eof()) {
          std::string item;
          ReadToken(is, false, &item);
          if (item == "</BuildVector>") break;
          std::vector<int32> v;
          if (!SplitStringToIntegers(item, ":", false, &v) || v.empty() || v.size() > 3) KALDI_ERR << "Error parsing <BuildVector>";
          if (v.size() == 1) frame_offsets.push_back(v[0]);
          else {
            const int32 lo = v[0], hi = v.back(), step = v.size() == 3 ? v[1] : 1;
            KALDI_ASSERT((lo <= hi && step > 0) || (lo >= hi && step < 0));
            if (v.size() == 2 || step > 0) { for (int32 j = lo; j <= hi; j += step) frame_offsets.push_back(j); }
            else { for (int32 j = lo; j <= hi; j += step) frame_offsets.push_back(j); }   // reference loop: runs only while j <= max
          }
        }
      } else {
        KALDI_ERR << "Unknown token " << token << ", a typo in config? (ReadVector|BuildVector)";
      }
      is >> std::ws;
    }
    frame_offsets_ = frame_offsets;
    KALDI_ASSERT(frame_offsets_.Dim() * InputDim() == OutputDim());
  }
  void ReadData(std::istream& is, bool binary) {
    std::vector<int32> fo;
    ReadIntegerVector(is, binary, &fo);
    frame_offsets_ = fo;
    KALDI_ASSERT(frame_offsets_.Dim() * InputDim() == OutputDim());
  }
  void WriteData(std::ostream& os, bool binary) const { WriteIntegerVector(os, binary, frame_offsets_.Host()); }
  std::string Info() const {
    std::ostringstream os;
    os << "\n  frame_offsets [ ";
    for (int32 v : frame_offsets_.Host()) os << v << " ";
    os << "]";
    return os.str();
  }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_splice_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), input_dim_, frame_offsets_.Data(), frame_offsets_.Dim()));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_splice_bwd(CuStream(), in_diff->Data(), in_diff->Stride(), out_diff.Data(), out_diff.Stride(), in.NumRows(), input_dim_,
                            frame_offsets_.Data(), frame_offsets_.Dim()));
  }
 protected:
  CuArrayInt frame_offsets_;
};

// ---- AddShift / Rescale (nnet-various.h): per-dimension shift / scale, used as fixed feature transforms ----
class AddShift : public UpdatableComponent {
 public:
  AddShift(int32 dim_in, int32 dim_out) : UpdatableComponent(dim_in, dim_out), shift_data_(dim_in), learn_rate_coef_(0.0f) {}
  Component* Copy() const { return new AddShift(*this); }
  ComponentType GetType() const { return kAddShift; }
  void InitData(std::istream& is) {
    float init_param = 0.0f;
    ProtoOptions po("(InitParam|LearnRateCoef)"); po.Float("<InitParam>", &init_param); po.Float("<LearnRateCoef>", &learn_rate_coef_); po.Parse(is);
    shift_data_.Set(init_param);
  }
  void ReadData(std::istream& is, bool binary) {
    if ('<' == Peek(is, binary)) { ExpectToken(is, binary, "<LearnRateCoef>"); ReadBasicType(is, binary, &learn_rate_coef_); }
    shift_data_.Read(is, binary);
    KALDI_ASSERT(shift_data_.Dim() == output_dim_);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<LearnRateCoef>"); WriteBasicType(os, binary, learn_rate_coef_);
    shift_data_.Write(os, binary);
  }
  int32 NumParams() const { return shift_data_.Dim(); }
  void GetParams(Vector<BaseFloat>* w) const { shift_data_.CopyToVec(w); }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* p) { p->clear(); p->push_back(std::make_pair(shift_data_.Data(), shift_data_.Dim())); }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    out->CopyFromMat(in);
    ASLP_OK(aslp_add_vec_to_rows(CuStream(), out->Data(), out->Stride(), out->NumRows(), out->NumCols(), shift_data_.Data(), 1.0f, 1.0f));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) { in_diff->CopyFromMat(out_diff); }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    const float lr = opts_.learn_rate * learn_rate_coef_;
    if (lr == 0.0f) return;
    ASLP_OK(aslp_col_sum(CuStream(), shift_data_.Data(), diff.Data(), diff.Stride(), diff.NumRows(), diff.NumCols(), -lr, 1.0f, 0.0f));
  }
 private:
  CuVector<BaseFloat> shift_data_;
  float learn_rate_coef_;
};

class Rescale : public UpdatableComponent {
 public:
  Rescale(int32 dim_in, int32 dim_out) : UpdatableComponent(dim_in, dim_out), scale_data_(dim_in), learn_rate_coef_(0.0f) {}
  Component* Copy() const { return new Rescale(*this); }
  ComponentType GetType() const { return kRescale; }
  void InitData(std::istream& is) {
    float init_param = 0.0f;
    ProtoOptions po("(InitParam|LearnRateCoef)"); po.Float("<InitParam>", &init_param); po.Float("<LearnRateCoef>", &learn_rate_coef_); po.Parse(is);
    scale_data_.Set(init_param);
  }
  void ReadData(std::istream& is, bool binary) {
    if ('<' == Peek(is, binary)) { ExpectToken(is, binary, "<LearnRateCoef>"); ReadBasicType(is, binary, &learn_rate_coef_); }
    scale_data_.Read(is, binary);
    KALDI_ASSERT(scale_data_.Dim() == output_dim_);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<LearnRateCoef>"); WriteBasicType(os, binary, learn_rate_coef_);
    scale_data_.Write(os, binary);
  }
  int32 NumParams() const { return scale_data_.Dim(); }
  void GetParams(Vector<BaseFloat>* w) const { scale_data_.CopyToVec(w); }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* p) { p->clear(); p->push_back(std::make_pair(scale_data_.Data(), scale_data_.Dim())); }
  // y = x * diag(scale): the eval-mode BatchNorm kernel with mean 0, inv_std = scale, scale 1, shift 0 would do; a dedicated
  // helper keeps it one pass: out = in, then out *= scale per column via bn_fwd_eval(scale=scale, shift=0, mean=0, inv_std=1)
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) { Apply(in, out); }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) { Apply(out_diff, in_diff); }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    const float lr = opts_.learn_rate * learn_rate_coef_;
    if (lr == 0.0f) return;
    ASLP_OK(aslp_col_dot(CuStream(), scale_data_.Data(), input.Data(), input.Stride(), diff.Data(), diff.Stride(), diff.NumRows(), diff.NumCols(), -lr, 1.0f, 0.0f));
  }
 private:
  void Apply(const CuMatrixBase<BaseFloat>& src, CuMatrixBase<BaseFloat>* dst) {
    if (zeros_.Dim() != scale_data_.Dim()) { zeros_.Resize(scale_data_.Dim(), kSetZero); ones_.Resize(scale_data_.Dim()); ones_.Set(1.0f); }
    ASLP_OK(aslp_bn_fwd_eval(CuStream(), dst->Data(), dst->Stride(), src.Data(), src.Stride(), src.NumRows(), src.NumCols(), scale_data_.Data(),
                             zeros_.Data(), zeros_.Data(), ones_.Data()));
  }
  CuVector<BaseFloat> scale_data_, zeros_, ones_;
  float learn_rate_coef_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
