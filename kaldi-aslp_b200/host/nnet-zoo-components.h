// nnet-zoo-components.h -- the smaller members of the component zoo, each one or two fused kernels of csrc/zoo.cu:
//   Dropout              src/aslp-nnet/nnet-activation.h:203-273
//   BlockSoftmax         src/aslp-nnet/nnet-activation.h:64-146
//   PnormComponent       src/aslp-nnet/nnet-activation.h:305-356
//   MaxoutComponent      src/aslp-nnet/nnet-activation.h:358-377   (no marker maps to it in the reference either: "<Maxout>" -> kPnormComponent)
//   LengthNormComponent  src/aslp-nnet/nnet-various.h:327-365
//   CopyComponent        src/aslp-nnet/nnet-various.h:186-316
#ifndef ASLP_HOST_NNET_ZOO_COMPONENTS_H_
#define ASLP_HOST_NNET_ZOO_COMPONENTS_H_
#include <algorithm>
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

// The mask comes from a counter-based device generator (Philox-4x32-10: no state to allocate or re-seed, one launch draws the mask
// and applies it).  ASLP_DROPOUT_HOST_RAND=1 draws it on the host from the C library's rand() in the reference CPU path's element
// order (cu-rand.cc:172-176: RandUniform() < retention, row-major) -- the mode the parity tests use, since a random mask can only be
// compared bit for bit when both sides consume the same random stream.
class Dropout : public Component {
 public:
  Dropout(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out), dropout_retention_(0.5f), calls_(0) {}
  Component* Copy() const { return new Dropout(*this); }
  ComponentType GetType() const { return kDropout; }
  void InitData(std::istream& is) {
    ProtoOptions po("(DropoutRetention)");
    po.Float("<DropoutRetention>", &dropout_retention_);
    po.Parse(is);
    KALDI_ASSERT(dropout_retention_ > 0.0 && dropout_retention_ <= 1.0);
  }
  void ReadData(std::istream& is, bool binary) {
    if ('<' == Peek(is, binary)) {
      ExpectToken(is, binary, "<DropoutRetention>");
      ReadBasicType(is, binary, &dropout_retention_);
    }
    KALDI_ASSERT(dropout_retention_ > 0.0 && dropout_retention_ <= 1.0);
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<DropoutRetention>");
    WriteBasicType(os, binary, dropout_retention_);
  }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    dropout_mask_.Resize(in.NumRows(), in.NumCols(), kUndefined);
    const char* hr = std::getenv("ASLP_DROPOUT_HOST_RAND");
    const bool host_rand = hr != nullptr && hr[0] == '1';
    if (host_rand) {
      Matrix<BaseFloat> m(in.NumRows(), in.NumCols(), kUndefined);
      for (int32 r = 0; r < m.NumRows(); r++)
        for (int32 c = 0; c < m.NumCols(); c++) m(r, c) = RandUniform() < dropout_retention_ ? 1.0f : 0.0f;
      dropout_mask_.CopyFromMat(m);
      ASLP_OK(aslp_mul_elements(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), dropout_mask_.Data(), dropout_mask_.Stride(),
                                in.NumRows(), in.NumCols(), 1.0f / dropout_retention_));
    } else {
      static const unsigned long long seed = static_cast<unsigned long long>(Rand()) * 2654435761ull + 0x5DEECE66Dull;    // follows srand()
      ASLP_OK(aslp_dropout_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), dropout_mask_.Data(), dropout_mask_.Stride(),
                               in.NumRows(), in.NumCols(), dropout_retention_, seed + static_cast<unsigned long long>(Id() + 1) * 0x9E3779B97F4A7C15ull, calls_++));
    }
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_mul_elements(CuStream(), in_diff->Data(), in_diff->Stride(), out_diff.Data(), out_diff.Stride(), dropout_mask_.Data(), dropout_mask_.Stride(),
                              out_diff.NumRows(), out_diff.NumCols(), 1.0f / dropout_retention_));
  }
  BaseFloat GetDropoutRetention() { return dropout_retention_; }
  void SetDropoutRetention(BaseFloat dr) {
    dropout_retention_ = dr;
    KALDI_ASSERT(dropout_retention_ > 0.0 && dropout_retention_ <= 1.0);
  }
 private:
  CuMatrix<BaseFloat> dropout_mask_;
  BaseFloat dropout_retention_;
  unsigned long long calls_;
};

class BlockSoftmax : public Component {
 public:
  BlockSoftmax(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new BlockSoftmax(*this); }
  ComponentType GetType() const { return kBlockSoftmax; }
  void InitData(std::istream& is) {
    std::string token, dims_str;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<BlockDims>") is >> dims_str;
      else KALDI_ERR << "Unknown token " << token << ", a typo in config?" << " (BlockDims)";
      is >> std::ws;
    }
    if (!SplitStringToIntegers(dims_str, ",:", false, &block_dims)) KALDI_ERR << "Invalid block-dims " << dims_str;
    SetOffsets();
  }
  void ReadData(std::istream& is, bool binary) { ReadIntegerVector(is, binary, &block_dims); SetOffsets(); }
  void WriteData(std::ostream& os, bool binary) const { WriteIntegerVector(os, binary, block_dims); }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    for (size_t bl = 0; bl < block_dims.size(); bl++) {
      CuSubMatrix<BaseFloat> in_bl = in.ColRange(block_offset[bl], block_dims[bl]), out_bl = out->ColRange(block_offset[bl], block_dims[bl]);
      ASLP_OK(aslp_softmax_rows(CuStream(), out_bl.Data(), out_bl.Stride(), in_bl.Data(), in_bl.Stride(), in_bl.NumRows(), in_bl.NumCols()));
    }
  }
  // the loss delivers y - t; a block whose row does not carry the target sums to one and is zeroed: diff * (1 - row sum)
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    for (size_t bl = 0; bl < block_dims.size(); bl++) {
      CuSubMatrix<BaseFloat> src = out_diff.ColRange(block_offset[bl], block_dims[bl]), dst = in_diff->ColRange(block_offset[bl], block_dims[bl]);
      ASLP_OK(aslp_rows_one_minus_sum(CuStream(), dst.Data(), dst.Stride(), src.Data(), src.Stride(), src.NumRows(), src.NumCols()));
    }
  }
  std::string Info() const {
    std::ostringstream os;
    os << "\n  softmax-dims [ ";
    for (int32 d : block_dims) os << d << " ";
    os << "]";
    return os.str();
  }
  std::vector<int32> block_dims, block_offset;
 private:
  void SetOffsets() {
    block_offset.assign(block_dims.size() + 1, 0);
    for (size_t i = 0; i < block_dims.size(); i++) block_offset[i + 1] = block_offset[i] + block_dims[i];
    KALDI_ASSERT(OutputDim() == block_offset.back());
  }
};

class PnormComponent : public Component {
 public:
  PnormComponent(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out), p(2.0f) {}
  Component* Copy() const { return new PnormComponent(*this); }
  ComponentType GetType() const { return kPnormComponent; }
  void InitData(std::istream& is) {
    ProtoOptions po("(P)");
    po.Float("<P>", &p);
    po.Parse(is);
    KALDI_ASSERT(p != 0);
  }
  void ReadData(std::istream& is, bool binary) { ExpectToken(is, binary, "<P>"); ReadBasicType(is, binary, &p); }
  void WriteData(std::ostream& os, bool binary) const { WriteToken(os, binary, "<P>"); WriteBasicType(os, binary, p); }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    KALDI_ASSERT(in.NumCols() % out->NumCols() == 0);
    ASLP_OK(aslp_group_pnorm_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), out->NumCols(), in.NumCols() / out->NumCols(), p));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_group_pnorm_bwd(CuStream(), in_diff->Data(), in_diff->Stride(), in.Data(), in.Stride(), out.Data(), out.Stride(), out_diff.Data(),
                                 out_diff.Stride(), in.NumRows(), out.NumCols(), in.NumCols() / out.NumCols(), p));
  }
 private:
  BaseFloat p;
};

class MaxoutComponent : public Component {
 public:
  MaxoutComponent(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) { KALDI_ASSERT(dim_in % dim_out == 0); }
  Component* Copy() const { return new MaxoutComponent(*this); }
  ComponentType GetType() const { return kMaxoutComponent; }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_group_max_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), in.NumRows(), out->NumCols(), in.NumCols() / out->NumCols()));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_group_max_bwd(CuStream(), in_diff->Data(), in_diff->Stride(), in.Data(), in.Stride(), out.Data(), out.Stride(), out_diff.Data(),
                               out_diff.Stride(), in.NumRows(), out.NumCols(), in.NumCols() / out.NumCols()));
  }
};

class LengthNormComponent : public Component {
 public:
  LengthNormComponent(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new LengthNormComponent(*this); }
  ComponentType GetType() const { return kLengthNormComponent; }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    if (row_scales_.Dim() != in.NumRows()) row_scales_.Resize(in.NumRows(), kUndefined);
    ASLP_OK(aslp_length_norm_fwd(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), row_scales_.Data(), in.NumRows(), in.NumCols()));
  }
  // the reference treats the scale as a constant: diff_by_x(s * x) = s (nnet-various.h:352-356)
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    ASLP_OK(aslp_mul_rows_vec(CuStream(), in_diff->Data(), in_diff->Stride(), out_diff.Data(), out_diff.Stride(), row_scales_.Data(), out_diff.NumRows(), out_diff.NumCols()));
  }
 private:
  CuVector<BaseFloat> row_scales_;
};

class CopyComponent : public Component {
 public:
  CopyComponent(int32 dim_in, int32 dim_out) : Component(dim_in, dim_out) {}
  Component* Copy() const { return new CopyComponent(*this); }
  ComponentType GetType() const { return kCopy; }
  void InitData(std::istream& is) {
    std::vector<int32> idx;
    std::string token;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<ReadVector>") {
        ReadIntegerVector(is, false, &idx);
      } else if (token == "<BuildVector>") {          // <BuildVector> 1:1:1000 1 2 3 1:10 </BuildVector>  [matlab indexing]
        while (!is.eof()) {
          std::string item;
          ReadToken(is, false, &item);
          if (item == "</BuildVector>") break;
          std::vector<int32> v;
          if (!SplitStringToIntegers(item, ":", false, &v) || v.empty() || v.size() > 3) KALDI_ERR << "Error parsing <BuildVector>";
          if (v.size() == 1) { idx.push_back(v[0]); continue; }
          const int32 lo = v[0], hi = v.back(), step = v.size() == 3 ? v[1] : 1;
          KALDI_ASSERT((lo <= hi && step > 0) || (lo >= hi && step < 0));
          for (int32 j = lo; j <= hi; j += step) idx.push_back(j);       // the reference's loop condition, also for negative steps
        }
      } else {
        KALDI_ERR << "Unknown token " << token << ", a typo in config?" << " (ReadVector|BuildVector)";
      }
      is >> std::ws;
    }
    for (int32& i : idx) --i;
    Set(idx);
  }
  void ReadData(std::istream& is, bool binary) {
    std::vector<int32> idx;
    ReadIntegerVector(is, binary, &idx);
    for (int32& i : idx) --i;
    Set(idx);
  }
  void WriteData(std::ostream& os, bool binary) const {
    std::vector<int32> idx(copy_from_indices_.Host());
    for (int32& i : idx) ++i;
    WriteIntegerVector(os, binary, idx);
  }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    ASLP_OK(aslp_copy_cols(CuStream(), out->Data(), out->Stride(), in.Data(), in.Stride(), copy_from_indices_.Data(), in.NumRows(), out->NumCols()));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    static bool warning_displayed = false;
    if (!warning_displayed) { KALDI_WARN << __func__ << "Not implemented!"; warning_displayed = true; }
    in_diff->SetZero();
  }
 private:
  void Set(const std::vector<int32>& idx) {
    for (int32 i : idx) KALDI_ASSERT(i >= 0 && i < InputDim());
    copy_from_indices_ = idx;
    KALDI_ASSERT(copy_from_indices_.Dim() == OutputDim());
  }
  CuArrayInt copy_from_indices_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
