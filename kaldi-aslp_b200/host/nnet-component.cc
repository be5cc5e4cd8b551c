#include "nnet-component.h"
#include <algorithm>
#include "nnet-activation.h"
#include "nnet-affine-transform.h"
#include "nnet-conv-pool.h"
#include "nnet-gru-streams.h"
#include "nnet-lstm-family.h"
#include "nnet-misc-components.h"
#include "nnet-zoo-components.h"

namespace kaldi {
namespace aslp_nnet {

// marker <-> type table of the components this build implements (reference table: nnet-component.cc:46-81)
const struct Component::key_value Component::kMarkerMap[] = {
    {Component::kSoftmax, "<Softmax>"},
    {Component::kSigmoid, "<Sigmoid>"},
    {Component::kTanh, "<Tanh>"},
    {Component::kReLU, "<ReLU>"},
    {Component::kSplice, "<Splice>"},
    {Component::kAddShift, "<AddShift>"},
    {Component::kRescale, "<Rescale>"},
    {Component::kAffineTransform, "<AffineTransform>"},
    {Component::kLinearTransform, "<LinearTransform>"},
    {Component::kLstmProjectedStreams, "<LstmProjectedStreams>"},
    {Component::kBLstmProjectedStreams, "<BLstmProjectedStreams>"},
    {Component::kBatchNormalization, "<BatchNormalization>"},
    {Component::kInputLayer, "<InputLayer>"},
    {Component::kOutputLayer, "<OutputLayer>"},
    {Component::kScaleLayer, "<ScaleLayer>"},
    {Component::kLstm, "<Lstm>"},
    {Component::kBLstm, "<BLstm>"},
    {Component::kRowConvolution, "<RowConvolution>"},
    {Component::kBLstmProjectedStreamsLC, "<BLstmProjectedStreamsLC>"},
    {Component::kGruStreams, "<GruStreams>"},
    {Component::kCompactFsmn, "<CompactFsmn>"},
    {Component::kConvolutionalComponent, "<ConvolutionalComponent>"},
    {Component::kMaxPoolingComponent, "<MaxPoolingComponent>"},
    {Component::kBlockSoftmax, "<BlockSoftmax>"},
    {Component::kDropout, "<Dropout>"},
    {Component::kLengthNormComponent, "<LengthNormComponent>"},
    {Component::kCopy, "<Copy>"},
    {Component::kLstmCifgProjectedStreams, "<LstmCifgProjectedStreams>"},
    {Component::kPnormComponent, "<Pnorm>"},
    {Component::kPnormComponent, "<Maxout>"},      // the reference's table maps BOTH markers to the p-norm type (nnet-component.cc:79-80); kept
};
static const int kNumMarkers = sizeof(Component::kMarkerMap) / sizeof(Component::kMarkerMap[0]);

const char* Component::TypeToMarker(ComponentType t) {
  for (int i = 0; i < kNumMarkers; i++)
    if (kMarkerMap[i].key == t) return kMarkerMap[i].value;
  KALDI_ERR << "Unknown type" << t;
  return NULL;
}

Component::ComponentType Component::MarkerToType(const std::string& s) {
  auto lower = [](std::string v) { std::transform(v.begin(), v.end(), v.begin(), ::tolower); return v; };
  const std::string want = lower(s);
  for (int i = 0; i < kNumMarkers; i++)
    if (want == lower(kMarkerMap[i].value)) return kMarkerMap[i].key;
  KALDI_ERR << "Unknown marker : '" << s << "' (this build covers the aslp-nnet training path; see DESIGN.md for the component list)";
  return kUnknown;
}

Component* Component::NewComponentOfType(ComponentType t, int32 in, int32 out) {
  switch (t) {
    case kAffineTransform: return new AffineTransform(in, out);
    case kLinearTransform: return new LinearTransform(in, out);
    case kLstmProjectedStreams: return new LstmProjectedStreams(in, out);
    case kBLstmProjectedStreams: return new BLstmProjectedStreams(in, out);
    case kSoftmax: return new Softmax(in, out);
    case kSigmoid: return new Sigmoid(in, out);
    case kTanh: return new Tanh(in, out);
    case kReLU: return new ReLU(in, out);
    case kSplice: return new Splice(in, out);
    case kAddShift: return new AddShift(in, out);
    case kRescale: return new Rescale(in, out);
    case kBatchNormalization: return new BatchNormalization(in, out);
    case kInputLayer: return new InputLayer(in, out);
    case kOutputLayer: return new OutputLayer(in, out);
    case kScaleLayer: return new ScaleLayer(in, out);
    case kLstm: return new Lstm(in, out);
    case kBLstm: return new BLstm(in, out);
    case kRowConvolution: return new RowConvolution(in, out);
    case kBLstmProjectedStreamsLC: return new BLstmProjectedStreamsLC(in, out);
    case kGruStreams: return new GruStreams(in, out);
    case kCompactFsmn: return new CompactFsmn(in, out);
    case kConvolutionalComponent: return new ConvolutionalComponent(in, out);
    case kMaxPoolingComponent: return new MaxPoolingComponent(in, out);
    case kBlockSoftmax: return new BlockSoftmax(in, out);
    case kDropout: return new Dropout(in, out);
    case kLengthNormComponent: return new LengthNormComponent(in, out);
    case kCopy: return new CopyComponent(in, out);
    case kLstmCifgProjectedStreams: return new LstmCifgProjectedStreams(in, out);
    case kPnormComponent: return new PnormComponent(in, out);
    case kMaxoutComponent: return new MaxoutComponent(in, out);
    default: KALDI_ERR << "Missing type: " << static_cast<int>(t);
  }
  return NULL;
}

// one proto line: <Marker> <InputDim> n <OutputDim> m [<Name> x <Input> a:off,b:off] <Key> value ...  (nnet-component.cc:211-285)
Component* Component::Init(const std::string& conf_line) {
  std::istringstream is(conf_line);
  std::string marker;
  int32 input_dim, output_dim;
  ReadToken(is, false, &marker);
  const ComponentType type = MarkerToType(marker);
  ExpectToken(is, false, "<InputDim>");
  ReadBasicType(is, false, &input_dim);
  ExpectToken(is, false, "<OutputDim>");
  ReadBasicType(is, false, &output_dim);
  Component* ans = NewComponentOfType(type, input_dim, output_dim);
  if (conf_line.find("<Name>") != std::string::npos) {          // graph nets: named component with named inputs
    std::string name, input_string;
    ExpectToken(is, false, "<Name>");
    ReadToken(is, false, &name);
    ExpectToken(is, false, "<Input>");
    ReadToken(is, false, &input_string);
    std::vector<std::string> parts, input_name;
    SplitStringToVector(input_string, ",", true, &parts);
    std::vector<int32> offset(parts.size(), 0);
    for (size_t i = 0; i < parts.size(); i++) {
      std::vector<std::string> field;
      SplitStringToVector(parts[i], ":", true, &field);
      KALDI_ASSERT(field.size() >= 1 && field.size() <= 2);
      if (field.size() == 2) ConvertStringToInteger(field[1], &offset[i]);
      input_name.push_back(field[0]);
    }
    ans->SetInputName(input_name);
    ans->SetName(name);
    ans->SetOffset(offset);
  }
  is >> std::ws;
  ans->InitData(is);
  return ans;
}

Component* Component::Read(std::istream& is, bool binary) {
  if (Peek(is, binary) == EOF) return NULL;
  std::string token;
  ReadToken(is, binary, &token);
  if (token == "<Nnet>") ReadToken(is, binary, &token);     // optional opening tag
  if (token == "</Nnet>") return NULL;
  int32 dim_out, dim_in, id;
  ReadBasicType(is, binary, &dim_out);
  ReadBasicType(is, binary, &dim_in);
  std::string name;
  if (Peek(is, binary) == '<') { ExpectToken(is, binary, "<Name>"); ReadToken(is, binary, &name); }
  std::vector<int32> input, offset;
  ReadBasicType(is, binary, &id);
  ReadIntegerVector(is, binary, &input);
  ReadIntegerVector(is, binary, &offset);
  KALDI_ASSERT(input.size() == offset.size());
  Component* ans = NewComponentOfType(MarkerToType(token), dim_in, dim_out);
  ans->ReadData(is, binary);
  ans->SetName(name);
  ans->SetId(id);
  ans->SetInput(input);
  ans->SetOffset(offset);
  return ans;
}

void Component::Write(std::ostream& os, bool binary) const {
  WriteToken(os, binary, TypeToMarker(GetType()));
  WriteBasicType(os, binary, OutputDim());
  WriteBasicType(os, binary, InputDim());
  if (!name_.empty()) { WriteToken(os, binary, "<Name>"); WriteToken(os, binary, name_); }
  WriteBasicType(os, binary, id_);
  WriteIntegerVector(os, binary, input_);
  WriteIntegerVector(os, binary, offset_);
  if (!binary) os << "\n";
  this->WriteData(os, binary);
}

void Component::WriteStandard(std::ostream& os, bool binary) const {
  WriteToken(os, binary, TypeToMarker(GetType()));
  WriteBasicType(os, binary, OutputDim());
  WriteBasicType(os, binary, InputDim());
  if (!binary) os << "\n";
  this->WriteData(os, binary);
}

void Component::Feedforward(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out) {
  if (input_dim_ != in.NumCols())
    KALDI_ERR << "Non-matching dims! " << TypeToMarker(GetType()) << " input-dim : " << input_dim_ << " data : " << in.NumCols();
  out->Resize(in.NumRows(), output_dim_, kUndefined);
  FeedforwardFnc(in, out);
}

void Component::Propagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out) {
  if (input_dim_ != in.NumCols())
    KALDI_ERR << "Non-matching dims! " << TypeToMarker(GetType()) << " input-dim : " << input_dim_ << " data : " << in.NumCols();
  out->Resize(in.NumRows(), output_dim_, kUndefined);
  PropagateFnc(in, out);
}

void Component::Backpropagate(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrix<BaseFloat>* in_diff) {
  if (output_dim_ != out_diff.NumCols())
    KALDI_ERR << "Non-matching output dims, component:" << output_dim_ << " data:" << out_diff.NumCols();
  if (in_diff == NULL) return;     // only nested-nnet components back-propagate without a target, none are on this path
  in_diff->Resize(out_diff.NumRows(), input_dim_, kUndefined);
  KALDI_ASSERT((in.NumRows() == out.NumRows()) && (in.NumRows() == out_diff.NumRows()) && (in.NumRows() == in_diff->NumRows()));
  KALDI_ASSERT(in.NumCols() == in_diff->NumCols());
  KALDI_ASSERT(out.NumCols() == out_diff.NumCols());
  BackpropagateFnc(in, out, out_diff, in_diff);
}

void ProtoOptions::Parse(std::istream& is) {
  std::string token;
  is >> std::ws;
  while (!is.eof()) {
    ReadToken(is, false, &token);
    bool found = false;
    for (auto& kv : f_) if (kv.first == token) { ReadBasicType(is, false, kv.second); found = true; break; }
    if (!found) for (auto& kv : i_) if (kv.first == token) { ReadBasicType(is, false, kv.second); found = true; break; }
    if (!found) KALDI_ERR << "Unknown token " << token << ", a typo in config? " << accepted_;
    is >> std::ws;
  }
}

void InitMatParam(CuMatrix<BaseFloat>* m, float scale) {
  Matrix<BaseFloat> h(m->NumRows(), m->NumCols());
  RandomState rs;                                        // MatrixBase::SetRandUniform (kaldi-matrix.cc:1190-1198)
  for (int32 r = 0; r < h.NumRows(); ++r)
    for (int32 c = 0; c < h.NumCols(); ++c) {
      float u = RandUniform(&rs);                        // uniform in [0, 1]
      u += -0.5f;                                        // Add(-0.5)
      u *= 2 * scale;                                    // Scale(2 * scale)
      h(r, c) = u;
    }
  *m = h;
}
void InitVecParam(CuVector<BaseFloat>* v, float scale) {
  Vector<BaseFloat> tmp(v->Dim());
  for (int32 i = 0; i < tmp.Dim(); i++) tmp(i) = (RandUniform() - 0.5) * 2 * scale;
  *v = tmp;
}
void CopyRowsToVec(const CuMatrixBase<BaseFloat>& m, float* dst) {
  Matrix<float> h;
  m.CopyToMat(&h);
  std::copy(h.Data(), h.Data() + static_cast<size_t>(h.NumRows()) * h.NumCols(), dst);
}

}  // namespace aslp_nnet
}  // namespace kaldi
