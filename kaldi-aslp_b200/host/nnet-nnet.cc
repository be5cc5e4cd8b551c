#include "nnet-nnet.h"
#include <map>
#include "nnet-activation.h"
#include "nnet-affine-transform.h"
#include "nnet-gru-streams.h"
#include "nnet-lstm-family.h"
#include "nnet-misc-components.h"
#include "nnet-zoo-components.h"

namespace kaldi {
namespace aslp_nnet {

Nnet::Nnet(const Nnet& other) { *this = other; }
Nnet& Nnet::operator=(const Nnet& other) {
  if (this == &other) return *this;
  Destroy();
  for (int32 i = 0; i < other.NumComponents(); i++) components_.push_back(other.GetComponent(i).Copy());
  SetTrainOptions(other.opts_);
  InitInputOutput();
  Check();
  return *this;
}
Nnet::~Nnet() { Destroy(); }

// A component whose only input is the full output of a component with no other consumer can read / write that
// buffer in place; everything else goes through the reference's zero + add assembly (nnet-nnet.cc:77-95,122-145).
static bool DirectInput(const std::vector<Component*>& comps, int32 i) {
  const Component* c = comps[i];
  if (c->GetType() == Component::kInputLayer) return false;
  const std::vector<int32>& in = c->GetInput();
  if (in.size() != 1 || c->GetOffset()[0] != 0) return false;
  if (comps[in[0]]->OutputDim() != c->InputDim()) return false;
  int32 consumers = 0;
  for (size_t k = 0; k < comps.size(); ++k) {
    if (comps[k]->GetType() == Component::kInputLayer) continue;
    for (int32 src : comps[k]->GetInput()) if (src == in[0]) ++consumers;
  }
  return consumers == 1;
}

// Epilogue fusion across component pairs.  Forward: Affine / Linear i followed by a Sigmoid / Tanh / ReLU that is
// its only consumer -> one product with the activation in its epilogue; the Affine's own output buffer is not written.
// Backward: the same kind of activation a in front of an Affine b -> b's backward product applies f'(y_a) in its epilogue
// and writes d(loss)/d(input of a); a's own out-diff buffer is not written.  Only for the few-tile shapes whose products
// take the split-K reduce pass (AffineTransform::SmallBatchShape), which is where a step is launch-bound and where the
// golden-net comparisons of every per-component buffer do not reach.  ASLP_FUSE_EPILOGUE=0 switches it off.
static bool FusionEnabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ASLP_FUSE_EPILOGUE"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
static int ActKind(const Component* c) {
  switch (c->GetType()) {
    case Component::kSigmoid: return ASLP_ACT_SIGMOID;
    case Component::kTanh: return ASLP_ACT_TANH;
    case Component::kReLU: return ASLP_ACT_RELU;
    default: return -1;
  }
}
static bool IsAffine(const Component* c) { return c->GetType() == Component::kAffineTransform || c->GetType() == Component::kLinearTransform; }
// activation `a` reads, in place, the whole output of an Affine that nobody else reads
static bool ActBehindAffine(const std::vector<Component*>& comps, const std::vector<int32>& outputs, int32 a) {
  if (a < 1 || a >= static_cast<int32>(comps.size()) || ActKind(comps[a]) < 0 || !DirectInput(comps, a)) return false;
  const int32 p = comps[a]->GetInput()[0];
  if (!IsAffine(comps[p])) return false;
  for (int32 o : outputs) if (o == p) return false;      // the pre-activation is a net output: it has to exist
  return true;
}

void Nnet::Propagate(const std::vector<const CuMatrixBase<BaseFloat>*>& in, std::vector<CuMatrix<BaseFloat>*>* out) {
  KALDI_ASSERT(NULL != out);
  KALDI_ASSERT(in.size() == input_.size());
  const int32 num_frame = in[0]->NumRows();
  for (size_t i = 0; i < input_.size(); i++) {
    CuMatrix<BaseFloat>& b = input_buf_[input_[i]];
    b.Resize(num_frame, components_[input_[i]]->InputDim(), kUndefined);
    b.CopyFromMat(*(in[i]));
  }
  for (int32 i = 0; i < NumComponents(); i++) {
    Component* c = components_[i];
    const CuMatrixBase<BaseFloat>* src = &input_buf_[i];
    if (c->GetType() != Component::kInputLayer) {
      if (DirectInput(components_, i)) {
        src = &output_buf_[c->GetInput()[0]];
      } else {
        const std::vector<int32>& input_idx = c->GetInput();
        const std::vector<int32>& offset = c->GetOffset();
        KALDI_ASSERT(input_idx.size() == offset.size());
        input_buf_[i].Resize(num_frame, c->InputDim(), kSetZero);
        for (size_t j = 0; j < input_idx.size(); j++) {
          const int32 out_len = components_[input_idx[j]]->OutputDim();
          CuSubMatrix<BaseFloat> dst = input_buf_[i].ColRange(offset[j], out_len);
          dst.AddMat(1.0, output_buf_[input_idx[j]]);
        }
      }
    }
    Timer tim;
    // Affine + activation in one product (the activation must be the next component in execution order)
    if (FusionEnabled() && IsAffine(c) && i + 1 < NumComponents() && ActBehindAffine(components_, output_, i + 1) &&
        components_[i + 1]->GetInput()[0] == i && static_cast<AffineTransform*>(c)->SmallBatchShape(num_frame, false)) {
      output_buf_[i].Resize(num_frame, c->OutputDim(), kUndefined);          // keeps its shape (Backpropagate checks it); never written
      output_buf_[i + 1].Resize(num_frame, components_[i + 1]->OutputDim(), kUndefined);
      static_cast<AffineTransform*>(c)->PropagateFused(*src, ActKind(components_[i + 1]), &output_buf_[i + 1]);
      propagate_time_[i].first = Component::TypeToMarker(c->GetType());
      propagate_time_[i].second += tim.Elapsed();
      propagate_time_[i + 1].first = Component::TypeToMarker(components_[i + 1]->GetType());
      ++i;
      continue;
    }
    c->Propagate(*src, &output_buf_[i]);
    propagate_time_[i].first = Component::TypeToMarker(c->GetType());
    propagate_time_[i].second += tim.Elapsed();
  }
  for (size_t i = 0; i < output_.size(); i++) *((*out)[i]) = output_buf_[output_[i]];
}

void Nnet::Backpropagate(const std::vector<const CuMatrixBase<BaseFloat>*>& out_diff, std::vector<CuMatrix<BaseFloat>*>* in_diff) {
  KALDI_ASSERT(out_diff.size() == output_.size());
  const int32 num_frame = out_diff[0]->NumRows();
  std::vector<char> direct(NumComponents(), 0), fed_direct(NumComponents(), 0), skip_bwd(NumComponents(), 0);
  for (int32 i = 0; i < NumComponents(); i++) {
    direct[i] = DirectInput(components_, i);
    if (direct[i]) fed_direct[components_[i]->GetInput()[0]] = 1;
  }
  for (int32 i = 0; i < NumComponents(); i++)     // buffers written in place by their single consumer need no zeroing
    output_diff_buf_[i].Resize(num_frame, components_[i]->OutputDim(), fed_direct[i] ? kUndefined : kSetZero);
  for (size_t i = 0; i < output_.size(); i++) output_diff_buf_[output_[i]].CopyFromMat(*(out_diff[i]));
  for (int32 i = NumComponents() - 1; i >= 0; i--) {
    Component* c = components_[i];
    const CuMatrixBase<BaseFloat>& cin = direct[i] ? static_cast<const CuMatrixBase<BaseFloat>&>(output_buf_[c->GetInput()[0]]) : input_buf_[i];
    CuMatrix<BaseFloat>* target = direct[i] ? &output_diff_buf_[c->GetInput()[0]] : &input_diff_buf_[i];
    Timer tim;
    if (skip_bwd[i]) continue;                            // an activation whose derivative the Affine behind it has applied
    // Affine whose input is an activation that reads an Affine in place: d(loss)/d(pre-activation) straight from this product
    const int32 a = direct[i] ? c->GetInput()[0] : -1;
    bool a_is_output = false;
    for (int32 o : output_) if (o == a) a_is_output = true;
    if (FusionEnabled() && IsAffine(c) && a >= 0 && a == i - 1 && direct[a] && !a_is_output && ActKind(components_[a]) >= 0 &&
        static_cast<AffineTransform*>(c)->SmallBatchShape(num_frame, true)) {
      CuMatrix<BaseFloat>* pre = &output_diff_buf_[components_[a]->GetInput()[0]];
      pre->Resize(num_frame, components_[a]->InputDim(), kUndefined);
      static_cast<AffineTransform*>(c)->BackpropagateFused(output_diff_buf_[i], output_buf_[a], ActKind(components_[a]), pre);
      skip_bwd[a] = 1;
    } else {
      c->Backpropagate(cin, output_buf_[i], output_diff_buf_[i], target);
    }
    // update inside backprop (:126-129).  (Running an Affine's weight half on the side stream under the next layer's backward
    // product was measured and dropped: cfg1 0.329 ms with it, 0.321 without -- a tcgen05 GEMM CTA takes a whole SM's shared
    // memory, so two products never share the chip; profiles/r02_config_bench.jsonl)
    if (c->IsUpdatable()) {
      dynamic_cast<UpdatableComponent*>(c)->Update(cin, output_diff_buf_[i]);
      if (update_observer_) update_observer_(i);
    }
    back_propagate_time_[i].first = Component::TypeToMarker(c->GetType());
    back_propagate_time_[i].second += tim.Elapsed();
    if (c->GetType() != Component::kInputLayer && !direct[i]) {
      const std::vector<int32>& input_idx = c->GetInput();
      const std::vector<int32>& offset = c->GetOffset();
      for (size_t j = 0; j < input_idx.size(); j++) {
        KALDI_ASSERT(input_idx[j] >= 0 && input_idx[j] <= NumComponents());
        const int32 out_len = components_[input_idx[j]]->OutputDim();
        output_diff_buf_[input_idx[j]].AddMat(1.0, input_diff_buf_[i].ColRange(offset[j], out_len));
      }
    }
  }
  CuJoin();      // weight gradients / updates that ran on the side stream are complete before anyone reads the weights
  if (NULL == in_diff) return;
  for (size_t i = 0; i < input_.size(); i++)
    if ((*in_diff)[i] != NULL) *((*in_diff)[i]) = input_diff_buf_[input_[i]];
}

bool Nnet::StepReplayable() const {
  if (input_.size() != 1 || output_.size() != 1 || update_observer_) return false;
  for (const Component* c : components_) {
    switch (c->GetType()) {
      case Component::kAffineTransform: case Component::kLinearTransform: case Component::kSigmoid: case Component::kTanh:
      case Component::kReLU: case Component::kSoftmax: case Component::kCompactFsmn: case Component::kSplice:
      case Component::kAddShift: case Component::kRescale: case Component::kInputLayer: case Component::kOutputLayer:
        break;
      default:
        return false;
    }
  }
  return true;
}

void Nnet::Feedforward(const std::vector<const CuMatrixBase<BaseFloat>*>& in, std::vector<CuMatrix<BaseFloat>*>* out) {
  KALDI_ASSERT(NULL != out);
  KALDI_ASSERT(in.size() == input_.size());
  const int32 num_frame = in[0]->NumRows();
  for (size_t i = 0; i < input_.size(); i++) {
    CuMatrix<BaseFloat>& b = input_buf_[input_[i]];
    b.Resize(num_frame, components_[input_[i]]->InputDim(), kUndefined);
    b.CopyFromMat(*(in[i]));
  }
  for (int32 i = 0; i < NumComponents(); i++) {
    Component* c = components_[i];
    const CuMatrixBase<BaseFloat>* src = &input_buf_[i];
    if (c->GetType() != Component::kInputLayer) {
      if (DirectInput(components_, i)) {
        src = &output_buf_[c->GetInput()[0]];
      } else {
        input_buf_[i].Resize(num_frame, c->InputDim(), kSetZero);
        for (size_t j = 0; j < c->GetInput().size(); j++) {
          const int32 s = c->GetInput()[j];
          CuSubMatrix<BaseFloat> dst = input_buf_[i].ColRange(c->GetOffset()[j], components_[s]->OutputDim());
          dst.AddMat(1.0, output_buf_[s]);
        }
      }
    }
    c->Feedforward(*src, &output_buf_[i]);
  }
  for (size_t i = 0; i < output_.size(); i++) *((*out)[i]) = output_buf_[output_[i]];
}

void Nnet::Propagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out) {
  KALDI_ASSERT(NULL != out);
  if (NumComponents() == 0) { (*out) = in; return; }
  KALDI_ASSERT(input_.size() == 1 && output_.size() == 1);
  std::vector<const CuMatrixBase<BaseFloat>*> in_vec(1, &in);
  std::vector<CuMatrix<BaseFloat>*> out_vec(1, out);
  Propagate(in_vec, &out_vec);
}
void Nnet::Backpropagate(const CuMatrixBase<BaseFloat>& out_diff, CuMatrix<BaseFloat>* in_diff) {
  if (NumComponents() == 0) { if (in_diff) (*in_diff) = out_diff; return; }
  KALDI_ASSERT(input_.size() == 1 && output_.size() == 1);
  std::vector<const CuMatrixBase<BaseFloat>*> od(1, &out_diff);
  std::vector<CuMatrix<BaseFloat>*> id(1, in_diff);
  Backpropagate(od, &id);
}
void Nnet::Feedforward(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out) {
  KALDI_ASSERT(NULL != out);
  if (NumComponents() == 0) { (*out) = in; return; }
  KALDI_ASSERT(input_.size() == 1 && output_.size() == 1);
  std::vector<const CuMatrixBase<BaseFloat>*> in_vec(1, &in);
  std::vector<CuMatrix<BaseFloat>*> out_vec(1, out);
  Feedforward(in_vec, &out_vec);
}

int32 Nnet::OutputDim() const { KALDI_ASSERT(!components_.empty()); return components_.back()->OutputDim(); }
int32 Nnet::InputDim() const { KALDI_ASSERT(!components_.empty()); return components_.front()->InputDim(); }
const Component& Nnet::GetComponent(int32 c) const { KALDI_ASSERT(static_cast<size_t>(c) < components_.size()); return *(components_[c]); }
Component& Nnet::GetComponent(int32 c) { KALDI_ASSERT(static_cast<size_t>(c) < components_.size()); return *(components_[c]); }

void Nnet::SetComponent(int32 c, Component* component) {
  KALDI_ASSERT(static_cast<size_t>(c) < components_.size());
  delete components_[c];
  components_[c] = component;
  InitInputOutput();
  Check();
}
void Nnet::AppendComponent(Component* comp) {
  components_.push_back(comp);
  for (int32 i = 0; i < NumComponents(); i++) { components_[i]->SetId(i); components_[i]->SetMonoInput(i - 1); }
  InitInputOutput();
}
void Nnet::AppendNnet(const Nnet& other) {
  for (int32 i = 0; i < other.NumComponents(); i++) AppendComponent(other.GetComponent(i).Copy());
  InitInputOutput();
  Check();
}
void Nnet::RemoveComponent(int32 c) {
  KALDI_ASSERT(c < NumComponents());
  Component* ptr = components_[c];
  components_.erase(components_.begin() + c);
  delete ptr;
  InitInputOutput();
  Check();
}

int32 Nnet::NumParams() const {
  int32 n = 0;
  for (Component* c : components_) if (c->IsUpdatable()) n += dynamic_cast<UpdatableComponent*>(c)->NumParams();
  return n;
}
void Nnet::GetParams(Vector<BaseFloat>* wei_copy) const {
  wei_copy->Resize(NumParams());
  int32 pos = 0;
  for (Component* c : components_) {
    if (!c->IsUpdatable()) continue;
    Vector<BaseFloat> p;
    dynamic_cast<UpdatableComponent*>(c)->GetParams(&p);
    for (int32 i = 0; i < p.Dim(); ++i) (*wei_copy)(pos + i) = p(i);
    pos += p.Dim();
  }
  KALDI_ASSERT(pos == NumParams());
}
void Nnet::GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
  KALDI_ASSERT(params != NULL);
  params->clear();
  for (Component* c : components_) {
    if (!c->IsUpdatable()) continue;
    std::vector<std::pair<BaseFloat*, int>> cp;
    dynamic_cast<UpdatableComponent*>(c)->GetGpuParams(&cp);
    params->insert(params->end(), cp.begin(), cp.end());
  }
}
void Nnet::GetAccStats(std::vector<double*>* acc_params, std::vector<std::pair<double*, int>>* data_params) {
  KALDI_ASSERT(acc_params != NULL && data_params != NULL);
  acc_params->clear();
  data_params->clear();
  for (Component* c : components_) {
    if (c->GetType() != Component::kBatchNormalization) continue;
    std::vector<std::pair<double*, int>> cp;
    acc_params->push_back(dynamic_cast<BatchNormalization*>(c)->GetAccStats(&cp));
    data_params->insert(data_params->end(), cp.begin(), cp.end());
  }
}

void Nnet::ResetLstmStreams(const std::vector<int32>& flags) {
  for (Component* c : components_) {
    if (LstmFamily* l = dynamic_cast<LstmFamily*>(c)) l->ResetLstmStreams(flags);      // no-op for the non-carrying types
    else if (LstmCifgProjectedStreams* q = dynamic_cast<LstmCifgProjectedStreams*>(c)) q->ResetLstmStreams(flags);
    else if (GruStreams* g = dynamic_cast<GruStreams*>(c)) g->ResetLstmStreams(flags);
  }
}
// The reference forwards SetSeqLengths to BLstmProjectedStreams, BLstm, LstmProjectedStreams, Lstm, RowConvolution,
// GruStreams (nnet-nnet.cc:492-523) but NOT to BLstmProjectedStreamsLC, although that class defines it "for whole
// sentence train" (lc.h:497-501); an LC net under a whole-utterance trainer would then run as ONE stream of T*S rows.
// This build forwards it to the LC type too (deviations); ASLP_STRICT_REFERENCE_QUIRKS=1 restores the omission.
void Nnet::SetSeqLengths(const std::vector<int32>& lens) {
  static const bool strict = getenv("ASLP_STRICT_REFERENCE_QUIRKS") != nullptr;
  for (Component* c : components_) {
    if (LstmFamily* l = dynamic_cast<LstmFamily*>(c)) {
      if (strict && c->GetType() == Component::kBLstmProjectedStreamsLC) continue;
      l->SetSeqLengths(lens);
    } else if (LstmCifgProjectedStreams* q = dynamic_cast<LstmCifgProjectedStreams*>(c)) q->SetSeqLengths(lens);
    else if (RowConvolution* r = dynamic_cast<RowConvolution*>(c)) r->SetSeqLengths(lens);
    else if (GruStreams* g = dynamic_cast<GruStreams*>(c)) g->SetSeqLengths(lens);
  }
}
// nnet-nnet.cc:454-464
void Nnet::SetDropoutRetention(BaseFloat r) {
  for (int32 c = 0; c < NumComponents(); c++) {
    if (GetComponent(c).GetType() == Component::kDropout) {
      Dropout& comp = dynamic_cast<Dropout&>(GetComponent(c));
      const BaseFloat r_old = comp.GetDropoutRetention();
      comp.SetDropoutRetention(r);
      KALDI_LOG << "Setting dropout-retention in component " << c << " from " << r_old << " to " << r;
    }
  }
}
void Nnet::SetChunkSize(int chunk_size) {
  for (Component* c : components_)
    if (c->GetType() == Component::kBLstmProjectedStreamsLC) dynamic_cast<LstmFamily*>(c)->SetChunkSize(chunk_size);
}

// simple nets get an InputLayer in front, ids 0..n and an OutputLayer behind (nnet-nnet.cc:534-559)
void Nnet::AutoComplete() {
  const int32 input_dim = components_[0]->InputDim();
  Component* in_comp = new InputLayer(input_dim, input_dim);
  in_comp->SetId(0);
  in_comp->SetMonoInput(-1);
  components_.insert(components_.begin(), in_comp);
  for (size_t i = 1; i < components_.size(); i++) {
    KALDI_ASSERT(components_[i]->Id() < 0);
    components_[i]->SetId(static_cast<int32>(i));
    components_[i]->SetMonoInput(static_cast<int32>(i) - 1);
  }
  const int32 n = static_cast<int32>(components_.size()), output_dim = components_[n - 1]->OutputDim();
  Component* out_comp = new OutputLayer(output_dim, output_dim);
  out_comp->SetId(n);
  out_comp->SetMonoInput(n - 1);
  components_.push_back(out_comp);
}

void Nnet::Init(const std::string& file) {
  Input in;
  in.OpenTextMode(file);
  std::istream& is = in.Stream();
  std::string conf_line, token;
  bool simple_net = true;
  while (std::getline(is, conf_line)) {
    if (conf_line.find_first_not_of(" \t\r") == std::string::npos) continue;
    KALDI_VLOG(1) << conf_line;
    std::istringstream ls(conf_line);
    ls >> std::ws >> token;
    if (token == "<NnetProto>" || token == "</NnetProto>") continue;
    if (token == "<StructureType>") {
      ls >> std::ws >> token;
      if (token == "graph") simple_net = false;
      else if (token == "simple") simple_net = true;
      else KALDI_ERR << "The net's structure must be simple or graph!";
      continue;
    }
    components_.push_back(Component::Init(conf_line + "\n"));
  }
  if (!simple_net) { AssignComponentId(components_); SortComponent(components_); }
  else AutoComplete();
  in.Close();
  InitInputOutput();
  Check();
}

void Nnet::Read(const std::string& file) {
  bool binary;
  Input in(file, &binary);
  Read(in.Stream(), binary);
  in.Close();
  if (NumComponents() == 0) KALDI_WARN << "The network '" << file << "' is empty.";
}
void Nnet::Read(std::istream& is, bool binary) {
  Component* comp;
  while (NULL != (comp = Component::Read(is, binary))) {
    const int id = comp->Id();
    if (id >= static_cast<int>(components_.size())) components_.resize(id + 1, NULL);
    if (components_[id] != NULL) KALDI_ERR << "Component id " << id << " already be taken, the id must be unique";
    components_[id] = comp;
  }
  opts_.learn_rate = 0.0;      // quirk kept: Read resets the learn rate, trainers set it afterwards (:632)
  InitInputOutput();
  Check();
}
void Nnet::Write(const std::string& file, bool binary) const {
  Output out(file, binary, true);
  Write(out.Stream(), binary);
  out.Close();
}
void Nnet::Write(std::ostream& os, bool binary) const {
  Check();
  WriteToken(os, binary, "<Nnet>");
  if (binary == false) os << std::endl;
  for (int32 i = 0; i < NumComponents(); i++) components_[i]->Write(os, binary);
  WriteToken(os, binary, "</Nnet>");
  if (binary == false) os << std::endl;
}
void Nnet::WriteStandard(const std::string& file, bool binary) const {
  Output out(file, binary, true);
  Write(out.Stream(), binary);          // (the reference's file overload also writes the full format, :684-688)
  out.Close();
}
void Nnet::WriteStandard(std::ostream& os, bool binary) const {
  Check();
  WriteToken(os, binary, "<Nnet>");
  if (binary == false) os << std::endl;
  for (int32 i = 0; i < NumComponents(); i++) {
    if (components_[i]->GetType() == Component::kInputLayer || components_[i]->GetType() == Component::kOutputLayer) continue;
    components_[i]->WriteStandard(os, binary);
  }
  WriteToken(os, binary, "</Nnet>");
  if (binary == false) os << std::endl;
}

std::string Nnet::Info() const {
  std::ostringstream ostr;
  ostr << "num-components " << NumComponents() << std::endl;
  ostr << "input-dim " << InputDim() << std::endl;
  ostr << "output-dim " << OutputDim() << std::endl;
  ostr << "number-of-parameters " << static_cast<float>(NumParams()) / 1e6 << " millions" << std::endl;
  for (int32 i = 0; i < NumComponents(); i++) {
    const Component* c = components_[i];
    ostr << "component " << i + 1 << " : " << Component::TypeToMarker(c->GetType()) << ", input-dim " << c->InputDim()
         << ", output-dim " << c->OutputDim() << ", id " << c->Id() << ", input ";
    for (size_t j = 0; j < c->GetInput().size(); j++) ostr << c->GetInput()[j] << ":" << c->GetOffset()[j] << ",";
    ostr << "  " << c->Info() << std::endl;
  }
  return ostr.str();
}
std::string Nnet::InfoGradient() const {
  std::ostringstream ostr;
  ostr << "### Gradient stats :\n";
  for (int32 i = 0; i < NumComponents(); i++)
    ostr << "Component " << i + 1 << " : " << Component::TypeToMarker(components_[i]->GetType()) << ", " << components_[i]->InfoGradient() << std::endl;
  return ostr.str();
}
std::string Nnet::InfoPropagate() const {
  std::ostringstream ostr;
  ostr << "### Forward propagation buffer content :\n";
  ostr << "[0] output of <Input> " << MomentStatistics(input_buf_[0]) << std::endl;
  for (int32 i = 0; i < NumComponents(); i++)
    ostr << "[" << 1 + i << "] output of " << Component::TypeToMarker(components_[i]->GetType()) << MomentStatistics(output_buf_[i]) << std::endl;
  return ostr.str();
}
std::string Nnet::InfoBackPropagate() const {
  std::ostringstream ostr;
  ostr << "### Backward propagation buffer content :\n";
  ostr << "[0] diff of <Input> " << MomentStatistics(output_diff_buf_[0]) << std::endl;
  for (int32 i = 0; i < NumComponents(); i++)
    ostr << "[" << 1 + i << "] diff-output of " << Component::TypeToMarker(components_[i]->GetType()) << MomentStatistics(output_diff_buf_[i]) << std::endl;
  return ostr.str();
}

void Nnet::Check() const {
  if (input_.size() < 1) KALDI_ERR << "Must have at least one InputLayer";
  if (output_.size() < 1) KALDI_ERR << "Must have at least one OutputLayer";
  for (int i = 0; i < NumComponents(); i++) {
    if (components_[i] == NULL) KALDI_ERR << "Component id must be consistant, but have no id " << i;
    if (components_[i]->Id() != i) KALDI_ERR << "Component id not equal index id, May be error in Read";
  }
  for (int i = 0; i < NumComponents(); i++) {
    if (components_[i]->GetType() == Component::kInputLayer) continue;
    const std::vector<int32>& input_idx = components_[i]->GetInput();
    const std::vector<int32>& offset = components_[i]->GetOffset();
    KALDI_ASSERT(input_idx.size() == offset.size());
    for (size_t j = 0; j < input_idx.size(); j++) {
      const int idx = input_idx[j];
      if (idx < 0 || idx >= NumComponents() || components_[idx]->Id() >= components_[i]->Id())
        KALDI_ERR << "Input id must be less than Component id, case  <Id> " << i << " <Input> " << idx;
      const int32 out_dim = components_[idx]->OutputDim();
      if (offset[j] + out_dim > components_[i]->InputDim())
        KALDI_ERR << "Component " << idx << " outputdim + offset must be less than offset " << offset[j] << " outdim " << out_dim
                  << " Component " << i << " inputdim";
    }
  }
  // NaN / Inf guard on the weights (:810-819)
  Vector<BaseFloat> weights;
  GetParams(&weights);
  const BaseFloat sum = weights.Sum();
  if (std::isinf(sum)) KALDI_ERR << "'inf' in network parameters (weight explosion, try lower learning rate?)";
  if (std::isnan(sum)) KALDI_ERR << "'nan' in network parameters (try lower learning rate?)";
}

void Nnet::Destroy() {
  for (Component* c : components_) delete c;
  components_.resize(0);
  input_buf_.resize(0); input_diff_buf_.resize(0); output_buf_.resize(0); output_diff_buf_.resize(0);
}

void Nnet::SetTrainOptions(const NnetTrainOptions& opts) {
  opts_ = opts;
  for (Component* c : components_) if (c->IsUpdatable()) dynamic_cast<UpdatableComponent*>(c)->SetTrainOptions(opts_);
}

void Nnet::InitInputOutput() {
  input_.clear();
  output_.clear();
  for (Component* c : components_) {
    if (c == NULL) continue;
    if (c->GetType() == Component::kInputLayer) input_.push_back(c->Id());
    else if (c->GetType() == Component::kOutputLayer) output_.push_back(c->Id());
  }
  const size_t n = components_.size();
  input_buf_.resize(n); output_buf_.resize(n); input_diff_buf_.resize(n); output_diff_buf_.resize(n);
  propagate_time_.assign(n, std::make_pair(std::string(), 0.0f));
  back_propagate_time_.assign(n, std::make_pair(std::string(), 0.0f));
}

void Nnet::GetComponentTime() {
  for (size_t i = 0; i < propagate_time_.size(); i++) {
    KALDI_LOG << propagate_time_[i].first << ": Propagate time " << propagate_time_[i].second << "s, Back-Propagate time "
              << back_propagate_time_[i].second << "s, total time " << propagate_time_[i].second + back_propagate_time_[i].second << "s";
    propagate_time_[i].second = 0.0;
    back_propagate_time_[i].second = 0.0;
  }
}

// graph nets: ids by a depth-first (LIFO work list) topological order over the <Name>/<Input> declarations, so that
// the numbering matches what the reference assigns to the same proto (nnet-nnet.cc:886-949)
void Nnet::AssignComponentId(std::vector<Component*>& comp) {
  const int32 n = static_cast<int32>(comp.size());
  std::vector<int32> pending(n);
  for (int32 i = 0; i < n; i++) {
    const std::vector<std::string>& in = comp[i]->GetInputName();
    pending[i] = (in.size() == 1 && in[0] == "-1") ? 0 : static_cast<int32>(in.size());
  }
  std::vector<std::string> work;
  for (int32 i = 0; i < n; i++) if (pending[i] == 0) work.push_back(comp[i]->GetName());
  int32 next_id = 0;
  while (!work.empty()) {
    const std::string name = work.back();
    work.pop_back();
    for (int32 i = 0; i < n; i++) {
      if (comp[i]->GetName() == name) comp[i]->SetId(next_id++);
      for (const std::string& dep : comp[i]->GetInputName()) {
        if (dep == comp[i]->GetName()) KALDI_ERR << "The input of component " << dep << "include itself, Please check it!";
        if (dep == name && --pending[i] == 0) work.push_back(comp[i]->GetName());
      }
    }
  }
  if (next_id != n) KALDI_ERR << "The graph has a cycle";
  std::map<std::string, int32> id_of;
  for (int32 i = 0; i < n; i++) id_of[comp[i]->GetName()] = comp[i]->GetId();
  for (int32 i = 0; i < n; i++) {
    const std::vector<std::string>& names = comp[i]->GetInputName();
    std::vector<int32> input(names.size(), 0);
    for (size_t j = 0; j < names.size(); j++) {
      if (names[j] == "-1") input[j] = -1;
      else { auto it = id_of.find(names[j]); if (it != id_of.end()) input[j] = it->second; }
    }
    comp[i]->SetInput(input);
  }
}
void Nnet::SortComponent(std::vector<Component*>& comp) {
  std::vector<Component*> sorted(comp.size());
  for (Component* c : comp) sorted[c->GetId()] = c;
  comp.swap(sorted);
}

}  // namespace aslp_nnet
}  // namespace kaldi
