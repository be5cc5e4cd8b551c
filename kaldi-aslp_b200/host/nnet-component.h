// nnet-component.h -- the aslp-nnet Component / UpdatableComponent interface, kept source-compatible
// (same method names, argument meaning and error behaviour as src/aslp-nnet/nnet-component.h:45-347) so the
// trainer mains and the parity tests read like the reference's.  The bodies differ: every PropagateFnc /
// BackpropagateFnc / Update is one or a few fused C-ABI calls (include/aslp_b200.h) on device-resident data.
#ifndef ASLP_HOST_NNET_COMPONENT_H_
#define ASLP_HOST_NNET_COMPONENT_H_
#include "matrix.h"
#include "nnet-trnopts.h"

namespace kaldi {
namespace aslp_nnet {

class Component {
 public:
  // numeric values as in the reference (nnet-component.h:50-103)
  typedef enum {
    kUnknown = 0x0,
    kUpdatableComponent = 0x0100, kAffineTransform, kLinearTransform, kConvolutionalComponent, kConvolutional2DComponent,
    kLstmProjectedStreams, kBLstmProjectedStreams,
    kActivationFunction = 0x0200, kSoftmax, kBlockSoftmax, kSigmoid, kTanh, kDropout, kReLU, kLengthNormComponent,
    kTranform = 0x0400, kRbm, kSplice, kCopy, kTranspose, kBlockLinearity, kAddShift, kRescale,
    kKlHmm = 0x0800, kSentenceAveragingComponent, kSimpleSentenceAveragingComponent, kAveragePoolingComponent,
    kAveragePooling2DComponent, kMaxPoolingComponent, kMaxPooling2DComponent, kFramePoolingComponent, kParallelComponent,
    kBatchNormalization = 0x0f00, kInputLayer, kOutputLayer, kScaleLayer, kLstm, kBLstm, kRowConvolution,
    kBLstmProjectedStreamsLC, kGruStreams, kLstmCifgProjectedStreams, kCompactFsmn, kPnormComponent, kMaxoutComponent
  } ComponentType;
  struct key_value { const ComponentType key; const char* value; };
  static const struct key_value kMarkerMap[];
  static const char* TypeToMarker(ComponentType t);
  static ComponentType MarkerToType(const std::string& s);   // case insensitive

  Component(int32 input_dim, int32 output_dim) : input_dim_(input_dim), output_dim_(output_dim), id_(-1) {}
  virtual ~Component() {}
  virtual Component* Copy() const = 0;
  virtual ComponentType GetType() const = 0;
  virtual bool IsUpdatable() const { return false; }

  int32 InputDim() const { return input_dim_; }
  int32 OutputDim() const { return output_dim_; }
  int32 Id() const { return id_; }
  int32 GetId() const { return id_; }
  void SetId(int id) { id_ = id; }
  void SetName(const std::string& name) { name_ = name; }
  const std::string& GetName() const { return name_; }
  const std::vector<int32>& GetInput() const { return input_; }
  void SetInput(const std::vector<int32>& input) { input_ = input; }
  void SetInputName(const std::vector<std::string>& n) { input_name_ = n; }
  const std::vector<std::string>& GetInputName() const { return input_name_; }
  void SetMonoInput(int id) { input_.assign(1, id); offset_.assign(1, 0); }
  const std::vector<int32>& GetOffset() const { return offset_; }
  void SetOffset(const std::vector<int32>& offset) { offset_ = offset; }

  // dim check, size `out` / `in_diff`, call the virtual (nnet-component.h:286-347).  The reference zeroes the
  // target first; here targets are sized without a memset and every *Fnc overwrites all of its output.
  virtual void Feedforward(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out);
  void Propagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out);
  void Backpropagate(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrix<BaseFloat>* in_diff);

  static Component* Init(const std::string& conf_line);
  static Component* Read(std::istream& is, bool binary);
  void Write(std::ostream& os, bool binary) const;
  void WriteStandard(std::ostream& os, bool binary) const;

  virtual std::string Info() const { return ""; }
  virtual std::string InfoGradient() const { return ""; }

 protected:
  virtual void FeedforwardFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) { PropagateFnc(in, out); }
  virtual void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) = 0;
  virtual void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) = 0;
  virtual void InitData(std::istream& is) {}
  virtual void ReadData(std::istream& is, bool binary) {}
  virtual void WriteData(std::ostream& os, bool binary) const {}

  int32 input_dim_, output_dim_, id_;
  std::string name_;
  std::vector<std::string> input_name_;
  std::vector<int32> input_, offset_;

 private:
  static Component* NewComponentOfType(ComponentType t, int32 input_dim, int32 output_dim);
};

class UpdatableComponent : public Component {
 public:
  UpdatableComponent(int32 input_dim, int32 output_dim) : Component(input_dim, output_dim) {}
  bool IsUpdatable() const { return true; }
  virtual int32 NumParams() const = 0;
  virtual void GetParams(Vector<BaseFloat>* params) const = 0;
  // (device pointer, element count INCLUDING row padding) per parameter tensor; pointers stay valid for the life of
  // the component -- the arena/packing contract of the aslp-parallel workers (nnet-component.h:264; bsp-worker.cc:14-16)
  virtual void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) = 0;
  virtual void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) = 0;
  virtual void SetTrainOptions(const NnetTrainOptions& opts) { opts_ = opts; }
  const NnetTrainOptions& GetTrainOptions() const { return opts_; }
  virtual void InitData(std::istream& is) = 0;
 protected:
  NnetTrainOptions opts_;
};

// ---- helpers shared by the component headers ----
// "<Key> value" option parsing of a proto line (Component::Init passes the rest of the line to InitData)
class ProtoOptions {
 public:
  explicit ProtoOptions(const char* accepted) : accepted_(accepted) {}
  void Float(const char* key, float* v) { f_.push_back(std::make_pair(std::string(key), v)); }
  void Int(const char* key, int32* v) { i_.push_back(std::make_pair(std::string(key), v)); }
  void Parse(std::istream& is);
 private:
  const char* accepted_;
  std::vector<std::pair<std::string, float*>> f_;
  std::vector<std::pair<std::string, int32*>> i_;
};
// uniform [-scale, scale] fills in the reference's element order and RNG streams
// (InitMatParam: CuMatrix::SetRandUniform -> MatrixBase::SetRandUniform with a fresh RandomState; InitVecParam: global Rand())
void InitMatParam(CuMatrix<BaseFloat>* m, float scale);
void InitVecParam(CuVector<BaseFloat>* v, float scale);
void CopyRowsToVec(const CuMatrixBase<BaseFloat>& m, float* dst);     // CopyRowsFromMat into a params super-vector (synchronises)

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
