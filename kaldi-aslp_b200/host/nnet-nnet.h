// nnet-nnet.h -- the Nnet graph executor, public API as src/aslp-nnet/nnet-nnet.h:38-174.
// Buffers input_buf_/output_buf_/input_diff_buf_/output_diff_buf_ per component as in the reference
// (nnet-nnet.cc:70-154), with Update() applied right after each component's Backpropagate (:126-129).
#ifndef ASLP_HOST_NNET_NNET_H_
#define ASLP_HOST_NNET_NNET_H_
#include <functional>
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class Nnet {
 public:
  Nnet() {}
  Nnet(const Nnet& other);
  Nnet& operator=(const Nnet& other);
  ~Nnet();

  void Propagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out);
  void Propagate(const std::vector<const CuMatrixBase<BaseFloat>*>& in, std::vector<CuMatrix<BaseFloat>*>* out);
  void Backpropagate(const CuMatrixBase<BaseFloat>& out_diff, CuMatrix<BaseFloat>* in_diff);
  void Backpropagate(const std::vector<const CuMatrixBase<BaseFloat>*>& out_diff, std::vector<CuMatrix<BaseFloat>*>* in_diff);
  void Feedforward(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out);
  void Feedforward(const std::vector<const CuMatrixBase<BaseFloat>*>& in, std::vector<CuMatrix<BaseFloat>*>* out);
  void GetComponentTime();

  int32 InputDim() const;
  int32 OutputDim() const;
  int32 NumInput() const { return static_cast<int32>(input_.size()); }
  int32 NumOutput() const { return static_cast<int32>(output_.size()); }
  int32 NumComponents() const { return static_cast<int32>(components_.size()); }
  const Component& GetComponent(int32 c) const;
  Component& GetComponent(int32 c);
  void SetComponent(int32 c, Component* component);
  void AppendComponent(Component* dynamically_allocated_comp);
  void AppendNnet(const Nnet& nnet_to_append);
  void RemoveComponent(int32 c);
  void RemoveLastComponent() { RemoveComponent(NumComponents() - 1); }

  // per-component buffers of the last Propagate / Backpropagate (the reference's public accessors return the legacy,
  // never-filled propagate_buf_; these return the live ones so parity tests can compare layer by layer)
  const std::vector<CuMatrix<BaseFloat>>& PropagateBuffer() const { return output_buf_; }
  const std::vector<CuMatrix<BaseFloat>>& BackpropagateBuffer() const { return output_diff_buf_; }   // d(loss)/d(output of component c)

  int32 NumParams() const;
  void GetParams(Vector<BaseFloat>* wei_copy) const;
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params);
  void GetAccStats(std::vector<double*>* acc_params, std::vector<std::pair<double*, int>>* data_params);

  void ResetLstmStreams(const std::vector<int32>& stream_reset_flag);
  void SetSeqLengths(const std::vector<int32>& sequence_lengths);
  void SetChunkSize(int chunk_size);
  void SetDropoutRetention(BaseFloat r);        // nnet-nnet.h:153

  void Init(const std::string& config_file);
  void Read(const std::string& file);
  void Read(std::istream& in, bool binary);
  void Write(const std::string& file, bool binary) const;
  void Write(std::ostream& out, bool binary) const;
  void WriteStandard(const std::string& file, bool binary) const;
  void WriteStandard(std::ostream& out, bool binary) const;

  std::string Info() const;
  std::string InfoGradient() const;
  std::string InfoPropagate() const;
  std::string InfoBackPropagate() const;
  void Check() const;
  void Destroy();

  // true when a training step of this net is pure device work that depends on nothing but its input buffers: every
  // component is of a kind that keeps no per-step state on the host (no stream-reset flags, sequence lengths, random
  // masks, running statistics) -- what XentTrainStep needs to record the step once and replay it (nnet-train-step.h)
  bool StepReplayable() const;
  // called by Backpropagate with a component's index right after that component's Update has been enqueued (its parameters
  // are final for this minibatch once the compute and side streams reach this point): the parallel workers start exchanging
  // the component's tensors while the layers below are still back-propagating (parallel.h, IWorker::BeginSynchronize).
  // A net with an observer is not step-replayable.  NULL removes it.
  void SetUpdateObserver(std::function<void(int32)> f) { update_observer_ = f; }
  void SetTrainOptions(const NnetTrainOptions& opts);
  const NnetTrainOptions& GetTrainOptions() const { return opts_; }
  void AutoComplete();
  void AssignComponentId(std::vector<Component*>& components);
  void SortComponent(std::vector<Component*>& components);

 private:
  void InitInputOutput();
  std::vector<Component*> components_;
  std::vector<int32> input_, output_;
  std::vector<std::pair<std::string, BaseFloat>> propagate_time_, back_propagate_time_;
  std::vector<CuMatrix<BaseFloat>> input_buf_, output_buf_, input_diff_buf_, output_diff_buf_;
  NnetTrainOptions opts_;
  std::function<void(int32)> update_observer_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
