// parallel.h -- the aslp-parallel worker interface over NCCL (NVLink 5 / NVSwitch) instead of host-staged MPI.
// Reference: src/aslp-parallel/itf.h:27-36 (IWorker), mpi-node.h:18-101 (MpiNode), bsp-worker.{h,cc},
// bmuf-worker.{h,cc}, sod-worker.{h,cc}, optimizer.h:172-232 (OptimizerOption).
// The parameter tensors stay in the components (non-owning GetGpuParams views, as in the reference); each
// Synchronize is: one int allreduce (frame count / termination protocol), ONE multi-tensor pack kernel, ONE
// ncclAllReduce of the packed arena, ONE multi-tensor apply kernel -- versus 44 x (scale, D2H, MPI_Allreduce, H2D).
#ifndef ASLP_HOST_PARALLEL_H_
#define ASLP_HOST_PARALLEL_H_
#include "matrix.h"
#include "parse-options.h"

namespace kaldi {

// replaces MpiNode: owns the communicator (ctor/dtor owned MPI_Init/Finalize there)
class NcclNode {
 public:
  NcclNode(const char nccl_id[128], int nranks, int rank);
  virtual ~NcclNode();
  int Rank() const { return rank_; }
  int NumNodes() const { return nranks_; }
  bool IsMainNode() const { return rank_ == 0; }
  void Barrier();
  void AllReduce(int* host_data, int n);                          // in-place SUM of host ints (mpi-node.h:69-73)
  void AllReduceDevice(float* dev, size_t n);
  // BatchNorm statistics at the end of an epoch: frame counters (host doubles) and the device fp64 running sums
  void ReduceAccStat(const std::vector<double*>& acc_params, const std::vector<std::pair<double*, int>>& data_params);
 protected:
  aslp_comm_t comm_;
  int rank_, nranks_;
};

// Process-launch bootstrap for the worker mains (what mpirun + MPI_Init did in the reference): rank and world size
// from RANK / WORLD_SIZE (torchrun's variables; OMPI_COMM_WORLD_RANK / _SIZE are honoured too), the ncclUniqueId
// through a file named by ASLP_NCCL_ID_FILE (rank 0 writes it atomically, the others wait for it).
struct WorkerBootstrap {
  char id[128];
  int rank, nranks;
  WorkerBootstrap();
};

namespace aslp_nnet { class Nnet; }

class IWorker : public NcclNode {
 public:
  IWorker(const char id[128], int nranks, int rank) : NcclNode(id, nranks, rank) { Reset(); }
  // the reference's workers take no launch arguments (MpiNode's constructor calls MPI_Init, mpi-node.h:21-27): bootstrap from the
  // environment instead, so that `new BspWorker()` / `new BmufWorker(momentum, learn_rate)` in an unmodified main keep working
  explicit IWorker(const WorkerBootstrap& b) : NcclNode(b.id, b.nranks, b.rank) { Reset(); }
  virtual ~IWorker();
  virtual void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  virtual bool Synchronize(int num_worker_samples) = 0;   // false when every rank is out of data
  virtual bool IsAsync() const { return false; }          // true: a parameter-server worker (Synchronize never says "all done")
  // a rank that has finished its shard keeps answering with zero frames so collectives stay matched (bsp-worker.cc:60-65)
  virtual void Stop() { KALDI_LOG << "Worker " << Rank() << "finished, waitting for others"; while (Synchronize(0)) {} }

  // ---- the exchange pipelined by layer (not in the reference, whose Synchronize blocks after the minibatch: same arithmetic,
  // same result per tensor).  InitParam(nnet) registers the net's tensors as InitParam(params) does and remembers which belong
  // to which component; every Synchronize then exchanges component by component, top layer first -- on EVERY rank, so all
  // ranks of a job must use the same form of InitParam.  A trainer that knows a synchronisation is due after the coming
  // minibatch calls BeginSynchronize(n) before it instead of Synchronize(n) after it: the frame counts are exchanged at once,
  // and each component's tensors as soon as Nnet::Backpropagate has enqueued its Update -- on the worker's own stream, while
  // the layers below are still back-propagating.  EndSynchronize() after the minibatch makes the compute stream wait for
  // the last exchange and returns what Synchronize(n) would have.  Only workers whose exchange does not need the global
  // frame count on the HOST up front can do this (CanOverlap(): BSP -- whose weight frames_r / frames_all is taken on the
  // device behind the count's all-reduce --, BMUF and SOD).
  virtual bool CanOverlap() const { return false; }
  void InitParam(aslp_nnet::Nnet* nnet);
  void BeginSynchronize(int num_worker_samples);
  bool EndSynchronize();
 protected:
  struct Segment { int component, first, count; size_t offset, length; };   // tensors [first, first + count) = arena [offset, offset + length)
  void Reset();
  bool AllFinished(int num_worker_samples, int* num_all);
  // pack, all-reduce and apply the tensors of one segment on stream `st` (the whole exchange when there are no segments)
  virtual void ExchangeSegment(aslp_stream_t st, const Segment& seg) { KALDI_ERR << "this worker has no segmented exchange"; }
  void ExchangeAll();                               // every segment on the compute stream, top component first
  virtual void AfterExchange() {}                   // once per synchronisation (SOD's step counter)
  void OnComponentUpdated(int component);
  aslp_tensor_ref_t* table_dev_;
  int ntensors_;
  size_t total_;            // packed arena length
  CuVector<BaseFloat> arena_;          // the all-reduce buffer
  std::vector<Segment> segments_;      // in component order; empty: one exchange of everything
  aslp_nnet::Nnet* nnet_;
  aslp_stream_t comm_stream_;
  void *ev_compute_, *ev_side_, *ev_done_, *ev_count_;
  int* count_host_;                    // page-locked: this rank's frame count in, the job's total out
  int* count_dev_;
  bool armed_;
  int begin_samples_;                  // the frame count BeginSynchronize was given
  std::vector<char> exchanged_;
};

class BspWorker : public IWorker {
 public:
  BspWorker(const char id[128], int nranks, int rank) : IWorker(id, nranks, rank) {}
  BspWorker() : IWorker(WorkerBootstrap()) {}                                  // bsp-worker.h:21
  bool Synchronize(int num_worker_samples);
  bool CanOverlap() const { return true; }
 protected:
  void ExchangeSegment(aslp_stream_t st, const Segment& seg);
 private:
  float factor_;
};

class BmufWorker : public IWorker {
 public:
  BmufWorker(const char id[128], int nranks, int rank, float momentum = 0.9f, float learn_rate = 1.0f)
      : IWorker(id, nranks, rank), momentum_(momentum), learn_rate_(learn_rate) {}
  // the reference's own signature and argument ORDER (bmuf-worker.h:31: learn rate first)
  explicit BmufWorker(float learn_rate = 1.0f, float momentum = 0.9f) : IWorker(WorkerBootstrap()), momentum_(momentum), learn_rate_(learn_rate) {}
  using IWorker::InitParam;
  void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  bool Synchronize(int num_worker_samples);
  bool CanOverlap() const { return true; }
 protected:
  void ExchangeSegment(aslp_stream_t st, const Segment& seg);
 private:
  float momentum_, learn_rate_;
  CuVector<BaseFloat> w_prev_, delta_prev_;
};

struct OptimizerOption {
  std::string solver;
  float lr, momentum, adagrad_lr, rmsprop_lr, adam_lr, adadelta_gamma, adam_beta1, adam_beta2;
  OptimizerOption() : solver("momentum"), lr(0.01f), momentum(0.9f), adagrad_lr(0.01f), rmsprop_lr(0.001f), adam_lr(0.001f),
                      adadelta_gamma(0.95f), adam_beta1(0.9f), adam_beta2(0.999f) {}
  void Register(OptionsItf* opts) {
    opts->Register("solver", &solver, "Optimizer solver(sgd | momentum | adagrad | adadelta | rmsprop | adam)");
    opts->Register("lr", &lr, "learning rate for (sgd | momentum) optimizer");
    opts->Register("sgd-momentum", &momentum, "momentum for (momentum) optimizer");
    opts->Register("adagrad-lr", &adagrad_lr, "learning rate for (adagrad) optimizer");
    opts->Register("rmsprop-lr", &rmsprop_lr, "learning rate for (rmsprop) optimizer");
    opts->Register("adam-lr", &adam_lr, "learning rate for (adam) optimizer");
    opts->Register("adadelta-gamma", &adadelta_gamma, "update factor for (adadelta) optimizer");
    opts->Register("adam-beta1", &adam_beta1, "update mean factor for (adam) optimizer");
    opts->Register("adam-beta2", &adam_beta2, "update variance factor for (adam) optimizer");
  }
};

class SodWorker : public IWorker {
 public:
  SodWorker(const char id[128], int nranks, int rank, const OptimizerOption& config) : IWorker(id, nranks, rank), config_(config), step_(1) {}
  explicit SodWorker(const OptimizerOption& config) : IWorker(WorkerBootstrap()), config_(config), step_(1) {}              // sod-worker.h
  using IWorker::InitParam;
  void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  bool Synchronize(int num_worker_samples);
  bool CanOverlap() const { return true; }
 protected:
  void ExchangeSegment(aslp_stream_t st, const Segment& seg);
  void AfterExchange() { ++step_; }
 private:
  OptimizerOption config_;
  int step_;
  CuVector<BaseFloat> w_prev_, s1_, s2_;
};

}  // namespace kaldi
#endif
