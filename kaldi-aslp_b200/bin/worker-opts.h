// worker-opts.h -- the data-parallel extension shared by the trainer mains: the flags of the reference worker mains
// (src/aslp-parallelbin/aslp-nnet-train-frame-worker.cc:60-80, aslp-nnet-train-lstm-stream-worker.cc:70-95,
// aslp-nnet-train-lc-blstm-streams-worker.cc:100-120: --worker-type, --alpha, --sync-period, --bmuf-momentum,
// --bmuf-learn-rate, optimizer options), worker construction after the model is read, the per-minibatch frame counter with
// its Synchronize, and the end-of-data protocol.  One rank per GPU (RANK / WORLD_SIZE / LOCAL_RANK, NCCL id through
// ASLP_NCCL_ID_FILE); with an empty --worker-type a main is the plain single-process trainer.
#ifndef ASLP_BIN_WORKER_OPTS_H_
#define ASLP_BIN_WORKER_OPTS_H_
#include <memory>
#include "nnet-nnet.h"
#include "parallel-async.h"
#include "parse-options.h"

namespace kaldi {

// End of a rank's data, as in the reference worker mains (aslp-nnet-train-frame-worker.cc:171-178): Stop() -- synchronous workers
// answer zero-frame syncs until every rank is out of data (bsp-worker.cc:60-65; frames trained since the last sync are NOT
// flushed first, as in the reference), asynchronous ones tell the server they are finished -- then the BatchNorm frame counters and
// fp64 running sums are all-reduced on EVERY rank (the server main joins the same collective), so that the model rank 0 writes
// carries the statistics of all shards.
inline void FinishWorker(IWorker* worker, aslp_nnet::Nnet* nnet) {
  worker->Stop();
  std::vector<double*> acc_params;
  std::vector<std::pair<double*, int>> data_params;
  nnet->GetAccStats(&acc_params, &data_params);
  worker->ReduceAccStat(acc_params, data_params);
}

// The per-minibatch frame counter with its synchronisation (frame-worker.cc:150-156), in two forms: the reference's -- count,
// then Synchronize once the period is exceeded -- and, for workers that can (IWorker::CanOverlap: bsp, bmuf, sod) and were
// registered with InitParam(nnet), the exchange pipelined by layer: the trainer says BEFORE Backpropagate how many frames the
// minibatch has, a synchronisation that this minibatch makes due is begun there and rides under the backward pass.
// Same moments of synchronisation, same arithmetic, same model afterwards.
struct SyncCounter {
  IWorker* worker;
  int32 sync_period, frames_since_sync;
  bool pipelined, begun;
  SyncCounter() : worker(nullptr), sync_period(25600), frames_since_sync(0), pipelined(false), begun(false) {}
  // registers the net's tensors; pipeline = false keeps the reference's blocking form
  void Attach(IWorker* w, aslp_nnet::Nnet* nnet, int32 period, bool pipeline) {
    worker = w; sync_period = period; frames_since_sync = 0; begun = false;
    pipelined = pipeline && w->CanOverlap();
    if (pipelined) {
      w->InitParam(nnet);
    } else {
      std::vector<std::pair<BaseFloat*, int>> params;
      nnet->GetGpuParams(&params);
      w->InitParam(params);
    }
  }
  void BeforeBackpropagate(int32 frames) {
    if (worker == nullptr || !pipelined || frames_since_sync + frames <= sync_period) return;
    worker->BeginSynchronize(frames_since_sync + frames);
    begun = true;
  }
  void Progress(int32 frames) {
    if (worker == nullptr) return;
    frames_since_sync += frames;
    if (begun) {
      KALDI_LOG << "Worker " << worker->Rank() << " synchronize once";
      worker->EndSynchronize();
      begun = false;
      frames_since_sync = 0;
    } else if (frames_since_sync > sync_period) {
      KALDI_LOG << "Worker " << worker->Rank() << " synchronize once";
      worker->Synchronize(frames_since_sync);
      frames_since_sync = 0;
    }
  }
};

struct WorkerOptions {
  std::string worker_type;
  float alpha, bmuf_momentum, bmuf_learn_rate;
  int32 sync_period;
  bool pipeline_sync;
  OptimizerOption optimizer_opts;
  std::unique_ptr<IWorker> worker;
  SyncCounter counter;
  aslp_nnet::Nnet* nnet_ = nullptr;
  WorkerOptions() : alpha(0.5f), bmuf_momentum(0.9f), bmuf_learn_rate(1.0f), sync_period(25600), pipeline_sync(true) {}
  void Register(ParseOptions* po) {
    po->Register("worker-type", &worker_type, "Worker type(bsp | bmuf | sod | easgd | asgd); empty: single process");
    po->Register("alpha", &alpha, "Moving rate alpha for easgd worker");
    po->Register("sync-period", &sync_period, "number frames for every synchronization");
    po->Register("bmuf-momentum", &bmuf_momentum, "bmuf block momentum");
    po->Register("bmuf-learn-rate", &bmuf_learn_rate, "bmuf block learning rate");
    po->Register("pipeline-sync", &pipeline_sync, "bsp | bmuf | sod: exchange each layer's tensors behind its Update, under the backward pass "
                 "(same result as the blocking exchange after the minibatch; every rank must use the same setting)");
    optimizer_opts.Register(po);
  }
  // before the model is read: one GPU per rank
  void SelectDevice() const {
    if (worker_type.empty()) return;
    if (const char* lr = std::getenv("LOCAL_RANK")) ASLP_OK(aslp_set_device(std::atoi(lr)));
  }
  void Create(aslp_nnet::Nnet* nnet, bool crossvalidate) {
    if (worker_type.empty() || crossvalidate) return;
    nnet_ = nnet;
    WorkerBootstrap boot;
    if (worker_type == "bsp") worker.reset(new BspWorker(boot.id, boot.nranks, boot.rank));
    else if (worker_type == "bmuf") worker.reset(new BmufWorker(boot.id, boot.nranks, boot.rank, bmuf_momentum, bmuf_learn_rate));
    else if (worker_type == "sod") worker.reset(new SodWorker(boot.id, boot.nranks, boot.rank, optimizer_opts));
    else if (worker_type == "easgd") worker.reset(new EasgdWorker(boot.id, boot.nranks, boot.rank, alpha));     // rank 0 runs aslp-nnet-train-server
    else if (worker_type == "asgd") worker.reset(new AsgdWorker(boot.id, boot.nranks, boot.rank));
    else KALDI_ERR << "Unsupported worker type: " << worker_type;
    counter.Attach(worker.get(), nnet, sync_period, pipeline_sync);
  }
  // before Backpropagate, with the frames of the minibatch that is about to be back-propagated (pipelined exchange only)
  void BeforeBackpropagate(int32 frames) { counter.BeforeBackpropagate(frames); }
  // after every minibatch (frame-worker.cc:150-156)
  void Progress(int32 frames) { counter.Progress(frames); }
  void Finish() { if (worker) FinishWorker(worker.get(), nnet_); }
  bool WritesModel() const { return !worker || worker->IsMainNode(); }
};

}  // namespace kaldi
#endif
