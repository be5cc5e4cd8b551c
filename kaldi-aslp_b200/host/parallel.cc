#include "parallel.h"
#include "nnet-nnet.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

namespace kaldi {

NcclNode::NcclNode(const char nccl_id[128], int nranks, int rank) : comm_(nullptr), rank_(rank), nranks_(nranks) {
  CuStream();      // device selected + stream before the communicator is created
  ASLP_OK(aslp_comm_init(&comm_, nccl_id, nranks, rank));
  // ncclCommInitRank is collective: once it returns on rank 0 every rank has read the id, so the rendezvous file can go -- a later
  // launch that reuses the path (the schedulers start one job per epoch) must never find this job's id
  if (rank == 0) if (const char* file = std::getenv("ASLP_NCCL_ID_FILE")) std::remove(file);
}
NcclNode::~NcclNode() { if (comm_ != nullptr) aslp_comm_destroy(comm_); }
void NcclNode::Barrier() { ASLP_OK(aslp_comm_barrier(comm_, CuStream())); }

void NcclNode::AllReduce(int* host_data, int n) {
  static int* dev = nullptr; static int cap = 0;
  if (n > cap) { if (dev) aslp_free(dev); ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&dev), sizeof(int) * (n + 4))); cap = n; }
  ASLP_OK(aslp_memcpy_h2d(CuStream(), dev, host_data, sizeof(int) * n));
  ASLP_OK(aslp_comm_allreduce_sum_i32(comm_, CuStream(), dev, n));
  ASLP_OK(aslp_memcpy_d2h(CuStream(), host_data, dev, sizeof(int) * n));
  CuSync();
}
void NcclNode::AllReduceDevice(float* dev, size_t n) { ASLP_OK(aslp_comm_allreduce_sum_f32(comm_, CuStream(), dev, n)); }

void NcclNode::ReduceAccStat(const std::vector<double*>& acc_params, const std::vector<std::pair<double*, int>>& data_params) {
  Barrier();
  if (!acc_params.empty()) {
    const int n = static_cast<int>(acc_params.size());
    double* dev = nullptr;
    ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&dev), sizeof(double) * n));
    std::vector<double> h(n);
    for (int i = 0; i < n; ++i) h[i] = *acc_params[i];
    ASLP_OK(aslp_memcpy_h2d(CuStream(), dev, h.data(), sizeof(double) * n));
    ASLP_OK(aslp_comm_allreduce_sum_f64(comm_, CuStream(), dev, n));
    ASLP_OK(aslp_memcpy_d2h(CuStream(), h.data(), dev, sizeof(double) * n));
    CuSync();
    for (int i = 0; i < n; ++i) *acc_params[i] = h[i];
    aslp_free(dev);
  }
  for (const auto& p : data_params) ASLP_OK(aslp_comm_allreduce_sum_f64(comm_, CuStream(), p.first, p.second));   // already device-resident
  CuSync();
}

void IWorker::Reset() {
  table_dev_ = nullptr; ntensors_ = 0; total_ = 0; nnet_ = nullptr; comm_stream_ = nullptr;
  ev_compute_ = ev_side_ = ev_done_ = ev_count_ = nullptr; count_host_ = nullptr; count_dev_ = nullptr; armed_ = false; begin_samples_ = 0;
}
IWorker::~IWorker() {
  if (armed_ && nnet_ != nullptr) nnet_->SetUpdateObserver(nullptr);
  if (comm_stream_ != nullptr) { aslp_stream_sync(comm_stream_); aslp_stream_destroy(comm_stream_); }
  aslp_event_destroy(ev_compute_); aslp_event_destroy(ev_side_); aslp_event_destroy(ev_done_); aslp_event_destroy(ev_count_);
  if (count_host_ != nullptr) aslp_free_host(count_host_);
  if (count_dev_ != nullptr) aslp_free(count_dev_);
  if (table_dev_ != nullptr) aslp_free(table_dev_);
}

void IWorker::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  std::vector<aslp_tensor_ref_t> table(params.size());
  size_t off = 0;
  for (size_t i = 0; i < params.size(); ++i) {
    table[i].ptr = params[i].first;
    table[i].offset = off;
    table[i].n = static_cast<size_t>(params[i].second);
    off += (table[i].n + 3) / 4 * 4;          // keep every tensor 16-byte aligned inside the arena
  }
  ntensors_ = static_cast<int>(params.size());
  total_ = off;
  segments_.clear();
  if (table_dev_ != nullptr) aslp_free(table_dev_);
  ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&table_dev_), sizeof(aslp_tensor_ref_t) * (table.size() + 1)));
  ASLP_OK(aslp_memcpy_h2d(CuStream(), table_dev_, table.data(), sizeof(aslp_tensor_ref_t) * table.size()));
  CuSync();
  arena_.Resize(static_cast<int32>(total_), kSetZero);
}

// the net's tensors in Nnet::GetGpuParams order, plus the component each run of them belongs to
void IWorker::InitParam(aslp_nnet::Nnet* nnet) {
  using aslp_nnet::Component; using aslp_nnet::UpdatableComponent;
  KALDI_ASSERT(nnet != nullptr);
  std::vector<std::pair<BaseFloat*, int>> params;
  std::vector<Segment> segs;
  size_t off = 0;
  for (int32 c = 0; c < nnet->NumComponents(); ++c) {
    Component& comp = nnet->GetComponent(c);
    if (!comp.IsUpdatable()) continue;
    std::vector<std::pair<BaseFloat*, int>> cp;
    dynamic_cast<UpdatableComponent&>(comp).GetGpuParams(&cp);
    if (cp.empty()) continue;
    Segment s;
    s.component = c; s.first = static_cast<int>(params.size()); s.count = static_cast<int>(cp.size()); s.offset = off;
    for (const auto& t : cp) off += (static_cast<size_t>(t.second) + 3) / 4 * 4;
    s.length = off - s.offset;
    segs.push_back(s);
    params.insert(params.end(), cp.begin(), cp.end());
  }
  InitParam(params);                 // the worker's own (virtual) registration: arena, previous model, optimizer state
  KALDI_ASSERT(off == total_);
  segments_ = segs;
  nnet_ = nnet;
}

bool IWorker::AllFinished(int num_worker_samples, int* num_all) {
  *num_all = num_worker_samples;
  AllReduce(num_all, 1);
  if (*num_all <= 0) { KALDI_LOG << "All worker finished their data"; return true; }
  return false;
}

void IWorker::ExchangeAll() {
  if (segments_.empty()) {
    Segment all; all.component = -1; all.first = 0; all.count = ntensors_; all.offset = 0; all.length = total_;
    ExchangeSegment(CuStream(), all);
  } else {
    for (size_t i = segments_.size(); i-- > 0;) ExchangeSegment(CuStream(), segments_[i]);
  }
  AfterExchange();
}

void IWorker::BeginSynchronize(int num_worker_samples) {
  KALDI_ASSERT(CanOverlap() && nnet_ != nullptr && !segments_.empty() && !armed_);
  KALDI_ASSERT(num_worker_samples > 0);        // a rank inside a minibatch has frames: the job's total cannot be "all finished"
  if (comm_stream_ == nullptr) {
    ASLP_OK(aslp_stream_create(&comm_stream_));
    ASLP_OK(aslp_malloc_host(reinterpret_cast<void**>(&count_host_), 64));
    ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&count_dev_), 64));
  }
  // the frame-count exchange of the termination protocol, as Synchronize issues it first -- here without the host waiting
  ASLP_OK(aslp_event_record(CuStream(), &ev_compute_));
  ASLP_OK(aslp_stream_wait_event(comm_stream_, ev_compute_));    // behind whatever an earlier (blocking) exchange left on the compute stream
  count_host_[0] = num_worker_samples;
  ASLP_OK(aslp_memcpy_h2d(comm_stream_, count_dev_, count_host_, sizeof(int)));
  ASLP_OK(aslp_comm_allreduce_sum_i32(comm_, comm_stream_, count_dev_, 1));
  ASLP_OK(aslp_memcpy_d2h(comm_stream_, count_host_ + 1, count_dev_, sizeof(int)));
  ASLP_OK(aslp_event_record(comm_stream_, &ev_count_));
  exchanged_.assign(segments_.size(), 0);
  armed_ = true;
  begin_samples_ = num_worker_samples;
  nnet_->SetUpdateObserver([this](int component) { OnComponentUpdated(component); });
}

void IWorker::OnComponentUpdated(int component) {
  if (!armed_) return;
  for (size_t i = 0; i < segments_.size(); ++i) {
    if (segments_[i].component != component || exchanged_[i]) continue;
    // Collectives must be issued in the same order on every rank: top component first, none skipped.
    for (size_t j = segments_.size(); j-- > i + 1;) KALDI_ASSERT(exchanged_[j]);
    // the Update may sit on the compute stream, on the side stream (weight-gradient tail of the recurrent layers), or both
    ASLP_OK(aslp_event_record(CuStream(), &ev_compute_));
    ASLP_OK(aslp_stream_wait_event(comm_stream_, ev_compute_));
    ASLP_OK(aslp_event_record(CuSideStream(), &ev_side_));
    ASLP_OK(aslp_stream_wait_event(comm_stream_, ev_side_));
    ExchangeSegment(comm_stream_, segments_[i]);
    exchanged_[i] = 1;
  }
}

bool IWorker::EndSynchronize() {
  KALDI_ASSERT(armed_);
  nnet_->SetUpdateObserver(nullptr);
  // components whose Update was not reported (a net driven without Nnet::Backpropagate): exchange them now, in order
  for (size_t i = segments_.size(); i-- > 0;) {
    if (exchanged_[i]) continue;
    ASLP_OK(aslp_event_record(CuStream(), &ev_compute_));
    ASLP_OK(aslp_stream_wait_event(comm_stream_, ev_compute_));
    ASLP_OK(aslp_event_record(CuSideStream(), &ev_side_));
    ASLP_OK(aslp_stream_wait_event(comm_stream_, ev_side_));
    ExchangeSegment(comm_stream_, segments_[i]);
    exchanged_[i] = 1;
  }
  armed_ = false;
  AfterExchange();
  ASLP_OK(aslp_event_record(comm_stream_, &ev_done_));
  ASLP_OK(aslp_stream_wait_event(CuStream(), ev_done_));         // the next Propagate reads the exchanged model
  ASLP_OK(aslp_stream_wait_event(CuSideStream(), ev_done_));
  ASLP_OK(aslp_event_sync(ev_count_));                            // long complete: it was issued before the minibatch
  return count_host_[1] > 0;
}

// frame-weighted MODEL average (bsp-worker.cc:33-58): w <- sum_r (frames_r / frames_all) * w_r
bool BspWorker::Synchronize(int num_worker_samples) {
  int num_all = 0;
  if (AllFinished(num_worker_samples, &num_all)) return false;
  factor_ = static_cast<float>(num_worker_samples) / num_all;
  KALDI_ASSERT(factor_ >= 0.0 && factor_ <= 1.0);
  ExchangeAll();
  return true;
}
void BspWorker::ExchangeSegment(aslp_stream_t st, const Segment& seg) {
  // pipelined: the job's frame total is still on the device, behind its all-reduce on this stream; the weight is the same
  // fp32 division either way, so the two forms give the same bits
  if (armed_) ASLP_OK(aslp_sync_pack_weighted(st, arena_.Data(), table_dev_ + seg.first, seg.count, begin_samples_, count_dev_));
  else ASLP_OK(aslp_sync_pack(st, arena_.Data(), table_dev_ + seg.first, seg.count, factor_));
  ASLP_OK(aslp_comm_allreduce_sum_f32(comm_, st, arena_.Data() + seg.offset, seg.length));
  ASLP_OK(aslp_sync_unpack(st, arena_.Data(), table_dev_ + seg.first, seg.count));
}

void BmufWorker::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  IWorker::InitParam(params);
  w_prev_.Resize(static_cast<int32>(total_), kSetZero);
  delta_prev_.Resize(static_cast<int32>(total_), kSetZero);
  ASLP_OK(aslp_sync_pack(CuStream(), w_prev_.Data(), table_dev_, ntensors_, 1.0f));      // prev = initial model
}
// block-wise model-update filtering (bmuf-worker.cc:37-68): G = SUM_r (w_r - w_prev);
// delta = mom * delta_prev + (1 - mom) * lr * G ; w = w_prev + delta ; w_prev = w ; delta_prev = delta
bool BmufWorker::Synchronize(int num_worker_samples) {
  int num_all = 0;
  if (AllFinished(num_worker_samples, &num_all)) return false;
  ExchangeAll();
  return true;
}
void BmufWorker::ExchangeSegment(aslp_stream_t st, const Segment& seg) {
  ASLP_OK(aslp_sync_pack_diff(st, arena_.Data(), table_dev_ + seg.first, seg.count, w_prev_.Data(), 1.0f));
  ASLP_OK(aslp_comm_allreduce_sum_f32(comm_, st, arena_.Data() + seg.offset, seg.length));
  ASLP_OK(aslp_sync_bmuf_apply_packed(st, table_dev_ + seg.first, seg.count, w_prev_.Data(), delta_prev_.Data(), arena_.Data(), momentum_, learn_rate_));
}

void SodWorker::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  IWorker::InitParam(params);
  w_prev_.Resize(static_cast<int32>(total_), kSetZero);
  s1_.Resize(static_cast<int32>(total_), kSetZero);
  s2_.Resize(static_cast<int32>(total_), kSetZero);
  ASLP_OK(aslp_sync_pack(CuStream(), w_prev_.Data(), table_dev_, ntensors_, 1.0f));
}
// "synchronously optimise the difference" (sod-worker.cc:37-61): G = SUM_r (w_prev - w_r) fed to the chosen optimizer
bool SodWorker::Synchronize(int num_worker_samples) {
  int num_all = 0;
  if (AllFinished(num_worker_samples, &num_all)) return false;
  ExchangeAll();
  return true;
}
void SodWorker::ExchangeSegment(aslp_stream_t st, const Segment& seg) {
  ASLP_OK(aslp_sync_pack_diff(st, arena_.Data(), table_dev_ + seg.first, seg.count, w_prev_.Data(), -1.0f));
  ASLP_OK(aslp_comm_allreduce_sum_f32(comm_, st, arena_.Data() + seg.offset, seg.length));
  int opt = ASLP_OPT_SGD; float lr = config_.lr, p1 = 0.f, p2 = 0.f;
  if (config_.solver == "sgd") { opt = ASLP_OPT_SGD; }
  else if (config_.solver == "momentum") { opt = ASLP_OPT_MOMENTUM; p1 = config_.momentum; }
  else if (config_.solver == "adagrad") { opt = ASLP_OPT_ADAGRAD; lr = config_.adagrad_lr; }
  else if (config_.solver == "rmsprop") { opt = ASLP_OPT_RMSPROP; lr = config_.rmsprop_lr; }
  else if (config_.solver == "adadelta") { opt = ASLP_OPT_ADADELTA; p1 = config_.adadelta_gamma; }
  else if (config_.solver == "adam") { opt = ASLP_OPT_ADAM; lr = config_.adam_lr; p1 = config_.adam_beta1; p2 = config_.adam_beta2; }
  else { KALDI_ERR << "Unknown solver type " << config_.solver; }
  ASLP_OK(aslp_sync_sod_apply_packed(st, opt, table_dev_ + seg.first, seg.count, arena_.Data(), s1_.Data(), s2_.Data(), w_prev_.Data(), lr, p1, p2, 1e-8f, step_));
}

WorkerBootstrap::WorkerBootstrap() : rank(0), nranks(1) {
  std::memset(id, 0, sizeof(id));
  auto env_int = [](const char* a, const char* b, int dflt) {
    const char* v = std::getenv(a);
    if (v == nullptr) v = std::getenv(b);
    return v != nullptr ? std::atoi(v) : dflt;
  };
  rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0);
  nranks = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
  KALDI_ASSERT(nranks >= 1 && rank >= 0 && rank < nranks);
  const char* file = std::getenv("ASLP_NCCL_ID_FILE");
  if (nranks > 1 && file == nullptr) KALDI_ERR << "WORLD_SIZE > 1 needs ASLP_NCCL_ID_FILE (a path every rank can reach) to pass the NCCL id";
  // A file left behind by an earlier launch (a job that died before its communicator was up) must not be taken for this one's:
  // the file carries a job nonce after the id -- ASLP_JOB_NONCE, else torchrun's TORCHELASTIC_RUN_ID, else MASTER_ADDR:MASTER_PORT,
  // which every rank of one launch shares -- and a rank keeps waiting while the nonce on disk is not its own.  Rank 0 removes the
  // file before writing and again once the communicator is up (NcclNode).  Without any of those variables only the removals
  // protect a reused path: give every launch its own path then.
  char nonce[64];
  std::memset(nonce, 0, sizeof(nonce));
  {
    std::string n;
    if (const char* v = std::getenv("ASLP_JOB_NONCE")) n = v;
    else if (const char* v = std::getenv("TORCHELASTIC_RUN_ID")) n = v;
    else if (const char* v = std::getenv("MASTER_PORT")) { const char* a = std::getenv("MASTER_ADDR"); n = std::string(a ? a : "") + ":" + v; }
    std::strncpy(nonce, n.c_str(), sizeof(nonce) - 1);
  }
  if (rank == 0) {
    ASLP_OK(aslp_comm_unique_id(id));
    if (file != nullptr) {
      std::remove(file);
      const std::string tmp = std::string(file) + ".tmp";
      { std::ofstream os(tmp, std::ios::binary); os.write(id, sizeof(id)); os.write(nonce, sizeof(nonce)); if (!os.good()) KALDI_ERR << "Cannot write " << tmp; }
      if (std::rename(tmp.c_str(), file) != 0) KALDI_ERR << "Cannot publish " << file;
    }
  } else {
    for (int tries = 0;; ++tries) {
      std::ifstream is(file, std::ios::binary);
      if (is.good()) {
        char got[64];
        is.read(id, sizeof(id));
        const bool have_id = is.gcount() == static_cast<std::streamsize>(sizeof(id));
        is.read(got, sizeof(got));
        if (have_id && is.gcount() == static_cast<std::streamsize>(sizeof(got)) && std::memcmp(got, nonce, sizeof(nonce)) == 0) break;
      }
      if (tries > 1200) KALDI_ERR << "Timed out waiting for " << file << " (a file with another launch's nonce does not count)";
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  }
}

}  // namespace kaldi
