// parallel-async.h -- the asynchronous parameter-server modes of aslp-parallel over NCCL point-to-point:
// EASGD (easgd-worker.{h,cc}, easgd-server.{h,cc}), ASGD (asgd-worker.{h,cc}, asgd-server.{h,cc}) and MASGD
// (masgd-server.{h,cc}, "LMASGD": one momentum buffer per worker; its worker is AsgdWorker).  Rank 0 is the server, as
// MpiNode::MainNode() is in the reference.
//
// What replaces what:
//  * MPI_Recv(MPI_ANY_SOURCE, kTagMsg) of the server loop -> a loopback TCP control channel (NCCL has no any-source
//    receive): every worker holds one connection to rank 0 and sends {kMsgSynchronize | kMsgFinished}; the server poll()s
//    them, so requests are served in ARRIVAL order exactly as the reference serves them;
//  * the per-tensor MPI_Send / MPI_Recv / MPI_Sendrecv through host staging buffers -> ONE ncclSend / ncclRecv pair of the
//    packed fp32 arena (multi-tensor pack / unpack kernels), device to device over NVLink;
//  * the per-tensor AddVec updates -> one axpby over the arena.
// The arithmetic is the reference's: EASGD moves worker and server towards each other by alpha using each other's
// PRE-update model; ASGD sends the accumulated delta w - w_prev, the server adds alpha * delta and answers with its model,
// with the optional periodic barrier every sync_period updates; MASGD filters each worker's deltas with its own momentum.
#ifndef ASLP_HOST_PARALLEL_ASYNC_H_
#define ASLP_HOST_PARALLEL_ASYNC_H_
#include "parallel.h"

namespace kaldi {

typedef enum { kMsgSynchronize = 0x00, kMsgFinished = 0x01 } MpiMsgType;      // itf.h:19-22

// port of the control channel: ASLP_CTRL_PORT, else MASTER_PORT + 1 (torchrun), else 29631
int CtrlPort();

class CtrlServer {            // rank 0
 public:
  CtrlServer(int port, int nworkers);
  ~CtrlServer();
  void RecvAny(int* worker_rank, int* msg_type);     // blocks; the MPI_Recv(ANY_SOURCE) of the server loops
 private:
  int listen_fd_;
  std::vector<int> fds_, ranks_;
};
class CtrlClient {            // ranks 1..N-1
 public:
  CtrlClient(int port, int rank);
  ~CtrlClient();
  void Send(int msg_type);
 private:
  int fd_;
};

class IServer : public NcclNode {             // itf.h:38-43
 public:
  IServer(const char id[128], int nranks) : NcclNode(id, nranks, 0), table_dev_(nullptr), ntensors_(0), total_(0), ctrl_(nullptr) {}
  virtual ~IServer();
  virtual void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  virtual void Run() = 0;
 protected:
  aslp_tensor_ref_t* table_dev_;
  int ntensors_;
  size_t total_;
  CuVector<BaseFloat> server_arena_, worker_arena_;
  CtrlServer* ctrl_;
};

class EasgdWorker : public IWorker {
 public:
  EasgdWorker(const char id[128], int nranks, int rank, float alpha = 0.5f);
  explicit EasgdWorker(float alpha);                           // easgd-worker.h: environment bootstrap (see IWorker)
  ~EasgdWorker();
  void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  bool Synchronize(int num_worker_samples);        // always true (easgd-worker.cc:66)
  bool IsAsync() const { return true; }
  void Stop();
 private:
  float alpha_;
  CuVector<BaseFloat> server_arena_;
  CtrlClient* ctrl_;
};
class EasgdServer : public IServer {
 public:
  EasgdServer(const char id[128], int nranks, float alpha = 0.5f) : IServer(id, nranks), alpha_(alpha) {}
  void Run();
  void Update(int worker_rank);
 private:
  float alpha_;
};

class AsgdWorker : public IWorker {
 public:
  AsgdWorker(const char id[128], int nranks, int rank);
  AsgdWorker();                                                // asgd-worker.h
  ~AsgdWorker();
  void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  bool Synchronize(int num_worker_samples);
  bool IsAsync() const { return true; }
  void Stop();
 private:
  CuVector<BaseFloat> w_prev_;
  CtrlClient* ctrl_;
};
// ASGD server; momentum >= 0 turns it into the MASGD server (per-worker momentum buffers, no alpha)
class AsgdServer : public IServer {
 public:
  AsgdServer(const char id[128], int nranks, float alpha = 1.0f, int sync_period = 1000, float masgd_momentum = -1.0f)
      : IServer(id, nranks), alpha_(alpha), sync_period_(sync_period), momentum_(masgd_momentum) {}
  void InitParam(const std::vector<std::pair<BaseFloat*, int>>& params);
  void Run();
  void Update(int worker_rank, int synchronized_count);
 private:
  void SendModel(int worker_rank);
  float alpha_;
  int sync_period_;
  float momentum_;
  std::vector<CuVector<BaseFloat>> diffs_;       // MASGD: one per worker
};

}  // namespace kaldi
#endif
