#include "parallel-async.h"
#include <arpa/inet.h>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>

namespace kaldi {

int CtrlPort() {
  if (const char* p = std::getenv("ASLP_CTRL_PORT")) return std::atoi(p);
  if (const char* p = std::getenv("MASTER_PORT")) return std::atoi(p) + 1;
  return 29631;
}

static void WriteAll(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n > 0) {
    const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) { if (errno == EINTR) continue; KALDI_ERR << "control channel: send failed: " << std::strerror(errno); }
    p += k; n -= static_cast<size_t>(k);
  }
}
static bool ReadAll(int fd, void* buf, size_t n) {        // false on orderly shutdown before the first byte
  char* p = static_cast<char*>(buf);
  size_t got = 0;
  while (got < n) {
    const ssize_t k = ::recv(fd, p + got, n - got, 0);
    if (k == 0) { if (got == 0) return false; KALDI_ERR << "control channel: peer closed in the middle of a message"; }
    if (k < 0) { if (errno == EINTR) continue; KALDI_ERR << "control channel: recv failed: " << std::strerror(errno); }
    got += static_cast<size_t>(k);
  }
  return true;
}

CtrlServer::CtrlServer(int port, int nworkers) : listen_fd_(-1) {
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  if (listen_fd_ < 0) KALDI_ERR << "control channel: socket() failed";
  int one = 1;
  ::setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
  addr.sin_port = htons(static_cast<uint16_t>(port));
  if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) KALDI_ERR << "control channel: cannot bind 127.0.0.1:" << port << ": " << std::strerror(errno);
  if (::listen(listen_fd_, nworkers + 4) != 0) KALDI_ERR << "control channel: listen failed";
  for (int i = 0; i < nworkers; ++i) {
    const int fd = ::accept(listen_fd_, nullptr, nullptr);
    if (fd < 0) KALDI_ERR << "control channel: accept failed: " << std::strerror(errno);
    ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    int32_t rank = -1;
    if (!ReadAll(fd, &rank, sizeof(rank))) KALDI_ERR << "control channel: worker closed before saying its rank";
    fds_.push_back(fd);
    ranks_.push_back(rank);
  }
}
CtrlServer::~CtrlServer() {
  for (int fd : fds_) if (fd >= 0) ::close(fd);
  if (listen_fd_ >= 0) ::close(listen_fd_);
}
void CtrlServer::RecvAny(int* worker_rank, int* msg_type) {
  for (;;) {
    std::vector<pollfd> pf;
    std::vector<size_t> which;
    for (size_t i = 0; i < fds_.size(); ++i)
      if (fds_[i] >= 0) { pollfd p; p.fd = fds_[i]; p.events = POLLIN; p.revents = 0; pf.push_back(p); which.push_back(i); }
    if (pf.empty()) KALDI_ERR << "control channel: every worker is gone but the server still expects messages";
    const int rc = ::poll(pf.data(), pf.size(), -1);
    if (rc < 0) { if (errno == EINTR) continue; KALDI_ERR << "control channel: poll failed"; }
    for (size_t k = 0; k < pf.size(); ++k) {
      if (!(pf[k].revents & (POLLIN | POLLHUP))) continue;
      int32_t msg = 0;
      if (!ReadAll(pf[k].fd, &msg, sizeof(msg))) { ::close(pf[k].fd); fds_[which[k]] = -1; continue; }   // closed after kMsgFinished
      *worker_rank = ranks_[which[k]];
      *msg_type = msg;
      return;
    }
  }
}

CtrlClient::CtrlClient(int port, int rank) : fd_(-1) {
  sockaddr_in addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
  addr.sin_port = htons(static_cast<uint16_t>(port));
  for (int tries = 0;; ++tries) {
    fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd_ < 0) KALDI_ERR << "control channel: socket() failed";
    if (::connect(fd_, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) == 0) break;
    ::close(fd_);
    fd_ = -1;
    if (tries > 1200) KALDI_ERR << "control channel: cannot reach the server at 127.0.0.1:" << port;
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
  }
  int one = 1;
  ::setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  const int32_t r = rank;
  WriteAll(fd_, &r, sizeof(r));
}
CtrlClient::~CtrlClient() { if (fd_ >= 0) ::close(fd_); }
void CtrlClient::Send(int msg_type) { const int32_t m = msg_type; WriteAll(fd_, &m, sizeof(m)); }

// ---------------------------------------------------------------- server base
IServer::~IServer() {
  delete ctrl_;
  if (table_dev_ != nullptr) aslp_free(table_dev_);
}
void IServer::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  std::vector<aslp_tensor_ref_t> table(params.size());
  size_t off = 0;
  for (size_t i = 0; i < params.size(); ++i) {
    table[i].ptr = params[i].first;
    table[i].offset = off;
    table[i].n = static_cast<size_t>(params[i].second);
    off += (table[i].n + 3) / 4 * 4;
  }
  ntensors_ = static_cast<int>(params.size());
  total_ = off;
  if (table_dev_ != nullptr) aslp_free(table_dev_);
  ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&table_dev_), sizeof(aslp_tensor_ref_t) * (table.size() + 1)));
  ASLP_OK(aslp_memcpy_h2d(CuStream(), table_dev_, table.data(), sizeof(aslp_tensor_ref_t) * table.size()));
  CuSync();
  server_arena_.Resize(static_cast<int32>(total_), kSetZero);
  worker_arena_.Resize(static_cast<int32>(total_), kSetZero);
  if (ctrl_ == nullptr) ctrl_ = new CtrlServer(CtrlPort(), NumNodes() - 1);
}

// ---------------------------------------------------------------- EASGD
EasgdWorker::EasgdWorker(const char id[128], int nranks, int rank, float alpha) : IWorker(id, nranks, rank), alpha_(alpha), ctrl_(nullptr) {
  KALDI_ASSERT(rank != 0);
}
EasgdWorker::EasgdWorker(float alpha) : IWorker(WorkerBootstrap()), alpha_(alpha), ctrl_(nullptr) { KALDI_ASSERT(Rank() != 0); }
EasgdWorker::~EasgdWorker() { delete ctrl_; }
void EasgdWorker::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  IWorker::InitParam(params);
  server_arena_.Resize(static_cast<int32>(total_), kSetZero);
  if (ctrl_ == nullptr) ctrl_ = new CtrlClient(CtrlPort(), Rank());
}
// easgd-worker.cc:37-66: signal, exchange models, x_worker = (1 - alpha) x_worker + alpha x_server
bool EasgdWorker::Synchronize(int num_worker_samples) {
  (void)num_worker_samples;
  ctrl_->Send(kMsgSynchronize);
  ASLP_OK(aslp_sync_pack(CuStream(), arena_.Data(), table_dev_, ntensors_, 1.0f));
  ASLP_OK(aslp_comm_sendrecv_f32(comm_, CuStream(), arena_.Data(), server_arena_.Data(), total_, 0));
  const int32 ld = static_cast<int32>((total_ + 3) / 4 * 4);
  ASLP_OK(aslp_axpby(CuStream(), arena_.Data(), ld, server_arena_.Data(), ld, 1, static_cast<int32>(total_), alpha_, 1.0f - alpha_));
  ASLP_OK(aslp_sync_unpack(CuStream(), arena_.Data(), table_dev_, ntensors_));
  return true;
}
void EasgdWorker::Stop() {
  ctrl_->Send(kMsgFinished);
  KALDI_LOG << "Worker " << Rank() << " finished";
}

void EasgdServer::Run() {             // easgd-server.cc:37-59
  int num_running_workers = NumNodes() - 1;
  while (num_running_workers > 0) {
    int msg_type = 0, worker_rank = 0;
    ctrl_->RecvAny(&worker_rank, &msg_type);
    KALDI_VLOG(2) << "Worker rank " << worker_rank << " Msg " << msg_type;
    switch (msg_type) {
      case kMsgFinished:
        num_running_workers--;
        KALDI_LOG << "Worker " << worker_rank << " Finished ";
        break;
      case kMsgSynchronize:
        Update(worker_rank);
        break;
      default:
        KALDI_WARN << "Unknown mpi msg type " << msg_type;
    }
  }
  CuSync();
  KALDI_LOG << "All worker finished";
}
// easgd-server.cc:61-85: x_server = (1 - alpha) x_server + alpha x_worker, both sides using the other's pre-update model
void EasgdServer::Update(int worker_rank) {
  ASLP_OK(aslp_sync_pack(CuStream(), server_arena_.Data(), table_dev_, ntensors_, 1.0f));
  ASLP_OK(aslp_comm_sendrecv_f32(comm_, CuStream(), server_arena_.Data(), worker_arena_.Data(), total_, worker_rank));
  const int32 ld = static_cast<int32>((total_ + 3) / 4 * 4);
  ASLP_OK(aslp_axpby(CuStream(), server_arena_.Data(), ld, worker_arena_.Data(), ld, 1, static_cast<int32>(total_), alpha_, 1.0f - alpha_));
  ASLP_OK(aslp_sync_unpack(CuStream(), server_arena_.Data(), table_dev_, ntensors_));
}

// ---------------------------------------------------------------- ASGD / MASGD
AsgdWorker::AsgdWorker(const char id[128], int nranks, int rank) : IWorker(id, nranks, rank), ctrl_(nullptr) { KALDI_ASSERT(rank != 0); }
AsgdWorker::AsgdWorker() : IWorker(WorkerBootstrap()), ctrl_(nullptr) { KALDI_ASSERT(Rank() != 0); }
AsgdWorker::~AsgdWorker() { delete ctrl_; }
void AsgdWorker::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  IWorker::InitParam(params);
  w_prev_.Resize(static_cast<int32>(total_), kSetZero);
  ASLP_OK(aslp_sync_pack(CuStream(), w_prev_.Data(), table_dev_, ntensors_, 1.0f));      // prev = initial model (asgd-worker.cc:20-24)
  if (ctrl_ == nullptr) ctrl_ = new CtrlClient(CtrlPort(), Rank());
}
// asgd-worker.cc:34-67: send the accumulated delta w - w_prev, receive the server's model, restart from it
bool AsgdWorker::Synchronize(int num_worker_samples) {
  (void)num_worker_samples;
  ctrl_->Send(kMsgSynchronize);
  ASLP_OK(aslp_sync_pack_diff(CuStream(), arena_.Data(), table_dev_, ntensors_, w_prev_.Data(), 1.0f));
  ASLP_OK(aslp_comm_send_f32(comm_, CuStream(), arena_.Data(), total_, 0));
  ASLP_OK(aslp_comm_recv_f32(comm_, CuStream(), w_prev_.Data(), total_, 0));     // may wait for the server's periodic barrier
  ASLP_OK(aslp_sync_unpack(CuStream(), w_prev_.Data(), table_dev_, ntensors_));
  return true;
}
void AsgdWorker::Stop() {
  ctrl_->Send(kMsgFinished);
  KALDI_LOG << "Worker " << Rank() << " finished";
}

void AsgdServer::InitParam(const std::vector<std::pair<BaseFloat*, int>>& params) {
  IServer::InitParam(params);
  if (momentum_ >= 0.0f) {
    diffs_.resize(NumNodes() - 1);
    for (CuVector<BaseFloat>& d : diffs_) d.Resize(static_cast<int32>(total_), kSetZero);
  }
}
void AsgdServer::SendModel(int worker_rank) {
  ASLP_OK(aslp_comm_send_f32(comm_, CuStream(), server_arena_.Data(), total_, worker_rank));
}
void AsgdServer::Run() {             // asgd-server.cc:35-77 (masgd-server.cc:41-91 is the same loop)
  int num_running_workers = NumNodes() - 1;
  int synchronized_count = 0;
  std::vector<int> waited_worker;
  ASLP_OK(aslp_sync_pack(CuStream(), server_arena_.Data(), table_dev_, ntensors_, 1.0f));
  while (num_running_workers > 0) {
    int msg_type = 0, worker_rank = 0;
    ctrl_->RecvAny(&worker_rank, &msg_type);
    KALDI_VLOG(2) << "Worker rank " << worker_rank << " Msg " << msg_type;
    switch (msg_type) {
      case kMsgFinished:
        num_running_workers--;
        KALDI_LOG << "Worker " << worker_rank << " Finished ";
        break;
      case kMsgSynchronize:
        ++synchronized_count;
        if (sync_period_ > 0 && synchronized_count >= sync_period_) waited_worker.push_back(worker_rank);
        Update(worker_rank, synchronized_count);
        break;
      default:
        KALDI_WARN << "Unknown mpi msg type " << msg_type;
    }
    // periodic barrier: once sync_period updates have been applied, workers are held until all of them have reported
    if (sync_period_ > 0 && synchronized_count >= sync_period_ && static_cast<int>(waited_worker.size()) == num_running_workers &&
        num_running_workers != 0) {
      for (size_t j = 0; j < waited_worker.size(); ++j) {
        KALDI_LOG << "Worker " << waited_worker[j] << " synchronized!";
        SendModel(waited_worker[j]);
      }
      synchronized_count = synchronized_count - sync_period_;
      waited_worker.clear();
    }
  }
  ASLP_OK(aslp_sync_unpack(CuStream(), server_arena_.Data(), table_dev_, ntensors_));
  CuSync();
  KALDI_LOG << "All worker finished";
}
// asgd-server.cc:80-102: w += alpha * delta;  masgd-server.cc:107-137 (LMASGD): d_k = momentum d_k + delta, w += d_k
void AsgdServer::Update(int worker_rank, int synchronized_count) {
  ASLP_OK(aslp_comm_recv_f32(comm_, CuStream(), worker_arena_.Data(), total_, worker_rank));
  const int32 ld = static_cast<int32>((total_ + 3) / 4 * 4), n = static_cast<int32>(total_);
  if (momentum_ >= 0.0f) {
    CuVector<BaseFloat>& d = diffs_[worker_rank - 1];
    ASLP_OK(aslp_axpby(CuStream(), d.Data(), ld, worker_arena_.Data(), ld, 1, n, 1.0f, momentum_));
    ASLP_OK(aslp_axpby(CuStream(), server_arena_.Data(), ld, d.Data(), ld, 1, n, 1.0f, 1.0f));
  } else {
    ASLP_OK(aslp_axpby(CuStream(), server_arena_.Data(), ld, worker_arena_.Data(), ld, 1, n, alpha_, 1.0f));
  }
  ASLP_OK(aslp_sync_unpack(CuStream(), server_arena_.Data(), table_dev_, ntensors_));     // the model the server main writes at the end
  if (synchronized_count < sync_period_ || sync_period_ <= 0) SendModel(worker_rank);
}

}  // namespace kaldi
