// Test infrastructure (oracle/): runs the reference's OWN synchronous workers -- BspWorker, BmufWorker, SodWorker, compiled from
// src/aslp-parallel/*.cc where they lie, against oracle/stub/mpi.h -- with N ranks as N threads, on a scripted sequence of
// synchronisations, and writes what every rank holds afterwards.  tests/test_cpu_oracle_pinning.py compares the restated formulas
// (oracle/aslp_oracle.py: bsp_sync, bmuf_sync, sod_optimize) with it; the NCCL workers are then tested against those.
//
// Second mode, the parameter-server workers and servers (EasgdWorker / EasgdServer, AsgdWorker / AsgdServer, MasgdServer): rank 0
// runs the server's Run(), ranks 1..N-1 a worker each; a scripted list of events says which worker trains (w += delta) and
// synchronises next, one at a time, so that the order in which the server sees them is fixed.
//
// usage: ref_worker_driver <bsp | bmuf | sod> <solver> <nranks> <bmuf_learn_rate> <bmuf_momentum> <in.bin> <out.bin>
//        ref_worker_driver async <easgd | asgd | masgd> <nranks> <alpha> <momentum> <sync_period> <in.bin> <out.bin>
// async in.bin : int32 ntensors, int32 size[ntensors], int32 nevents, float w0[total]; per event: int32 worker_rank, float delta[total]
// async out.bin: per event: float w_worker[total] after Synchronize; at the end: float w_server[total]
// in.bin  : int32 ntensors, int32 size[ntensors], int32 nsteps, float w0[total];
//           then per step, per rank: int32 frames, float delta[total]   (the rank's local training since the last sync: w += delta)
// out.bin : per step, per rank: int32 keep_going, float w[total] after Synchronize(frames)
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "aslp-parallel/bsp-worker.h"
#include "aslp-parallel/bmuf-worker.h"
#include "aslp-parallel/sod-worker.h"
#include "aslp-parallel/easgd-worker.h"
#include "aslp-parallel/easgd-server.h"
#include "aslp-parallel/asgd-worker.h"
#include "aslp-parallel/asgd-server.h"
#include "aslp-parallel/masgd-server.h"
#include <deque>

// ---- the N-threads-as-N-ranks MPI stand-in (declared in stub/mpi.h)
static int g_nranks = 1;
static thread_local int t_rank = 0;
static std::mutex g_mu;
static std::condition_variable g_cv;
static int g_arrived = 0, g_generation = 0;
static std::vector<void*> g_bufs;

int MPI_Init(int*, char***) { return 0; }
int MPI_Finalize() { return 0; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = t_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_nranks; return 0; }
template <typename T>
static void reduce_in_rank_order(int count) {
  std::vector<T> sum(static_cast<T*>(g_bufs[0]), static_cast<T*>(g_bufs[0]) + count);
  for (int r = 1; r < g_nranks; ++r) {
    const T* p = static_cast<const T*>(g_bufs[r]);
    for (int i = 0; i < count; ++i) sum[i] = sum[i] + p[i];
  }
  for (int r = 0; r < g_nranks; ++r) std::memcpy(g_bufs[r], sum.data(), sizeof(T) * count);
}
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE) std::abort();
  std::unique_lock<std::mutex> lk(g_mu);
  g_bufs[t_rank] = recvbuf;
  if (++g_arrived == g_nranks) {
    if (type == MPI_INT) reduce_in_rank_order<int>(count);
    else if (type == MPI_FLOAT) reduce_in_rank_order<float>(count);
    else if (type == MPI_DOUBLE) reduce_in_rank_order<double>(count);
    else std::abort();
    g_arrived = 0;
    ++g_generation;
    g_cv.notify_all();
  } else {
    const int gen = g_generation;
    g_cv.wait(lk, [&] { return g_generation != gen; });
  }
  return 0;
}
int MPI_Barrier(MPI_Comm) { int dummy = 0; return MPI_Allreduce(MPI_IN_PLACE, &dummy, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD); }

// ---- point to point
struct Msg { int src, tag; std::vector<char> data; };
static std::vector<std::deque<Msg> > g_box;           // one mailbox per destination rank, guarded by g_mu / g_cv
static size_t type_size(MPI_Datatype t) { return t == MPI_INT || t == MPI_FLOAT || t == MPI_UNSIGNED ? 4 : (t == MPI_CHAR ? 1 : 8); }
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm) {
  Msg m;
  m.src = t_rank; m.tag = tag;
  m.data.assign(static_cast<const char*>(buf), static_cast<const char*>(buf) + count * type_size(type));
  std::unique_lock<std::mutex> lk(g_mu);
  g_box[dest].push_back(std::move(m));
  g_cv.notify_all();
  return 0;
}
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm, MPI_Status* status) {
  std::unique_lock<std::mutex> lk(g_mu);
  std::deque<Msg>& box = g_box[t_rank];
  for (;;) {
    for (size_t i = 0; i < box.size(); ++i) {
      if ((source == MPI_ANY_SOURCE || box[i].src == source) && (tag == MPI_ANY_TAG || box[i].tag == tag)) {
        if (box[i].data.size() > count * type_size(type)) std::abort();      // MPI_ERR_TRUNCATE
        std::memcpy(buf, box[i].data.data(), box[i].data.size());
        if (status != nullptr) { status->MPI_SOURCE = box[i].src; status->MPI_TAG = box[i].tag; status->MPI_ERROR = 0; }
        box.erase(box.begin() + i);
        return 0;
      }
    }
    g_cv.wait(lk);
  }
}
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status) {
  MPI_Send(sendbuf, sendcount, sendtype, dest, sendtag, comm);
  return MPI_Recv(recvbuf, recvcount, recvtype, source, recvtag, comm, status);
}

using namespace kaldi;

static int async_main(int argc, char** argv) {
  if (argc != 9) { std::fprintf(stderr, "usage: %s async <easgd|asgd|masgd> <nranks> <alpha> <momentum> <sync_period> <in.bin> <out.bin>\n", argv[0]); return 2; }
  const std::string kind = argv[2];
  g_nranks = std::atoi(argv[3]);
  const float alpha = std::atof(argv[4]), momentum = std::atof(argv[5]);
  const int sync_period = std::atoi(argv[6]);
  g_bufs.assign(g_nranks, nullptr);
  g_box.assign(g_nranks, std::deque<Msg>());
  FILE* f = std::fopen(argv[7], "rb");
  if (f == nullptr) return 3;
  int ntensors = 0, nevents = 0;
  if (std::fread(&ntensors, 4, 1, f) != 1) return 3;
  std::vector<int> sizes(ntensors);
  if (std::fread(sizes.data(), 4, ntensors, f) != static_cast<size_t>(ntensors)) return 3;
  if (std::fread(&nevents, 4, 1, f) != 1) return 3;
  int total = 0;
  for (int s : sizes) total += s;
  std::vector<float> w0(total);
  if (std::fread(w0.data(), 4, total, f) != static_cast<size_t>(total)) return 3;
  std::vector<int> who(nevents);
  std::vector<float> delta(static_cast<size_t>(nevents) * total);
  for (int e = 0; e < nevents; ++e) {
    if (std::fread(&who[e], 4, 1, f) != 1) return 3;
    if (std::fread(&delta[static_cast<size_t>(e) * total], 4, total, f) != static_cast<size_t>(total)) return 3;
  }
  std::fclose(f);
  std::vector<float> out(static_cast<size_t>(nevents) * total), server_final(total);
  std::mutex turn_mu;
  std::condition_variable turn_cv;
  int turn = 0;                                          // index of the event whose worker may go

  auto make_params = [&](std::vector<float>& w) {
    std::vector<std::pair<BaseFloat*, int> > params;
    int off = 0;
    for (int s : sizes) { params.push_back(std::make_pair(w.data() + off, s)); off += s; }
    return params;
  };
  auto server_main = [&]() {
    t_rank = 0;
    std::vector<float> w(w0);
    IServer* server = nullptr;
    if (kind == "easgd") server = new EasgdServer(alpha);
    else if (kind == "asgd") server = new AsgdServer(alpha, sync_period);
    else server = new MasgdServer(sync_period, momentum);
    server->InitParam(make_params(w));
    server->Run();
    server_final = w;
    delete server;
  };
  auto worker_main = [&](int rank) {
    t_rank = rank;
    std::vector<float> w(w0);
    IWorker* worker = nullptr;
    if (kind == "easgd") worker = new EasgdWorker(alpha);
    else worker = new AsgdWorker();                     // the MASGD server is served by ASGD workers
    worker->InitParam(make_params(w));
    for (int e = 0; e < nevents; ++e) {
      if (who[e] != rank) continue;
      { std::unique_lock<std::mutex> lk(turn_mu); turn_cv.wait(lk, [&] { return turn == e; }); }
      const float* d = &delta[static_cast<size_t>(e) * total];
      for (int i = 0; i < total; ++i) w[i] += d[i];
      worker->Synchronize(1);
      std::memcpy(&out[static_cast<size_t>(e) * total], w.data(), sizeof(float) * total);
      { std::unique_lock<std::mutex> lk(turn_mu); ++turn; turn_cv.notify_all(); }
    }
    { std::unique_lock<std::mutex> lk(turn_mu); turn_cv.wait(lk, [&] { return turn == nevents; }); }
    worker->Stop();
    delete worker;
  };
  std::vector<std::thread> th;
  th.emplace_back(server_main);
  for (int r = 1; r < g_nranks; ++r) th.emplace_back(worker_main, r);
  for (auto& t : th) t.join();
  f = std::fopen(argv[8], "wb");
  if (f == nullptr) return 4;
  std::fwrite(out.data(), 4, out.size(), f);
  std::fwrite(server_final.data(), 4, total, f);
  std::fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "async") return async_main(argc, argv);
  if (argc != 8) { std::fprintf(stderr, "usage: %s <bsp|bmuf|sod> <solver> <nranks> <bmuf_lr> <bmuf_momentum> <in.bin> <out.bin>\n", argv[0]); return 2; }
  const std::string kind = argv[1], solver = argv[2];
  g_nranks = std::atoi(argv[3]);
  const float bmuf_lr = std::atof(argv[4]), bmuf_mom = std::atof(argv[5]);
  g_bufs.assign(g_nranks, nullptr);
  FILE* f = std::fopen(argv[6], "rb");
  if (f == nullptr) return 3;
  int ntensors = 0, nsteps = 0;
  if (std::fread(&ntensors, 4, 1, f) != 1) return 3;
  std::vector<int> sizes(ntensors);
  if (std::fread(sizes.data(), 4, ntensors, f) != static_cast<size_t>(ntensors)) return 3;
  if (std::fread(&nsteps, 4, 1, f) != 1) return 3;
  int total = 0;
  for (int s : sizes) total += s;
  std::vector<float> w0(total);
  if (std::fread(w0.data(), 4, total, f) != static_cast<size_t>(total)) return 3;
  std::vector<int> frames(static_cast<size_t>(nsteps) * g_nranks);
  std::vector<float> delta(static_cast<size_t>(nsteps) * g_nranks * total);
  for (int s = 0; s < nsteps; ++s)
    for (int r = 0; r < g_nranks; ++r) {
      if (std::fread(&frames[s * g_nranks + r], 4, 1, f) != 1) return 3;
      if (std::fread(&delta[(static_cast<size_t>(s) * g_nranks + r) * total], 4, total, f) != static_cast<size_t>(total)) return 3;
    }
  std::fclose(f);
  std::vector<int> keep(static_cast<size_t>(nsteps) * g_nranks);
  std::vector<float> out(static_cast<size_t>(nsteps) * g_nranks * total);
  OptimizerOption opt;                                   // the reference's defaults (optimizer.h:172-232), solver chosen by name
  opt.solver = solver;

  auto rank_main = [&](int rank) {
    t_rank = rank;
    std::vector<float> w(w0);                            // one flat buffer; the tensors are consecutive runs of it
    std::vector<std::pair<BaseFloat*, int> > params;
    int off = 0;
    for (int s : sizes) { params.push_back(std::make_pair(w.data() + off, s)); off += s; }
    IWorker* worker = nullptr;
    if (kind == "bsp") worker = new BspWorker();
    else if (kind == "bmuf") worker = new BmufWorker(bmuf_lr, bmuf_mom);
    else worker = new SodWorker(opt);
    worker->InitParam(params);
    for (int s = 0; s < nsteps; ++s) {
      const float* d = &delta[(static_cast<size_t>(s) * g_nranks + rank) * total];
      for (int i = 0; i < total; ++i) w[i] += d[i];
      keep[s * g_nranks + rank] = worker->Synchronize(frames[s * g_nranks + rank]) ? 1 : 0;
      std::memcpy(&out[(static_cast<size_t>(s) * g_nranks + rank) * total], w.data(), sizeof(float) * total);
    }
    delete worker;
  };
  std::vector<std::thread> th;
  for (int r = 0; r < g_nranks; ++r) th.emplace_back(rank_main, r);
  for (auto& t : th) t.join();

  f = std::fopen(argv[7], "wb");
  if (f == nullptr) return 4;
  for (int s = 0; s < nsteps; ++s)
    for (int r = 0; r < g_nranks; ++r) {
      std::fwrite(&keep[s * g_nranks + r], 4, 1, f);
      std::fwrite(&out[(static_cast<size_t>(s) * g_nranks + r) * total], 4, total, f);
    }
  std::fclose(f);
  return 0;
}
