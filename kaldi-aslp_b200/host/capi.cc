// capi.cc -- C handle API of include/aslp_nnet_c.h over the C++ host layer.  Exceptions stop here.
#include <cstring>
#include "../../include/aslp_nnet_c.h"
#include "cu-workspace.h"
#include "nnet-loss.h"
#include "nnet-nnet.h"
#include "nnet-train-step.h"
#include "parallel.h"
#include "parallel-async.h"

using namespace kaldi;
using namespace kaldi::aslp_nnet;

static thread_local std::string g_err;

#define CAPI_BEGIN try {
#define CAPI_END                                                   \
  return 0;                                                        \
  } catch (const std::exception& e) { g_err = e.what(); return 1; } \
  catch (...) { g_err = "unknown exception"; return 1; }

static Nnet* N(aslp_nnet_t n) { if (n == nullptr) KALDI_ERR << "null Nnet handle"; return static_cast<Nnet*>(n); }

static void CopyOut(const std::string& s, char* buf, size_t bytes) {
  if (buf == nullptr || bytes == 0) return;
  const size_t k = std::min(bytes - 1, s.size());
  memcpy(buf, s.data(), k);
  buf[k] = '\0';
}

// reusable device staging for host inputs / outputs
static CuMatrix<BaseFloat> g_in, g_out, g_diff, g_indiff, g_loss_diff;
static XentTrainStep g_xent_step;
long long aslp_nnet_step_replays(void) { return g_xent_step.Replays(); }

extern "C" {

const char* aslp_nnet_last_error(void) { return g_err.c_str(); }
int aslp_nnet_select_device(int dev) { CAPI_BEGIN CuSelectDevice(dev); CAPI_END }
int aslp_nnet_srand(int seed) { srand(seed); return 0; }
int aslp_nnet_set_gemm_precision(int precision) { SetGemmPrecision(precision); return 0; }
int aslp_nnet_device_sync(void) { CAPI_BEGIN CuSync(); CAPI_END }
unsigned long long aslp_nnet_launch_count(void) { return aslp_launch_count(); }
long long aslp_nnet_step_replays(void);
static void* g_events[16] = {nullptr};
int aslp_nnet_event_record(int slot) {
  CAPI_BEGIN
  if (slot < 0 || slot >= 16) KALDI_ERR << "event slot out of range";
  ASLP_OK(aslp_event_record(CuStream(), &g_events[slot]));
  CAPI_END
}
int aslp_nnet_event_elapsed_ms(int a, int b, float* ms) {
  CAPI_BEGIN
  if (a < 0 || a >= 16 || b < 0 || b >= 16 || !g_events[a] || !g_events[b]) KALDI_ERR << "event slot not recorded";
  ASLP_OK(aslp_event_elapsed_ms(g_events[a], g_events[b], ms));
  CAPI_END
}
int aslp_nnet_pinned_alloc(void** host_ptr, size_t bytes) { CAPI_BEGIN CuStream(); ASLP_OK(aslp_malloc_host(host_ptr, bytes)); CAPI_END }
int aslp_nnet_pinned_free(void* host_ptr) { CAPI_BEGIN ASLP_OK(aslp_free_host(host_ptr)); CAPI_END }

int aslp_nnet_init(const char* proto_file, aslp_nnet_t* out) { CAPI_BEGIN Nnet* n = new Nnet(); try { n->Init(proto_file); } catch (...) { delete n; throw; } *out = n; CAPI_END }
int aslp_nnet_read(const char* model_file, aslp_nnet_t* out) { CAPI_BEGIN Nnet* n = new Nnet(); try { n->Read(model_file); } catch (...) { delete n; throw; } *out = n; CAPI_END }
int aslp_nnet_write(aslp_nnet_t n, const char* file, int binary) { CAPI_BEGIN N(n)->Write(file, binary != 0); CAPI_END }
int aslp_nnet_destroy(aslp_nnet_t n) { CAPI_BEGIN delete static_cast<Nnet*>(n); CAPI_END }
int aslp_nnet_input_dim(aslp_nnet_t n, int* dim) { CAPI_BEGIN *dim = N(n)->InputDim(); CAPI_END }
int aslp_nnet_output_dim(aslp_nnet_t n, int* dim) { CAPI_BEGIN *dim = N(n)->OutputDim(); CAPI_END }
int aslp_nnet_num_components(aslp_nnet_t n, int* count) { CAPI_BEGIN *count = N(n)->NumComponents(); CAPI_END }
int aslp_nnet_num_params(aslp_nnet_t n, int* count) { CAPI_BEGIN *count = N(n)->NumParams(); CAPI_END }
int aslp_nnet_info(aslp_nnet_t n, char* buf, size_t bytes) { CAPI_BEGIN CopyOut(N(n)->Info(), buf, bytes); CAPI_END }
int aslp_nnet_get_params(aslp_nnet_t n, float* host_out, int count) {
  CAPI_BEGIN
  Vector<BaseFloat> p;
  N(n)->GetParams(&p);
  if (p.Dim() != count) KALDI_ERR << "GetParams: the net has " << p.Dim() << " parameters, buffer holds " << count;
  memcpy(host_out, p.Data(), sizeof(float) * count);
  CAPI_END
}
int aslp_nnet_set_train_options(aslp_nnet_t n, float lr, float mmt, float l2, float l1) {
  CAPI_BEGIN
  NnetTrainOptions o; o.learn_rate = lr; o.momentum = mmt; o.l2_penalty = l2; o.l1_penalty = l1;
  N(n)->SetTrainOptions(o);
  CAPI_END
}
int aslp_nnet_set_seq_lengths(aslp_nnet_t n, const int* lengths, int count) { CAPI_BEGIN N(n)->SetSeqLengths(std::vector<int32>(lengths, lengths + count)); CAPI_END }
int aslp_nnet_reset_streams(aslp_nnet_t n, const int* flags, int count) { CAPI_BEGIN N(n)->ResetLstmStreams(std::vector<int32>(flags, flags + count)); CAPI_END }
int aslp_nnet_set_chunk_size(aslp_nnet_t n, int chunk_size) { CAPI_BEGIN N(n)->SetChunkSize(chunk_size); CAPI_END }

int aslp_nnet_propagate(aslp_nnet_t n, const float* host_in, int rows, int cols, float* host_out) {
  CAPI_BEGIN
  g_in.Resize(rows, cols, kUndefined);
  g_in.CopyFromHost(host_in, cols);
  N(n)->Propagate(g_in, &g_out);
  if (host_out != nullptr) g_out.CopyToHost(host_out, g_out.NumCols());
  CuSync();
  CAPI_END
}
int aslp_nnet_feedforward(aslp_nnet_t n, const float* host_in, int rows, int cols, float* host_out) {
  CAPI_BEGIN
  g_in.Resize(rows, cols, kUndefined);
  g_in.CopyFromHost(host_in, cols);
  N(n)->Feedforward(g_in, &g_out);
  if (host_out != nullptr) g_out.CopyToHost(host_out, g_out.NumCols());
  CuSync();
  CAPI_END
}
int aslp_nnet_backpropagate(aslp_nnet_t n, const float* host_out_diff, int rows, int cols, float* host_in_diff) {
  CAPI_BEGIN
  g_diff.Resize(rows, cols, kUndefined);
  g_diff.CopyFromHost(host_out_diff, cols);
  N(n)->Backpropagate(g_diff, &g_indiff);
  if (host_in_diff != nullptr) g_indiff.CopyToHost(host_in_diff, g_indiff.NumCols());
  CuSync();
  CAPI_END
}
static void CopyBuf(const CuMatrix<BaseFloat>& m, float* host_out, int rows, int cols) {
  if (m.NumRows() != rows || m.NumCols() != cols) KALDI_ERR << "buffer is " << m.NumRows() << " x " << m.NumCols() << ", asked for " << rows << " x " << cols;
  m.CopyToHost(host_out, cols);
  CuSync();
}
int aslp_nnet_component_output(aslp_nnet_t n, int c, float* host_out, int rows, int cols) { CAPI_BEGIN CopyBuf(N(n)->PropagateBuffer().at(c), host_out, rows, cols); CAPI_END }
int aslp_nnet_component_out_diff(aslp_nnet_t n, int c, float* host_out, int rows, int cols) { CAPI_BEGIN CopyBuf(N(n)->BackpropagateBuffer().at(c), host_out, rows, cols); CAPI_END }

// the frame-level objectives share one handle type: the handle is a LossItf* (Xent | Mse | MultiTaskLoss)
int aslp_xent_create(aslp_xent_t* out) { CAPI_BEGIN *out = static_cast<LossItf*>(new Xent()); CAPI_END }
int aslp_loss_create(const char* objective, aslp_xent_t* out) {
  CAPI_BEGIN
  const std::string obj(objective);
  if (obj == "xent") *out = static_cast<LossItf*>(new Xent());
  else if (obj == "mse") *out = static_cast<LossItf*>(new Mse());
  else if (obj.compare(0, 9, "multitask") == 0) { MultiTaskLoss* m = new MultiTaskLoss(); m->InitFromString(obj); *out = static_cast<LossItf*>(m); }
  else KALDI_ERR << "Unknown objective function code : " << obj;
  CAPI_END
}
int aslp_xent_destroy(aslp_xent_t x) { CAPI_BEGIN delete static_cast<LossItf*>(x); CAPI_END }
int aslp_xent_report(aslp_xent_t x, char* buf, size_t bytes, double stats5[5]) {
  CAPI_BEGIN
  LossItf* l = static_cast<LossItf*>(x);
  CopyOut(l->Report(), buf, bytes);
  if (stats5 != nullptr) {
    Xent* xe = dynamic_cast<Xent*>(l);
    stats5[0] = l->AvgLoss(); stats5[1] = xe ? xe->Frames() : 0; stats5[2] = xe ? xe->Correct() : 0; stats5[3] = 0; stats5[4] = 0;
  }
  CAPI_END
}
int aslp_warpctc_create(aslp_warpctc_t* out) { CAPI_BEGIN *out = new WarpCtc(); CAPI_END }
int aslp_warpctc_destroy(aslp_warpctc_t c) { CAPI_BEGIN delete static_cast<WarpCtc*>(c); CAPI_END }
int aslp_warpctc_rejected(aslp_warpctc_t c, int* n) { CAPI_BEGIN *n = static_cast<WarpCtc*>(c)->NumRejected(); CAPI_END }
int aslp_warpctc_report(aslp_warpctc_t c, char* buf, size_t bytes) { CAPI_BEGIN CopyOut(static_cast<WarpCtc*>(c)->Report(), buf, bytes); CAPI_END }

static const CuMatrixBase<BaseFloat>& StageFeatures(const float* features, int on_device, int rows, int cols, CuSubMatrix<BaseFloat>* view) {
  if (on_device) { *view = CuSubMatrix<BaseFloat>(const_cast<float*>(features), rows, cols, (cols + 3) / 4 * 4); return *view; }
  g_in.Resize(rows, cols, kUndefined);
  g_in.CopyFromHost(features, cols);
  return g_in;
}

int aslp_train_step_xent(aslp_nnet_t n, aslp_xent_t x, const float* features, int on_device, int rows, int cols, const int* targets,
                         const float* frame_mask) {
  CAPI_BEGIN
  CuSubMatrix<BaseFloat> view(nullptr, 0, 0, 0);
  const CuMatrixBase<BaseFloat>& in = StageFeatures(features, on_device, rows, cols, &view);
  Posterior post(rows);
  Vector<BaseFloat> fw(rows);
  for (int r = 0; r < rows; ++r) { post[r].push_back(std::make_pair(targets[r], 1.0f)); fw(r) = frame_mask ? frame_mask[r] : 1.0f; }
  Xent* xent = dynamic_cast<Xent*>(static_cast<LossItf*>(x));
  if (xent != nullptr) {                                // the trainers' step (nnet-train-step.h): recorded and replayed when it repeats
    g_xent_step.Run(N(n), xent, in, fw, post);
  } else {
    N(n)->Propagate(in, &g_out);
    static_cast<LossItf*>(x)->Eval(fw, g_out, post, &g_loss_diff);
    N(n)->Backpropagate(g_loss_diff, NULL);
  }
  CAPI_END
}

int aslp_train_step_ctc(aslp_nnet_t n, aslp_warpctc_t c, const float* features, int on_device, int rows, int cols, const int* frame_num_utt,
                        int nseq, const int* flat_labels, const int* label_lengths, float norm_learn_rate, int with_error_rate, float* costs_out) {
  CAPI_BEGIN
  Nnet* net = N(n);
  WarpCtc* ctc = static_cast<WarpCtc*>(c);
  std::vector<int32> lens(frame_num_utt, frame_num_utt + nseq);
  std::vector<std::vector<int32>> labels(nseq);
  std::vector<std::string> keys(nseq);
  int off = 0, valid_frames = 0;
  for (int s = 0; s < nseq; ++s) {
    labels[s].assign(flat_labels + off, flat_labels + off + label_lengths[s]);
    off += label_lengths[s];
    keys[s] = "utt" + ToString(s);
    valid_frames += lens[s];
  }
  net->SetSeqLengths(lens);                                           // aslp-nnet-train-warp-ctc-streams.cc:175
  if (norm_learn_rate > 0.0f) {                                       // :177-178 learn_rate = norm_lr / valid frames
    NnetTrainOptions o = net->GetTrainOptions();
    o.learn_rate = norm_learn_rate / valid_frames;
    net->SetTrainOptions(o);
  }
  CuSubMatrix<BaseFloat> view(nullptr, 0, 0, 0);
  const CuMatrixBase<BaseFloat>& in = StageFeatures(features, on_device, rows, cols, &view);
  net->Propagate(in, &g_out);                                         // :182
  ctc->Eval(keys, lens, g_out, labels, &g_loss_diff);                 // :187
  if (with_error_rate) ctc->ErrorRate(lens, g_out, labels);           // :190
  net->Backpropagate(g_loss_diff, NULL);                              // :194
  if (costs_out != nullptr) memcpy(costs_out, ctc->LastCosts().data(), sizeof(float) * nseq);
  CAPI_END
}

int aslp_eesenctc_create(aslp_eesenctc_t* out) { CAPI_BEGIN *out = new Ctc(); CAPI_END }
int aslp_eesenctc_destroy(aslp_eesenctc_t c) { CAPI_BEGIN delete static_cast<Ctc*>(c); CAPI_END }
int aslp_eesenctc_report(aslp_eesenctc_t c, char* buf, size_t bytes) { CAPI_BEGIN CopyOut(static_cast<Ctc*>(c)->Report(), buf, bytes); CAPI_END }

int aslp_train_step_ctc_eesen(aslp_nnet_t n, aslp_eesenctc_t c, const float* features, int on_device, int rows, int cols,
                              const int* frame_num_utt, int nseq, const int* flat_labels, const int* label_lengths, float norm_learn_rate,
                              int with_error_rate, float* obj_out) {
  CAPI_BEGIN
  Nnet* net = N(n);
  Ctc* ctc = static_cast<Ctc*>(c);
  std::vector<int32> lens(frame_num_utt, frame_num_utt + nseq);
  std::vector<std::vector<int32>> labels(nseq);
  std::vector<std::string> keys(nseq);
  int off = 0, valid_frames = 0;
  for (int s = 0; s < nseq; ++s) {
    labels[s].assign(flat_labels + off, flat_labels + off + label_lengths[s]);
    off += label_lengths[s];
    keys[s] = "utt" + ToString(s);
    valid_frames += lens[s];
  }
  net->SetSeqLengths(lens);
  if (norm_learn_rate > 0.0f) {
    NnetTrainOptions o = net->GetTrainOptions();
    o.learn_rate = norm_learn_rate / valid_frames;
    net->SetTrainOptions(o);
  }
  CuSubMatrix<BaseFloat> view(nullptr, 0, 0, 0);
  const CuMatrixBase<BaseFloat>& in = StageFeatures(features, on_device, rows, cols, &view);
  net->Propagate(in, &g_out);
  ctc->EvalParallel(keys, lens, g_out, labels, &g_loss_diff);
  if (with_error_rate) ctc->ErrorRateMSeq(lens, g_out, labels);
  net->Backpropagate(g_loss_diff, NULL);
  if (obj_out != nullptr) memcpy(obj_out, ctc->LastObj().data(), sizeof(float) * nseq);
  CAPI_END
}

int aslp_nnet_upload(const float* host, int rows, int cols, float** device_out, int* stride_out) {
  CAPI_BEGIN
  const int stride = (cols + 3) / 4 * 4;
  void* p = nullptr;
  CuStream();
  ASLP_OK(aslp_malloc(&p, sizeof(float) * (static_cast<size_t>(rows) * stride + 4)));
  ASLP_OK(aslp_memset(CuStream(), p, 0, sizeof(float) * (static_cast<size_t>(rows) * stride + 4)));
  ASLP_OK(aslp_memcpy2d_h2d(CuStream(), p, sizeof(float) * stride, host, sizeof(float) * cols, sizeof(float) * cols, rows));
  CuSync();
  *device_out = static_cast<float*>(p);
  if (stride_out) *stride_out = stride;
  CAPI_END
}
int aslp_nnet_free_device(float* p) { CAPI_BEGIN ASLP_OK(aslp_free(p)); CAPI_END }

int aslp_worker_create(const char* type, const char nccl_id[128], int nranks, int rank, float bmuf_momentum, float bmuf_learn_rate,
                       const char* sod_solver, aslp_worker_t* out) {
  CAPI_BEGIN
  const std::string t(type);
  IWorker* w = nullptr;
  if (t == "bsp") w = new BspWorker(nccl_id, nranks, rank);
  else if (t == "bmuf") w = new BmufWorker(nccl_id, nranks, rank, bmuf_momentum, bmuf_learn_rate);
  else if (t == "sod") { OptimizerOption o; if (sod_solver && *sod_solver) o.solver = sod_solver; w = new SodWorker(nccl_id, nranks, rank, o); }
  else if (t == "easgd") w = new EasgdWorker(nccl_id, nranks, rank, bmuf_learn_rate);      // bmuf_learn_rate carries alpha
  else if (t == "asgd" || t == "masgd") w = new AsgdWorker(nccl_id, nranks, rank);       // the MASGD worker is the ASGD worker
  else KALDI_ERR << "Unsupported worker type: " << t << " (bsp | bmuf | sod | easgd | asgd | masgd)";
  *out = w;
  CAPI_END
}
int aslp_worker_init_param(aslp_worker_t w, aslp_nnet_t n) {
  CAPI_BEGIN
  std::vector<std::pair<BaseFloat*, int>> params;
  N(n)->GetGpuParams(&params);
  static_cast<IWorker*>(w)->InitParam(params);
  CAPI_END
}
int aslp_worker_init_param_by_component(aslp_worker_t w, aslp_nnet_t n) { CAPI_BEGIN static_cast<IWorker*>(w)->InitParam(N(n)); CAPI_END }
int aslp_worker_can_overlap(aslp_worker_t w, int* yes) { CAPI_BEGIN *yes = static_cast<IWorker*>(w)->CanOverlap() ? 1 : 0; CAPI_END }
int aslp_worker_begin_synchronize(aslp_worker_t w, int num_frames) { CAPI_BEGIN static_cast<IWorker*>(w)->BeginSynchronize(num_frames); CAPI_END }
int aslp_worker_end_synchronize(aslp_worker_t w, int* keep_going) { CAPI_BEGIN const bool k = static_cast<IWorker*>(w)->EndSynchronize(); if (keep_going) *keep_going = k ? 1 : 0; CAPI_END }
int aslp_worker_synchronize(aslp_worker_t w, int num_frames, int* keep_going) { CAPI_BEGIN const bool k = static_cast<IWorker*>(w)->Synchronize(num_frames); if (keep_going) *keep_going = k ? 1 : 0; CAPI_END }
int aslp_worker_stop(aslp_worker_t w) { CAPI_BEGIN static_cast<IWorker*>(w)->Stop(); CAPI_END }
int aslp_worker_reduce_acc_stat(aslp_worker_t w, aslp_nnet_t n) {
  CAPI_BEGIN
  std::vector<double*> acc;
  std::vector<std::pair<double*, int>> data;
  N(n)->GetAccStats(&acc, &data);
  static_cast<IWorker*>(w)->ReduceAccStat(acc, data);
  CAPI_END
}
int aslp_worker_destroy(aslp_worker_t w) { CAPI_BEGIN delete static_cast<IWorker*>(w); CAPI_END }

int aslp_server_create(const char* type, const char nccl_id[128], int nranks, float alpha, int sync_period, float momentum, aslp_server_t* out) {
  CAPI_BEGIN
  const std::string t(type);
  IServer* sv = nullptr;
  if (t == "easgd") sv = new EasgdServer(nccl_id, nranks, alpha);
  else if (t == "asgd") sv = new AsgdServer(nccl_id, nranks, alpha, sync_period, -1.0f);
  else if (t == "masgd") sv = new AsgdServer(nccl_id, nranks, 1.0f, sync_period, momentum);
  else KALDI_ERR << "Unsupported server type: " << t << " (easgd | asgd | masgd)";
  *out = sv;
  CAPI_END
}
int aslp_server_init_param(aslp_server_t sv, aslp_nnet_t n) {
  CAPI_BEGIN
  std::vector<std::pair<BaseFloat*, int>> params;
  N(n)->GetGpuParams(&params);
  static_cast<IServer*>(sv)->InitParam(params);
  CAPI_END
}
int aslp_server_run(aslp_server_t sv) { CAPI_BEGIN static_cast<IServer*>(sv)->Run(); CAPI_END }
int aslp_server_destroy(aslp_server_t sv) { CAPI_BEGIN delete static_cast<IServer*>(sv); CAPI_END }

/* control channel alone (no GPU): the MPI_Recv(MPI_ANY_SOURCE) replacement of the async servers */
int aslp_ctrl_server_create(int port, int nworkers, void** out) { CAPI_BEGIN *out = new CtrlServer(port, nworkers); CAPI_END }
int aslp_ctrl_server_recv_any(void* sv, int* worker_rank, int* msg_type) { CAPI_BEGIN static_cast<CtrlServer*>(sv)->RecvAny(worker_rank, msg_type); CAPI_END }
int aslp_ctrl_server_destroy(void* sv) { CAPI_BEGIN delete static_cast<CtrlServer*>(sv); CAPI_END }
int aslp_ctrl_client_create(int port, int rank, void** out) { CAPI_BEGIN *out = new CtrlClient(port, rank); CAPI_END }
int aslp_ctrl_client_send(void* c, int msg_type) { CAPI_BEGIN static_cast<CtrlClient*>(c)->Send(msg_type); CAPI_END }
int aslp_ctrl_client_destroy(void* c) { CAPI_BEGIN delete static_cast<CtrlClient*>(c); CAPI_END }

}  // extern "C"
