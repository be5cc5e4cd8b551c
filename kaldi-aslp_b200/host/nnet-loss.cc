#include "nnet-loss.h"
#include <algorithm>
#include "../../include/ctc.h"
#include "cu-workspace.h"

namespace kaldi {
namespace aslp_nnet {

// ------------------------------------------------------------------ Xent
Xent::Xent() : stats_dev_(nullptr), frames_(0), correct_(0), loss_(0), entropy_(0), likelyhood_(0), frames_progress_(0),
               stage_rows_(0), stage_pad_(0), stage_next_(0) {
  for (double& b : base_) b = 0;
  for (int i = 0; i < kStageSlots; ++i) { stage_host_[i] = nullptr; stage_event_[i] = nullptr; stage_cap_[i] = 0; }
}
Xent::~Xent() {
  if (stats_dev_ != nullptr) aslp_free(stats_dev_);
  for (int i = 0; i < kStageSlots; ++i) {
    if (stage_event_[i] != nullptr) { aslp_event_sync(stage_event_[i]); aslp_event_destroy(stage_event_[i]); }
    if (stage_host_[i] != nullptr) aslp_free_host(stage_host_[i]);
  }
}

static void EnsureStats(double** p) {
  if (*p == nullptr) {
    ASLP_OK(aslp_malloc(reinterpret_cast<void**>(p), 5 * sizeof(double)));
    ASLP_OK(aslp_memset(CuStream(), *p, 0, 5 * sizeof(double)));
  }
}

void Xent::Fetch() {
  if (stats_dev_ == nullptr) return;
  double h[5];
  ASLP_OK(aslp_memcpy_d2h(CuStream(), h, stats_dev_, sizeof(h)));
  CuSync();
  loss_ = h[0]; entropy_ = h[1]; likelyhood_ = h[2]; correct_ = h[3]; frames_ = h[4];
}

void Xent::Progress(double num_frames) {
  static const int32 progress_step = 3600 * 100;   // 1h of frames (nnet-loss.cc:137)
  frames_progress_ += num_frames;
  if (frames_progress_ > progress_step) {
    Fetch();
    const double df = frames_ - base_[4];
    KALDI_LOG << "ProgressLoss[last " << static_cast<int>(df / 100 / 3600) << "h of " << static_cast<int>(frames_ / 100 / 3600) << "h]: "
              << (likelyhood_ - base_[2]) / df << " (Likelyhood) " << ((loss_ - base_[0]) - (entropy_ - base_[1])) / df << " (Xent)";
    base_[0] = loss_; base_[1] = entropy_; base_[2] = likelyhood_; base_[3] = correct_; base_[4] = frames_;
    frames_progress_ = 0;
  }
}

void Xent::Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& targets, CuMatrix<BaseFloat>* diff) {
  KALDI_ASSERT(net_out.NumCols() == targets.NumCols() && net_out.NumRows() == targets.NumRows());
  KALDI_ASSERT(net_out.NumRows() == frame_weights.Dim());
  KALDI_ASSERT(KALDI_ISFINITE(frame_weights.Sum()));
  EnsureStats(&stats_dev_);
  frame_w_dev_ = frame_weights;
  diff->Resize(net_out.NumRows(), net_out.NumCols(), kUndefined);
  ASLP_OK(aslp_xent_dense(CuStream(), diff->Data(), diff->Stride(), net_out.Data(), net_out.Stride(), targets.Data(), targets.Stride(),
                          net_out.NumRows(), net_out.NumCols(), frame_w_dev_.Data(), stats_dev_));
  Progress(frame_weights.Sum());
}

void Xent::Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& post, CuMatrix<BaseFloat>* diff) {
  const int32 num_frames = net_out.NumRows(), num_pdf = net_out.NumCols();
  KALDI_ASSERT(num_frames == static_cast<int32>(post.size()));
  KALDI_ASSERT(num_frames == frame_weights.Dim());
  bool sparse = true;
  for (const auto& f : post) if (f.size() > 1) { sparse = false; break; }
  if (!sparse) {                                   // PosteriorToMatrix (nnet-utils.h) then the dense path
    PosteriorToMatrix(post, num_pdf, &tgt_mat_);
    Eval(frame_weights, net_out, tgt_mat_, diff);
    return;
  }
  const double nf = StageSparse(frame_weights, post, num_pdf);
  KALDI_ASSERT(nf >= 0.0);
  LaunchSparse(net_out, diff);
  Progress(nf);
}

double Xent::StageSparse(const VectorBase<BaseFloat>& frame_weights, const Posterior& post, int32 num_pdf) {
  const int32 num_frames = static_cast<int32>(post.size());
  KALDI_ASSERT(num_frames == frame_weights.Dim());
  for (const auto& f : post) if (f.size() > 1) return -1.0;
  EnsureStats(&stats_dev_);
  const int32 pad = (num_frames + 3) / 4 * 4;
  // a ring of page-locked slots: the upload is asynchronous, so the slot of step n is written again only after the copy
  // of step n - kStageSlots has finished (its event) -- which also bounds how far the host runs ahead of the device
  const int slot = static_cast<int>(stage_next_++ % kStageSlots);
  if (stage_event_[slot] != nullptr) ASLP_OK(aslp_event_sync(stage_event_[slot]));
  if (stage_cap_[slot] < static_cast<size_t>(3) * pad) {
    if (stage_host_[slot] != nullptr) ASLP_OK(aslp_free_host(stage_host_[slot]));
    void* p = nullptr;
    ASLP_OK(aslp_malloc_host(&p, sizeof(float) * 3 * pad));
    stage_host_[slot] = static_cast<float*>(p);
    stage_cap_[slot] = static_cast<size_t>(3) * pad;
  }
  float* h = stage_host_[slot];
  int32* hi = reinterpret_cast<int32*>(h);
  double nf = 0;
  for (int32 t = 0; t < num_frames; ++t) {
    int32 id = 0;
    float w = 0.f;
    if (!post[t].empty()) {
      if (post[t][0].first >= num_pdf) KALDI_ERR << "Posterior has pdf-id " << post[t][0].first << " but the net has " << num_pdf << " outputs";
      id = post[t][0].first;
      w = post[t][0].second;
    }
    hi[t] = id; h[pad + t] = w; h[2 * pad + t] = frame_weights(t);
    nf += frame_weights(t) * w;
  }
  for (int32 t = num_frames; t < pad; ++t) { hi[t] = 0; h[pad + t] = 0.f; h[2 * pad + t] = 0.f; }
  if (stage_dev_.Dim() != 3 * pad) stage_dev_.Resize(3 * pad, kUndefined);
  stage_rows_ = num_frames; stage_pad_ = pad;
  ASLP_OK(aslp_memcpy_h2d(CuStream(), stage_dev_.Data(), h, sizeof(float) * 3 * pad));
  ASLP_OK(aslp_event_record(CuStream(), &stage_event_[slot]));
  return nf;
}

void Xent::LaunchSparse(const CuMatrixBase<BaseFloat>& net_out, CuMatrix<BaseFloat>* diff) {
  const int32 num_frames = net_out.NumRows(), num_pdf = net_out.NumCols();
  KALDI_ASSERT(num_frames == stage_rows_);
  diff->Resize(num_frames, num_pdf, kUndefined);
  ASLP_OK(aslp_xent_sparse(CuStream(), diff->Data(), diff->Stride(), net_out.Data(), net_out.Stride(), num_frames, num_pdf,
                           reinterpret_cast<const int32*>(stage_dev_.Data()), stage_dev_.Data() + stage_pad_, stage_dev_.Data() + 2 * stage_pad_, stats_dev_));
}

BaseFloat Xent::AvgLoss() { Fetch(); return (loss_ - entropy_) / frames_; }

std::string Xent::Report() {         // line formats the schedulers grep (nnet-loss.cc:175-200)
  Fetch();
  std::ostringstream oss;
  if (0 == frames_) {
    oss << "AvgLoss: " << 0.0 << " (Xent), " << "Likelyhood: " << 0.0 << " " << "Frame: " << 0 << std::endl;
    oss << "FRAME_ACCURACY >> " << 0.0 << "% <<" << std::endl;
  } else {
    oss << "AvgLoss: " << (loss_ - entropy_) / frames_ << " (Xent), " << "Likelyhood: " << (likelyhood_) / frames_ << " " << "Frame: " << frames_ << std::endl;
    if (correct_ >= 0.0) oss << "FRAME_ACCURACY >> " << 100.0 * correct_ / frames_ << "% <<" << std::endl;
  }
  return oss.str();
}

void PosteriorToMatrix(const Posterior& post, int32 num_cols, CuMatrix<BaseFloat>* mat) {
  const int32 num_rows = static_cast<int32>(post.size());
  Matrix<BaseFloat> m(num_rows, num_cols);
  for (int32 t = 0; t < num_rows; ++t)
    for (const auto& pr : post[t]) {
      if (pr.first >= num_cols) KALDI_ERR << "Out-of-bound Posterior element with index " << pr.first << ", higher than number of columns " << num_cols;
      m(t, pr.first) += pr.second;
    }
  *mat = m;
}

// ------------------------------------------------------------------ Mse
Mse::Mse() : loss_dev_(nullptr), frames_(0), frames_progress_(0), loss_base_(0), num_tgt_(0) {}
Mse::~Mse() { if (loss_dev_ != nullptr) aslp_free(loss_dev_); }

double Mse::Loss() {
  if (loss_dev_ == nullptr) return 0.0;
  double h = 0;
  ASLP_OK(aslp_memcpy_d2h(CuStream(), &h, loss_dev_, sizeof(h)));
  CuSync();
  return h;
}

void Mse::Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const CuMatrixBase<BaseFloat>& target, CuMatrix<BaseFloat>* diff) {
  KALDI_ASSERT(net_out.NumCols() == target.NumCols());
  KALDI_ASSERT(net_out.NumRows() == target.NumRows());
  KALDI_ASSERT(net_out.NumRows() == frame_weights.Dim());
  KALDI_ASSERT(KALDI_ISFINITE(frame_weights.Sum()));
  const int32 num_frames = frame_weights.Sum();           // truncated to an integer, as in the reference (nnet-loss.cc:219)
  KALDI_ASSERT(num_frames >= 0.0);
  if (loss_dev_ == nullptr) {
    ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&loss_dev_), sizeof(double)));
    ASLP_OK(aslp_memset(CuStream(), loss_dev_, 0, sizeof(double)));
  }
  frame_weights_ = frame_weights;
  num_tgt_ = net_out.NumCols();
  diff->Resize(net_out.NumRows(), net_out.NumCols(), kUndefined);
  ASLP_OK(aslp_mse(CuStream(), diff->Data(), diff->Stride(), net_out.Data(), net_out.Stride(), target.Data(), target.Stride(), net_out.NumRows(),
                   net_out.NumCols(), frame_weights_.Data(), loss_dev_));
  frames_ += num_frames;
  static const int32 progress_step = 3600 * 100;          // 1h
  frames_progress_ += num_frames;
  if (frames_progress_ > progress_step) {
    const double loss = Loss();
    KALDI_LOG << "ProgressLoss[last " << static_cast<int>(frames_progress_ / 100 / 3600) << "h of " << static_cast<int>(frames_ / 100 / 3600) << "h]: "
              << (loss - loss_base_) / frames_progress_ << " (Mse)";
    loss_base_ = loss;
    frames_progress_ = 0;
  }
}

void Mse::Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& post, CuMatrix<BaseFloat>* diff) {
  KALDI_ASSERT(net_out.NumRows() == static_cast<int32>(post.size()));
  PosteriorToMatrix(post, net_out.NumCols(), &tgt_mat_);
  Eval(frame_weights, net_out, tgt_mat_, diff);
}

BaseFloat Mse::AvgLoss() { return Loss() / frames_; }

std::string Mse::Report() {
  const double loss = Loss();
  const BaseFloat root_mean_square = sqrt(loss / frames_ / num_tgt_);
  std::ostringstream oss;
  oss << "AvgLoss: " << loss / frames_ << " (Mse), " << "[RMS " << root_mean_square << ", frames " << frames_ << "]" << std::endl;
  return oss.str();
}

// ------------------------------------------------------------------ MultiTaskLoss
MultiTaskLoss::~MultiTaskLoss() { for (LossItf* l : loss_vec_) delete l; }

void MultiTaskLoss::InitFromString(const std::string& s) {
  std::vector<std::string> v;
  std::string cur;
  for (char ch : s) { if (ch == ',' || ch == ':') { v.push_back(cur); cur.clear(); } else cur.push_back(ch); }
  v.push_back(cur);
  KALDI_ASSERT((v.size() - 1) % 3 == 0);     // triplets
  KALDI_ASSERT(v[0] == "multitask");
  for (size_t i = 1; i + 2 < v.size(); i += 3) {
    if (v[i] == "xent") loss_vec_.push_back(new Xent());
    else if (v[i] == "mse") loss_vec_.push_back(new Mse());
    else KALDI_ERR << "Unknown objective function code : " << v[i];
    char* end = nullptr;
    const long dim = strtol(v[i + 1].c_str(), &end, 10);
    if (end == v[i + 1].c_str() || *end != '\0') KALDI_ERR << "Cannot convert 'dim' " << v[i + 1] << " to integer!";
    loss_dim_.push_back(static_cast<int32>(dim));
    const double w = strtod(v[i + 2].c_str(), &end);
    if (end == v[i + 2].c_str() || *end != '\0') KALDI_ERR << "Cannot convert 'weight' " << v[i + 2] << " to integer!";
    KALDI_ASSERT(w >= 0.0);
    loss_weights_.push_back(static_cast<BaseFloat>(w));
  }
  loss_dim_offset_.assign(loss_dim_.size() + 1, 0);
  for (size_t i = 1; i <= loss_dim_.size(); i++) loss_dim_offset_[i] = loss_dim_offset_[i - 1] + loss_dim_[i - 1];
  KALDI_ASSERT(loss_vec_.size() > 0);
}

void MultiTaskLoss::Eval(const VectorBase<BaseFloat>& frame_weights, const CuMatrixBase<BaseFloat>& net_out, const Posterior& post, CuMatrix<BaseFloat>* diff) {
  const int32 num_frames = net_out.NumRows(), num_output = net_out.NumCols();
  KALDI_ASSERT(num_frames == static_cast<int32>(post.size()));
  KALDI_ASSERT(num_output == loss_dim_offset_.back());
  PosteriorToMatrix(post, num_output, &tgt_mat_);
  diff->Resize(num_frames, num_output);
  CuMatrix<BaseFloat> diff_aux;
  for (size_t i = 0; i < loss_vec_.size(); i++) {
    loss_vec_[i]->Eval(frame_weights, net_out.ColRange(loss_dim_offset_[i], loss_dim_[i]), tgt_mat_.ColRange(loss_dim_offset_[i], loss_dim_[i]), &diff_aux);
    diff_aux.Scale(loss_weights_[i]);
    diff->ColRange(loss_dim_offset_[i], loss_dim_[i]).CopyFromMat(diff_aux);
  }
}

std::string MultiTaskLoss::Report() {
  const BaseFloat overall_loss = AvgLoss();
  std::ostringstream oss;
  oss << "MultiTaskLoss, with " << loss_vec_.size() << " parallel loss functions." << std::endl;
  for (size_t i = 0; i < loss_vec_.size(); i++) oss << "Loss " << i + 1 << ", " << loss_vec_[i]->Report() << std::endl;
  oss << "Loss (OVERALL), " << "AvgLoss: " << overall_loss << " (MultiTaskLoss), " << "weights [ ";
  for (BaseFloat w : loss_weights_) oss << w << " ";
  oss << "], values [ ";
  for (LossItf* l : loss_vec_) oss << l->AvgLoss() << " ";
  oss << "]" << std::endl;
  return oss.str();
}

BaseFloat MultiTaskLoss::AvgLoss() {
  BaseFloat ans(0.0);
  for (size_t i = 0; i < loss_vec_.size(); i++) {
    BaseFloat val = loss_weights_[i] * loss_vec_[i]->AvgLoss();
    if (!KALDI_ISFINITE(val)) {
      KALDI_WARN << "Loss " << i + 1 << ", has bad objective function value '" << val << "', using 0.0 instead.";
      val = 0.0;
    }
    ans += val;
  }
  return ans;
}

// ------------------------------------------------------------------ WarpCtc
WarpCtc::WarpCtc()
    : frames_(0), sequences_num_(0), ref_num_(0), error_num_(0), frames_progress_(0), ref_num_progress_(0), error_num_progress_(0),
      sequences_progress_(0), obj_progress_(0.0), report_step_(100), obj_(0), loss_sum_(0), loss_square_sum_(0), loss_sum_bak_(0),
      loss_square_sum_bak_(0), normal_num_(0), stat_period_(500) {}

void WarpCtc::SetUseGpu(bool use_gpu) {
  if (!use_gpu) KALDI_ERR << "WarpCtc: this build has no CPU path (--use-gpu=no is the reference oracle, not the product)";
}

void WarpCtc::Eval(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out,
                   const std::vector<std::vector<int32>>& labels, CuMatrix<BaseFloat>* diff) {
  KALDI_ASSERT(diff != NULL);
  // the C API wants activations with row stride == alphabet size (include/ctc.h)
  if (net_out.Stride() != net_out.NumCols())
    KALDI_ERR << "WarpCtc: the output dimension (" << net_out.NumCols() << ") must be a multiple of 4 so that rows are dense, as on the reference CPU path";
  diff->Resize(net_out.NumRows(), net_out.NumCols(), kSetZero);      // rows past an utterance's end stay zero
  const int minibatch = static_cast<int>(frame_num_utt.size());
  KALDI_ASSERT(minibatch > 0 && net_out.NumRows() % minibatch == 0);
  std::vector<int> flat_labels, label_lengths;
  for (int i = 0; i < minibatch; i++) {
    const std::vector<int>& l = labels[i];
    for (size_t j = 0; j < l.size(); j++) KALDI_ASSERT(l[j] < net_out.NumCols());
    flat_labels.insert(flat_labels.end(), l.begin(), l.end());
    label_lengths.push_back(static_cast<int>(l.size()));
  }
  if (flat_labels.empty()) flat_labels.push_back(0);
  const int alphabet_size = net_out.NumCols();
  costs_.assign(minibatch, 0);
  ctcComputeInfo info;
  info.loc = CTC_GPU;
  info.stream = reinterpret_cast<CUstream>(CuStream());
  size_t bytes = 0;
  if (get_workspace_size(label_lengths.data(), frame_num_utt.data(), alphabet_size, minibatch, info, &bytes) != CTC_STATUS_SUCCESS)
    KALDI_ERR << "Error in get_workspace_size";
  void* ws = CuWorkspace(bytes);               // persistent workspace: no per-minibatch malloc/free
  const ctcStatus_t rc = compute_ctc_loss(net_out.Data(), diff->Data(), flat_labels.data(), label_lengths.data(), frame_num_utt.data(),
                                          alphabet_size, minibatch, costs_.data(), ws, info);
  if (rc != CTC_STATUS_SUCCESS) KALDI_ERR << "Error: compute_ctc_loss, stat = " << ctcGetStatusString(rc);
  StatAndAverageLossCheck(utt, frame_num_utt, costs_, diff);          // WARP_CTC_GRAD_CHECK == WARP_CTC_AVG_LOSS_CHECK (warp-ctc.h:25)
  ASLP_OK(aslp_clamp(CuStream(), diff->Data(), diff->Stride(), diff->NumRows(), diff->NumCols(), -1.0f, 1.0f));
  if (sequences_progress_ >= report_step_) {
    KALDI_LOG << "Progress " << sequences_num_ << " sequences (" << frames_ / (100.0 * 3600) << "Hr):"
              << " Obj(log[Pzx]) = " << obj_progress_ / sequences_progress_ << " Obj(frame) = " << obj_progress_ / frames_progress_
              << " TokenAcc = " << 100.0 * (1.0 - error_num_progress_ / ref_num_progress_) << " %";
    sequences_progress_ = 0; frames_progress_ = 0; obj_progress_ = 0.0; error_num_progress_ = 0; ref_num_progress_ = 0;
  }
}

// running mean / sigma of the per-frame loss over a sliding window of stat_period_ utterances; an utterance whose
// loss is non-finite, outside 6 sigma, or outside (0, 3000) has its diff rows zeroed (warp-ctc.cc:288-365)
void WarpCtc::StatAndAverageLossCheck(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt,
                                      const std::vector<float>& pzx_host, CuMatrix<BaseFloat>* diff) {
  const int32 num_sequence = static_cast<int32>(frame_num_utt.size());
  for (int s = 0; s < num_sequence; s++) {
    const double loss_per_frame = pzx_host[s] / frame_num_utt[s];
    if (normal_num_ < stat_period_ / 2) {
      normal_num_++;
      loss_sum_ += loss_per_frame; loss_sum_bak_ += loss_per_frame;
      loss_square_sum_ += loss_per_frame * loss_per_frame; loss_square_sum_bak_ += loss_per_frame * loss_per_frame;
      obj_ += pzx_host[s]; obj_progress_ += pzx_host[s];
    } else {
      const double mean = loss_sum_ / normal_num_, sigma = sqrt(loss_square_sum_ / normal_num_);
      if (KALDI_ISFINITE(pzx_host[s]) && (loss_per_frame >= (mean - 6 * sigma) && loss_per_frame <= (mean + 6 * sigma)) &&
          (pzx_host[s] > 0 && pzx_host[s] < 3000)) {
        normal_num_++;
        loss_sum_ += loss_per_frame;
        loss_square_sum_ += loss_per_frame * loss_per_frame;
        obj_ += pzx_host[s]; obj_progress_ += pzx_host[s];
        if (normal_num_ == stat_period_) {
          loss_sum_ -= loss_sum_bak_; loss_square_sum_ -= loss_square_sum_bak_;
          loss_sum_bak_ = loss_sum_; loss_square_sum_bak_ = loss_square_sum_;
          normal_num_ = stat_period_ / 2;
        }
      } else {
        KALDI_WARN << "Sequences " << utt[s] << " obj is abnormal(sum " << pzx_host[s] << " per_frame " << loss_per_frame << " mean "
                   << loss_sum_ / normal_num_ << " sigma " << loss_square_sum_ / normal_num_ << "), drop it's diff and stat";
        ++rejected_num_;
        // rows t*num_sequence + s, t < frames: one strided zero fill instead of a host loop of per-row SetZero
        ASLP_OK(aslp_memset2d(CuStream(), diff->Data() + static_cast<size_t>(s) * diff->Stride(), sizeof(float) * diff->Stride() * num_sequence, 0,
                              sizeof(float) * diff->NumCols(), frame_num_utt[s]));
      }
    }
    frames_ += frame_num_utt[s];
    frames_progress_ += frame_num_utt[s];
  }
  const double grad_sum = diff->Sum();
  if (!KALDI_ISFINITE(grad_sum)) {
    KALDI_WARN << "DIFF FINITE: nan or inf ocurred in the diff, ignore";
    diff->SetZero();
  }
  sequences_progress_ += num_sequence;
  sequences_num_ += num_sequence;
}

int32 LevenshteinEditDistance(const std::vector<int32>& ref, const std::vector<int32>& hyp, int32* ins, int32* del, int32* sub) {
  // standard DP with operation counts (util/edit-distance-inl.h); ties prefer substitution, then deletion, then insertion
  struct Cell { int32 ins, del, sub, total; };
  const size_t R = ref.size(), Hn = hyp.size();
  std::vector<Cell> prev(R + 1), cur(R + 1);
  for (size_t i = 0; i <= R; ++i) prev[i] = Cell{0, static_cast<int32>(i), 0, static_cast<int32>(i)};
  for (size_t j = 1; j <= Hn; ++j) {
    cur[0] = Cell{prev[0].ins + 1, prev[0].del, prev[0].sub, prev[0].total + 1};
    for (size_t i = 1; i <= R; ++i) {
      const int32 s = prev[i - 1].total + (ref[i - 1] == hyp[j - 1] ? 0 : 1);
      const int32 d = cur[i - 1].total + 1, in = prev[i].total + 1;
      if (s <= d && s <= in) { cur[i] = prev[i - 1]; if (ref[i - 1] != hyp[j - 1]) cur[i].sub++; cur[i].total = s; }
      else if (d <= in) { cur[i] = cur[i - 1]; cur[i].del++; cur[i].total = d; }
      else { cur[i] = prev[i]; cur[i].ins++; cur[i].total = in; }
    }
    prev.swap(cur);
  }
  if (ins) *ins = prev[R].ins;
  if (del) *del = prev[R].del;
  if (sub) *sub = prev[R].sub;
  return prev[R].total;
}

// greedy decode: per-frame arg-max -> collapse repeats -> drop blank (0) -> edit distance (warp-ctc.cc:487-526)
void WarpCtc::ErrorRate(const std::vector<int>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out, std::vector<std::vector<int>>& label) {
  const int32 rows = net_out.NumRows();
  int32* idx_dev = static_cast<int32*>(CuWorkspace(sizeof(int32) * (rows + 4)));
  ASLP_OK(aslp_row_argmax(CuStream(), idx_dev, net_out.Data(), net_out.Stride(), rows, net_out.NumCols()));
  std::vector<int32> data(rows);
  ASLP_OK(aslp_memcpy_d2h(CuStream(), data.data(), idx_dev, sizeof(int32) * rows));
  CuSync();
  const int32 num_seq = static_cast<int32>(frame_num_utt.size());
  for (int32 s = 0; s < num_seq; s++) {
    const int32 num_frame = frame_num_utt[s];
    std::vector<int32> hyp;
    int32 last = -1;
    for (int32 f = 0; f < num_frame; f++) {
      const int32 v = data[f * num_seq + s];
      if (f == 0 || v != last) { if (v != 0) hyp.push_back(v); }
      last = v;
    }
    int32 ins, del, sub;
    const int32 err = LevenshteinEditDistance(label[s], hyp, &ins, &del, &sub);
    error_num_ += err; ref_num_ += static_cast<int32>(label[s].size());
    error_num_progress_ += err; ref_num_progress_ += static_cast<int32>(label[s].size());
  }
}

std::string WarpCtc::Report() {
  std::ostringstream oss;
  oss << " Obj(log[Pzx]) = " << obj_ / sequences_num_ << " Obj(frame) = " << obj_ / frames_ << " TOKEN_ACCURACY >> "
      << 100.0 * (1.0 - error_num_ / ref_num_) << " % <<";
  return oss.str();
}


// ------------------------------------------------------------------ Eesen CTC
Ctc::Ctc()
    : frames_(0), sequences_num_(0), ref_num_(0), error_num_(0), frames_progress_(0), ref_num_progress_(0), error_num_progress_(0),
      sequences_progress_(0), obj_progress_(0.0), report_step_(100), obj_(0), loss_sum_(0), loss_square_sum_(0), loss_sum_bak_(0),
      loss_square_sum_bak_(0), normal_num_(0), stat_period_(100) {}

void Ctc::EvalParallel(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out,
                       std::vector<std::vector<int32>>& label, CuMatrix<BaseFloat>* diff) {
  KALDI_ASSERT(diff != NULL);
  diff->Resize(net_out.NumRows(), net_out.NumCols(), kUndefined);      // the kernel writes every element
  const int32 num_sequence = static_cast<int32>(frame_num_utt.size());
  const int32 num_frames = net_out.NumRows();
  KALDI_ASSERT(num_sequence > 0 && num_frames % num_sequence == 0);
  const int32 num_frames_per_sequence = num_frames / num_sequence;
  int32 max_label_len = 0;
  for (int32 s = 0; s < num_sequence; s++) max_label_len = std::max<int32>(max_label_len, static_cast<int32>(label[s].size()));
  // label expansion with blanks, padded with -1 (ctc-loss.cc:133-150)
  const int32 exp_len_labels = 2 * max_label_len + 1;
  std::vector<int32> label_expand(static_cast<size_t>(num_sequence) * exp_len_labels, -1);
  for (int32 s = 0; s < num_sequence; s++) {
    const std::vector<int32>& l = label[s];
    for (size_t j = 0; j < l.size(); j++) {
      if (l[j] >= net_out.NumCols()) KALDI_ERR << "label gt outdim " << l[j] << " " << net_out.NumCols();
      label_expand[s * exp_len_labels + 2 * j] = 0;
      label_expand[s * exp_len_labels + 2 * j + 1] = l[j];
    }
    label_expand[s * exp_len_labels + 2 * l.size()] = 0;
  }
  labels_dev_ = label_expand;
  seq_len_dev_ = frame_num_utt;
  pzx_dev_.Resize(num_sequence, kUndefined);
  const size_t wsb = aslp_ctc_eesen_workspace_bytes(num_frames_per_sequence, num_sequence, net_out.NumCols(), exp_len_labels);
  ASLP_OK(aslp_ctc_eesen(CuStream(), diff->Data(), diff->Stride(), net_out.Data(), net_out.Stride(), num_frames_per_sequence, num_sequence,
                         net_out.NumCols(), labels_dev_.Data(), exp_len_labels, seq_len_dev_.Data(), pzx_dev_.Data(), CuWorkspace(wsb), wsb));
  Vector<float> pzx_host;
  pzx_dev_.CopyToVec(&pzx_host);
  pzx_.resize(num_sequence);
  for (int32 s = 0; s < num_sequence; s++) pzx_[s] = -pzx_host(s);                      // pzx.Scale(-1): the objective
  StatAndAverageLossCheck(utt, frame_num_utt, pzx_, diff);
  ASLP_OK(aslp_clamp(CuStream(), diff->Data(), diff->Stride(), diff->NumRows(), diff->NumCols(), -1.0f, 1.0f));
  if (sequences_progress_ >= report_step_) {
    KALDI_LOG << "Progress " << sequences_num_ << " sequences (" << frames_ / (100.0 * 3600) << "Hr):"
              << " Obj(log[Pzx]) = " << obj_progress_ / sequences_progress_ << " Obj(frame) = " << obj_progress_ / frames_progress_
              << " TokenAcc = " << 100.0 * (1.0 - error_num_progress_ / ref_num_progress_) << " %";
    sequences_progress_ = 0; frames_progress_ = 0; obj_progress_ = 0.0; error_num_progress_ = 0; ref_num_progress_ = 0;
  }
}

void Ctc::Eval(const CuMatrixBase<BaseFloat>& net_out, const std::vector<int32>& label, CuMatrix<BaseFloat>* diff) {
  std::vector<std::string> utt(1, "utt");
  std::vector<int32> frames(1, net_out.NumRows());
  std::vector<std::vector<int32>> labels(1, label);
  EvalParallel(utt, frames, net_out, labels, diff);
}

// ctc-loss.cc:229-296: unlike the warp-ctc variant the warm-up half of the window only takes finite losses in (0, 3000)
void Ctc::StatAndAverageLossCheck(const std::vector<std::string>& utt, const std::vector<int32>& frame_num_utt,
                                  const std::vector<float>& pzx_host, CuMatrix<BaseFloat>* diff) {
  const int32 num_sequence = static_cast<int32>(frame_num_utt.size());
  for (int s = 0; s < num_sequence; s++) {
    const double loss_per_frame = pzx_host[s] / frame_num_utt[s];
    if (normal_num_ < stat_period_ / 2) {
      if (KALDI_ISFINITE(pzx_host[s]) && pzx_host[s] > 0 && pzx_host[s] < 3000) {
        normal_num_++;
        loss_sum_ += loss_per_frame; loss_sum_bak_ += loss_per_frame;
        loss_square_sum_ += loss_per_frame * loss_per_frame; loss_square_sum_bak_ += loss_per_frame * loss_per_frame;
        obj_ += pzx_host[s]; obj_progress_ += pzx_host[s];
      }
    } else {
      const double mean = loss_sum_ / normal_num_, sigma = sqrt(loss_square_sum_ / normal_num_);
      if (KALDI_ISFINITE(pzx_host[s]) && (loss_per_frame >= (mean - 6 * sigma) && loss_per_frame <= (mean + 6 * sigma)) &&
          (pzx_host[s] > 0 && pzx_host[s] < 3000)) {
        normal_num_++;
        loss_sum_ += loss_per_frame;
        loss_square_sum_ += loss_per_frame * loss_per_frame;
        obj_ += pzx_host[s]; obj_progress_ += pzx_host[s];
        if (normal_num_ == stat_period_) {
          loss_sum_ -= loss_sum_bak_; loss_square_sum_ -= loss_square_sum_bak_;
          loss_sum_bak_ = loss_sum_; loss_square_sum_bak_ = loss_square_sum_;
          normal_num_ = stat_period_ / 2;
        }
      } else {
        KALDI_WARN << "Sequences " << utt[s] << " obj is abnormal(sum " << pzx_host[s] << " per_frame " << loss_per_frame << " mean "
                   << loss_sum_ / normal_num_ << " sigma " << loss_square_sum_ / normal_num_ << "), drop it's diff and stat";
        ASLP_OK(aslp_memset2d(CuStream(), diff->Data() + static_cast<size_t>(s) * diff->Stride(), sizeof(float) * diff->Stride() * num_sequence, 0,
                              sizeof(float) * diff->NumCols(), frame_num_utt[s]));
      }
    }
    frames_ += frame_num_utt[s];
    frames_progress_ += frame_num_utt[s];
  }
  const double grad_sum = diff->Sum();
  if (!KALDI_ISFINITE(grad_sum)) {
    KALDI_WARN << "DIFF FINITE: nan or inf ocurred in the diff, ignore";
    diff->SetZero();
  }
  sequences_progress_ += num_sequence;
  sequences_num_ += num_sequence;
}

void Ctc::ErrorRateMSeq(const std::vector<int>& frame_num_utt, const CuMatrixBase<BaseFloat>& net_out, std::vector<std::vector<int>>& label) {
  const int32 rows = net_out.NumRows();
  int32* idx_dev = static_cast<int32*>(CuWorkspace(sizeof(int32) * (rows + 4)));
  ASLP_OK(aslp_row_argmax(CuStream(), idx_dev, net_out.Data(), net_out.Stride(), rows, net_out.NumCols()));
  std::vector<int32> data(rows);
  ASLP_OK(aslp_memcpy_d2h(CuStream(), data.data(), idx_dev, sizeof(int32) * rows));
  CuSync();
  const int32 num_seq = static_cast<int32>(frame_num_utt.size());
  for (int32 s = 0; s < num_seq; s++) {
    std::vector<int32> hyp;
    int32 last = -1;
    for (int32 f = 0; f < frame_num_utt[s]; f++) {          // greedy path, repetitions collapsed, blanks removed (ctc-loss.cc:385-424)
      const int32 v = data[f * num_seq + s];
      if (f == 0 || v != last) { if (v != 0) hyp.push_back(v); }
      last = v;
    }
    int32 ins, del, sub;
    const int32 err = LevenshteinEditDistance(label[s], hyp, &ins, &del, &sub);
    error_num_ += err; ref_num_ += static_cast<int32>(label[s].size());
    error_num_progress_ += err; ref_num_progress_ += static_cast<int32>(label[s].size());
  }
}

std::string Ctc::Report() {
  std::ostringstream oss;
  oss << " Obj(log[Pzx]) = " << obj_ / sequences_num_ << " Obj(frame) = " << obj_ / frames_ << " TOKEN_ACCURACY >> "
      << 100.0 * (1.0 - error_num_ / ref_num_) << " % <<";
  return oss.str();
}

}  // namespace aslp_nnet
}  // namespace kaldi
