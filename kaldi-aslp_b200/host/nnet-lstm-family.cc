#include "nnet-lstm-family.h"
#include "cu-workspace.h"
#include <memory>

namespace kaldi {
namespace aslp_nnet {

LstmFamily::LstmFamily(int32 input_dim, int32 output_dim, const Traits& tr)
    : UpdatableComponent(input_dim, output_dim), tr_(tr), ncell_(0), nrecur_(0), nstream_(0), chunk_size_(0),
      clip_gradient_(0.0f), d_(tr.ndirs), per_utt_reset_(false) {
  // Lstm: ncell = OutputDim; BLstm: OutputDim/2; projected: nrecur = OutputDim (1 dir) or OutputDim/2 (2 dirs)
  if (tr_.projected) nrecur_ = output_dim / tr_.ndirs;
  else ncell_ = output_dim / tr_.ndirs;
}

void LstmFamily::AllocCorr() {
  for (Dir& d : d_) {
    d.w_gifo_x_corr.Resize(4 * ncell_, input_dim_, kSetZero);
    d.w_gifo_r_corr.Resize(4 * ncell_, RecDim(), kSetZero);
    d.bias_corr.Resize(4 * ncell_, kSetZero);
    d.peep_i_corr.Resize(ncell_, kSetZero);
    d.peep_f_corr.Resize(ncell_, kSetZero);
    d.peep_o_corr.Resize(ncell_, kSetZero);
    if (tr_.projected) d.w_r_m_corr.Resize(nrecur_, ncell_, kSetZero);
  }
}

void LstmFamily::InitData(std::istream& is) {
  float param_scale = 0.02f;
  ProtoOptions po(tr_.has_celldim_token ? "(CellDim|ClipGradient|ParamScale)" : "(ClipGradient|ParamScale)");
  if (tr_.has_celldim_token) po.Int("<CellDim>", &ncell_);
  po.Float("<ClipGradient>", &clip_gradient_);
  po.Float("<ParamScale>", &param_scale);
  po.Parse(is);
  KALDI_ASSERT(ncell_ > 0);
  // the reference's order: all matrices (per direction x, r, [rm]), then the biases, then the peepholes
  for (Dir& d : d_) {
    d.w_gifo_x.Resize(4 * ncell_, input_dim_, kUndefined);
    d.w_gifo_r.Resize(4 * ncell_, RecDim(), kUndefined);
    InitMatParam(&d.w_gifo_x, param_scale);
    InitMatParam(&d.w_gifo_r, param_scale);
    if (tr_.projected) { d.w_r_m.Resize(nrecur_, ncell_, kUndefined); InitMatParam(&d.w_r_m, param_scale); }
  }
  for (Dir& d : d_) { d.bias.Resize(4 * ncell_, kUndefined); InitVecParam(&d.bias, param_scale); }
  for (Dir& d : d_) {
    d.peep_i.Resize(ncell_, kUndefined); d.peep_f.Resize(ncell_, kUndefined); d.peep_o.Resize(ncell_, kUndefined);
    InitVecParam(&d.peep_i, param_scale); InitVecParam(&d.peep_f, param_scale); InitVecParam(&d.peep_o, param_scale);
  }
  AllocCorr();
  KALDI_ASSERT(clip_gradient_ >= 0.0);
}

void LstmFamily::ReadData(std::istream& is, bool binary) {
  if (tr_.has_celldim_token) { ExpectToken(is, binary, "<CellDim>"); ReadBasicType(is, binary, &ncell_); }
  ExpectToken(is, binary, "<ClipGradient>");
  ReadBasicType(is, binary, &clip_gradient_);
  for (Dir& d : d_) {
    d.w_gifo_x.Read(is, binary);
    d.w_gifo_r.Read(is, binary);
    d.bias.Read(is, binary);
    d.peep_i.Read(is, binary);
    d.peep_f.Read(is, binary);
    d.peep_o.Read(is, binary);
    if (tr_.projected) d.w_r_m.Read(is, binary);
    KALDI_ASSERT(d.w_gifo_x.NumRows() == 4 * ncell_ && d.w_gifo_x.NumCols() == input_dim_);
    KALDI_ASSERT(d.w_gifo_r.NumRows() == 4 * ncell_ && d.w_gifo_r.NumCols() == RecDim());
  }
  AllocCorr();    // momentum buffers are not part of the file: zeroed on read
}

void LstmFamily::WriteData(std::ostream& os, bool binary) const {
  if (tr_.has_celldim_token) { WriteToken(os, binary, "<CellDim>"); WriteBasicType(os, binary, ncell_); }
  WriteToken(os, binary, "<ClipGradient>");
  WriteBasicType(os, binary, clip_gradient_);
  for (const Dir& d : d_) {
    d.w_gifo_x.Write(os, binary);
    d.w_gifo_r.Write(os, binary);
    d.bias.Write(os, binary);
    d.peep_i.Write(os, binary);
    d.peep_f.Write(os, binary);
    d.peep_o.Write(os, binary);
    if (tr_.projected) d.w_r_m.Write(os, binary);
  }
}

int32 LstmFamily::NumParams() const {
  const Dir& d = d_[0];
  int32 n = d.w_gifo_x.NumRows() * d.w_gifo_x.NumCols() + d.w_gifo_r.NumRows() * d.w_gifo_r.NumCols() + d.bias.Dim() + 3 * ncell_;
  if (tr_.projected) n += d.w_r_m.NumRows() * d.w_r_m.NumCols();
  return tr_.ndirs * n;
}

void LstmFamily::GetParams(Vector<BaseFloat>* wei_copy) const {
  wei_copy->Resize(NumParams());
  float* p = wei_copy->Data();
  auto vec = [&](const CuVector<BaseFloat>& v) { Vector<float> h; v.CopyToVec(&h); for (int32 i = 0; i < h.Dim(); ++i) *p++ = h(i); };
  auto mat = [&](const CuMatrix<BaseFloat>& m) { CopyRowsToVec(m, p); p += static_cast<size_t>(m.NumRows()) * m.NumCols(); };
  for (const Dir& d : d_) {
    mat(d.w_gifo_x); mat(d.w_gifo_r); vec(d.bias); vec(d.peep_i); vec(d.peep_f); vec(d.peep_o);
    if (tr_.projected) mat(d.w_r_m);
  }
}

void LstmFamily::GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
  params->clear();
  for (Dir& d : d_) {
    params->push_back(std::make_pair(d.w_gifo_x.Data(), d.w_gifo_x.NumRows() * d.w_gifo_x.Stride()));
    params->push_back(std::make_pair(d.w_gifo_r.Data(), d.w_gifo_r.NumRows() * d.w_gifo_r.Stride()));
    params->push_back(std::make_pair(d.bias.Data(), d.bias.Dim()));
    params->push_back(std::make_pair(d.peep_i.Data(), d.peep_i.Dim()));
    params->push_back(std::make_pair(d.peep_f.Data(), d.peep_f.Dim()));
    params->push_back(std::make_pair(d.peep_o.Data(), d.peep_o.Dim()));
    if (tr_.projected) params->push_back(std::make_pair(d.w_r_m.Data(), d.w_r_m.NumRows() * d.w_r_m.Stride()));
  }
}

std::string LstmFamily::Info() const {
  std::string s;
  const char* tag[2] = {"f_", "b_"};
  for (int i = 0; i < tr_.ndirs; ++i) {
    const Dir& d = d_[i];
    const std::string pre = tr_.ndirs == 2 ? tag[i] : "";
    s += "\n  " + pre + "w_gifo_x_  " + MomentStatistics(d.w_gifo_x) + "\n  " + pre + "w_gifo_r_  " + MomentStatistics(d.w_gifo_r) +
         "\n  " + pre + "bias_  " + MomentStatistics(d.bias) + "\n  " + pre + "peephole_i_c_  " + MomentStatistics(d.peep_i) +
         "\n  " + pre + "peephole_f_c_  " + MomentStatistics(d.peep_f) + "\n  " + pre + "peephole_o_c_  " + MomentStatistics(d.peep_o);
    if (tr_.projected) s += "\n  " + pre + "w_r_m_  " + MomentStatistics(d.w_r_m);
  }
  return s;
}
std::string LstmFamily::InfoGradient() const {
  std::string s;
  const char* tag[2] = {"f_", "b_"};
  for (int i = 0; i < tr_.ndirs; ++i) {
    const Dir& d = d_[i];
    const std::string pre = tr_.ndirs == 2 ? tag[i] : "";
    s += "\n  " + pre + "w_gifo_x_corr_  " + MomentStatistics(d.w_gifo_x_corr) + "\n  " + pre + "w_gifo_r_corr_  " + MomentStatistics(d.w_gifo_r_corr) +
         "\n  " + pre + "bias_corr_  " + MomentStatistics(d.bias_corr);
    if (tr_.projected) s += "\n  " + pre + "w_r_m_corr_  " + MomentStatistics(d.w_r_m_corr);
  }
  return s;
}

void LstmFamily::ResetLstmStreams(const std::vector<int32>& stream_reset_flag) {
  if (!tr_.carry_state) return;
  if (nstream_ == 0) {
    nstream_ = static_cast<int32>(stream_reset_flag.size());
    prev_state_.Resize(nstream_, Width(), kSetZero);
    KALDI_LOG << "Running training with " << nstream_ << " streams.";
  }
  KALDI_ASSERT(prev_state_.NumRows() == static_cast<int32>(stream_reset_flag.size()));
  for (size_t s = 0; s < stream_reset_flag.size(); ++s)
    if (stream_reset_flag[s] == 1) prev_state_.RowRange(static_cast<int32>(s), 1).SetZero();
}

void LstmFamily::SetSeqLengths(const std::vector<int32>& sequence_lengths) {
  if (tr_.use_seq_lengths) {
    sequence_lengths_ = sequence_lengths;
    seq_len_dev_ = sequence_lengths;
    nstream_ = static_cast<int32>(sequence_lengths.size());
  } else {
    // whole-sentence training through a state-carrying component: streams restart from zero (lc.h:497-501)
    nstream_ = static_cast<int32>(sequence_lengths.size());
    prev_state_.Resize(nstream_, Width(), kSetZero);
  }
}

bool LstmFamily::FoldProjection() const {
  if (!tr_.projected) return false;
  // ASLP_LSTM_FOLD_PROJECTION=0/1 forces the choice (tests run both); by default fold unless it would more than
  // double the recurrent FMAs per step (4C*C folded vs 4C*R + R*C)
  const char* env = std::getenv("ASLP_LSTM_FOLD_PROJECTION");
  if (env != nullptr && env[0] != '\0') return env[0] != '0';
  return ncell_ <= 2 * nrecur_;
}

void LstmFamily::FillDirArgs(void* arr_v, int T, int S, bool bwd) {
  aslp_lstm_dir_t* arr = static_cast<aslp_lstm_dir_t*>(arr_v);
  const bool fold = FoldProjection();
  for (int i = 0; i < tr_.ndirs; ++i) {
    Dir& d = d_[i];
    aslp_lstm_dir_t& a = arr[i];
    a.T = T; a.S = S; a.C = ncell_; a.R = (tr_.projected && !fold) ? nrecur_ : 0;
    a.reverse = i;                                  // direction 0 walks t = 1..T, direction 1 walks t = T..1
    a.buf = d.prop.Data(); a.ldb = d.prop.Stride();
    a.dbuf = bwd ? d.back.Data() : nullptr; a.lddb = bwd ? d.back.Stride() : 0;
    if (fold) { a.w_r = d.w_fused.Data(); a.ldwr = d.w_fused.Stride(); a.w_rm = nullptr; a.ldwrm = 0; }
    else {
      a.w_r = d.w_gifo_r.Data(); a.ldwr = d.w_gifo_r.Stride();
      a.w_rm = tr_.projected ? d.w_r_m.Data() : nullptr; a.ldwrm = tr_.projected ? d.w_r_m.Stride() : 0;
    }
    a.peep_i = d.peep_i.Data(); a.peep_f = d.peep_f.Data(); a.peep_o = d.peep_o.Data();
    a.seq_len_dev = (tr_.use_seq_lengths && i == 1) ? seq_len_dev_.Data() : nullptr;
    a.cell_clip = 50.0f;
  }
}

void LstmFamily::PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
  if (nstream_ == 0) {
    if (tr_.use_seq_lengths) KALDI_ERR << "SetSeqLengths must be called before Propagate for " << TypeToMarker(GetType());
    per_utt_reset_ = true;          // nnet-forward: one stream, state reset per utterance
    nstream_ = 1;
    prev_state_.Resize(nstream_, Width(), kSetZero);
    KALDI_LOG << "Running nnet-forward with per-utterance LSTM-state reset";
  }
  if (per_utt_reset_ && tr_.carry_state) prev_state_.SetZero();
  KALDI_ASSERT(nstream_ > 0 && in.NumRows() % nstream_ == 0);
  const int32 S = nstream_, T = in.NumRows() / S, C = ncell_, W = Width();
  aslp_stream_t st = CuStream();
  for (int i = 0; i < tr_.ndirs; ++i) {
    Dir& d = d_[i];
    // rows [S,(T+1)S) are fully overwritten (GEMM -> gifo, kernel -> the rest); only the boundary blocks need zeroing
    d.prop.Resize((T + 2) * S, W, kUndefined);
    d.prop.RowRange(0, S).SetZero();
    d.prop.RowRange((T + 1) * S, S).SetZero();
    if (i == 0 && tr_.carry_state) d.prop.RowRange(0, S).CopyFromMat(prev_state_);
    // x -> g,i,f,o for the whole chunk, bias fused in the epilogue (lc.h:552-555, :632-645)
    CuSubMatrix<BaseFloat> gifo = d.prop.Range(S, T * S, 0, 4 * C);
    const size_t wsb = aslp_gemm_workspace_bytes(T * S, 4 * C, input_dim_);      // few-tile shapes (short BPTT chunks) split K
    ASLP_OK(aslp_gemm(st, 0, 1, T * S, 4 * C, input_dim_, 1.0f, in.Data(), in.Stride(), d.w_gifo_x.Data(), d.w_gifo_x.Stride(), 0.0f,
                      gifo.Data(), gifo.Stride(), d.bias.Data(), 0.0f, GemmPrecision(), wsb ? CuWorkspace(wsb) : nullptr, wsb));
  }
  const bool fold = FoldProjection();
  if (fold) {
    for (Dir& d : d_) {
      d.w_fused.Resize(4 * C, C, kUndefined);
      ASLP_OK(aslp_gemm(st, 0, 0, 4 * C, C, nrecur_, 1.0f, d.w_gifo_r.Data(), d.w_gifo_r.Stride(), d.w_r_m.Data(), d.w_r_m.Stride(), 0.0f,
                        d.w_fused.Data(), d.w_fused.Stride(), nullptr, 0.0f, GemmPrecision(), nullptr, 0));
    }
    if (tr_.carry_state) {
      // the carried r(0) was projected with the W_r_m of the PREVIOUS minibatch (the weights have been updated since), so
      // the first step cannot go through W': feed r(0) W_gifo_r^T into the first step's pre-activations and hide m(0)
      Dir& d = d_[0];
      CuSubMatrix<BaseFloat> r0 = d.prop.Range(0, S, 7 * C, nrecur_), gifo1 = d.prop.Range(S, S, 0, 4 * C), m0 = d.prop.Range(0, S, 6 * C, C);
      ASLP_OK(aslp_gemm(st, 0, 1, S, 4 * C, nrecur_, 1.0f, r0.Data(), r0.Stride(), d.w_gifo_r.Data(), d.w_gifo_r.Stride(), 1.0f,
                        gifo1.Data(), gifo1.Stride(), nullptr, 0.0f, GemmPrecision(), nullptr, 0));
      m0.SetZero();
    }
  }
  aslp_lstm_dir_t args[2];
  FillDirArgs(args, T, S, false);
  const size_t wsb = aslp_lstm_workspace_bytes(T, S, C, (tr_.projected && !fold) ? nrecur_ : 0, tr_.ndirs, 0);
  ASLP_OK(aslp_lstm_seq_fwd(st, args, tr_.ndirs, CuWorkspace(wsb), wsb));
  if (fold) {
    // r(t) = m(t) W_r_m^T for the whole chunk at once
    for (Dir& d : d_) {
      CuSubMatrix<BaseFloat> ym = d.prop.Range(S, T * S, 6 * C, C), yr = d.prop.Range(S, T * S, 7 * C, nrecur_);
      ASLP_OK(aslp_gemm(st, 0, 1, T * S, nrecur_, C, 1.0f, ym.Data(), ym.Stride(), d.w_r_m.Data(), d.w_r_m.Stride(), 0.0f,
                        yr.Data(), yr.Stride(), nullptr, 0.0f, GemmPrecision(), nullptr, 0));
    }
  }
  if (tr_.carry_state) {
    const int32 row = tr_.lc ? chunk_size_ : T;      // lc.h:629 vs nnet-lstm-projected-streams.h:432
    KALDI_ASSERT(row >= 0 && row <= T);
    prev_state_.CopyFromMat(d_[0].prop.RowRange(row * S, S));
  }
  const int32 O = OutPerDir(), col = tr_.projected ? 7 * C : 6 * C;
  for (int i = 0; i < tr_.ndirs; ++i) {
    CuSubMatrix<BaseFloat> dst = out->ColRange(i * O, O);
    dst.CopyFromMat(d_[i].prop.Range(S, T * S, col, O));
  }
}

void LstmFamily::BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
  const int32 S = nstream_, T = in.NumRows() / S, C = ncell_, W = Width();
  const int32 O = OutPerDir(), ocol = tr_.projected ? 7 * C : 6 * C;
  aslp_stream_t st = CuStream();
  for (int i = 0; i < tr_.ndirs; ++i) {
    Dir& d = d_[i];
    d.back.Resize((T + 2) * S, W, kUndefined);
    d.back.RowRange(0, S).SetZero();
    d.back.RowRange((T + 1) * S, S).SetZero();
    CuSubMatrix<BaseFloat> od = d.back.Range(S, T * S, ocol, O);
    od.CopyFromMat(out_diff.ColRange(i * O, O));
  }
  const bool fold = FoldProjection();
  const int prec = GemmPrecision();
  if (fold) {
    // the out_diff share of d_m for every step: out_diff_r W_r_m, placed where the kernel expects out_diff (m columns)
    for (Dir& d : d_) {
      CuSubMatrix<BaseFloat> od = d.back.Range(S, T * S, 7 * C, nrecur_), odm = d.back.Range(S, T * S, 6 * C, C);
      ASLP_OK(aslp_gemm(st, 0, 0, T * S, C, nrecur_, 1.0f, od.Data(), od.Stride(), d.w_r_m.Data(), d.w_r_m.Stride(), 0.0f,
                        odm.Data(), odm.Stride(), nullptr, 0.0f, prec, nullptr, 0));
    }
  }
  aslp_lstm_dir_t args[2];
  FillDirArgs(args, T, S, true);
  const size_t wsb = aslp_lstm_workspace_bytes(T, S, C, (tr_.projected && !fold) ? nrecur_ : 0, tr_.ndirs, 1);
  ASLP_OK(aslp_lstm_seq_bwd(st, args, tr_.ndirs, CuWorkspace(wsb), wsb));
  if (fold) {
    // d_r(t) = out_diff(t) + dgifo(successor of t) W_gifo_r: direction 0's successor is t+1, direction 1's is t-1;
    // the boundary blocks of the derivative buffer are zero
    for (int i = 0; i < tr_.ndirs; ++i) {
      Dir& d = d_[i];
      CuSubMatrix<BaseFloat> dg_next = d.back.Range(i == 0 ? 2 * S : 0, T * S, 0, 4 * C), dr = d.back.Range(S, T * S, 7 * C, nrecur_);
      ASLP_OK(aslp_gemm(st, 0, 0, T * S, nrecur_, 4 * C, 1.0f, dg_next.Data(), dg_next.Stride(), d.w_gifo_r.Data(), d.w_gifo_r.Stride(), 1.0f,
                        dr.Data(), dr.Stride(), nullptr, 0.0f, prec, nullptr, 0));
    }
  }

  const float mmt = opts_.momentum, clip = clip_gradient_;
  for (int i = 0; i < tr_.ndirs; ++i) {
    Dir& d = d_[i];
    CuSubMatrix<BaseFloat> dgifo = d.back.Range(S, T * S, 0, 4 * C);
    // g,i,f,o -> x (lc.h:963-965): second direction accumulates
    const size_t wsb = aslp_gemm_workspace_bytes(T * S, input_dim_, 4 * C);
    ASLP_OK(aslp_gemm(st, 0, 0, T * S, input_dim_, 4 * C, 1.0f, dgifo.Data(), dgifo.Stride(), d.w_gifo_x.Data(), d.w_gifo_x.Stride(),
                      i == 0 ? 0.0f : 1.0f, in_diff->Data(), in_diff->Stride(), nullptr, 0.0f, prec, wsb ? CuWorkspace(wsb) : nullptr, wsb));
  }
  if (skip_wgrad_) { async_tail_ = false; return; }
  // Everything below (weight gradients, then Update) is off the critical path of the backward pass: nothing downstream reads
  // these weights or corr buffers before the next Propagate.  It goes to the side stream so that it overlaps the backward
  // recurrence of the layer below, which is latency-bound and leaves most SMs idle (Nnet::Backpropagate joins at its end).
  async_tail_ = CuAsyncEnabled();
  if (async_tail_) CuFork();
  std::unique_ptr<CuStreamScope> side_scope;
  if (async_tail_) side_scope.reset(new CuStreamScope(CuSideStream()));
  st = CuStream();
  for (int i = 0; i < tr_.ndirs; ++i) {
    Dir& d = d_[i];
    // rows of the forward buffers that held the recurrent input / previous cell of each step:
    // direction 0 read t-1 (rows [0,T*S)), direction 1 read t+1 (rows [2S,(T+2)S))   (lc.h:981-1000 vs :1022-1040)
    const int32 prev0 = i == 0 ? 0 : 2 * S;
    CuSubMatrix<BaseFloat> dgifo = d.back.Range(S, T * S, 0, 4 * C);
    const size_t gws = 64u << 20;
    void* ws = CuWorkspace(gws);
    ASLP_OK(aslp_gemm(st, 1, 0, 4 * C, input_dim_, T * S, 1.0f, dgifo.Data(), dgifo.Stride(), in.Data(), in.Stride(), mmt,
                      d.w_gifo_x_corr.Data(), d.w_gifo_x_corr.Stride(), nullptr, clip, prec, ws, gws));
    CuSubMatrix<BaseFloat> rec_prev = d.prop.Range(prev0, T * S, ocol, RecDim());
    ASLP_OK(aslp_gemm(st, 1, 0, 4 * C, RecDim(), T * S, 1.0f, dgifo.Data(), dgifo.Stride(), rec_prev.Data(), rec_prev.Stride(), mmt,
                      d.w_gifo_r_corr.Data(), d.w_gifo_r_corr.Stride(), nullptr, clip, prec, ws, gws));
    ASLP_OK(aslp_col_sum(st, d.bias_corr.Data(), dgifo.Data(), dgifo.Stride(), T * S, 4 * C, 1.0f, mmt, clip));
    CuSubMatrix<BaseFloat> c_prev = d.prop.Range(prev0, T * S, 4 * C, C), c_cur = d.prop.Range(S, T * S, 4 * C, C);
    CuSubMatrix<BaseFloat> di = d.back.Range(S, T * S, C, C), df = d.back.Range(S, T * S, 2 * C, C), d_out = d.back.Range(S, T * S, 3 * C, C);
    ASLP_OK(aslp_col_dot(st, d.peep_i_corr.Data(), di.Data(), di.Stride(), c_prev.Data(), c_prev.Stride(), T * S, C, 1.0f, mmt, clip));
    ASLP_OK(aslp_col_dot(st, d.peep_f_corr.Data(), df.Data(), df.Stride(), c_prev.Data(), c_prev.Stride(), T * S, C, 1.0f, mmt, clip));
    ASLP_OK(aslp_col_dot(st, d.peep_o_corr.Data(), d_out.Data(), d_out.Stride(), c_cur.Data(), c_cur.Stride(), T * S, C, 1.0f, mmt, clip));
    if (tr_.projected) {
      CuSubMatrix<BaseFloat> dr = d.back.Range(S, T * S, 7 * C, nrecur_), ym = d.prop.Range(S, T * S, 6 * C, C);
      ASLP_OK(aslp_gemm(st, 1, 0, nrecur_, C, T * S, 1.0f, dr.Data(), dr.Stride(), ym.Data(), ym.Stride(), mmt,
                        d.w_r_m_corr.Data(), d.w_r_m_corr.Stride(), nullptr, clip, prec, ws, gws));
    }
  }
}

void LstmFamily::Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
  // plain -lr * corr: no per-component coefficient, no L2 (lc.h:1085-1110)
  const float lr = opts_.learn_rate;
  std::unique_ptr<CuStreamScope> side_scope;
  if (async_tail_) side_scope.reset(new CuStreamScope(CuSideStream()));     // stays behind this component's weight-gradient GEMMs
  aslp_stream_t st = CuStream();
  auto mat = [&](CuMatrix<BaseFloat>& w, const CuMatrix<BaseFloat>& c) { ASLP_OK(aslp_axpby(st, w.Data(), w.Stride(), c.Data(), c.Stride(), w.NumRows(), w.NumCols(), -lr, 1.0f)); };
  auto vec = [&](CuVector<BaseFloat>& w, const CuVector<BaseFloat>& c) { ASLP_OK(aslp_axpby(st, w.Data(), (w.Dim() + 3) / 4 * 4, c.Data(), (c.Dim() + 3) / 4 * 4, 1, w.Dim(), -lr, 1.0f)); };
  for (Dir& d : d_) {
    mat(d.w_gifo_x, d.w_gifo_x_corr); mat(d.w_gifo_r, d.w_gifo_r_corr); vec(d.bias, d.bias_corr);
    vec(d.peep_i, d.peep_i_corr); vec(d.peep_f, d.peep_f_corr); vec(d.peep_o, d.peep_o_corr);
    if (tr_.projected) mat(d.w_r_m, d.w_r_m_corr);
  }
}


// ------------------------------------------------------------------ LstmCifgProjectedStreams
LstmCifgProjectedStreams::LstmCifgProjectedStreams(int32 input_dim, int32 output_dim)
    : UpdatableComponent(input_dim, output_dim), ncell_(0), nrecur_(output_dim), clip_gradient_(0.0f),
      engine_(input_dim, output_dim, LstmFamily::Traits{kLstmProjectedStreams, 1, true, true, true, false, false}) {
  engine_.SetSkipWeightGradients(true);
}

void LstmCifgProjectedStreams::AllocCorr() {
  w_gfo_x_corr_.Resize(3 * ncell_, input_dim_, kSetZero);
  w_gfo_r_corr_.Resize(3 * ncell_, nrecur_, kSetZero);
  bias_corr_.Resize(3 * ncell_, kSetZero);
  peephole_f_c_corr_.Resize(ncell_, kSetZero);
  peephole_o_c_corr_.Resize(ncell_, kSetZero);
  w_r_m_corr_.Resize(nrecur_, ncell_, kSetZero);
}

void LstmCifgProjectedStreams::SizeEngine() {
  engine_.ncell_ = ncell_;
  engine_.clip_gradient_ = clip_gradient_;
  LstmFamily::Dir& d = engine_.d_[0];
  d.w_gifo_x.Resize(4 * ncell_, input_dim_, kUndefined);
  d.w_gifo_r.Resize(4 * ncell_, nrecur_, kUndefined);
  d.w_r_m.Resize(nrecur_, ncell_, kUndefined);
  d.bias.Resize(4 * ncell_, kUndefined);
  d.peep_i.Resize(ncell_, kUndefined); d.peep_f.Resize(ncell_, kUndefined); d.peep_o.Resize(ncell_, kUndefined);
}

void LstmCifgProjectedStreams::InitData(std::istream& is) {
  float param_scale = 0.02f;
  ProtoOptions po("(CellDim|ClipGradient|ParamScale)");
  po.Int("<CellDim>", &ncell_);
  po.Float("<ClipGradient>", &clip_gradient_);
  po.Float("<ParamScale>", &param_scale);
  po.Parse(is);
  KALDI_ASSERT(ncell_ > 0);
  w_gfo_x_.Resize(3 * ncell_, input_dim_, kUndefined);
  w_gfo_r_.Resize(3 * ncell_, nrecur_, kUndefined);
  w_r_m_.Resize(nrecur_, ncell_, kUndefined);
  InitMatParam(&w_gfo_x_, param_scale);
  InitMatParam(&w_gfo_r_, param_scale);
  InitMatParam(&w_r_m_, param_scale);
  bias_.Resize(3 * ncell_, kUndefined);
  peephole_f_c_.Resize(ncell_, kUndefined);
  peephole_o_c_.Resize(ncell_, kUndefined);
  InitVecParam(&bias_, param_scale);
  InitVecParam(&peephole_f_c_, param_scale);
  InitVecParam(&peephole_o_c_, param_scale);
  AllocCorr();
  SizeEngine();
  KALDI_ASSERT(clip_gradient_ >= 0.0);
}

void LstmCifgProjectedStreams::ReadData(std::istream& is, bool binary) {
  ExpectToken(is, binary, "<CellDim>");
  ReadBasicType(is, binary, &ncell_);
  ExpectToken(is, binary, "<ClipGradient>");
  ReadBasicType(is, binary, &clip_gradient_);
  w_gfo_x_.Read(is, binary);
  w_gfo_r_.Read(is, binary);
  bias_.Read(is, binary);
  peephole_f_c_.Read(is, binary);
  peephole_o_c_.Read(is, binary);
  w_r_m_.Read(is, binary);
  KALDI_ASSERT(w_gfo_x_.NumRows() == 3 * ncell_ && w_gfo_x_.NumCols() == input_dim_);
  KALDI_ASSERT(w_gfo_r_.NumRows() == 3 * ncell_ && w_gfo_r_.NumCols() == nrecur_);
  AllocCorr();
  SizeEngine();
}

void LstmCifgProjectedStreams::WriteData(std::ostream& os, bool binary) const {
  WriteToken(os, binary, "<CellDim>");
  WriteBasicType(os, binary, ncell_);
  WriteToken(os, binary, "<ClipGradient>");
  WriteBasicType(os, binary, clip_gradient_);
  w_gfo_x_.Write(os, binary);
  w_gfo_r_.Write(os, binary);
  bias_.Write(os, binary);
  peephole_f_c_.Write(os, binary);
  peephole_o_c_.Write(os, binary);
  w_r_m_.Write(os, binary);
}

int32 LstmCifgProjectedStreams::NumParams() const {
  return w_gfo_x_.NumRows() * w_gfo_x_.NumCols() + w_gfo_r_.NumRows() * w_gfo_r_.NumCols() + bias_.Dim() + peephole_f_c_.Dim() + peephole_o_c_.Dim() +
         w_r_m_.NumRows() * w_r_m_.NumCols();
}

void LstmCifgProjectedStreams::GetParams(Vector<BaseFloat>* wei_copy) const {
  wei_copy->Resize(NumParams());
  float* p = wei_copy->Data();
  auto vec = [&](const CuVector<BaseFloat>& v) { Vector<float> h; v.CopyToVec(&h); for (int32 i = 0; i < h.Dim(); ++i) *p++ = h(i); };
  auto mat = [&](const CuMatrix<BaseFloat>& m) { CopyRowsToVec(m, p); p += static_cast<size_t>(m.NumRows()) * m.NumCols(); };
  mat(w_gfo_x_); mat(w_gfo_r_); vec(bias_); vec(peephole_f_c_); vec(peephole_o_c_); mat(w_r_m_);
}

void LstmCifgProjectedStreams::GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
  params->clear();
  params->push_back(std::make_pair(w_gfo_x_.Data(), w_gfo_x_.NumRows() * w_gfo_x_.Stride()));
  params->push_back(std::make_pair(w_gfo_r_.Data(), w_gfo_r_.NumRows() * w_gfo_r_.Stride()));
  params->push_back(std::make_pair(bias_.Data(), bias_.Dim()));
  params->push_back(std::make_pair(peephole_f_c_.Data(), peephole_f_c_.Dim()));
  params->push_back(std::make_pair(peephole_o_c_.Data(), peephole_o_c_.Dim()));
  params->push_back(std::make_pair(w_r_m_.Data(), w_r_m_.NumRows() * w_r_m_.Stride()));
}

std::string LstmCifgProjectedStreams::Info() const {
  return std::string("  ") + "\n  w_gfo_x_  " + MomentStatistics(w_gfo_x_) + "\n  w_gfo_r_  " + MomentStatistics(w_gfo_r_) + "\n  bias_  " + MomentStatistics(bias_) +
         "\n  peephole_f_c_  " + MomentStatistics(peephole_f_c_) + "\n  peephole_o_c_  " + MomentStatistics(peephole_o_c_) + "\n  w_r_m_  " + MomentStatistics(w_r_m_);
}
std::string LstmCifgProjectedStreams::InfoGradient() const {
  return std::string("  ") + "\n  Gradients:" + "\n  w_gfo_x_corr_  " + MomentStatistics(w_gfo_x_corr_) + "\n  w_gfo_r_corr_  " + MomentStatistics(w_gfo_r_corr_) +
         "\n  bias_corr_  " + MomentStatistics(bias_corr_) + "\n  peephole_f_c_corr_  " + MomentStatistics(peephole_f_c_corr_) +
         "\n  peephole_o_c_corr_  " + MomentStatistics(peephole_o_c_corr_) + "\n  w_r_m_corr_  " + MomentStatistics(w_r_m_corr_);
}

void LstmCifgProjectedStreams::PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
  aslp_stream_t st = CuStream();
  LstmFamily::Dir& d = engine_.d_[0];
  const int32 C = ncell_;
  // four-gate view of the three-gate parameters: rows [g | -f | f | o]
  ASLP_OK(aslp_cifg_expand(st, d.w_gifo_x.Data(), d.w_gifo_x.Stride(), w_gfo_x_.Data(), w_gfo_x_.Stride(), C, input_dim_));
  ASLP_OK(aslp_cifg_expand(st, d.w_gifo_r.Data(), d.w_gifo_r.Stride(), w_gfo_r_.Data(), w_gfo_r_.Stride(), C, nrecur_));
  ASLP_OK(aslp_cifg_expand(st, d.bias.Data(), 1, bias_.Data(), 1, C, 1));
  const int32 ldv = (C + 3) / 4 * 4;
  ASLP_OK(aslp_axpby(st, d.peep_i.Data(), ldv, peephole_f_c_.Data(), ldv, 1, C, -1.0f, 0.0f));
  ASLP_OK(aslp_axpby(st, d.peep_f.Data(), ldv, peephole_f_c_.Data(), ldv, 1, C, 1.0f, 0.0f));
  ASLP_OK(aslp_axpby(st, d.peep_o.Data(), ldv, peephole_o_c_.Data(), ldv, 1, C, 1.0f, 0.0f));
  d.w_r_m.CopyFromMat(w_r_m_);
  engine_.PropagateFnc(in, out);
}

void LstmCifgProjectedStreams::BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff,
                                                CuMatrixBase<BaseFloat>* in_diff) {
  engine_.BackpropagateFnc(in, out, out_diff, in_diff);        // recurrence, d_r, in_diff = dgifo W4_x (= DGFO w_gfo_x_)
  const int32 S = engine_.nstream_, T = in.NumRows() / S, C = ncell_;
  const LstmFamily::Dir& d = engine_.d_[0];
  aslp_stream_t st = CuStream();
  const int prec = GemmPrecision();
  const float mmt = opts_.momentum, clip = clip_gradient_;
  dgfo_.Resize(T * S, 3 * C, kUndefined);
  CuSubMatrix<BaseFloat> dgifo = d.back.Range(S, T * S, 0, 4 * C);
  ASLP_OK(aslp_cifg_compact(st, dgfo_.Data(), dgfo_.Stride(), dgifo.Data(), dgifo.Stride(), T * S, C));
  const size_t gws = 64u << 20;
  void* ws = CuWorkspace(gws);
  // nnet-lstm-couple-if-projected-streams.h:593-640: momentum as beta of the product, then the elementwise clip
  ASLP_OK(aslp_gemm(st, 1, 0, 3 * C, input_dim_, T * S, 1.0f, dgfo_.Data(), dgfo_.Stride(), in.Data(), in.Stride(), mmt, w_gfo_x_corr_.Data(),
                    w_gfo_x_corr_.Stride(), nullptr, clip, prec, ws, gws));
  CuSubMatrix<BaseFloat> r_prev = d.prop.Range(0, T * S, 7 * C, nrecur_);
  ASLP_OK(aslp_gemm(st, 1, 0, 3 * C, nrecur_, T * S, 1.0f, dgfo_.Data(), dgfo_.Stride(), r_prev.Data(), r_prev.Stride(), mmt, w_gfo_r_corr_.Data(),
                    w_gfo_r_corr_.Stride(), nullptr, clip, prec, ws, gws));
  ASLP_OK(aslp_col_sum(st, bias_corr_.Data(), dgfo_.Data(), dgfo_.Stride(), T * S, 3 * C, 1.0f, mmt, clip));
  CuSubMatrix<BaseFloat> c_prev = d.prop.Range(0, T * S, 4 * C, C), c_cur = d.prop.Range(S, T * S, 4 * C, C);
  CuSubMatrix<BaseFloat> df = dgfo_.ColRange(C, C), d_out = dgfo_.ColRange(2 * C, C);
  ASLP_OK(aslp_col_dot(st, peephole_f_c_corr_.Data(), df.Data(), df.Stride(), c_prev.Data(), c_prev.Stride(), T * S, C, 1.0f, mmt, clip));
  ASLP_OK(aslp_col_dot(st, peephole_o_c_corr_.Data(), d_out.Data(), d_out.Stride(), c_cur.Data(), c_cur.Stride(), T * S, C, 1.0f, mmt, clip));
  CuSubMatrix<BaseFloat> dr = d.back.Range(S, T * S, 7 * C, nrecur_), ym = d.prop.Range(S, T * S, 6 * C, C);
  ASLP_OK(aslp_gemm(st, 1, 0, nrecur_, C, T * S, 1.0f, dr.Data(), dr.Stride(), ym.Data(), ym.Stride(), mmt, w_r_m_corr_.Data(), w_r_m_corr_.Stride(),
                    nullptr, clip, prec, ws, gws));
}

void LstmCifgProjectedStreams::Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
  const float lr = opts_.learn_rate;
  aslp_stream_t st = CuStream();
  auto mat = [&](CuMatrix<BaseFloat>& w, const CuMatrix<BaseFloat>& c) { ASLP_OK(aslp_axpby(st, w.Data(), w.Stride(), c.Data(), c.Stride(), w.NumRows(), w.NumCols(), -lr, 1.0f)); };
  auto vec = [&](CuVector<BaseFloat>& w, const CuVector<BaseFloat>& c) { ASLP_OK(aslp_axpby(st, w.Data(), (w.Dim() + 3) / 4 * 4, c.Data(), (c.Dim() + 3) / 4 * 4, 1, w.Dim(), -lr, 1.0f)); };
  mat(w_gfo_x_, w_gfo_x_corr_); mat(w_gfo_r_, w_gfo_r_corr_); vec(bias_, bias_corr_);
  vec(peephole_f_c_, peephole_f_c_corr_); vec(peephole_o_c_, peephole_o_c_corr_); mat(w_r_m_, w_r_m_corr_);
}

}  // namespace aslp_nnet
}  // namespace kaldi
