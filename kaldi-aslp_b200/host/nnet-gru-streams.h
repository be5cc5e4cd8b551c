// nnet-gru-streams.h -- GruStreams over the persistent GRU recurrence kernel (aslp_gru_seq_fwd/bwd).
// Reference: src/aslp-nnet/nnet-gru-streams.h:238-450.  Buffers keep the reference layout
// [(T+2)S, 5H] = [z r m g h]; state carried from row T; wgrads with momentum and clip fused in GEMM epilogues.
#ifndef ASLP_HOST_NNET_GRU_STREAMS_H_
#define ASLP_HOST_NNET_GRU_STREAMS_H_
#include "cu-workspace.h"
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class GruStreams : public UpdatableComponent {
 public:
  GruStreams(int32 input_dim, int32 output_dim) : UpdatableComponent(input_dim, output_dim), nstream_(0), clip_gradient_(0.0f), per_utt_reset_(false) {}
  Component* Copy() const { return new GruStreams(*this); }
  ComponentType GetType() const { return kGruStreams; }

  void InitData(std::istream& is) {
    float param_scale = 0.02f;
    ProtoOptions po("(ClipGradient|ParamScale)");
    po.Float("<ClipGradient>", &clip_gradient_); po.Float("<ParamScale>", &param_scale); po.Parse(is);
    const int32 H = output_dim_;
    w_zrm_x_.Resize(3 * H, input_dim_, kUndefined); w_zr_h_.Resize(2 * H, H, kUndefined); w_m_g_.Resize(H, H, kUndefined);
    InitMatParam(&w_zrm_x_, param_scale); InitMatParam(&w_zr_h_, param_scale); InitMatParam(&w_m_g_, param_scale);
    bias_.Resize(3 * H, kUndefined);
    InitVecParam(&bias_, param_scale);
    AllocCorr();
  }
  void ReadData(std::istream& is, bool binary) {
    ExpectToken(is, binary, "<ClipGradient>"); ReadBasicType(is, binary, &clip_gradient_);
    w_zrm_x_.Read(is, binary); w_zr_h_.Read(is, binary); w_m_g_.Read(is, binary); bias_.Read(is, binary);
    AllocCorr();
  }
  void WriteData(std::ostream& os, bool binary) const {
    WriteToken(os, binary, "<ClipGradient>"); WriteBasicType(os, binary, clip_gradient_);
    w_zrm_x_.Write(os, binary); w_zr_h_.Write(os, binary); w_m_g_.Write(os, binary); bias_.Write(os, binary);
  }
  int32 NumParams() const {
    return w_zrm_x_.NumRows() * w_zrm_x_.NumCols() + w_zr_h_.NumRows() * w_zr_h_.NumCols() + w_m_g_.NumRows() * w_m_g_.NumCols() + bias_.Dim();
  }
  void GetParams(Vector<BaseFloat>* w) const {
    w->Resize(NumParams());
    float* p = w->Data();
    for (const CuMatrix<BaseFloat>* m : {&w_zrm_x_, &w_zr_h_, &w_m_g_}) { CopyRowsToVec(*m, p); p += static_cast<size_t>(m->NumRows()) * m->NumCols(); }
    Vector<float> b; bias_.CopyToVec(&b);
    for (int32 i = 0; i < b.Dim(); ++i) *p++ = b(i);
  }
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params) {
    params->clear();
    params->push_back(std::make_pair(w_zrm_x_.Data(), w_zrm_x_.NumRows() * w_zrm_x_.Stride()));
    params->push_back(std::make_pair(w_zr_h_.Data(), w_zr_h_.NumRows() * w_zr_h_.Stride()));
    params->push_back(std::make_pair(w_m_g_.Data(), w_m_g_.NumRows() * w_m_g_.Stride()));
    params->push_back(std::make_pair(bias_.Data(), bias_.Dim()));
  }
  std::string Info() const {
    return std::string("  ") + "\n  w_zrm_x_  " + MomentStatistics(w_zrm_x_) + "\n  w_zr_h_  " + MomentStatistics(w_zr_h_) +
           "\n  w_m_g_  " + MomentStatistics(w_m_g_) + "\n  bias_  " + MomentStatistics(bias_);
  }

  void ResetLstmStreams(const std::vector<int32>& flag) {
    if (nstream_ == 0) {
      nstream_ = static_cast<int32>(flag.size());
      prev_state_.Resize(nstream_, 5 * output_dim_, kSetZero);
      KALDI_LOG << "Running training with " << nstream_ << " streams.";
    }
    KALDI_ASSERT(prev_state_.NumRows() == static_cast<int32>(flag.size()));
    for (size_t s = 0; s < flag.size(); ++s) if (flag[s] == 1) prev_state_.RowRange(static_cast<int32>(s), 1).SetZero();
  }
  void SetSeqLengths(const std::vector<int32>& l) { nstream_ = static_cast<int32>(l.size()); prev_state_.Resize(nstream_, 5 * output_dim_, kSetZero); }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {
    if (nstream_ == 0) {
      per_utt_reset_ = true; nstream_ = 1;
      prev_state_.Resize(nstream_, 5 * output_dim_, kSetZero);
      KALDI_LOG << "Runing nnet-forward with per-utterance GRU-state reset";
    }
    if (per_utt_reset_) prev_state_.SetZero();
    KALDI_ASSERT(in.NumRows() % nstream_ == 0);
    const int32 S = nstream_, T = in.NumRows() / S, H = output_dim_;
    aslp_stream_t st = CuStream();
    prop_.Resize((T + 2) * S, 5 * H, kUndefined);
    prop_.RowRange(0, S).CopyFromMat(prev_state_);
    prop_.RowRange((T + 1) * S, S).SetZero();
    CuSubMatrix<BaseFloat> zrm = prop_.Range(S, T * S, 0, 3 * H);
    ASLP_OK(aslp_gemm(st, 0, 1, T * S, 3 * H, input_dim_, 1.0f, in.Data(), in.Stride(), w_zrm_x_.Data(), w_zrm_x_.Stride(), 0.0f, zrm.Data(), zrm.Stride(),
                      bias_.Data(), 0.0f, GemmPrecision(), nullptr, 0));
    aslp_gru_t g = MakeArgs(T, S, false);
    const size_t wsb = aslp_gru_workspace_bytes(T, S, H, 0);
    ASLP_OK(aslp_gru_seq_fwd(st, &g, CuWorkspace(wsb), wsb));
    out->CopyFromMat(prop_.Range(S, T * S, 4 * H, H));
    prev_state_.CopyFromMat(prop_.RowRange(T * S, S));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {
    const int32 S = nstream_, T = in.NumRows() / S, H = output_dim_;
    aslp_stream_t st = CuStream();
    back_.Resize((T + 2) * S, 5 * H, kUndefined);
    back_.RowRange(0, S).SetZero();
    back_.RowRange((T + 1) * S, S).SetZero();
    CuSubMatrix<BaseFloat> dh = back_.Range(S, T * S, 4 * H, H);
    dh.CopyFromMat(out_diff);
    aslp_gru_t g = MakeArgs(T, S, true);
    const size_t wsb = aslp_gru_workspace_bytes(T, S, H, 1);
    ASLP_OK(aslp_gru_seq_bwd(st, &g, CuWorkspace(wsb), wsb));
    const float mmt = opts_.momentum, clip = clip_gradient_;
    const int prec = GemmPrecision();
    CuSubMatrix<BaseFloat> dzrm = back_.Range(S, T * S, 0, 3 * H), dzr = back_.Range(S, T * S, 0, 2 * H), dm = back_.Range(S, T * S, 2 * H, H);
    ASLP_OK(aslp_gemm(st, 0, 0, T * S, input_dim_, 3 * H, 1.0f, dzrm.Data(), dzrm.Stride(), w_zrm_x_.Data(), w_zrm_x_.Stride(), 0.0f, in_diff->Data(),
                      in_diff->Stride(), nullptr, 0.0f, prec, nullptr, 0));
    const size_t gws = 64u << 20;
    void* ws = CuWorkspace(gws);
    ASLP_OK(aslp_gemm(st, 1, 0, 3 * H, input_dim_, T * S, 1.0f, dzrm.Data(), dzrm.Stride(), in.Data(), in.Stride(), mmt, w_zrm_x_corr_.Data(),
                      w_zrm_x_corr_.Stride(), nullptr, clip, prec, ws, gws));
    ASLP_OK(aslp_col_sum(st, bias_corr_.Data(), dzrm.Data(), dzrm.Stride(), T * S, 3 * H, 1.0f, mmt, clip));
    CuSubMatrix<BaseFloat> h_prev = prop_.Range(0, T * S, 4 * H, H), yg = prop_.Range(S, T * S, 3 * H, H);
    ASLP_OK(aslp_gemm(st, 1, 0, 2 * H, H, T * S, 1.0f, dzr.Data(), dzr.Stride(), h_prev.Data(), h_prev.Stride(), mmt, w_zr_h_corr_.Data(),
                      w_zr_h_corr_.Stride(), nullptr, clip, prec, ws, gws));
    ASLP_OK(aslp_gemm(st, 1, 0, H, H, T * S, 1.0f, dm.Data(), dm.Stride(), yg.Data(), yg.Stride(), mmt, w_m_g_corr_.Data(), w_m_g_corr_.Stride(),
                      nullptr, clip, prec, ws, gws));
  }
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) {
    const float lr = opts_.learn_rate;
    w_zrm_x_.AddMat(-lr, w_zrm_x_corr_); w_zr_h_.AddMat(-lr, w_zr_h_corr_); w_m_g_.AddMat(-lr, w_m_g_corr_);
    const int32 ld = (bias_.Dim() + 3) / 4 * 4;
    ASLP_OK(aslp_axpby(CuStream(), bias_.Data(), ld, bias_corr_.Data(), ld, 1, bias_.Dim(), -lr, 1.0f));
  }

 private:
  void AllocCorr() {
    const int32 H = output_dim_;
    w_zrm_x_corr_.Resize(3 * H, input_dim_, kSetZero); w_zr_h_corr_.Resize(2 * H, H, kSetZero); w_m_g_corr_.Resize(H, H, kSetZero);
    bias_corr_.Resize(3 * H, kSetZero);
  }
  aslp_gru_t MakeArgs(int32 T, int32 S, bool bwd) {
    aslp_gru_t g;
    g.T = T; g.S = S; g.H = output_dim_;
    g.buf = prop_.Data(); g.ldb = prop_.Stride();
    g.dbuf = bwd ? back_.Data() : nullptr; g.lddb = bwd ? back_.Stride() : 0;
    g.w_zr_h = w_zr_h_.Data(); g.ldwzr = w_zr_h_.Stride();
    g.w_m_g = w_m_g_.Data(); g.ldwmg = w_m_g_.Stride();
    return g;
  }
  int32 nstream_;
  BaseFloat clip_gradient_;
  bool per_utt_reset_;
  CuMatrix<BaseFloat> w_zrm_x_, w_zr_h_, w_m_g_, w_zrm_x_corr_, w_zr_h_corr_, w_m_g_corr_;
  CuVector<BaseFloat> bias_, bias_corr_;
  CuMatrix<BaseFloat> prop_, back_, prev_state_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
