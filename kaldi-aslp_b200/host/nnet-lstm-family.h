// nnet-lstm-family.h -- Lstm, BLstm, LstmProjectedStreams, BLstmProjectedStreams and BLstmProjectedStreamsLC
// as ONE parameterised implementation over the persistent recurrence kernel (aslp_lstm_seq_fwd/bwd).
// Reference classes and what distinguishes them:
//   Lstm                     nnet-recurrent-component.{h,cc}:28-491   1 dir, no projection, state carried from row T
//   BLstm                    nnet-recurrent-component.cc:556-1418     2 dirs, no projection, per-stream lengths
//   LstmProjectedStreams     nnet-lstm-projected-streams.h            1 dir, projection R = OutputDim, state from row T
//   BLstmProjectedStreams    nnet-blstm-projected-streams.h           2 dirs, R = OutputDim/2, lengths, no carry
//   BLstmProjectedStreamsLC  nnet-blstm-projected-streams-lc.h        2 dirs, R = OutputDim/2, forward state carried
//                                                                     from row chunk_size (:629), backward from zero
// Per minibatch: 1 tcgen05 GEMM (+bias) per direction for x*W_x^T, ONE persistent launch for all T steps of all
// directions, then the chunk wgrad GEMMs with momentum and clip fused in their epilogues.
// Projected variants fold the projection into the recurrence (W' = W_gifo_r W_r_m, gates(t) += W' m(t-1)), which
// halves the number of cross-SM exchanges per time step; r(t) = W_r_m m(t) and d_r(t) = out_diff + W_gifo_r^T dgifo(t+1)
// then come from bulk GEMMs over the whole chunk (same sums as lc.h:664-668 / :872-884, re-associated).
#ifndef ASLP_HOST_NNET_LSTM_FAMILY_H_
#define ASLP_HOST_NNET_LSTM_FAMILY_H_
#include "nnet-component.h"

namespace kaldi {
namespace aslp_nnet {

class LstmFamily : public UpdatableComponent {
 public:
  struct Traits {
    ComponentType type;
    int ndirs;
    bool projected;        // has W_r_m
    bool has_celldim_token;
    bool carry_state;      // forward-direction state carried across minibatches
    bool lc;               // carried state taken from row chunk_size_ instead of row T
    bool use_seq_lengths;  // backward-direction rows past the stream length are zeroed
  };
  LstmFamily(int32 input_dim, int32 output_dim, const Traits& tr);
  Component* Copy() const { return new LstmFamily(*this); }
  ComponentType GetType() const { return tr_.type; }

  void InitData(std::istream& is);
  void ReadData(std::istream& is, bool binary);
  void WriteData(std::ostream& os, bool binary) const;
  int32 NumParams() const;
  void GetParams(Vector<BaseFloat>* wei_copy) const;
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params);
  std::string Info() const;
  std::string InfoGradient() const;

  // recurrent extras reached from Nnet by dynamic_cast (nnet-nnet.cc:467-532)
  void ResetLstmStreams(const std::vector<int32>& stream_reset_flag);
  void SetSeqLengths(const std::vector<int32>& sequence_lengths);
  void SetChunkSize(int32 chunk_size) { chunk_size_ = chunk_size; }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out);
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff);
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff);

  // the weight-gradient tail of BackpropagateFnc is skipped: the owner computes its own (LstmCifgProjectedStreams)
  void SetSkipWeightGradients(bool v) { skip_wgrad_ = v; }

  // test / tooling access
  const CuMatrix<BaseFloat>& PropagateBuf(int dir) const { return d_[dir].prop; }
  const CuMatrix<BaseFloat>& BackpropagateBuf(int dir) const { return d_[dir].back; }

 private:
  friend class LstmCifgProjectedStreams;
  struct Dir {
    CuMatrix<BaseFloat> w_gifo_x, w_gifo_r, w_r_m;
    CuVector<BaseFloat> bias, peep_i, peep_f, peep_o;
    CuMatrix<BaseFloat> w_gifo_x_corr, w_gifo_r_corr, w_r_m_corr;
    CuVector<BaseFloat> bias_corr, peep_i_corr, peep_f_corr, peep_o_corr;
    CuMatrix<BaseFloat> prop, back;
    CuMatrix<BaseFloat> w_fused;      // W_gifo_r * W_r_m [4C, C]: the projection folded into the recurrence (rebuilt every Propagate)
  };
  void AllocCorr();
  int32 Width() const { return 7 * ncell_ + nrecur_; }
  int32 RecDim() const { return tr_.projected ? nrecur_ : ncell_; }   // recurrent input dim of W_gifo_r
  int32 OutPerDir() const { return tr_.projected ? nrecur_ : ncell_; }
  void FillDirArgs(void* arr, int T, int S, bool bwd);
  bool FoldProjection() const;          // run the recurrence on m(t-1) through W_gifo_r*W_r_m, r(t) / d_r(t) by bulk GEMMs

  Traits tr_;
  int32 ncell_, nrecur_, nstream_, chunk_size_;
  BaseFloat clip_gradient_;
  std::vector<Dir> d_;
  CuMatrix<BaseFloat> prev_state_;                 // [S, 7C+R] forward-direction carried state
  std::vector<int32> sequence_lengths_;
  CuArrayInt seq_len_dev_;
  bool async_tail_ = false;             // the last BackpropagateFnc put its weight gradients on the side stream
  bool per_utt_reset_;                  // nnet-forward mode: 1 stream, state reset every call (the reference's function-local static)
  bool skip_wgrad_ = false;
};

// thin named types so that factory / dynamic_cast code reads like the reference
#define ASLP_LSTM_TYPE(Name, ...)                                                             \
  class Name : public LstmFamily {                                                            \
   public:                                                                                    \
    Name(int32 in, int32 out) : LstmFamily(in, out, Traits{__VA_ARGS__}) {}                   \
    Component* Copy() const { return new Name(*this); }                                       \
  }
//                                   type                      dirs proj  celldim carry  lc     lengths
ASLP_LSTM_TYPE(Lstm,                    kLstm,                    1, false, false, true,  false, false);
ASLP_LSTM_TYPE(BLstm,                   kBLstm,                   2, false, false, false, false, true);
ASLP_LSTM_TYPE(LstmProjectedStreams,    kLstmProjectedStreams,    1, true,  true,  true,  false, false);
ASLP_LSTM_TYPE(BLstmProjectedStreams,   kBLstmProjectedStreams,   2, true,  true,  false, false, true);
ASLP_LSTM_TYPE(BLstmProjectedStreamsLC, kBLstmProjectedStreamsLC, 2, true,  true,  true,  true,  false);
#undef ASLP_LSTM_TYPE


// LstmCifgProjectedStreams (src/aslp-nnet/nnet-lstm-couple-if-projected-streams.h): projected LSTM whose input gate is coupled to the
// forget gate, i = 1 - f, with three gate blocks [g f o] and two peepholes.  Since 1 - sigmoid(x) = sigmoid(-x), the coupled cell IS
// the four-gate cell with W_i = -W_f, bias_i = -bias_f, peephole_i = -peephole_f: the component keeps the reference's three-gate
// parameters (file format, GetParams / GetGpuParams order, gradients, clipping) and drives the persistent recurrence of LstmFamily
// with the expanded four-gate view, rebuilt from the parameters at every Propagate.  The three-gate derivative the weight gradients
// need is d_f(coupled) = d_f - d_i of the four-gate cell (nnet-lstm-couple-if-projected-streams.h:566-569).
class LstmCifgProjectedStreams : public UpdatableComponent {
 public:
  LstmCifgProjectedStreams(int32 input_dim, int32 output_dim);
  Component* Copy() const { return new LstmCifgProjectedStreams(*this); }
  ComponentType GetType() const { return kLstmCifgProjectedStreams; }
  void InitData(std::istream& is);
  void ReadData(std::istream& is, bool binary);
  void WriteData(std::ostream& os, bool binary) const;
  int32 NumParams() const;
  void GetParams(Vector<BaseFloat>* wei_copy) const;
  void GetGpuParams(std::vector<std::pair<BaseFloat*, int>>* params);
  std::string Info() const;
  std::string InfoGradient() const;
  void ResetLstmStreams(const std::vector<int32>& stream_reset_flag) { engine_.ResetLstmStreams(stream_reset_flag); }
  void SetSeqLengths(const std::vector<int32>& sequence_lengths) { engine_.SetSeqLengths(sequence_lengths); }
  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out);
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out, const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff);
  void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff);
 private:
  void SizeEngine();
  void AllocCorr();
  int32 ncell_, nrecur_;
  BaseFloat clip_gradient_;
  CuMatrix<BaseFloat> w_gfo_x_, w_gfo_r_, w_r_m_, w_gfo_x_corr_, w_gfo_r_corr_, w_r_m_corr_;
  CuVector<BaseFloat> bias_, peephole_f_c_, peephole_o_c_, bias_corr_, peephole_f_c_corr_, peephole_o_c_corr_;
  CuMatrix<BaseFloat> dgfo_;         // [T*S, 3C] three-gate derivatives of the last backward pass
  LstmFamily engine_;                // four-gate recurrence (LstmProjectedStreams traits), weight gradients off
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
