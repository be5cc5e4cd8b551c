// nnet-randomizer.h -- frame-level shuffling for the frame trainers, same options, capacity rules and shuffle order
// as src/aslp-nnet/nnet-randomizer.{h,cc}, plus FrameDataReader (src/aslp-nnet/data-reader.cc:34-182, single
// feature / single target form).  The frame matrix stays on the device; the shuffle is one row-gather launch.
#ifndef ASLP_HOST_NNET_RANDOMIZER_H_
#define ASLP_HOST_NNET_RANDOMIZER_H_
#include <cstdlib>
#include "batch-feeder.h"
#include "matrix.h"
#include "nnet-loss.h"
#include "parse-options.h"
#include "table.h"

namespace kaldi {
namespace aslp_nnet {

struct NnetDataRandomizerOptions {
  int32 randomizer_size, randomizer_seed, minibatch_size;
  NnetDataRandomizerOptions() : randomizer_size(32768), randomizer_seed(777), minibatch_size(256) {}
  void Register(OptionsItf* opts) {
    opts->Register("randomizer-size", &randomizer_size, "Capacity of randomizer, length of concatenated utterances which are used for frame-level shuffling (in frames, affects memory consumption, max 8000000).");
    opts->Register("randomizer-seed", &randomizer_seed, "Seed value for srand, sets fixed order of frame-level shuffling");
    opts->Register("minibatch-size", &minibatch_size, "Size of a minibatch.");
  }
};

class RandomizerMask {
 public:
  RandomizerMask() {}
  explicit RandomizerMask(const NnetDataRandomizerOptions& conf) { Init(conf); }      // nnet-randomizer.h:74
  void Init(const NnetDataRandomizerOptions& conf) {
    KALDI_LOG << "Seeding by srand with : " << conf.randomizer_seed;
    srand(conf.randomizer_seed);
  }
  // the reference calls std::random_shuffle(begin, end) (nnet-randomizer.cc:41), which libstdc++ implements as
  // "for i = 1..n-1: swap(v[i], v[rand() % (i+1)])" on the C library's rand(); restated so that the frame order is
  // the same as the reference trainer's for the same --randomizer-seed
  const std::vector<int32>& Generate(int32 mask_size) {
    mask_.resize(mask_size);
    for (int32 i = 0; i < mask_size; ++i) mask_[i] = i;
    for (int32 i = 1; i < mask_size; ++i) {
      const int32 j = std::rand() % (i + 1);
      if (i != j) std::swap(mask_[i], mask_[j]);
    }
    return mask_;
  }
 private:
  std::vector<int32> mask_;
};

class MatrixRandomizer {
 public:
  MatrixRandomizer() : data_begin_(0), data_end_(0) {}
  explicit MatrixRandomizer(const NnetDataRandomizerOptions& conf) : data_begin_(0), data_end_(0) { Init(conf); }
  void Init(const NnetDataRandomizerOptions& conf) { conf_ = conf; }
  void AddData(const CuMatrixBase<BaseFloat>& m) {
    if (data_.NumCols() == 0) data_.Resize(conf_.randomizer_size, m.NumCols());
    if (data_begin_ > 0) {
      KALDI_ASSERT(data_begin_ <= data_end_);
      const int32 leftover = data_end_ - data_begin_;
      KALDI_ASSERT(leftover < data_begin_);
      if (leftover > 0) data_.RowRange(0, leftover).CopyFromMat(data_.RowRange(data_begin_, leftover));
      data_begin_ = 0; data_end_ = leftover;
      data_.RowRange(leftover, data_.NumRows() - leftover).SetZero();
    }
    if (data_.NumRows() < data_end_ + m.NumRows()) {
      CuMatrix<BaseFloat> aux(data_);
      data_.Resize(data_end_ + m.NumRows() + 1000, data_.NumCols());
      data_.RowRange(0, aux.NumRows()).CopyFromMat(aux);
    }
    data_.RowRange(data_end_, m.NumRows()).CopyFromMat(m);
    data_end_ += m.NumRows();
  }
  bool IsFull() const { return data_begin_ == 0 && data_end_ > conf_.randomizer_size; }
  int32 NumFrames() const { return data_end_; }
  void Randomize(const std::vector<int32>& mask) {
    KALDI_ASSERT(data_begin_ == 0 && data_end_ > 0 && data_end_ == static_cast<int32>(mask.size()));
    data_aux_ = data_;
    mask_dev_ = mask;
    ASLP_OK(aslp_copy_rows(CuStream(), data_.Data(), data_.Stride(), data_aux_.Data(), data_aux_.Stride(), mask_dev_.Data(),
                           static_cast<int>(mask.size()), data_.NumCols()));
  }
  bool Done() const { return data_end_ - data_begin_ < conf_.minibatch_size; }
  void Next() { data_begin_ += conf_.minibatch_size; }
  const CuMatrixBase<BaseFloat>& Value() {
    KALDI_ASSERT(data_end_ - data_begin_ >= conf_.minibatch_size);
    minibatch_.Resize(conf_.minibatch_size, data_.NumCols(), kUndefined);
    minibatch_.CopyFromMat(data_.RowRange(data_begin_, conf_.minibatch_size));
    return minibatch_;
  }
 private:
  CuMatrix<BaseFloat> data_, data_aux_, minibatch_;
  CuArrayInt mask_dev_;
  int32 data_begin_, data_end_;
  NnetDataRandomizerOptions conf_;
};

template <typename T>
class StdVectorRandomizer {
 public:
  StdVectorRandomizer() : data_begin_(0), data_end_(0) {}
  explicit StdVectorRandomizer(const NnetDataRandomizerOptions& conf) : data_begin_(0), data_end_(0) { Init(conf); }
  void Init(const NnetDataRandomizerOptions& conf) { conf_ = conf; }
  void AddData(const std::vector<T>& v) {
    if (data_.size() == 0) data_.resize(conf_.randomizer_size);
    if (data_begin_ > 0) {
      KALDI_ASSERT(data_begin_ <= data_end_);
      const int32 leftover = data_end_ - data_begin_;
      KALDI_ASSERT(leftover < data_begin_);
      if (leftover > 0) std::copy(data_.begin() + data_begin_, data_.begin() + data_begin_ + leftover, data_.begin());
      data_begin_ = 0; data_end_ = leftover;
    }
    if (data_.size() < data_end_ + v.size()) data_.resize(data_end_ + v.size() + 1000);
    std::copy(v.begin(), v.end(), data_.begin() + data_end_);
    data_end_ += static_cast<int32>(v.size());
  }
  bool IsFull() const { return data_begin_ == 0 && data_end_ > conf_.randomizer_size; }
  int32 NumFrames() const { return data_end_; }
  void Randomize(const std::vector<int32>& mask) {
    KALDI_ASSERT(data_begin_ == 0 && data_end_ > 0 && data_end_ == static_cast<int32>(mask.size()));
    std::vector<T> aux(data_);
    for (size_t i = 0; i < mask.size(); ++i) data_.at(i) = aux.at(mask.at(i));
  }
  bool Done() const { return data_end_ - data_begin_ < conf_.minibatch_size; }
  void Next() { data_begin_ += conf_.minibatch_size; }
  const std::vector<T>& Value() {
    KALDI_ASSERT(data_end_ - data_begin_ >= conf_.minibatch_size);
    minibatch_.assign(data_.begin() + data_begin_, data_.begin() + data_begin_ + conf_.minibatch_size);
    return minibatch_;
  }
 private:
  std::vector<T> data_, minibatch_;
  int32 data_begin_, data_end_;
  NnetDataRandomizerOptions conf_;
};
typedef StdVectorRandomizer<std::vector<std::pair<int32, BaseFloat>>> PosteriorRandomizer;
typedef StdVectorRandomizer<int32> Int32VectorRandomizer;

// per-frame weights (nnet-randomizer.h:105-138): the same begin / end arithmetic on a Vector<BaseFloat>
class VectorRandomizer {
 public:
  VectorRandomizer() {}
  explicit VectorRandomizer(const NnetDataRandomizerOptions& conf) : r_(conf) {}
  void Init(const NnetDataRandomizerOptions& conf) { r_.Init(conf); }
  void AddData(const Vector<BaseFloat>& v) { r_.AddData(std::vector<BaseFloat>(v.Data(), v.Data() + v.Dim())); }
  bool IsFull() const { return r_.IsFull(); }
  int32 NumFrames() const { return r_.NumFrames(); }
  void Randomize(const std::vector<int32>& mask) { r_.Randomize(mask); }
  bool Done() const { return r_.Done(); }
  void Next() { r_.Next(); }
  const Vector<BaseFloat>& Value() {
    const std::vector<BaseFloat>& m = r_.Value();
    minibatch_.Resize(static_cast<int32>(m.size()), kUndefined);
    for (size_t i = 0; i < m.size(); ++i) minibatch_(static_cast<int32>(i)) = m[i];
    return minibatch_;
  }
 private:
  StdVectorRandomizer<BaseFloat> r_;
  Vector<BaseFloat> minibatch_;
};

// features + frame targets -> shuffled minibatches (data-reader.cc:62-182)
// The begin / end arithmetic of MatrixRandomizer / StdVectorRandomizer without the data (nnet-randomizer.cc:60-134): which
// utterance closes a refill depends only on frame counts, so the batch feeder can group the utterances of the NEXT refill
// while the minibatches of the current one train.  tests/test_cpu_batch_feeder.py replays random tables against the real
// StdVectorRandomizer.
struct RandomizerReplay {
  int32 begin, end, randomizer_size, minibatch_size;
  RandomizerReplay(int32 randomizer_size_, int32 minibatch_size_) : begin(0), end(0), randomizer_size(randomizer_size_), minibatch_size(minibatch_size_) {}
  bool IsFull() const { return begin == 0 && end > randomizer_size; }
  void AddData(int32 rows) {
    if (begin > 0) { end -= begin; begin = 0; }          // the leftover moves to the front
    end += rows;
  }
  void ConsumeMinibatches() { while (end - begin >= minibatch_size) begin += minibatch_size; }     // Done() / Next() until used up
};

class FrameDataReaderSingle {
 public:
  FrameDataReaderSingle(const std::string& feature_rspecifier, const std::string& targets_rspecifier, const NnetDataRandomizerOptions& rand_opts)
      : feature_reader_(feature_rspecifier), targets_reader_(targets_rspecifier), read_done_(false), sim_(rand_opts.randomizer_size, rand_opts.minibatch_size),
        sim_read_done_(false), feeder_([this](Block* b) { return FillBlock(b); }, /*attach_device=*/false) {
    feature_randomizer_.Init(rand_opts);
    targets_randomizer_.Init(rand_opts);
    randomizer_mask_.Init(rand_opts);
  }
  bool Done() { return read_done_ && feature_randomizer_.Done(); }
  bool ReadData(const CuMatrixBase<BaseFloat>** feat, const Posterior** targets) {
    if (Done()) KALDI_ERR << "Already read done";
    if (feature_randomizer_.Done()) FillRandomizer();
    if (!Done()) {                         // even after a refill there may be less than one minibatch left
      *feat = &feature_randomizer_.Value();
      feature_randomizer_.Next();
      *targets = &targets_randomizer_.Value();
      targets_randomizer_.Next();
      return true;
    }
    return false;
  }
 private:
  // What one FillRandomizer call of the reference reads (data-reader.cc:118-160): the utterances up to the one that makes
  // the randomizer full, or to the end of the table.  The feeder thread (batch-feeder.h) prepares the NEXT refill while the
  // minibatches of the current one train: it replays the randomizer's begin / end arithmetic (which utterance closes a
  // refill depends only on frame counts) and concatenates the utterances into one page-locked block, so a refill costs one
  // H2D copy instead of one allocation + synchronous pageable copy per utterance.
  struct Block {
    PinnedMatrix feats;                    // rows of all utterances of the refill, in reading order
    std::vector<int32> utt_rows;
    std::vector<Posterior> targets;
    bool read_done = false;
  };
  bool FillBlock(Block* b) {               // feeder thread
    if (sim_read_done_) return false;
    b->feats.Resize(0, 0, kUndefined);
    b->utt_rows.clear(); b->targets.clear(); b->read_done = false;
    while (true) {
      if (sim_.IsFull()) break;
      if (feature_reader_.Done()) { b->read_done = sim_read_done_ = true; break; }
      const std::string utt = feature_reader_.Key();
      if (!targets_reader_.HasKey(utt)) {
        KALDI_WARN << utt << ", missing targets";
      } else {
        const Matrix<BaseFloat>& mat = feature_reader_.Value();
        const Posterior& targets = targets_reader_.Value(utt);
        if (static_cast<int32>(targets.size()) != mat.NumRows()) KALDI_ERR << "feature and target dim must match";
        b->feats.AppendRows(mat.Data(), mat.NumRows(), mat.NumCols());
        b->utt_rows.push_back(mat.NumRows());
        b->targets.push_back(targets);
        sim_.AddData(mat.NumRows());
      }
      feature_reader_.Next();
    }
    sim_.ConsumeMinibatches();                 // the minibatches of this refill
    return true;
  }
  void FillRandomizer() {
    Block* b = feeder_.Next();
    KALDI_ASSERT(b != nullptr);
    if (b->utt_rows.empty() && b->read_done) {
      // The table ended exactly at the previous refill's boundary: nothing to add, and less than one minibatch is left over.
      // The reference runs Randomize() here and dies on its data_begin_ == 0 assertion (nnet-randomizer.cc:70); this build
      // ends the epoch instead (Done() is now true).
      read_done_ = true;
      feeder_.Release(b);
      return;
    }
    if (b->feats.NumRows() > 0) {
      block_dev_.Resize(b->feats.NumRows(), b->feats.NumCols(), kUndefined);
      block_dev_.CopyFromHost(b->feats.Data(), b->feats.Stride());                          // asynchronous: the block is page-locked
    }
    int32 row = 0;
    for (size_t i = 0; i < b->utt_rows.size(); ++i) {                                       // the same AddData sequence as utterance by utterance
      feature_randomizer_.AddData(block_dev_.RowRange(row, b->utt_rows[i]));
      targets_randomizer_.AddData(b->targets[i]);
      row += b->utt_rows[i];
    }
    if (b->read_done) read_done_ = true;
    feeder_.Release(b);
    const std::vector<int32>& mask = randomizer_mask_.Generate(feature_randomizer_.NumFrames());
    feature_randomizer_.Randomize(mask);
    targets_randomizer_.Randomize(mask);
  }
  SequentialBaseFloatMatrixReader feature_reader_;
  RandomAccessPosteriorReader targets_reader_;
  MatrixRandomizer feature_randomizer_;
  PosteriorRandomizer targets_randomizer_;
  RandomizerMask randomizer_mask_;
  bool read_done_;
  RandomizerReplay sim_;                   // the feeder thread's replay of the randomizer's data_begin_ / data_end_
  bool sim_read_done_;
  CuMatrix<BaseFloat> block_dev_;
  BatchFeeder<Block> feeder_;              // last member: its thread uses everything above
};

// Several feature streams and several target streams shuffled by ONE mask (the multi-input / multi-output nets of
// aslp-nnet-train-frame-mimo.cc; data-reader.cc:18-125): every stream has its own table and randomizer, the utterances must
// come in the same order in every feature table, an utterance needs targets in every target table, and all streams of an
// utterance must have the same number of frames.  Reads utterance by utterance on the calling thread (no feeder).
class FrameDataReaderMulti {
 public:
  FrameDataReaderMulti(const std::vector<std::string>& feature_rspecifiers, const std::vector<std::string>& targets_rspecifiers,
                       const NnetDataRandomizerOptions& rand_opts) : read_done_(false) {
    KALDI_ASSERT(!feature_rspecifiers.empty() && !targets_rspecifiers.empty());
    for (const std::string& r : feature_rspecifiers) {
      feature_readers_.emplace_back(new SequentialBaseFloatMatrixReader(r));
      feature_randomizers_.emplace_back(new MatrixRandomizer());
      feature_randomizers_.back()->Init(rand_opts);
    }
    for (const std::string& r : targets_rspecifiers) {
      targets_readers_.emplace_back(new RandomAccessPosteriorReader(r));
      targets_randomizers_.emplace_back(new PosteriorRandomizer());
      targets_randomizers_.back()->Init(rand_opts);
    }
    randomizer_mask_.Init(rand_opts);
  }
  bool Done() { return read_done_ && feature_randomizers_[0]->Done(); }
  void ReadData(std::vector<const CuMatrixBase<BaseFloat>*>* input, std::vector<const Posterior*>* output) {
    if (Done()) KALDI_ERR << "Already read done";
    if (feature_randomizers_[0]->Done()) FillRandomizer();
    input->resize(feature_randomizers_.size());
    output->resize(targets_randomizers_.size());
    for (size_t i = 0; i < feature_randomizers_.size(); ++i) { (*input)[i] = &feature_randomizers_[i]->Value(); feature_randomizers_[i]->Next(); }
    for (size_t i = 0; i < targets_randomizers_.size(); ++i) { (*output)[i] = &targets_randomizers_[i]->Value(); targets_randomizers_[i]->Next(); }
  }
 private:
  void FillRandomizer() {
    while (true) {
      if (feature_randomizers_[0]->IsFull()) break;
      if (feature_readers_[0]->Done()) {
        for (size_t i = 1; i < feature_readers_.size(); ++i) KALDI_ASSERT(feature_readers_[i]->Done());
        read_done_ = true;
        break;
      }
      const std::string utt = feature_readers_[0]->Key();
      for (size_t i = 1; i < feature_readers_.size(); ++i)
        if (utt != feature_readers_[i]->Key())
          KALDI_ERR << "all feature not in the same order[0] " << utt << "[" << i << "] " << feature_readers_[i]->Key();
      bool all_have_target = true;
      for (auto& tr : targets_readers_)
        if (!tr->HasKey(utt)) { KALDI_WARN << utt << ", missing targets"; all_have_target = false; }
      if (all_have_target) {
        int32 num_frame = 0;
        for (size_t i = 0; i < feature_readers_.size(); ++i) {
          const Matrix<BaseFloat>& mat = feature_readers_[i]->Value();
          if (i == 0) num_frame = mat.NumRows();
          else if (mat.NumRows() != num_frame) KALDI_ERR << "all feature dim not equal";
          feature_randomizers_[i]->AddData(CuMatrix<BaseFloat>(mat));
        }
        for (size_t i = 0; i < targets_readers_.size(); ++i) {
          const Posterior& targets = targets_readers_[i]->Value(utt);
          if (static_cast<int32>(targets.size()) != num_frame) KALDI_ERR << "feature and target dim must match";
          targets_randomizers_[i]->AddData(targets);
        }
      }
      for (auto& fr : feature_readers_) fr->Next();
    }
    const std::vector<int32>& mask = randomizer_mask_.Generate(feature_randomizers_[0]->NumFrames());
    for (auto& r : feature_randomizers_) r->Randomize(mask);
    for (auto& r : targets_randomizers_) r->Randomize(mask);
  }
  std::vector<std::unique_ptr<SequentialBaseFloatMatrixReader>> feature_readers_;
  std::vector<std::unique_ptr<RandomAccessPosteriorReader>> targets_readers_;
  std::vector<std::unique_ptr<MatrixRandomizer>> feature_randomizers_;
  std::vector<std::unique_ptr<PosteriorRandomizer>> targets_randomizers_;
  RandomizerMask randomizer_mask_;
  bool read_done_;
};

// The reference's one class with both constructors (data-reader.h:24-47): a single feature / target pair takes the
// feeder-backed reader above, lists take the multi-stream one.
class FrameDataReader {
 public:
  FrameDataReader(const std::string& feature_rspecifier, const std::string& targets_rspecifier, const NnetDataRandomizerOptions& rand_opts)
      : single_(new FrameDataReaderSingle(feature_rspecifier, targets_rspecifier, rand_opts)) {}
  FrameDataReader(const std::vector<std::string>& feature_rspecifiers, const std::vector<std::string>& targets_rspecifiers,
                  const NnetDataRandomizerOptions& rand_opts) : multi_(new FrameDataReaderMulti(feature_rspecifiers, targets_rspecifiers, rand_opts)) {}
  bool Done() { return single_ ? single_->Done() : multi_->Done(); }
  bool ReadData(const CuMatrixBase<BaseFloat>** feat, const Posterior** targets) {
    if (single_) return single_->ReadData(feat, targets);
    std::vector<const CuMatrixBase<BaseFloat>*> in;
    std::vector<const Posterior*> out;
    multi_->ReadData(&in, &out);
    *feat = in[0]; *targets = out[0];
    return true;
  }
  void ReadData(std::vector<const CuMatrixBase<BaseFloat>*>* input, std::vector<const Posterior*>* output) {
    if (multi_) { multi_->ReadData(input, output); return; }
    input->resize(1); output->resize(1);
    if (!single_->ReadData(&(*input)[0], &(*output)[0])) KALDI_ERR << "Already read done";
  }
 private:
  std::unique_ptr<FrameDataReaderSingle> single_;
  std::unique_ptr<FrameDataReaderMulti> multi_;
};


// multi-stream truncated-BPTT batches with target delay (data-reader.h:49-98, data-reader.cc:184-341)
struct SequenceDataReaderOptions {
  int32 batch_size, num_stream, drop_len, skip_width, targets_delay, length_tolerance;
  double frame_limit;
  SequenceDataReaderOptions() : batch_size(20), num_stream(100), drop_len(0), skip_width(1), targets_delay(5), length_tolerance(5), frame_limit(100000) {}
  void Register(OptionsItf* opts) {
    opts->Register("batch-size", &batch_size, "--LSTM-- BPTT batch_size");
    opts->Register("num-stream", &num_stream, "--LSTM-- BPTT multistream training");
    opts->Register("drop-len", &drop_len, "if Sentence frame length greater than drop_len,then drop it, default(0, no drop)");
    opts->Register("skip-width", &skip_width, "num of frame for one skip(default 1, no skip)");
    opts->Register("targets-delay", &targets_delay, "--LSTM-- BPTT targets delay");
    opts->Register("length-tolerance", &length_tolerance, "Allowed length difference of features/targets (frames),for the whole utterance training");
    opts->Register("frame-limit", &frame_limit, "Max number of frames to be processed for whole utterance training");
  }
};

class SequenceDataReader {
 public:
  SequenceDataReader(const std::string& feature_rspecifier, const std::string& targets_rspecifier, const SequenceDataReaderOptions& opts)
      : opts_(opts), read_done_(false), feature_reader_(feature_rspecifier), target_reader_(targets_rspecifier),
        curt_(opts.num_stream, 0), lent_(opts.num_stream, 0), new_utt_flags_(opts.num_stream, 0), keys_(opts.num_stream),
        feats_(opts.num_stream), targets_(opts.num_stream) {}
  bool Done() { return read_done_ && feature_reader_.Done(); }
  const std::vector<int32>& GetNewUttFlags() const { return new_utt_flags_; }
  // feat is left untouched once every stream is exhausted (the reference's FillBatchBuff skips the copy, data-reader.cc:297-322:
  // the trainer then runs one more minibatch on the previous features with an all-zero mask -- kept)
  void ReadData(CuMatrix<BaseFloat>* feat, Posterior* target, Vector<BaseFloat>* frame_mask) {
    if (Done()) KALDI_ERR << "Already read done!";
    AddNewUtt();
    if (FillBatchBuff(&host_, target, frame_mask)) *feat = host_;
  }
  // the same minibatch into a page-locked host matrix (the batch feeder's slot); false: every stream is exhausted and
  // `feat` was not written (the caller keeps its previous device matrix, as above)
  bool ReadDataHost(PinnedMatrix* feat, Posterior* target, Vector<BaseFloat>* frame_mask) {
    if (Done()) KALDI_ERR << "Already read done!";
    AddNewUtt();
    return FillBatchBuff(feat, target, frame_mask);
  }
 private:
  void AddNewUtt() {
    for (int32 s = 0; s < opts_.num_stream; s++) {
      if (curt_[s] < lent_[s]) { new_utt_flags_[s] = 0; continue; }
      while (!feature_reader_.Done()) {
        const std::string key = feature_reader_.Key();
        const Matrix<BaseFloat>& mat = feature_reader_.Value();
        if (opts_.drop_len > 0 && mat.NumRows() > opts_.drop_len) { KALDI_WARN << key << ", too long, droped"; feature_reader_.Next(); continue; }
        if (!target_reader_.HasKey(key)) { KALDI_WARN << key << ", missing targets"; feature_reader_.Next(); continue; }
        const Posterior& target = target_reader_.Value(key);
        if (mat.NumRows() != static_cast<int32>(target.size())) {
          KALDI_WARN << key << ", length miss-match between feats and targers, skip";
          feature_reader_.Next();
          continue;
        }
        if (opts_.skip_width > 1) {
          const int32 skip_len = (mat.NumRows() - 1) / opts_.skip_width + 1;
          feats_[s].Resize(skip_len, mat.NumCols());
          targets_[s].assign(skip_len, Posterior::value_type());
          for (int32 i = 0; i < skip_len; i++) {
            std::copy(mat.RowData(i * opts_.skip_width), mat.RowData(i * opts_.skip_width) + mat.NumCols(), feats_[s].RowData(i));
            targets_[s][i] = target[i * opts_.skip_width];
          }
        } else {
          feats_[s] = mat;
          targets_[s] = target;
        }
        keys_[s] = key;
        curt_[s] = 0;
        lent_[s] = feats_[s].NumRows();
        new_utt_flags_[s] = 1;
        feature_reader_.Next();
        break;
      }
    }
  }
  template <class HostMat>
  bool FillBatchBuff(HostMat* host, Posterior* target, Vector<BaseFloat>* frame_mask) {
    const int32 num_stream = opts_.num_stream, batch_size = opts_.batch_size, delay = opts_.targets_delay;
    for (int32 s = 0; s < num_stream; s++) {
      if (curt_[s] < lent_[s]) { read_done_ = false; break; }
      read_done_ = true;
    }
    const int32 feat_dim = feats_[0].NumCols();
    if (!read_done_) host->Resize(batch_size * num_stream, feat_dim, kSetZero);
    target->assign(batch_size * num_stream, Posterior::value_type());
    frame_mask->Resize(batch_size * num_stream);
    if (read_done_) return false;
    for (int32 t = 0; t < batch_size; t++) {
      for (int32 s = 0; s < num_stream; s++) {
        const int32 row = t * num_stream + s;
        if (lent_[s] == 0) { curt_[s]++; continue; }      // a stream that never received an utterance (the reference indexes [-1] here)
        if (curt_[s] < lent_[s]) { (*frame_mask)(row) = 1; (*target)[row] = targets_[s][curt_[s]]; }
        else { (*frame_mask)(row) = 0; (*target)[row] = targets_[s][lent_[s] - 1]; }
        const int32 src = (curt_[s] + delay < lent_[s]) ? curt_[s] + delay : lent_[s] - 1;    // shifted by the target delay, padded with the last frame
        std::copy(feats_[s].RowData(src), feats_[s].RowData(src) + feat_dim, host->RowData(row));
        curt_[s]++;
      }
    }
    return true;
  }
  SequenceDataReaderOptions opts_;
  bool read_done_;
  SequentialBaseFloatMatrixReader feature_reader_;
  RandomAccessPosteriorReader target_reader_;
  std::vector<int32> curt_, lent_, new_utt_flags_;
  std::vector<std::string> keys_;
  std::vector<Matrix<BaseFloat>> feats_;
  std::vector<Posterior> targets_;
  Matrix<BaseFloat> host_;
};

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
