// batch-feeder.h -- double-buffered minibatch feeding for the trainer mains.
// Reference: the trainers read, filter and pack a minibatch on the one host thread that also drives the device, and copy it
// from pageable memory (src/aslp-nnetbin/aslp-nnet-train-blstm-streams-lc.cc:185-268, aslp-nnet-train-warp-ctc-streams.cc:
// 116-175, src/aslp-nnet/data-reader.cc:200-324).  Here a feeder thread runs the SAME sequential reading / packing code
// (so curt / lent / new_utt_flags, skipped-utterance counters and the packed rows are what the reference computes, in the
// same order) one or two minibatches ahead into page-locked slots; the consumer issues one asynchronous H2D copy per
// minibatch and hands the slot back once the copy has completed on the device.
#ifndef ASLP_HOST_BATCH_FEEDER_H_
#define ASLP_HOST_BATCH_FEEDER_H_
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "matrix.h"

namespace kaldi {

// Batch: default-constructible, not copied.  fill(Batch*) runs on the feeder thread and returns false when there is no
// further minibatch (that slot is then not delivered).  ASLP_FEEDER_DEPTH (default 2) slots; 0 runs fill() inline on the
// calling thread (no thread, same results).
template <class Batch>
class BatchFeeder {
 public:
  typedef std::function<bool(Batch*)> FillFn;
  // attach_device: the feeder thread launches device work of its own (a feature transform) and needs a stream.
  // track_copies = false: Release() hands the slot straight back (the consumer made no asynchronous device copy out of it;
  // also what the host-only unit test of the hand-over logic uses, tests/test_cpu_batch_feeder.py)
  BatchFeeder(FillFn fill, bool attach_device, bool track_copies = true)
      : fill_(fill), attach_(attach_device), track_(track_copies), depth_(Depth()), done_(false), stop_(false) {
    const int n = depth_ > 0 ? depth_ : 1;
    for (int i = 0; i < n; ++i) slots_.emplace_back(new Slot());
    if (depth_ > 0) {
      for (auto& s : slots_) free_.push_back(s.get());
      thread_ = std::thread([this] { Run(); });
    }
  }
  ~BatchFeeder() {
    Join();
    for (auto& s : slots_) if (s->copied != nullptr) { aslp_event_destroy(s->copied); s->copied = nullptr; }
  }
  // stops the feeder thread (after the fill() it may be in); what fill() wrote through captured references -- e.g. the
  // skipped-utterance counters, which also count utterances skipped after the last delivered minibatch -- is then safe to read
  void Join() {
    if (thread_.joinable()) {
      { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
      cv_.notify_all();
      thread_.join();
    }
  }
  // the next minibatch in reading order, or nullptr at the end of the data; an exception thrown by fill() is rethrown here
  Batch* Next() {
    if (depth_ == 0) {
      Slot* s = slots_[0].get();
      WaitCopied(s);
      return fill_(&s->batch) ? &s->batch : nullptr;
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (ready_.empty()) {
      if (error_) std::rethrow_exception(error_);
      return nullptr;
    }
    Slot* s = ready_.front();
    ready_.pop_front();
    return &s->batch;
  }
  // the consumer has issued its last read of the slot's host memory on the current stream (the H2D copy): the slot is
  // refilled only after that copy has completed on the device
  void Release(Batch* b) {
    Slot* s = FindSlot(b);
    if (track_) {
      ASLP_OK(aslp_event_record(CuStream(), &s->copied));
      s->pending = true;
    }
    if (depth_ == 0) return;
    { std::lock_guard<std::mutex> lk(mu_); free_.push_back(s); }
    cv_.notify_all();
  }
  int depth() const { return depth_; }

 private:
  struct Slot {
    Batch batch;
    void* copied = nullptr;      // CUDA event: the consumer's copy out of this slot
    bool pending = false;
  };
  static int Depth() {
    const char* e = std::getenv("ASLP_FEEDER_DEPTH");
    if (e == nullptr || *e == '\0') return 2;
    const int d = std::atoi(e);
    return d < 0 ? 0 : (d > 8 ? 8 : d);
  }
  Slot* FindSlot(Batch* b) {
    for (auto& s : slots_) if (&s->batch == b) return s.get();
    KALDI_ERR << "BatchFeeder::Release: not a batch of this feeder";
    return nullptr;
  }
  static void WaitCopied(Slot* s) {
    if (s->pending) { ASLP_OK(aslp_event_sync(s->copied)); s->pending = false; }
  }
  void Run() {
    try {
      if (attach_) CuThreadAttach();
      else if (track_) CuThreadUseDevice();      // pinned allocations and event waits of this thread belong to the process's GPU
      for (;;) {
        Slot* s = nullptr;
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [this] { return !free_.empty() || stop_; });
          if (stop_) break;
          s = free_.front();
          free_.pop_front();
        }
        WaitCopied(s);
        const bool more = fill_(&s->batch);
        {
          std::lock_guard<std::mutex> lk(mu_);
          if (more) ready_.push_back(s); else done_ = true;
        }
        cv_.notify_all();
        if (!more) break;
      }
    } catch (...) {
      std::lock_guard<std::mutex> lk(mu_);
      error_ = std::current_exception();
      done_ = true;
      cv_.notify_all();
    }
    if (attach_) CuThreadDetach();
  }

  FillFn fill_;
  bool attach_, track_;
  int depth_;
  std::vector<std::unique_ptr<Slot>> slots_;
  std::deque<Slot*> free_, ready_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool done_, stop_;
  std::exception_ptr error_;
  std::thread thread_;
};

}  // namespace kaldi
#endif
