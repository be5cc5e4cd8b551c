// Host-only check of BatchFeeder's hand-over logic (kaldi-aslp_b200/host/batch-feeder.h): delivery order, slot reuse, the
// end-of-data signal, Join() making the fill lambda's captured state readable, and an exception thrown on the feeder thread
// surfacing in Next().  No device: track_copies = false.  Driven by tests/test_cpu_batch_feeder.py with ASLP_FEEDER_DEPTH.
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <thread>
#include "batch-feeder.h"

struct Batch { int seq = -1; std::vector<int> payload; };

int main(int argc, char** argv) {
  using kaldi::BatchFeeder;
  const int n = argc > 1 ? std::atoi(argv[1]) : 50;
  const int fail_at = argc > 2 ? std::atoi(argv[2]) : -1;
  int produced = 0, skipped_after_last = 0;
  auto fill = [&](Batch* b) -> bool {
    if (produced == fail_at) throw std::runtime_error("fill failed on purpose");
    if (produced == n) { skipped_after_last = 7; return false; }   // e.g. utterances skipped after the last minibatch
    b->seq = produced;
    b->payload.assign(100 + produced % 5, produced);
    if (produced % 7 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(2));   // a slow read now and then
    ++produced;
    return true;
  };
  try {
    BatchFeeder<Batch> feeder(fill, /*attach_device=*/false, /*track_copies=*/false);
    int next = 0;
    long long sum = 0;
    while (Batch* b = feeder.Next()) {
      if (b->seq != next) { std::printf("FAIL order: got %d want %d\n", b->seq, next); return 2; }
      for (int v : b->payload) if (v != next) { std::printf("FAIL payload of %d\n", next); return 2; }
      if (static_cast<int>(b->payload.size()) != 100 + next % 5) { std::printf("FAIL size of %d\n", next); return 2; }
      sum += b->seq;
      if (next % 5 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));     // a slow consumer now and then
      feeder.Release(b);
      ++next;
    }
    feeder.Join();
    if (next != n || skipped_after_last != 7) { std::printf("FAIL end: delivered %d of %d, tail %d\n", next, n, skipped_after_last); return 2; }
    std::printf("OK depth=%d delivered=%d sum=%lld\n", feeder.depth(), next, sum);
    return 0;
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 3;
  }
}
