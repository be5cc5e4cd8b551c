// Host-only check of RandomizerReplay (kaldi-aslp_b200/host/nnet-randomizer.h): the utterance grouping the batch feeder
// predicts for the frame trainer's randomizer refills must be the grouping the reference's loop produces when it asks the
// real randomizer (data-reader.cc:86-125: IsFull() before every utterance, Randomize, minibatches until Done()).  The
// reference side runs on StdVectorRandomizer<int>, which shares MatrixRandomizer's begin / end arithmetic and needs no device.
#include <cstdio>
#include <vector>
#include "nnet-randomizer.h"

using namespace kaldi;
using namespace kaldi::aslp_nnet;

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int main() {
  unsigned seed = 12345;
  int trials = 0, boundary_cases = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    NnetDataRandomizerOptions conf;
    conf.minibatch_size = 4 + lcg(&seed) % 60;
    conf.randomizer_size = conf.minibatch_size + lcg(&seed) % 400;
    const int n = 1 + lcg(&seed) % 60;
    std::vector<int> len(n);
    for (int i = 0; i < n; ++i) len[i] = 1 + lcg(&seed) % 150;
    if (trial % 10 == 0) for (int i = 0; i < n; ++i) len[i] = conf.minibatch_size * (1 + i % 3);   // tables that end on refill boundaries

    std::vector<std::vector<int>> want, got;
    {   // the reference loop on the real randomizer
      StdVectorRandomizer<int> r;
      r.Init(conf);
      int next = 0;
      bool read_done = false;
      while (!(read_done && r.Done())) {
        if (r.Done()) {
          std::vector<int> grp;
          while (true) {
            if (r.IsFull()) break;
            if (next == n) { read_done = true; break; }
            r.AddData(std::vector<int>(len[next], next));
            grp.push_back(next++);
          }
          want.push_back(grp);
          if (grp.empty()) { ++boundary_cases; break; }         // nothing added: the reference asserts in Randomize here, this build ends the epoch
          std::vector<int32> mask(r.NumFrames());
          for (size_t i = 0; i < mask.size(); ++i) mask[i] = static_cast<int32>(i);
          r.Randomize(mask);
        }
        while (!r.Done()) r.Next();
      }
    }
    {   // the feeder's prediction (FrameDataReader::FillBlock)
      RandomizerReplay sim(conf.randomizer_size, conf.minibatch_size);
      int next = 0;
      bool read_done = false;
      while (!read_done) {
        std::vector<int> grp;
        while (true) {
          if (sim.IsFull()) break;
          if (next == n) { read_done = true; break; }
          sim.AddData(len[next]);
          grp.push_back(next++);
        }
        sim.ConsumeMinibatches();
        got.push_back(grp);
      }
    }
    if (want != got) {
      std::printf("FAIL trial %d: %zu refills predicted, %zu in the reference loop (minibatch %d, randomizer %d, %d utterances)\n", trial, got.size(),
                  want.size(), conf.minibatch_size, conf.randomizer_size, n);
      return 2;
    }
    ++trials;
  }
  std::printf("OK trials=%d boundary_cases=%d\n", trials, boundary_cases);
  return 0;
}
