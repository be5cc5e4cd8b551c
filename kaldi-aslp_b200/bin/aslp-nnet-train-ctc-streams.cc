// aslp-nnet-train-ctc-streams -- the Eesen-CTC variant of the multi-stream CTC trainer (src/aslp-nnetbin/aslp-nnet-train-ctc-streams.cc):
// same loop as aslp-nnet-train-warp-ctc-streams with Ctc::EvalParallel / ErrorRateMSeq as the loss (see ctc-streams-main.h).
#include "ctc-streams-main.h"

namespace {
struct EesenCtcAdapter {
  kaldi::aslp_nnet::Ctc c;
  void SetReportStep(int s) { c.SetReportStep(s); }
  void Eval(const std::vector<std::string>& k, const std::vector<kaldi::int32>& f, const kaldi::CuMatrixBase<kaldi::BaseFloat>& o,
            std::vector<std::vector<kaldi::int32>>& l, kaldi::CuMatrix<kaldi::BaseFloat>* d) { c.EvalParallel(k, f, o, l, d); }
  void ErrorRate(const std::vector<int>& f, const kaldi::CuMatrixBase<kaldi::BaseFloat>& o, std::vector<std::vector<int>>& l) { c.ErrorRateMSeq(f, o, l); }
  std::string Report() { return c.Report(); }
};
}  // namespace

int main(int argc, char* argv[]) {
  return kaldi::aslp_nnet::CtcStreamsMain<EesenCtcAdapter>(argc, argv,
      "Perform one iteration of CTC training by SGD.\n"
      "The updates are done per-utternace and by processing multiple utterances in parallel.\n"
      "\n"
      "Usage: aslp-nnet-train-ctc-streams [options] <feature-rspecifier> <labels-rspecifier> <model-in> [<model-out>]\n"
      "e.g.: \n"
      "aslp-nnet-train-ctc-streams scp:feature.scp ark:labels.ark nnet.init nnet.iter1\n");
}
