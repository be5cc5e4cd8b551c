// aslp-nnet-init -- same command line as src/aslp-nnetbin/aslp-nnet-init.cc:30-72 (the recipes also call it
// aslp-nnet-initialize): srand(seed), Nnet::Init(prototype), write the model.
#include "nnet-nnet.h"
#include "parse-options.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Initialize Neural Network parameters according to a prototype (aslp_nnet).\n"
        "Usage:  aslp-nnet-initialize [options] <nnet-prototype-in> <nnet-out>\n"
        "e.g.:\n"
        " aslp-nnet-initialize --binary=false nnet.proto nnet.init\n";
    ParseOptions po(usage);
    bool binary_write = true;
    po.Register("binary", &binary_write, "Write output in binary mode");
    int32 seed = 777;
    po.Register("seed", &seed, "Seed for random number generator");
    po.Read(argc, argv);
    if (po.NumArgs() != 2) { po.PrintUsage(); return 1; }
    const std::string nnet_config_in_filename = po.GetArg(1), nnet_out_filename = po.GetArg(2);
    std::srand(seed);
    Nnet nnet;
    nnet.Init(nnet_config_in_filename);
    nnet.Write(nnet_out_filename, binary_write);
    KALDI_LOG << "Written initialized model to " << nnet_out_filename;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what() << '\n';
    return -1;
  }
}
