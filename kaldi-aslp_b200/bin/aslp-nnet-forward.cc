// aslp-nnet-forward -- src/aslp-nnetbin/aslp-nnet-forward.cc; body in forward-main.h
#include "forward-main.h"
int main(int argc, char* argv[]) { return kaldi::aslp_nnet::ForwardMain(argc, argv, /*split_skip=*/false); }
