// aslp-nnet-forward-skip -- src/aslp-nnetbin/aslp-nnet-forward-skip.cc (one pass of the net per skip offset, outputs scattered
// back so that every frame has its own posterior); body in forward-main.h
#include "forward-main.h"
int main(int argc, char* argv[]) { return kaldi::aslp_nnet::ForwardMain(argc, argv, /*split_skip=*/true); }
