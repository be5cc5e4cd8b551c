// Include-path shim: the reference spells this header "aslp-nnet/warp-ctc.h" (src/aslp-nnet/warp-ctc.h); here it is host/nnet-loss.h.
#include "../../host/nnet-loss.h"
