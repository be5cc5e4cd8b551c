// Include-path shim: the reference spells this header "aslp-nnet/nnet-loss.h" (src/aslp-nnet/nnet-loss.h); here it is host/nnet-loss.h.
#include "../../host/nnet-loss.h"
