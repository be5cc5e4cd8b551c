// Include-path shim: the reference spells this header "aslp-nnet/ctc-loss.h" (src/aslp-nnet/ctc-loss.h); here it is host/nnet-loss.h.
#include "../../host/nnet-loss.h"
