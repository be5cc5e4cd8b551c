// Include-path shim: the reference spells this header "aslp-nnet/nnet-nnet.h" (src/aslp-nnet/nnet-nnet.h); here it is host/nnet-nnet.h.
#include "../../host/nnet-nnet.h"
