// Include-path shim: the reference spells this header "aslp-nnet/nnet-trnopts.h" (src/aslp-nnet/nnet-trnopts.h); here it is host/nnet-trnopts.h.
#include "../../host/nnet-trnopts.h"
