// Include-path shim: the reference spells this header "aslp-nnet/nnet-pdf-prior.h" (src/aslp-nnet/nnet-pdf-prior.h); here it is host/nnet-pdf-prior.h.
#include "../../host/nnet-pdf-prior.h"
