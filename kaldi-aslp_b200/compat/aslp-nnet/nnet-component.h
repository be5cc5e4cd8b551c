// Include-path shim: the reference spells this header "aslp-nnet/nnet-component.h" (src/aslp-nnet/nnet-component.h); here it is host/nnet-component.h.
#include "../../host/nnet-component.h"
