// Include-path shim: the reference spells this header "aslp-nnet/data-reader.h" (src/aslp-nnet/data-reader.h); here it is host/nnet-randomizer.h.
#include "../../host/nnet-randomizer.h"
