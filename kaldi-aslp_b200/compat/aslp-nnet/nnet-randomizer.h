// Include-path shim: the reference spells this header "aslp-nnet/nnet-randomizer.h" (src/aslp-nnet/nnet-randomizer.h); here it is host/nnet-randomizer.h.
#include "../../host/nnet-randomizer.h"
