// Include-path shim: the reference spells this header "aslp-parallel/itf.h" (src/aslp-parallel/itf.h); here it is host/parallel.h.
#include "../../host/parallel.h"
