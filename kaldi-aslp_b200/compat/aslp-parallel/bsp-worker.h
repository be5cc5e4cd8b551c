// Include-path shim: the reference spells this header "aslp-parallel/bsp-worker.h" (src/aslp-parallel/bsp-worker.h); here it is host/parallel.h.
#include "../../host/parallel.h"
