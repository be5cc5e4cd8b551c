// Include-path shim: the reference spells this header "aslp-parallel/sod-worker.h" (src/aslp-parallel/sod-worker.h); here it is host/parallel.h.
#include "../../host/parallel.h"
