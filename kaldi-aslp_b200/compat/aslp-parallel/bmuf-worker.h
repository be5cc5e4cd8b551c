// Include-path shim: the reference spells this header "aslp-parallel/bmuf-worker.h" (src/aslp-parallel/bmuf-worker.h); here it is host/parallel.h.
#include "../../host/parallel.h"
