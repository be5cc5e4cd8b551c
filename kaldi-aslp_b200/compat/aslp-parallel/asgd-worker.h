// Include-path shim: the reference spells this header "aslp-parallel/asgd-worker.h" (src/aslp-parallel/asgd-worker.h); here it is host/parallel-async.h.
#include "../../host/parallel-async.h"
