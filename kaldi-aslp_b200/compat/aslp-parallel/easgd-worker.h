// Include-path shim: the reference spells this header "aslp-parallel/easgd-worker.h" (src/aslp-parallel/easgd-worker.h); here it is host/parallel-async.h.
#include "../../host/parallel-async.h"
