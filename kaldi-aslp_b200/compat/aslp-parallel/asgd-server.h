// Include-path shim: the reference spells this header "aslp-parallel/asgd-server.h" (src/aslp-parallel/asgd-server.h); here it is host/parallel-async.h.
#include "../../host/parallel-async.h"
