// Include-path shim: the reference spells this header "aslp-parallel/easgd-server.h" (src/aslp-parallel/easgd-server.h); here it is host/parallel-async.h.
#include "../../host/parallel-async.h"
