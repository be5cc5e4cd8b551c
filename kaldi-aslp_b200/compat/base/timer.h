// Include-path shim: the reference spells this header "base/timer.h" (src/base/timer.h); here it is host/base.h.
#include "../../host/base.h"
