// Include-path shim: the reference spells this header "fstext/fstext-lib.h" (src/fstext/fstext-lib.h); here it is host/base.h.
#include "../../host/base.h"
