// Include-path shim: the reference spells this header "util/common-utils.h" (src/util/common-utils.h); here it is host/table.h.
#include "../../host/table.h"
