// Include-path shim: the reference spells this header "base/kaldi-common.h" (src/base/kaldi-common.h); here it is host/base.h.
#include "../../host/base.h"
