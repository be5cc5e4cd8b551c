// Include-path shim: the reference spells this header "aslp-cudamatrix/cu-device.h" (src/aslp-cudamatrix/cu-device.h); here it is host/matrix.h.
#include "../../host/matrix.h"
