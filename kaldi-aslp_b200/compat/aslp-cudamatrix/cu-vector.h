// Include-path shim: the reference spells this header "aslp-cudamatrix/cu-vector.h" (src/aslp-cudamatrix/cu-vector.h); here it is host/matrix.h.
#include "../../host/matrix.h"
