// Include-path shim: the reference spells this header "aslp-cudamatrix/cu-matrix.h" (src/aslp-cudamatrix/cu-matrix.h); here it is host/matrix.h.
#include "../../host/matrix.h"
