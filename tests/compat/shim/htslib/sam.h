/* shim: htslib/sam.h -> the htslib-compatible iterator tier of this repository */
#include "b200_htslib_compat.h"
