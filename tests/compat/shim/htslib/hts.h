/* shim: the pileup callers only need what b200_htslib_compat.h declares */
#include "b200_htslib_compat.h"
