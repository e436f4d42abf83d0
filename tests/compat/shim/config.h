/* shim for compiling UNMODIFIED reference sources against include/b200_htslib_compat.h (tests only) */
