#ifndef HTSLIB_COMPAT_KSTRING_H
#define HTSLIB_COMPAT_KSTRING_H
#include "hts.h"
#endif
