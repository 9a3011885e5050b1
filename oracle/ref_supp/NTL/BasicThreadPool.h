/* oracle/ref_supp/NTL/BasicThreadPool.h -- TEST INFRASTRUCTURE: see ZZ.h (single stand-in header). */
#include "ZZ.h"
