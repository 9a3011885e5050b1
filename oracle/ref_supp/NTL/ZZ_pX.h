/* oracle/ref_supp/NTL/ZZ_pX.h -- TEST INFRASTRUCTURE: see ZZ.h (single stand-in header). */
#include "ZZ.h"
