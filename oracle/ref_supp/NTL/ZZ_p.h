/* oracle/ref_supp/NTL/ZZ_p.h -- TEST INFRASTRUCTURE: see ZZ.h (single stand-in header). */
#include "ZZ.h"
