// include/dropin/params.hpp -- stands in for the reference's params.hpp (KYBER_K selection :8-10, MPCITH_* :12-37); the macros
// are defined by ../kosk_dropin.hpp.  Select the parameter set with -DKYBER_K=2|3|4 (default 2, as in the reference).
#ifndef KOSK_DROPIN_PARAMS_HPP
#define KOSK_DROPIN_PARAMS_HPP
#include "../kosk_dropin.hpp"
#endif
