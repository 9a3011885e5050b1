// include/dropin/mlwe_verifier.hpp -- stands in for the reference's mlwe_verifier.hpp (verify(), :14-15); see ../kosk_dropin.hpp.
#ifndef KOSK_DROPIN_MLWE_VERIFIER_HPP
#define KOSK_DROPIN_MLWE_VERIFIER_HPP
#include "mlwe_prover.hpp"
#endif
