// include/dropin/mlwe_prover.hpp -- stands in for the reference's mlwe_prover.hpp (types :34-75, functions :50-55, :77-99) when a
// program written against the reference is compiled with -Iinclude/dropin and linked with libkosk_b200.so.
// Everything lives in ../kosk_dropin.hpp; NTL is not needed (the reference's prover never calls it).
#ifndef KOSK_DROPIN_MLWE_PROVER_HPP
#define KOSK_DROPIN_MLWE_PROVER_HPP
#include "../kosk_dropin.hpp"
#include <time.h>
namespace NTL {}      // so that the reference's `using namespace NTL;` (main.cpp:11) still compiles
#endif
