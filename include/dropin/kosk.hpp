// include/dropin/kosk.hpp -- stands in for the reference's kosk.hpp (kyber_keypair, kyber_keygen, kyber_verifiable_keygen,
// kyber_kosk_verify; :13-24); see ../kosk_dropin.hpp.
#ifndef KOSK_DROPIN_KOSK_HPP
#define KOSK_DROPIN_KOSK_HPP
#include "mlwe_prover.hpp"
#include "mlwe_verifier.hpp"
#endif
