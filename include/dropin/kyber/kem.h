/* include/dropin/kyber/kem.h -- placeholder for the reference's kyber/kem.h.  The reference includes it inside extern "C"
 * (main.cpp:3-6); crypto_kem_enc / crypto_kem_dec and KYBER_SSBYTES / KYBER_CIPHERTEXTBYTES are provided by kosk_dropin.hpp,
 * which the program reaches through mlwe_prover.hpp / kosk.hpp. */
#ifndef KOSK_DROPIN_KYBER_KEM_H
#define KOSK_DROPIN_KYBER_KEM_H
#include <stdint.h>
#endif
