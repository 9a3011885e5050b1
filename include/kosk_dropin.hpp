// kosk_dropin.hpp -- source-level drop-in for the reference's kosk.hpp (reference kosk.hpp:13-24).
//
// A program written against the reference API
//     kyber_keypair kp; uint8_t pi[MPCITH_PROOF_SIZE];
//     kyber_verifiable_keygen(&kp, pi);                 // kosk.cpp:72-86
//     bool ok = kyber_kosk_verify(pi, kp.pk);           // kosk.cpp:88-117
// compiles unchanged against this header (same type name, function names, signatures, KYBER_K selection by
// macro as in params.hpp:8-10 / kyber/params.h:4-6) and links against libkosk_b200.so instead of the
// reference's objects.  Randomness: the reference pulls from the OS through randombytes()
// (kyber/randombytes.c:43-57); here 32 bytes from getrandom() seed the KOSK counter-mode DRBG
// (include/kosk_b200.h).  kosk_dropin_set_seed() injects a fixed seed for reproducible runs.
#ifndef KOSK_DROPIN_HPP
#define KOSK_DROPIN_HPP

#include <stdint.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>
#include "kosk_b200.h"

#ifndef KYBER_K
#define KYBER_K 2 /* Change this for different security strengths (params.hpp:8-10) */
#endif
#if KYBER_K != 2 && KYBER_K != 3 && KYBER_K != 4
#error "KYBER_K must be in {2,3,4}"
#endif

#ifndef KYBER_PUBLICKEYBYTES
#define KYBER_SYMBYTES 32
#define KYBER_POLYVECBYTES (KYBER_K * 384)
#define KYBER_PUBLICKEYBYTES (KYBER_POLYVECBYTES + KYBER_SYMBYTES)
#define KYBER_SECRETKEYBYTES (2 * KYBER_POLYVECBYTES + 3 * KYBER_SYMBYTES)
#endif
#ifndef MPCITH_PROOF_SIZE
#define MPCITH_PROOF_SIZE ((size_t)(KYBER_K == 2 ? 664340 : KYBER_K == 3 ? 680980 : 744148)) /* sizeof(mpcith_proof) */
#endif

typedef struct {
    uint8_t pk[KYBER_PUBLICKEYBYTES];
    uint8_t sk[KYBER_SECRETKEYBYTES];
} kyber_keypair;

namespace kosk_dropin_detail {
struct State { kosk_b200_ctx *ctx; bool have_seed; uint8_t seed[32]; };
inline State &state()
{
    static State s = {nullptr, false, {0}};
    if (!s.ctx) {
        const char *dev = getenv("KOSK_B200_DEVICE");
        if (kosk_b200_create(&s.ctx, KYBER_K, dev ? atoi(dev) : 0, 64) != KOSK_OK) {
            fprintf(stderr, "kosk_b200: %s\n", kosk_b200_last_error());
            abort();                                  /* the reference aborts on RNG failure; there is no CPU fallback */
        }
    }
    return s;
}
}  // namespace kosk_dropin_detail

/* Fix the seed consumed by the next kyber_verifiable_keygen call (testing / reproducibility). */
inline void kosk_dropin_set_seed(const uint8_t seed[32])
{
    kosk_dropin_detail::State &s = kosk_dropin_detail::state();
    memcpy(s.seed, seed, 32); s.have_seed = true;
}

inline void kyber_verifiable_keygen(kyber_keypair *keypair, uint8_t *pi)
{
    kosk_dropin_detail::State &s = kosk_dropin_detail::state();
    uint8_t seed[32];
    if (s.have_seed) { memcpy(seed, s.seed, 32); s.have_seed = false; }
    else if (getrandom(seed, 32, 0) != 32) abort();
    if (kosk_b200_verifiable_keygen(s.ctx, seed, keypair->pk, keypair->sk, pi) != KOSK_OK) {
        fprintf(stderr, "kosk_b200: %s\n", kosk_b200_last_error());
        abort();
    }
}

inline bool kyber_kosk_verify(const uint8_t *pi, const uint8_t *pk)
{
    const int r = kosk_b200_kosk_verify(kosk_dropin_detail::state().ctx, pi, pk);
    if (r < 0) { fprintf(stderr, "kosk_b200: %s\n", kosk_b200_last_error()); abort(); }
    return r == 1;
}

#endif
