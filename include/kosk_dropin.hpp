// kosk_dropin.hpp -- source-level drop-in for the reference's kosk.hpp (reference kosk.hpp:13-24).
//
// A program written against the reference API
//     kyber_keypair kp; uint8_t pi[MPCITH_PROOF_SIZE];
//     kyber_verifiable_keygen(&kp, pi);                 // kosk.cpp:72-86
//     bool ok = kyber_kosk_verify(pi, kp.pk);           // kosk.cpp:88-117
// compiles unchanged against this header (same type name, function names, signatures, KYBER_K selection by
// macro as in params.hpp:8-10 / kyber/params.h:4-6) and links against libkosk_b200.so instead of the
// reference's objects.  The lower-level API main.cpp:16-59 uses (prepare_randomness, prepare_range_proof, kyber_keygen,
// prove, verify on mlwe_inst / mpcith_randomness / mpcith_range_proof / mpcith_proof; mlwe_prover.hpp:34-99,
// mlwe_verifier.hpp:14-15) and the KEM calls of main.cpp:98-113 (crypto_kem_enc / crypto_kem_dec) are here too, so the
// reference's whole main.cpp compiles unchanged against include/dropin/ (headers of the reference's names that forward
// here).  Randomness: the reference pulls from the OS through one global randombytes() (kyber/randombytes.c:43-57); here
// 32 bytes from getrandom() seed the context's KOSK counter-mode DRBG (include/kosk_b200.h) and every function consumes
// exactly the calls its reference counterpart makes.  kosk_dropin_set_seed() injects a fixed seed for reproducible runs.
#ifndef KOSK_DROPIN_HPP
#define KOSK_DROPIN_HPP

#include <stdint.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>
#include "kosk_b200.h"

/* KOSK_DROPIN_SHIM (defined only by csrc/dropin_shim.cpp): the functions below are emitted as the exported, non-inline symbols
 * of libkosk_kyber{512,768,1024}.so, with the reference's own (C++-mangled) names, so that OBJECT files compiled against the
 * reference's headers link against the shim unchanged (binary drop-in).  Everywhere else they are inline (source drop-in). */
#ifdef KOSK_DROPIN_SHIM
#define KOSK_DROPIN_API __attribute__((visibility("default")))
#else
#define KOSK_DROPIN_API inline
#endif

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
#ifndef KYBER_N
#define KYBER_N 256
#define KYBER_Q 3329
#endif
#ifndef KYBER_SSBYTES
#define KYBER_SSBYTES 32
#endif
#ifndef KYBER_ETA1
#define KYBER_ETA1 (KYBER_K == 2 ? 3 : 2)
#endif
#ifndef KYBER_CIPHERTEXTBYTES
#define KYBER_CIPHERTEXTBYTES (KYBER_K == 4 ? 1568 : KYBER_K * 320 + 128)
#endif
#ifndef MPCITH_N          /* params.hpp:12-37: identical for every KYBER_K except MPCITH_V */
#define MPCITH_N 1454
#define MPCITH_T 150
#define MPCITH_L KYBER_N
#define MPCITH_K 70
#define MPCITH_V (KYBER_K * 2)
#endif

typedef struct {
    uint8_t pk[KYBER_PUBLICKEYBYTES];
    uint8_t sk[KYBER_SECRETKEYBYTES];
} kyber_keypair;

/* ---- the reference's struct types (same names, member names and memory layout) ---- */
#ifndef KOSK_DROPIN_NO_KYBER_TYPES     /* define when the including program already has kyber/poly.h and kyber/polyvec.h */
typedef struct { int16_t coeffs[KYBER_N]; } poly;                 /* kyber/poly.h */
typedef struct { poly vec[KYBER_K]; } polyvec;                    /* kyber/polyvec.h */
#endif
typedef struct { size_t len; uint16_t share_x[MPCITH_N]; uint16_t share_y[MPCITH_N]; } share_vec;     /* ss.hpp:33-37 */
typedef struct { size_t len; uint16_t secret[KYBER_N]; } secret_vec;                                     /* ss.hpp:39-42 */
typedef struct { polyvec A[KYBER_K], t; polyvec s, e; } mlwe_inst;                                       /* mlwe_prover.hpp:34-37 */
#define KOSK_DROPIN_F (MPCITH_K + MPCITH_V + 1)
#define KOSK_DROPIN_E (KYBER_ETA1 * 2 + 1)
#define KOSK_DROPIN_M (KYBER_ETA1 * 2)
#define KOSK_DROPIN_R (MPCITH_N - MPCITH_T)
typedef struct {                                                                                          /* mlwe_prover.hpp:39-44 */
    uint16_t f[KOSK_DROPIN_F][KYBER_N], NTT_f[KOSK_DROPIN_F][KYBER_N];
    share_vec f_shares[KOSK_DROPIN_F], NTT_f_shares[KOSK_DROPIN_F];
} mpcith_randomness;
typedef struct { share_vec s_eta_shares[KYBER_K][KOSK_DROPIN_E], e_eta_shares[KYBER_K][KOSK_DROPIN_E]; } mpcith_range_proof;   /* :46-49 */
typedef struct {                                                                                          /* mlwe_prover.hpp:57-75 */
    uint16_t f_shares[MPCITH_T][KOSK_DROPIN_F], NTT_f_shares[MPCITH_T][KOSK_DROPIN_F];
    uint16_t beta_shares[KOSK_DROPIN_R][MPCITH_K], gamma_shares[KOSK_DROPIN_R][MPCITH_K];
    uint8_t Tcomm[KOSK_DROPIN_R][KYBER_SYMBYTES];
    uint16_t I[MPCITH_T];
    uint16_t s_shares[MPCITH_T][KYBER_K], e_shares[MPCITH_T][KYBER_K], t_shares[KOSK_DROPIN_R][KYBER_K];
    uint16_t NTT_s_shares[MPCITH_T][KYBER_K], NTT_e_shares[MPCITH_T][KYBER_K];
    uint16_t NTT_Ar_shares[MPCITH_T][KYBER_K], NTT_As_shares[MPCITH_T][KYBER_K];
    uint16_t sr_shares[KOSK_DROPIN_R][KYBER_K], er_shares[KOSK_DROPIN_R][KYBER_K];
    uint16_t s_eta_shares[KOSK_DROPIN_R][KYBER_K][KOSK_DROPIN_E], e_eta_shares[KOSK_DROPIN_R][KYBER_K][KOSK_DROPIN_E];
    uint16_t s_sub_eta_shares[MPCITH_T][KYBER_K][KOSK_DROPIN_E], e_sub_eta_shares[MPCITH_T][KYBER_K][KOSK_DROPIN_E];
    uint16_t z_s_ddeg_shares[MPCITH_T][KYBER_K][KOSK_DROPIN_M], z_e_ddeg_shares[MPCITH_T][KYBER_K][KOSK_DROPIN_M];
    uint16_t u_s_2ddeg_shares[KOSK_DROPIN_R][KYBER_K][KOSK_DROPIN_M], u_e_2ddeg_shares[KOSK_DROPIN_R][KYBER_K][KOSK_DROPIN_M];
    uint8_t comm[KOSK_DROPIN_R][KYBER_SYMBYTES];
} mpcith_proof;
#ifndef MPCITH_PROOF_SIZE
#define MPCITH_PROOF_SIZE sizeof(mpcith_proof)
#define MPCITH_PRE_RANDOMNESS_SIZE (sizeof(mpcith_randomness) + sizeof(mpcith_range_proof))
#endif
static_assert(sizeof(mpcith_proof) == (KYBER_K == 2 ? 664340 : KYBER_K == 3 ? 680980 : 744148), "mpcith_proof layout (SURVEY Appendix B)");
static_assert(sizeof(share_vec) == 5824 && sizeof(mlwe_inst) == (KYBER_K * KYBER_K + 3 * KYBER_K) * 512, "struct layout");

namespace kosk_dropin_detail {
struct State { kosk_b200_ctx *ctx; };
inline void die() { fprintf(stderr, "kosk_b200: %s\n", kosk_b200_last_error()); abort(); }   /* the reference aborts on RNG failure; there is no CPU fallback */
inline void reseed(kosk_b200_ctx *ctx)
{
    uint8_t seed[32];
    if (getrandom(seed, 32, 0) != 32) abort();
    if (kosk_b200_rng_reset(ctx, seed) != KOSK_OK) die();
}
inline State make_state()
{
    State s = {nullptr};
    const char *dev = getenv("KOSK_B200_DEVICE");
    if (kosk_b200_create(&s.ctx, KYBER_K, dev ? atoi(dev) : 0, 64) != KOSK_OK) die();
    reseed(s.ctx);
    return s;
}
/* One process-wide context and DRBG stand in for the reference's stack state and OS RNG.  Thread safety: the static is initialised
 * once (C++11), every C-ABI call takes the context's lock, and each function below is ONE such call, so concurrent callers are
 * serialised, never interleaved (the reference's functions are re-entrant; these are merely thread-safe).  The DRBG addresses its
 * calls with a 32-bit counter (include/kosk_b200.h): long before it could wrap, a fresh seed is drawn from the OS and the
 * counter restarts, so no randomness is ever replayed. */
inline State &state()
{
    static State s = make_state();
    if (kosk_b200_rng_calls(s.ctx) > (1u << 30)) reseed(s.ctx);
    return s;
}
inline void ok(int rc) { if (rc != KOSK_OK) die(); }
}  // namespace kosk_dropin_detail

/* Re-seed the DRBG that stands in for the reference's global randombytes() and restart its call counter (testing /
 * reproducibility): everything after this call is a deterministic function of `seed`. */
KOSK_DROPIN_API void kosk_dropin_set_seed(const uint8_t seed[32])
{
    kosk_dropin_detail::ok(kosk_b200_rng_reset(kosk_dropin_detail::state().ctx, seed));
}

KOSK_DROPIN_API void kyber_verifiable_keygen(kyber_keypair *keypair, uint8_t *pi)
{
    kosk_dropin_detail::ok(kosk_b200_verifiable_keygen_rng(kosk_dropin_detail::state().ctx, keypair->pk, keypair->sk, pi));
}

KOSK_DROPIN_API bool kyber_kosk_verify(const uint8_t *pi, const uint8_t *pk)
{
    const int r = kosk_b200_kosk_verify(kosk_dropin_detail::state().ctx, pi, pk);
    if (r < 0) { fprintf(stderr, "kosk_b200: %s\n", kosk_b200_last_error()); abort(); }
    return r == 1;
}

/* ---- struct-level API (mlwe_prover.hpp:77-99, mlwe_verifier.hpp:14-15, kosk.hpp:17-18) ---- */
KOSK_DROPIN_API void prepare_randomness(mpcith_randomness *rand) { kosk_dropin_detail::ok(kosk_b200_prepare_randomness(kosk_dropin_detail::state().ctx, rand)); }
KOSK_DROPIN_API void prepare_range_proof(mpcith_range_proof *eta_shares) { kosk_dropin_detail::ok(kosk_b200_prepare_range_proof(kosk_dropin_detail::state().ctx, eta_shares)); }
KOSK_DROPIN_API void kyber_keygen(kyber_keypair *keypair, mlwe_inst *raw_key) { kosk_dropin_detail::ok(kosk_b200_keygen(kosk_dropin_detail::state().ctx, keypair->pk, keypair->sk, raw_key)); }
KOSK_DROPIN_API void prove(mpcith_proof *pi, const mlwe_inst *mlwe, const mpcith_randomness *rand, const mpcith_range_proof *eta_share)
{
    kosk_dropin_detail::ok(kosk_b200_prove(kosk_dropin_detail::state().ctx, reinterpret_cast<uint8_t *>(pi), mlwe, rand, eta_share));
}
KOSK_DROPIN_API bool verify(const mpcith_proof *pi, const mlwe_inst *mlwe)
{
    const int r = kosk_b200_verify(kosk_dropin_detail::state().ctx, reinterpret_cast<const uint8_t *>(pi), mlwe);
    if (r < 0) kosk_dropin_detail::die();
    return r == 1;
}
KOSK_DROPIN_API void encode_mpcith_proof(uint8_t *buf, const mpcith_proof *pi) { memcpy(buf, pi, sizeof(mpcith_proof)); }     /* mlwe_prover.cpp:540-543 */
KOSK_DROPIN_API void decode_mpcith_proof(mpcith_proof *pi, const uint8_t *buf) { memcpy(pi, buf, sizeof(mpcith_proof)); }     /* :545-630, field by field = the same bytes */

/* ---- Kyber KEM on the generated keys (kyber/kem.h:29-33; main.cpp:98-113) ---- */
KOSK_DROPIN_API int kosk_dropin_kem_keypair(uint8_t *pk, uint8_t *sk) { kosk_dropin_detail::ok(kosk_b200_kem_keypair(kosk_dropin_detail::state().ctx, pk, sk)); return 0; }
KOSK_DROPIN_API int kosk_dropin_kem_enc(uint8_t *ct, uint8_t *ss, const uint8_t *pk) { kosk_dropin_detail::ok(kosk_b200_kem_enc(kosk_dropin_detail::state().ctx, ct, ss, pk)); return 0; }
KOSK_DROPIN_API int kosk_dropin_kem_dec(uint8_t *ss, const uint8_t *ct, const uint8_t *sk) { kosk_dropin_detail::ok(kosk_b200_kem_dec(kosk_dropin_detail::state().ctx, ss, ct, sk)); return 0; }
#ifndef crypto_kem_enc      /* the reference namespaces these through macros as well (kyber/kem.h:26-33) */
#define crypto_kem_keypair kosk_dropin_kem_keypair
#define crypto_kem_enc kosk_dropin_kem_enc
#define crypto_kem_dec kosk_dropin_kem_dec
#endif

#endif
