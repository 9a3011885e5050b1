// dropin_shim.cpp -- binary drop-in: libkosk_kyber512.so / libkosk_kyber768.so / libkosk_kyber1024.so (one per KYBER_K, like the
// reference's one-binary-per-K build, CMakeLists.txt:12-52).  Each exports, with the reference's own linkage,
//   * the C++ symbols of kosk.hpp:17-23, mlwe_prover.hpp:77-99 and mlwe_verifier.hpp:14-15:
//       _Z23kyber_verifiable_keygenP13kyber_keypairPh, _Z17kyber_kosk_verifyPKhS0_, _Z12kyber_keygenP13kyber_keypairP9mlwe_inst,
//       prepare_randomness, prepare_range_proof, prove, verify, encode_mpcith_proof, decode_mpcith_proof
//   * the C symbols kyber/kem.h:20-33 maps crypto_kem_* to: pqcrystals_kyber{512,768,1024}_ref_{keypair,keypair_derand,enc,enc_derand,dec}
// so that an object file compiled against the reference's OWN headers links against the shim instead of the reference's objects.
// All work happens in libkosk_b200.so (the shim depends on it); this file is only the symbol layer over include/kosk_dropin.hpp.
#define KOSK_DROPIN_SHIM 1
#include "../../include/kosk_dropin.hpp"

#if KYBER_K == 2
#define KOSK_KEM_NS(s) pqcrystals_kyber512_ref_##s
#elif KYBER_K == 3
#define KOSK_KEM_NS(s) pqcrystals_kyber768_ref_##s
#else
#define KOSK_KEM_NS(s) pqcrystals_kyber1024_ref_##s
#endif

extern "C" {
KOSK_DROPIN_API int KOSK_KEM_NS(keypair)(uint8_t *pk, uint8_t *sk) { return kosk_dropin_kem_keypair(pk, sk); }
KOSK_DROPIN_API int KOSK_KEM_NS(enc)(uint8_t *ct, uint8_t *ss, const uint8_t *pk) { return kosk_dropin_kem_enc(ct, ss, pk); }
KOSK_DROPIN_API int KOSK_KEM_NS(dec)(uint8_t *ss, const uint8_t *ct, const uint8_t *sk) { return kosk_dropin_kem_dec(ss, ct, sk); }
KOSK_DROPIN_API int KOSK_KEM_NS(keypair_derand)(uint8_t *pk, uint8_t *sk, const uint8_t *coins)
{
    kosk_dropin_detail::ok(kosk_b200_kem_keypair_derand_batch(kosk_dropin_detail::state().ctx, 1, coins, pk, sk));
    return 0;
}
KOSK_DROPIN_API int KOSK_KEM_NS(enc_derand)(uint8_t *ct, uint8_t *ss, const uint8_t *pk, const uint8_t *coins)
{
    kosk_dropin_detail::ok(kosk_b200_kem_enc_derand_batch(kosk_dropin_detail::state().ctx, 1, pk, coins, ct, ss));
    return 0;
}
}
