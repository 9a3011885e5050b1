/* oracle/kosk_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * Plain-C restatement of the reference's KOSK path with KYBER_K as a run-time argument:
 *   kyber_verifiable_keygen  (reference kosk.cpp:72-86)   -> kosk_oracle_verifiable_keygen
 *   kyber_kosk_verify        (reference kosk.cpp:88-117)  -> kosk_oracle_verify
 * plus the component functions the kernel-level parity tests compare against.
 * Parity pinning: validated bit-for-bit against oracle/_ref (the unmodified reference sources
 * compiled with the supplements in oracle/ref_supp) and against tests/golden fixtures
 * generated from oracle/_ref; see tests/test_oracle_cpu.py.
 */
#ifndef KOSK_ORACLE_H
#define KOSK_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KO_N 1454      /* MPCITH_N  (params.hpp:12) */
#define KO_T 150       /* MPCITH_T  (params.hpp:13) */
#define KO_R (KO_N - KO_T)
#define KO_L 256
#define KO_MK 70       /* MPCITH_K  (params.hpp:15) */
#define KO_D1 407      /* DEG_D+1   (ss.hpp:56)     */
#define KO_D2 813      /* DEG_2D+1  (ss.hpp:57)     */
#define KO_Q 3329

typedef struct {
    int k, eta, F, E, M;
    size_t pk_bytes, sk_bytes, proof_bytes;
    /* byte offsets of the 24 fields of struct mpcith_proof (mlwe_prover.hpp:57-75), SURVEY App. B */
    size_t o_f, o_Tf, o_beta, o_gamma, o_Tcomm, o_I, o_s, o_e, o_t, o_NTTs, o_NTTe, o_NTTAr, o_NTTAs,
           o_sr, o_er, o_seta, o_eeta, o_ssub, o_esub, o_zs, o_ze, o_us, o_ue, o_comm;
} ko_layout;

int  kosk_oracle_layout(int k, ko_layout *L);

void kosk_oracle_verifiable_keygen(int k, const uint8_t seed[32], int rng_mode,
                                   uint8_t *pk, uint8_t *sk, uint8_t *pi);
int  kosk_oracle_verify(int k, const uint8_t *pi, const uint8_t *pk);

/* intermediates of the last kosk_oracle_verifiable_keygen on this thread (kernel-level parity) */
typedef struct {
    uint16_t alpha[78];
    uint8_t  fs1_digest[32], fs2_digest[32];
    uint16_t I[KO_T];
    uint16_t first_share[KO_N];    /* f_shares[0].share_y */
    uint16_t first_secret[KO_L];   /* f[0] */
    uint16_t first_ntt[KO_L];      /* NTT_f[0] */
    uint8_t  tcomm0[32], view0[32];
} ko_trace;
const ko_trace *kosk_oracle_last_trace(void);

/* components */
void ko_share_ddeg(uint16_t shares[KO_N], const uint16_t y[KO_D1]);          /* ss.cpp:76-99  */
void ko_recon_ddeg(uint16_t secret[KO_L], const uint16_t shares[KO_D1]);     /* ss.cpp:37-54  */
void ko_recon_2ddeg(uint16_t secret[KO_L], const uint16_t shares[KO_D2]);    /* ss.cpp:56-73  */
void ko_ntt(uint16_t r[256]);        /* canonical in/out; kyber/ntt.c:80-95 + poly.c:261-265 */
void ko_basemul_acc(int k, uint16_t r[256], const uint16_t *a, const uint16_t *b); /* polyvec.c:202-214 + poly.c:307-313 */
void ko_gen_matrix(int k, uint16_t *A, const uint8_t seed[32]);              /* indcpa.c:168-193, A[i][j][256] */
void ko_sha3_256(uint8_t h[32], const uint8_t *in, size_t n);
void ko_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t n);
void ko_shake128(uint8_t *out, size_t outlen, const uint8_t *in, size_t n);
void ko_sha3_512(uint8_t h[64], const uint8_t *in, size_t n);
void ko_keccak_f1600(uint64_t a[25]);
void ko_randombytes_at(const uint8_t seed[32], uint32_t call, uint8_t *out, size_t n);
#ifdef __cplusplus
}
#endif
#endif
