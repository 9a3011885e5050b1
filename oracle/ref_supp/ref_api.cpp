/* oracle/ref_supp/ref_api.cpp -- TEST INFRASTRUCTURE. extern "C" driver around the UNMODIFIED
 * reference entry points (kosk.hpp:17-24), compiled once per KYBER_K into oracle/_ref/.
 * Runs each call on a thread with a large stack (the reference keeps multi-MB structs on the
 * stack: mlwe_prover.cpp:103, kosk.cpp:77-84) and under the deterministic randombytes of
 * oracle/ok_rng.c. */
#include "kosk.hpp"
extern "C" {
#include "kyber/kem.h"
}
#include "../ok_rng.h"
#include <pthread.h>
#include <time.h>

namespace {
struct job { int op; const uint8_t *seed; int mode; uint8_t *pk, *sk, *pi; const uint8_t *cpi, *cpk; int ok;
             uint8_t *rand_img, *eta_img, *inst_img; const uint8_t *cinst; };
void *run(void *p)
{
    job *j = (job *)p;
    if (j->op == 0) {
        kosk_rng_reset(j->seed, j->mode);
        kyber_keypair *kp = new kyber_keypair;
        kyber_verifiable_keygen(kp, j->pi);
        memcpy(j->pk, kp->pk, KYBER_PUBLICKEYBYTES);
        memcpy(j->sk, kp->sk, KYBER_SECRETKEYBYTES);
        delete kp;
    } else if (j->op == 1) {
        j->ok = kyber_kosk_verify(j->cpi, j->cpk) ? 1 : 0;
    } else if (j->op == 2) {          /* the struct-level sequence of main.cpp:16-59 */
        kosk_rng_reset(j->seed, j->mode);
        mpcith_randomness *rnd = new mpcith_randomness; mpcith_range_proof *eta = new mpcith_range_proof;
        mlwe_inst *inst = new mlwe_inst; kyber_keypair *kp = new kyber_keypair; mpcith_proof *pi = new mpcith_proof;
        memset(rnd, 0, sizeof *rnd); memset(eta, 0, sizeof *eta);      /* share_vec.len is never written by the reference */
        prepare_randomness(rnd);
        prepare_range_proof(eta);
        kyber_keygen(kp, inst);
        prove(pi, inst, rnd, eta);
        j->ok = verify(pi, inst) ? 1 : 0;
        memcpy(j->rand_img, rnd, sizeof *rnd); memcpy(j->eta_img, eta, sizeof *eta); memcpy(j->inst_img, inst, sizeof *inst);
        memcpy(j->pk, kp->pk, KYBER_PUBLICKEYBYTES); memcpy(j->sk, kp->sk, KYBER_SECRETKEYBYTES);
        encode_mpcith_proof(j->pi, pi);
        delete rnd; delete eta; delete inst; delete kp; delete pi;
    } else if (j->op == 4) {          /* prove() alone on caller-supplied structs, DRBG positioned at call number j->mode */
        kosk_rng_reset(j->seed, KOSK_RNG_COUNTER);
        uint8_t junk[1];
        for (int i = 0; i < j->mode; i++) randombytes(junk, 1);
        mpcith_randomness *rnd = new mpcith_randomness; mpcith_range_proof *eta = new mpcith_range_proof;
        mlwe_inst *inst = new mlwe_inst; mpcith_proof *pi = new mpcith_proof;
        memcpy(rnd, j->rand_img, sizeof *rnd); memcpy(eta, j->eta_img, sizeof *eta); memcpy(inst, j->cinst, sizeof *inst);
        prove(pi, inst, rnd, eta);
        encode_mpcith_proof(j->pi, pi);
        delete rnd; delete eta; delete inst; delete pi;
    } else {                          /* verify() on a raw instance */
        mpcith_proof *pi = new mpcith_proof; mlwe_inst *inst = new mlwe_inst;
        decode_mpcith_proof(pi, j->cpi); memcpy(inst, j->cinst, sizeof *inst);
        j->ok = verify(pi, inst) ? 1 : 0;
        delete pi; delete inst;
    }
    return 0;
}
void big_stack(job *j)
{
    pthread_attr_t a; pthread_attr_init(&a); pthread_attr_setstacksize(&a, (size_t)256 << 20);
    pthread_t t; pthread_create(&t, &a, run, j); pthread_join(t, 0); pthread_attr_destroy(&a);
}
}

extern "C" {
int ref_kyber_k(void) { return KYBER_K; }
size_t ref_pk_bytes(void) { return KYBER_PUBLICKEYBYTES; }
size_t ref_sk_bytes(void) { return KYBER_SECRETKEYBYTES; }
size_t ref_proof_bytes(void) { return MPCITH_PROOF_SIZE; }
void ref_verifiable_keygen(const uint8_t seed[32], int rng_mode, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    job j = {0, seed, rng_mode, pk, sk, pi, 0, 0, 0, 0, 0, 0, 0}; big_stack(&j);
}
int ref_kosk_verify(const uint8_t *pi, const uint8_t *pk)
{
    job j = {1, 0, 0, 0, 0, 0, pi, pk, 0, 0, 0, 0, 0}; big_stack(&j); return j.ok;
}
uint32_t ref_rng_calls(void) { return 0; }
size_t ref_inst_bytes(void) { return sizeof(mlwe_inst); }
size_t ref_randomness_bytes(void) { return sizeof(mpcith_randomness); }
size_t ref_range_proof_bytes(void) { return sizeof(mpcith_range_proof); }
size_t ref_share_vec_bytes(void) { return sizeof(share_vec); }
/* main.cpp:16-59 in the reference's own call order: prepare_randomness, prepare_range_proof, kyber_keygen, prove, verify */
int ref_struct_sequence(const uint8_t seed[32], int rng_mode, uint8_t *rand_img, uint8_t *eta_img, uint8_t *inst_img,
                        uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    job j = {2, seed, rng_mode, pk, sk, pi, 0, 0, 0, rand_img, eta_img, inst_img, 0}; big_stack(&j); return j.ok;
}
/* Kyber KEM of the reference (kyber/kem.c:76-169); small stack frames, no helper thread needed */
size_t ref_ct_bytes(void) { return KYBER_CIPHERTEXTBYTES; }
void ref_kem_enc_derand(uint8_t *ct, uint8_t *ss, const uint8_t *pk, const uint8_t *coins) { crypto_kem_enc_derand(ct, ss, pk, coins); }
void ref_kem_dec(uint8_t *ss, const uint8_t *ct, const uint8_t *sk) { crypto_kem_dec(ss, ct, sk); }
void ref_kem_keypair_derand(uint8_t *pk, uint8_t *sk, const uint8_t *coins) { crypto_kem_keypair_derand(pk, sk, coins); }
/* crypto_kem_keypair under the counter DRBG positioned at call number `call` */
void ref_kem_keypair_at(const uint8_t seed[32], uint32_t call, uint8_t *pk, uint8_t *sk)
{
    kosk_rng_reset(seed, KOSK_RNG_COUNTER);
    uint8_t junk[1];
    for (uint32_t i = 0; i < call; i++) randombytes(junk, 1);
    crypto_kem_keypair(pk, sk);
}
/* crypto_kem_enc under the counter DRBG positioned at call number `call` */
void ref_kem_enc_at(const uint8_t seed[32], uint32_t call, uint8_t *ct, uint8_t *ss, const uint8_t *pk)
{
    kosk_rng_reset(seed, KOSK_RNG_COUNTER);
    uint8_t junk[1];
    for (uint32_t i = 0; i < call; i++) randombytes(junk, 1);
    crypto_kem_enc(ct, ss, pk);
}
/* prove() (mlwe_prover.cpp:81-538) on caller-supplied struct images with the counter DRBG positioned at call number `call` */
void ref_prove_struct_at(const uint8_t seed[32], int call, const uint8_t *inst_img, uint8_t *rand_img, uint8_t *eta_img, uint8_t *pi)
{
    job j = {4, seed, call, 0, 0, pi, 0, 0, 0, rand_img, eta_img, 0, inst_img}; big_stack(&j);
}
int ref_verify_struct(const uint8_t *pi, const uint8_t *inst_img)
{
    job j = {3, 0, 0, 0, 0, 0, pi, 0, 0, 0, 0, 0, inst_img}; big_stack(&j); return j.ok;
}
}
