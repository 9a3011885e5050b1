/* oracle/ref_supp/ref_api.cpp -- TEST INFRASTRUCTURE. extern "C" driver around the UNMODIFIED
 * reference entry points (kosk.hpp:17-24), compiled once per KYBER_K into oracle/_ref/.
 * Runs each call on a thread with a large stack (the reference keeps multi-MB structs on the
 * stack: mlwe_prover.cpp:103, kosk.cpp:77-84) and under the deterministic randombytes of
 * oracle/ok_rng.c. */
#include "kosk.hpp"
#include "../ok_rng.h"
#include <pthread.h>
#include <time.h>

namespace {
struct job { int op; const uint8_t *seed; int mode; uint8_t *pk, *sk, *pi; const uint8_t *cpi, *cpk; int ok; };
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
    } else {
        j->ok = kyber_kosk_verify(j->cpi, j->cpk) ? 1 : 0;
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
    job j = {0, seed, rng_mode, pk, sk, pi, 0, 0, 0}; big_stack(&j);
}
int ref_kosk_verify(const uint8_t *pi, const uint8_t *pk)
{
    job j = {1, 0, 0, 0, 0, 0, pi, pk, 0}; big_stack(&j); return j.ok;
}
uint32_t ref_rng_calls(void) { return 0; }
}
