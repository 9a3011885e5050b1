/* examples/c_abi_demo.c -- the C ABI (include/kosk_b200.h) from plain C: batch prove, batch verify, tamper check.
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -Lmpcith_kyber_kosk_b200 -lkosk_b200 -Wl,-rpath,$PWD/mpcith_kyber_kosk_b200 -o c_abi_demo
 * usage: c_abi_demo <kyber_k> <n>      prints one line "k=.. n=.. accepted=.. tampered_rejected=.. fnv=<digest of all proofs>" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kosk_b200.h"

int main(int argc, char **argv)
{
    const int k = argc > 1 ? atoi(argv[1]) : 2;
    const size_t n = argc > 2 ? (size_t)atoi(argv[2]) : 4;
    kosk_b200_ctx *ctx = NULL;
    if (kosk_b200_create(&ctx, k, 0, 64) != KOSK_OK) { fprintf(stderr, "create: %s\n", kosk_b200_last_error()); return 2; }
    const size_t npk = kosk_b200_pk_bytes(k), nsk = kosk_b200_sk_bytes(k), npi = kosk_b200_proof_bytes(k);
    uint8_t *seeds = calloc(n, 32), *pk = malloc(n * npk), *sk = malloc(n * nsk), *pi = malloc(n * npi), *ok = malloc(n);
    for (size_t i = 0; i < n; i++) { seeds[32 * i] = (uint8_t)(i + 1); seeds[32 * i + 1] = 0xC0; }   /* fixed seeds: reproducible */
    if (kosk_b200_prove_batch(ctx, n, seeds, pk, sk, pi) != KOSK_OK) { fprintf(stderr, "prove: %s\n", kosk_b200_last_error()); return 3; }
    if (kosk_b200_verify_batch(ctx, n, pi, pk, ok) != KOSK_OK) { fprintf(stderr, "verify: %s\n", kosk_b200_last_error()); return 4; }
    size_t acc = 0; for (size_t i = 0; i < n; i++) acc += ok[i];
    unsigned long long h = 14695981039346656037ULL;
    for (size_t i = 0; i < n * npi; i++) { h ^= pi[i]; h *= 1099511628211ULL; }
    pi[17] ^= 4;                                            /* corrupt the first proof */
    const int rej = kosk_b200_kosk_verify(ctx, pi, pk) == 0;
    printf("k=%d n=%zu accepted=%zu tampered_rejected=%d fnv=%016llx launches=%llu\n", k, n, acc, rej, h,
           (unsigned long long)kosk_b200_kernel_launches(ctx));
    kosk_b200_destroy(ctx);
    free(seeds); free(pk); free(sk); free(pi); free(ok);
    return (acc == n && rej) ? 0 : 1;
}
