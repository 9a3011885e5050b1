// examples/main_dropin.cpp -- the KOSK part of the reference's demo (reference main.cpp:67-95) written against the
// reference API names, compiled against include/kosk_dropin.hpp and linked with libkosk_b200.so:
//   g++ -std=c++11 -O2 -DKYBER_K=2 -Iinclude examples/main_dropin.cpp -Lmpcith_kyber_kosk_b200 -lkosk_b200 \
//       -Wl,-rpath,$PWD/mpcith_kyber_kosk_b200 -o main_dropin
// With KOSK_SEED_HEX=<64 hex digits> the run is deterministic and prints FNV-1a-64 digests of pk / sk / proof.
#include <time.h>
#include <vector>
#include "kosk_dropin.hpp"

static uint64_t fnv1a(const uint8_t *p, size_t n)
{
    uint64_t h = 14695981039346656037ULL;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}
static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main()
{
    printf("=== kyber KOSK (drop-in, Kyber%d) ===\n", 256 * KYBER_K);
    const char *hex = getenv("KOSK_SEED_HEX");
    if (hex && strlen(hex) == 64) {
        uint8_t seed[32];
        for (int i = 0; i < 32; i++) { unsigned v; sscanf(hex + 2 * i, "%2x", &v); seed[i] = (uint8_t)v; }
        kosk_dropin_set_seed(seed);
    }
    kyber_keypair kp;
    std::vector<uint8_t> pi(MPCITH_PROOF_SIZE);
    double t0 = now();
    kyber_verifiable_keygen(&kp, pi.data());
    double t1 = now();
    printf(">>> kyber_verifiable_keygen time used: %f s (first call includes context creation)\n", t1 - t0);
    t0 = now();
    kyber_verifiable_keygen(&kp, pi.data());       // second call: steady-state single-proof latency, fresh OS seed
    t1 = now();
    printf(">>> kyber_verifiable_keygen (warm) time used: %f s\n", t1 - t0);
    if (hex && strlen(hex) == 64) {                // redo the deterministic one so that the digests below are reproducible
        uint8_t seed[32];
        for (int i = 0; i < 32; i++) { unsigned v; sscanf(hex + 2 * i, "%2x", &v); seed[i] = (uint8_t)v; }
        kosk_dropin_set_seed(seed);
        kyber_verifiable_keygen(&kp, pi.data());
    }
    t0 = now();
    bool res = kyber_kosk_verify(pi.data(), kp.pk);
    t1 = now();
    printf(">>> kyber_kosk_verify time used: %f s\n", t1 - t0);
    printf(res ? "[result] kosk verify success\n" : "[result] kosk verify failed\n");
    pi[0] ^= 1;
    printf(kyber_kosk_verify(pi.data(), kp.pk) ? "[tamper] accepted (BAD)\n" : "[tamper] rejected\n");
    pi[0] ^= 1;
    printf("[proof size] %zu kilobytes\n", (size_t)MPCITH_PROOF_SIZE / 1024);
    printf("[digest] pk=%016llx sk=%016llx proof=%016llx\n", (unsigned long long)fnv1a(kp.pk, sizeof kp.pk),
           (unsigned long long)fnv1a(kp.sk, sizeof kp.sk), (unsigned long long)fnv1a(pi.data(), pi.size()));
    return res ? 0 : 1;
}
