/* oracle/ok_rng.h -- TEST INFRASTRUCTURE. Deterministic replacement for the reference's
 * randombytes() (reference kyber/randombytes.c:43-57 reads the OS RNG).
 *
 * Definition shared by the CPU oracle(s) and the CUDA product ("KOSK counter-mode DRBG"):
 *   call number c (0-based, reset for every proof) returns
 *       SHAKE256( seed[32] || LE32(c) ) [0 : outlen]
 * so every call is an independent sponge and the GPU can expand all of them in parallel.
 * KOSK_RNG_STREAM reproduces the survey probe's RNG (one continuous SHAKE256 stream over
 * LE64(seed)) and exists only to re-check the survey's FNV anchors.
 */
#ifndef OK_RNG_H
#define OK_RNG_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
enum { KOSK_RNG_COUNTER = 0, KOSK_RNG_STREAM = 1 };
void kosk_rng_reset(const uint8_t seed[32], int mode);
uint32_t kosk_rng_calls(void);
uint64_t kosk_rng_bytes(void);
void randombytes(uint8_t *out, size_t outlen);
#ifdef __cplusplus
}
#endif
#endif
