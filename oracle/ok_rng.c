/* oracle/ok_rng.c -- TEST INFRASTRUCTURE. See ok_rng.h. */
#include "ok_rng.h"
#include "ok_keccak.h"

static __thread uint8_t g_seed[32];
static __thread int g_mode;
static __thread uint32_t g_calls;
static __thread uint64_t g_bytes;
static __thread ok_sponge g_stream;

void kosk_rng_reset(const uint8_t seed[32], int mode)
{
    memcpy(g_seed, seed, 32); g_mode = mode; g_calls = 0; g_bytes = 0;
    if (mode == KOSK_RNG_STREAM) {           /* survey probe: absorb LE64(seed), squeeze forever */
        ok_sponge_init(&g_stream, 136);
        ok_sponge_absorb(&g_stream, seed, 8);
        ok_sponge_finalize(&g_stream, 0x1F);
    }
}
uint32_t kosk_rng_calls(void) { return g_calls; }
uint64_t kosk_rng_bytes(void) { return g_bytes; }

void randombytes(uint8_t *out, size_t outlen)
{
    if (g_mode == KOSK_RNG_STREAM) {
        ok_sponge_squeeze(&g_stream, out, outlen);
    } else {
        uint8_t in[36];
        memcpy(in, g_seed, 32);
        in[32] = (uint8_t)g_calls; in[33] = (uint8_t)(g_calls >> 8);
        in[34] = (uint8_t)(g_calls >> 16); in[35] = (uint8_t)(g_calls >> 24);
        ok_shake256(out, outlen, in, 36);
    }
    g_calls++; g_bytes += outlen;
}
