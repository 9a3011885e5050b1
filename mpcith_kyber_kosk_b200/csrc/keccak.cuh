// keccak.cuh -- register-resident Keccak-f[1600] and the sponge modes the KOSK path uses.
// Replaces reference kyber/fips202.c:82-344 (KeccakF1600_StatePermute), :461-485 (absorb_once),
// :723-774 (shake256 / sha3_256 / sha3_512) and kyber/symmetric-shake.c:18-51.
// All 25 lanes live in registers: every lane index below is a compile-time constant after unrolling.
#pragma once
#include "kosk_common.cuh"

namespace kosk {

__constant__ uint64_t c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

__device__ __forceinline__ uint64_t rol64(uint64_t x, int n)
{
    // two funnel shifts on the 32-bit halves (SHF.L.W); n is a compile-time constant at every call site
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32), rl, rh;
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n < 32) { rl = __funnelshift_l(hi, lo, n); rh = __funnelshift_l(lo, hi, n); }
    else        { rl = __funnelshift_l(lo, hi, n - 32); rh = __funnelshift_l(hi, lo, n - 32); }
    return ((uint64_t)rh << 32) | rl;
}

template <int UNROLL = 1>
__device__ __forceinline__ void keccak_f1600(uint64_t (&a)[25])
{
#pragma unroll UNROLL
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            uint64_t d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
#pragma unroll
            for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
        }
        // rho + pi: B[y][2x+3y] = rol(A[x][y], rho[x][y])   (FIPS-202 3.2.2-3.2.3)
#define KOSK_RP(x, y, n) b[(y) + 5 * ((2 * (x) + 3 * (y)) % 5)] = rol64(a[(x) + 5 * (y)], n);
        KOSK_RP(0, 0, 0)  KOSK_RP(1, 0, 1)  KOSK_RP(2, 0, 62) KOSK_RP(3, 0, 28) KOSK_RP(4, 0, 27)
        KOSK_RP(0, 1, 36) KOSK_RP(1, 1, 44) KOSK_RP(2, 1, 6)  KOSK_RP(3, 1, 55) KOSK_RP(4, 1, 20)
        KOSK_RP(0, 2, 3)  KOSK_RP(1, 2, 10) KOSK_RP(2, 2, 43) KOSK_RP(3, 2, 25) KOSK_RP(4, 2, 39)
        KOSK_RP(0, 3, 41) KOSK_RP(1, 3, 45) KOSK_RP(2, 3, 15) KOSK_RP(3, 3, 21) KOSK_RP(4, 3, 8)
        KOSK_RP(0, 4, 18) KOSK_RP(1, 4, 2)  KOSK_RP(2, 4, 61) KOSK_RP(3, 4, 56) KOSK_RP(4, 4, 14)
#undef KOSK_RP
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; x++) a[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & b[(x + 2) % 5 + y]);
        a[0] ^= c_keccak_rc[r];
    }
}

__device__ __forceinline__ void keccak_zero(uint64_t (&a)[25])
{
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
}

// BE16 pair number i (0..3) of a squeezed lane -> ((b[2i] << 8) | b[2i+1]), reference ss.cpp:8, mlwe_prover.cpp:12,141,456
__device__ __forceinline__ uint32_t lane_be16(uint64_t lane, int i)
{
    uint32_t w = (uint32_t)(lane >> (16 * i)) & 0xFFFFu;
    return ((w & 0xFF) << 8) | (w >> 8);
}

// Counter-mode DRBG standing in for randombytes() (definition in include/kosk_b200.h):
// call c of a proof returns SHAKE256(seed[32] || LE32(c)).  Leaves the state after the first permutation.
__device__ __forceinline__ void drbg_begin(uint64_t (&a)[25], const uint64_t seed[4], uint32_t call)
{
    keccak_zero(a);
    a[0] = seed[0]; a[1] = seed[1]; a[2] = seed[2]; a[3] = seed[3];
    a[4] = (uint64_t)call | (0x1FULL << 32);
    a[16] = 0x8000000000000000ULL;         // rate 136 = lanes 0..16
    keccak_f1600(a);
}

// SHAKE256(key[32] || nonce) (kyber_shake256_prf, symmetric-shake.c:43-51); state after first permutation
__device__ __forceinline__ void prf_begin(uint64_t (&a)[25], const uint64_t key[4], uint8_t nonce)
{
    keccak_zero(a);
    a[0] = key[0]; a[1] = key[1]; a[2] = key[2]; a[3] = key[3];
    a[4] = (uint64_t)nonce | (0x1FULL << 8);
    a[16] = 0x8000000000000000ULL;
    keccak_f1600(a);
}

// ---- byte-addressed sponge with the state in local memory: only for the tiny, branchy keygen path ----
struct ByteSponge {
    uint64_t s[25];
    int rate, pos;
    __device__ void init(int r) { for (int i = 0; i < 25; i++) s[i] = 0; rate = r; pos = 0; }
    __device__ void permute() { uint64_t a[25]; for (int i = 0; i < 25; i++) a[i] = s[i]; keccak_f1600(a); for (int i = 0; i < 25; i++) s[i] = a[i]; }
    __device__ void absorb(const uint8_t *in, int n)
    {
        for (int i = 0; i < n; i++) { s[pos >> 3] ^= (uint64_t)in[i] << (8 * (pos & 7)); if (++pos == rate) { permute(); pos = 0; } }
    }
    __device__ void finalize(uint8_t dom)
    {
        s[pos >> 3] ^= (uint64_t)dom << (8 * (pos & 7)); s[(rate - 1) >> 3] ^= 0x80ULL << (8 * ((rate - 1) & 7)); pos = rate;
    }
    __device__ uint8_t next()
    {
        if (pos == rate) { permute(); pos = 0; }
        uint8_t b = (uint8_t)(s[pos >> 3] >> (8 * (pos & 7))); pos++; return b;
    }
};

}  // namespace kosk
