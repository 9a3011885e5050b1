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

// the same constants for fully unrolled permutations: folded into LOP3 immediates (a __constant__ array may be rewritten by the host, so
// every use of it stays an LDC)
struct KeccakRc { uint64_t v[24]; };
__device__ constexpr KeccakRc c_keccak_rc_imm = {{
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL}};

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
        // theta: A[x][y] ^= C[x-1] ^ rol(C[x+1], 1) as one three-input LOP3 per word half (no separate D: 10 LOP3 fewer per round)
        uint64_t c1[5];
#pragma unroll
        for (int x = 0; x < 5; x++) c1[x] = rol64(c[x], 1);
#pragma unroll
        for (int x = 0; x < 5; x++) {
#pragma unroll
            for (int y = 0; y < 25; y += 5) a[x + y] = a[x + y] ^ c[(x + 4) % 5] ^ c1[(x + 1) % 5];
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

// ---------------------------------------------------------------------------------------------
// Warp-cooperative Keccak-f[1600]: lane t < 25 of a warp holds state word A[x][y], t = x + 5y; theta / pi / chi
// exchange words with warp shuffles.  Used for the two strictly sequential Fiat-Shamir sponges (343 permutations
// each, mlwe_prover.cpp:131-135, :445-449), where one proof per warp cuts the latency ~5x against one per thread.
struct WarpKeccak {
    int t, x, y, rho, src_pi, src_c1, src_c2, l5, l10, l15, l20, lm1, lp1;
    __device__ __forceinline__ void init()
    {
        const int rho_tab[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        const int lane = threadIdx.x & 31;
        t = lane < 25 ? lane : 0; x = t % 5; y = t / 5;
        int r = 0;
#pragma unroll
        for (int i = 0; i < 25; i++) if (i == t) r = rho_tab[i];
        rho = r;
        src_pi = lane < 25 ? ((3 * y + x) % 5) + 5 * x : lane;      // B[x][y] = rol(A[(x+3y)%5][x], ..)
        l5 = lane < 25 ? (t + 5) % 25 : lane; l10 = lane < 25 ? (t + 10) % 25 : lane;
        l15 = lane < 25 ? (t + 15) % 25 : lane; l20 = lane < 25 ? (t + 20) % 25 : lane;
        lm1 = lane < 25 ? (x + 4) % 5 + 5 * y : lane; lp1 = lane < 25 ? (x + 1) % 5 + 5 * y : lane;
        // chi reads B[x+1][y] and B[x+2][y]; fetch them straight from their pre-pi source lanes (one shuffle stage less)
        const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
        src_c1 = lane < 25 ? ((3 * y + x1) % 5) + 5 * x1 : lane;
        src_c2 = lane < 25 ? ((3 * y + x2) % 5) + 5 * x2 : lane;
    }
    static __device__ __forceinline__ uint64_t shfl(uint64_t v, int src)
    {
        const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        return ((uint64_t)hi << 32) | lo;
    }
    __device__ __forceinline__ uint64_t permute(uint64_t a) const
    {
        const bool lane0 = (threadIdx.x & 31) == 0;
        // fully unrolled: the sponge is latency-bound on one warp, and across round boundaries the independent low / high word chains
        // overlap; measured 2.52 -> 2.09 us per permutation alone on an SM, 2.91 -> 2.33 us with 1024 sponges in flight
        // (tools/exp/sponge_round_bench.cu: unroll 2 / 4 / 8 give 2.30 / 2.15 / 2.09 us)
#pragma unroll
        for (int r = 0; r < 24; r++) {
            // theta: column parity (REDUX.XOR over per-column masks was measured 10x slower than four shuffles)
            const uint64_t c = a ^ shfl(a, l5) ^ shfl(a, l10) ^ shfl(a, l15) ^ shfl(a, l20);
            a ^= shfl(c, lm1) ^ rol64(shfl(c, lp1), 1);
            uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);           // rho: rotate left by a per-lane amount
            if (rho & 32) { const uint32_t tmp = lo; lo = hi; hi = tmp; }
            const uint32_t nh = __funnelshift_l(lo, hi, rho), nl = __funnelshift_l(hi, lo, rho);
            const uint64_t ar = ((uint64_t)nh << 32) | nl;
            const uint64_t b = shfl(ar, src_pi), b1 = shfl(ar, src_c1), b2 = shfl(ar, src_c2);   // pi, fused with chi's two neighbour reads
            a = b ^ (~b1 & b2);                                            // chi
            if (lane0) a ^= c_keccak_rc_imm.v[r];                          // iota
        }
        return a;
    }
    // SHA3-256 of the 1454 x 32-byte digest rows of one proof; result: lanes 0..3 hold the digest words
    __device__ __forceinline__ uint64_t tree_hash(const u8 *rows) const
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(rows);
        const int lane = threadIdx.x & 31;
        constexpr int NFULL = TREE_BYTES / 136;        // 342 full rate blocks + 16 bytes
        uint64_t a = 0, nxt = lane < 17 ? src[lane] : 0;
#pragma unroll 1
        for (int blk = 0; blk < NFULL; blk++) {
            a ^= nxt;
            nxt = 0;
            if (blk + 1 < NFULL) { if (lane < 17) nxt = src[(blk + 1) * 17 + lane]; }
            else if (lane < 2) nxt = src[NFULL * 17 + lane];
            a = permute(a);
        }
        a ^= nxt;
        if (lane == 2) a ^= 0x06ULL;
        if (lane == 16) a ^= 0x8000000000000000ULL;
        return permute(a);
    }
    // SHAKE256(digest || 0x01) (kyber_shake256_prf with nonce 1): state after the first permutation
    __device__ __forceinline__ uint64_t prf1(uint64_t digest_state) const
    {
        const int lane = threadIdx.x & 31;
        uint64_t a = lane < 4 ? digest_state : 0;
        if (lane == 4) a = 1ULL | (0x1FULL << 8);
        if (lane == 16) a = 0x8000000000000000ULL;
        return permute(a);
    }
};

}  // namespace kosk
