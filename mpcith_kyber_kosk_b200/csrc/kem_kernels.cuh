// kem_kernels.cuh -- Kyber KEM encapsulation / decapsulation for keys made by the KOSK keygen: the step
// main.cpp:98-113 runs right after kyber_verifiable_keygen.  Replaces reference kyber/kem.c:76-169 (crypto_kem_enc_derand,
// crypto_kem_enc, crypto_kem_dec), kyber/indcpa.c:260-336 (indcpa_enc / indcpa_dec, gen_at = gen_matrix transposed :168-193),
// kyber/poly.c:21-221 (compress / decompress / frommsg / tomsg, getnoise_eta1/eta2), kyber/polyvec.c:16-138, kyber/ntt.c:106-126
// (invntt) and kyber_shake256_rkprf (symmetric-shake.c:64-73).  One CTA of 128 threads per encapsulation / decapsulation,
// everything in shared memory, residues canonical in [0, q): every value that reaches a ciphertext or message byte is a
// function of the coefficient mod q only, so any exact representation gives the reference's bytes.
// Included by kosk_b200.cu after the context definition.
#pragma once

namespace kosk {

__constant__ u16 c_zeta_inv[128];       // inverses of c_zeta (plain residues)

struct KemDims { int k, eta1, du, dv, pvc, pc, ct_bytes; };
KOSK_HD KemDims kem_dims(int k)
{
    KemDims d; d.k = k; d.eta1 = (k == 2) ? 3 : 2; d.du = (k == 4) ? 11 : 10; d.dv = (k == 4) ? 5 : 4;
    d.pvc = k * 32 * d.du; d.pc = 32 * d.dv; d.ct_bytes = d.pvc + d.pc;       // kyber/params.h:29-41,53
    return d;
}

// inverse of ntt256_block: Gentleman-Sande butterflies with the inverse twiddles, then 128^-1 = 3303 (kyber/ntt.c:106-126)
__device__ __forceinline__ void invntt256_block(u16 *p, int tid)
{
    for (int len = 2; len <= 128; len <<= 1) {
        const int grp = tid / len, j = grp * 2 * len + (tid % len);
        const uint32_t zi = c_zeta_inv[128 / len + grp];
        const uint32_t a = p[j], b = p[j + len];
        p[j] = (u16)gf_add(a, b);
        p[j + len] = (u16)gf_mul(gf_sub(a, b), zi);
        __syncthreads();
    }
    p[tid] = (u16)gf_mul(p[tid], 3303); p[tid + 128] = (u16)gf_mul(p[tid + 128], 3303);
    __syncthreads();
}

template <int K>
struct KemSmem {
    u16 At[K * K][256];                       // A^T in the NTT domain (gen_at)
    u16 sp[K][256], ep[K][256], b[K][256], t[K][256];
    u16 epp[256], v[256];
    uint64_t blk[K * K > 2 * K + 1 ? K * K : 2 * K + 1][24];   // per-thread sponge output blocks of the samplers
    uint32_t mw[8];                           // message bits (dec)
    u8 m[32], coins[32], kr[64], hpk[32];
    u8 ct[K * 352 + 160];
    int fail;
};

// compress_d(x) = round(2^d x / q) mod 2^d with the reference's multiply-shift constants (poly.c:29-33,:48-52, polyvec.c:29-35,:57-63)
__device__ __forceinline__ uint32_t compress_d(uint32_t x, int d)
{
    if (d == 4) return ((((x << 4) + 1665) * 80635u) >> 28) & 0xF;
    if (d == 5) return ((((x << 5) + 1664) * 40318u) >> 27) & 0x1F;
    if (d == 10) return (uint32_t)(((((uint64_t)x << 10) + 1665) * 1290167ull) >> 32) & 0x3FF;
    return (uint32_t)(((((uint64_t)x << 11) + 1664) * 645084ull) >> 31) & 0x7FF;
}
__device__ __forceinline__ uint32_t decompress_d(uint32_t t, int d) { return (t * (uint32_t)Q + (1u << (d - 1))) >> d; }

// little-endian bit packing of `cnt` d-bit values into bytes (the byte patterns of poly_compress / polyvec_compress)
__device__ __forceinline__ void pack_bits(u8 *dst, const uint32_t *t, int cnt, int d)
{
    uint64_t acc = 0; int have = 0, o = 0;
    for (int i = 0; i < cnt; i++) {
        acc |= (uint64_t)t[i] << have; have += d;
        while (have >= 8) { dst[o++] = (u8)acc; acc >>= 8; have -= 8; }
    }
}
__device__ __forceinline__ void unpack_bits(uint32_t *t, const u8 *src, int cnt, int d)
{
    uint64_t acc = 0; int have = 0, o = 0;
    for (int i = 0; i < cnt; i++) {
        while (have < d) { acc |= (uint64_t)src[o++] << have; have += 8; }
        t[i] = (uint32_t)acc & ((1u << d) - 1); acc >>= d; have -= d;
    }
}

// hash_g(a[32] || b[32]) = SHA3-512 of 64 bytes, warp-cooperative (one state word per lane); all 32 lanes of a warp call it,
// out[64] is written by lanes 0..7
__device__ __forceinline__ void warp_hash_g(const WarpKeccak &wk, const u8 *a, const u8 *b, u8 *out)
{
    const int lane = threadIdx.x & 31;
    uint64_t w = 0;
    if (lane < 8) { const u8 *src = lane < 4 ? a + 8 * lane : b + 8 * (lane - 4); for (int q = 0; q < 8; q++) w |= (uint64_t)src[q] << (8 * q); }
    if (lane == 8) w = 0x06ULL | 0x8000000000000000ULL;              // rate 72: pad starts and ends in word 8
    w = wk.permute(w);
    if (lane < 8) for (int q = 0; q < 8; q++) out[8 * lane + q] = (u8)(w >> (8 * q));
}
// hash_h(pk) = SHA3-256 of 384K + 32 bytes (a whole number of 8-byte words), warp-cooperative; out[32] written by lanes 0..3
template <int K>
__device__ __forceinline__ void warp_hash_h_pk(const WarpKeccak &wk, const u8 *pk, u8 *out)
{
    const int lane = threadIdx.x & 31;
    constexpr int NB = (384 * K + 32) / 136, REMW = ((384 * K + 32) % 136) / 8;
    auto word = [&](int i) -> uint64_t { uint64_t v = 0; for (int q = 0; q < 8; q++) v |= (uint64_t)pk[8 * i + q] << (8 * q); return v; };
    uint64_t a = 0;
#pragma unroll 1
    for (int blk = 0; blk < NB; blk++) { if (lane < 17) a ^= word(blk * 17 + lane); a = wk.permute(a); }
    if (lane < REMW) a ^= word(NB * 17 + lane);
    if (lane == REMW) a ^= 0x06ULL;
    if (lane == 16) a ^= 0x8000000000000000ULL;
    a = wk.permute(a);
    if (lane < 4) for (int q = 0; q < 8; q++) out[8 * lane + q] = (u8)(a >> (8 * q));
}

// part of indcpa_enc that needs only the public key: A^T = gen_at(seed) and t-hat = polyvec_frombytes(pk) (indcpa.c:273-276)
template <int K>
__device__ __forceinline__ void enc_public(KemSmem<K> &S, const u8 *pk, int tid, int first_thread)
{
    const int g = tid - first_thread;
    if (g >= 0 && g < K * K) {                     // transposed: xof_absorb(seed, i, j), indcpa.c:176-179
        uint64_t sd[4];
        for (int w = 0; w < 4; w++) { uint64_t v = 0; for (int q = 0; q < 8; q++) v |= (uint64_t)pk[384 * K + 8 * w + q] << (8 * q); sd[w] = v; }
        xof_rej_uniform(S.At[g], sd, (uint32_t)(g / K), (uint32_t)(g % K), S.blk[g]);
    }
    for (int i = 0; i < K; i++) {
        const u8 *a = pk + 384 * i + 3 * tid;
        S.t[i][2 * tid] = (u16)(((a[0] | ((uint32_t)a[1] << 8)) & 0xFFF) % Q);          // raw 12-bit values act mod q in basemul
        S.t[i][2 * tid + 1] = (u16)((((a[1] >> 4) | ((uint32_t)a[2] << 4)) & 0xFFF) % Q);
    }
}

// the rest of indcpa_enc (indcpa.c:278-303): needs S.m, S.coins, S.At, S.t; leaves the ciphertext in S.ct
template <int K>
__device__ __forceinline__ void enc_secret(KemSmem<K> &S, int tid)
{
    constexpr int ETA1 = (K == 2) ? 3 : 2, DU = (K == 4) ? 11 : 10, DV = (K == 4) ? 5 : 4;
    if (tid < 2 * K + 1) {                         // getnoise_eta1 x K, getnoise_eta2 x (K + 1), nonces 0..2K (indcpa.c:278-282)
        uint64_t key[4];
        for (int w = 0; w < 4; w++) { uint64_t v = 0; for (int q = 0; q < 8; q++) v |= (uint64_t)S.coins[8 * w + q] << (8 * q); key[w] = v; }
        if (tid < K) prf_cbd<ETA1>(S.sp[tid], key, (u8)tid, S.blk[tid]);
        else if (tid < 2 * K) prf_cbd<2>(S.ep[tid - K], key, (u8)tid, S.blk[tid]);
        else prf_cbd<2>(S.epp, key, (u8)tid, S.blk[tid]);
    }
    __syncthreads();
    for (int i = 0; i < K; i++) ntt256_block(S.sp[i], tid);
    for (int i = 0; i <= K; i++) {                 // b_i = A^T[i] o sp ; i == K: v = t o sp
        uint32_t a0 = 0, a1 = 0;
        for (int j = 0; j < K; j++) {
            const u16 *row = i < K ? S.At[i * K + j] : S.t[j];
            uint32_t r0, r1;
            basemul_pair(r0, r1, row[2 * tid], row[2 * tid + 1], S.sp[j][2 * tid], S.sp[j][2 * tid + 1], tid);
            a0 = gf_add(a0, r0); a1 = gf_add(a1, r1);
        }
        u16 *dst = i < K ? S.b[i] : S.v;
        dst[2 * tid] = (u16)a0; dst[2 * tid + 1] = (u16)a1;
    }
    __syncthreads();
    for (int i = 0; i < K; i++) invntt256_block(S.b[i], tid);
    invntt256_block(S.v, tid);
    for (int c = tid; c < 256; c += 128) {
        for (int i = 0; i < K; i++) S.b[i][c] = (u16)gf_add(S.b[i][c], S.ep[i][c]);
        const uint32_t mbit = (S.m[c >> 3] >> (c & 7)) & 1;                            // poly_frommsg: (q+1)/2 per set bit
        S.v[c] = (u16)gf_add(gf_add(S.v[c], S.epp[c]), mbit ? (Q + 1) / 2 : 0);
    }
    __syncthreads();
    // pack_ciphertext: polyvec_compress(b) || poly_compress(v)
    constexpr int GU = (DU == 10) ? 4 : 8, NGU = 256 / GU, BU = GU * DU / 8;
    for (int g = tid; g < K * NGU; g += 128) {
        const int i = g / NGU, q0 = (g % NGU) * GU;
        uint32_t t[8];
        for (int x = 0; x < GU; x++) t[x] = compress_d(S.b[i][q0 + x], DU);
        pack_bits(S.ct + (size_t)i * 32 * DU + (g % NGU) * BU, t, GU, DU);
    }
    if (tid < 32) {
        uint32_t t[8];
        for (int x = 0; x < 8; x++) t[x] = compress_d(S.v[8 * tid + x], DV);
        pack_bits(S.ct + K * 32 * DU + tid * DV, t, 8, DV);
    }
    __syncthreads();
}

// crypto_kem_enc_derand (kem.c:76-97); coins = the 32 random bytes of crypto_kem_enc (kem.c:114-122).  With `seeds` set the
// coins are drawn inside the kernel as randombytes() call number calls[b] of the KOSK DRBG seeded with seeds[b].
template <int K>
__global__ void __launch_bounds__(128) k_kem_enc(const u8 *__restrict__ pks, const u8 *__restrict__ coins, const u8 *__restrict__ seeds,
                                                 uint32_t call, u8 *__restrict__ cts, u8 *__restrict__ sss, int n)
{
    __shared__ KemSmem<K> S;
    const KemDims d = kem_dims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    const u8 *pk = pks + (size_t)(384 * K + 32) * b;
    if (tid < 32) {                          // warp 0: coins, hash_h(pk), hash_g(m || H(pk)), one state word per lane
        WarpKeccak wk; wk.init();
        if (seeds) {
            uint64_t a = 0;
            if (tid < 4) for (int j = 0; j < 8; j++) a |= (uint64_t)seeds[32 * (size_t)b + 8 * tid + j] << (8 * j);
            if (tid == 4) a = (uint64_t)call | (0x1FULL << 32);
            if (tid == 16) a = 0x8000000000000000ULL;
            a = wk.permute(a);
            if (tid < 4) for (int j = 0; j < 8; j++) S.m[8 * tid + j] = (u8)(a >> (8 * j));
        } else S.m[tid] = coins[32 * (size_t)b + tid];
        warp_hash_h_pk<K>(wk, pk, S.hpk);
        __syncwarp();
        warp_hash_g(wk, S.m, S.hpk, S.kr);
        __syncwarp();
        S.coins[tid] = S.kr[32 + tid];
    }
    enc_public<K>(S, pk, tid, 32);           // A^T on the second warp while warp 0 hashes
    __syncthreads();
    enc_secret<K>(S, tid);
    for (int i = tid; i < d.ct_bytes; i += 128) cts[(size_t)d.ct_bytes * b + i] = S.ct[i];
    if (tid < 32) sss[32 * (size_t)b + tid] = S.kr[tid];
}

// crypto_kem_dec (kem.c:139-169) with indcpa_dec (indcpa.c:318-336)
template <int K>
__global__ void __launch_bounds__(128) k_kem_dec(const u8 *__restrict__ cts, const u8 *__restrict__ sks, u8 *__restrict__ sss, int n)
{
    __shared__ KemSmem<K> S;
    constexpr int DU = (K == 4) ? 11 : 10, DV = (K == 4) ? 5 : 4;
    const KemDims d = kem_dims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    const u8 *ct = cts + (size_t)d.ct_bytes * b, *sk = sks + (size_t)(768 * K + 96) * b, *pk = sk + 384 * K;
    // unpack_ciphertext + unpack_sk
    constexpr int GU = (DU == 10) ? 4 : 8, NGU = 256 / GU, BU = GU * DU / 8;
    for (int g = tid; g < K * NGU; g += 128) {
        const int i = g / NGU, q0 = (g % NGU) * GU;
        uint32_t t[8];
        unpack_bits(t, ct + (size_t)i * 32 * DU + (g % NGU) * BU, GU, DU);
        for (int x = 0; x < GU; x++) S.b[i][q0 + x] = (u16)decompress_d(t[x], DU);
    }
    if (tid < 32) {
        uint32_t t[8];
        unpack_bits(t, ct + K * 32 * DU + tid * DV, 8, DV);
        for (int x = 0; x < 8; x++) S.v[8 * tid + x] = (u16)decompress_d(t[x], DV);
    }
    for (int i = 0; i < K; i++) {
        const u8 *a = sk + 384 * i + 3 * tid;
        S.sp[i][2 * tid] = (u16)(((a[0] | ((uint32_t)a[1] << 8)) & 0xFFF) % Q);
        S.sp[i][2 * tid + 1] = (u16)((((a[1] >> 4) | ((uint32_t)a[2] << 4)) & 0xFFF) % Q);
    }
    if (tid < 8) S.mw[tid] = 0;
    enc_public<K>(S, pk, tid, 32);           // for the re-encryption; independent of the message
    __syncthreads();
    for (int i = 0; i < K; i++) ntt256_block(S.b[i], tid);
    {
        uint32_t a0 = 0, a1 = 0;
        for (int j = 0; j < K; j++) {
            uint32_t r0, r1;
            basemul_pair(r0, r1, S.sp[j][2 * tid], S.sp[j][2 * tid + 1], S.b[j][2 * tid], S.b[j][2 * tid + 1], tid);
            a0 = gf_add(a0, r0); a1 = gf_add(a1, r1);
        }
        S.epp[2 * tid] = (u16)a0; S.epp[2 * tid + 1] = (u16)a1;
    }
    __syncthreads();
    invntt256_block(S.epp, tid);
    for (int c = tid; c < 256; c += 128) {
        // poly_tomsg on the centered representative (poly_reduce), in the reference's 32-bit arithmetic (poly.c:208-219)
        uint32_t t = (uint32_t)gf_center(gf_sub(S.v[c], S.epp[c]));
        t <<= 1; t += 1665; t *= 80635u; t >>= 28; t &= 1;
        // no branch on the message bit: a warp covers 32 consecutive coefficients, the ballot is their message word
        const uint32_t word = __ballot_sync(0xffffffffu, t != 0);
        if ((tid & 31) == 0) S.mw[c >> 5] = word;
    }
    __syncthreads();
    if (tid < 32) {                          // hash_g(m' || H(pk) from sk), warp-cooperative
        WarpKeccak wk; wk.init();
        S.m[tid] = (u8)(S.mw[tid >> 2] >> (8 * (tid & 3)));
        __syncwarp();
        warp_hash_g(wk, S.m, sk + 768 * K + 32, S.kr);
        __syncwarp();
        S.coins[tid] = S.kr[32 + tid];
        if (tid == 0) S.fail = 0;
    }
    __syncthreads();
    enc_secret<K>(S, tid);
    int bad = 0;
    for (int i = tid; i < d.ct_bytes; i += 128) bad |= S.ct[i] != ct[i];
    atomicOr(&S.fail, bad);                  // unconditional: no path depends on whether the re-encryption matched
    __syncthreads();
    // Implicit rejection (kem.c:157-166): rkprf(z, ct) = SHAKE256(z || ct) (symmetric-shake.c:64-73) is ALWAYS computed, warp-cooperatively,
    // and the output is selected with a mask like the reference's verify() + cmov_int: accepted and rejected ciphertexts do the same work.
    if (tid < 32) {
        WarpKeccak wk; wk.init();
        constexpr int CTB = K * 32 * DU + 32 * DV, NW = 4 + CTB / 8, NB = NW / 17, REM = NW % 17;
        static_assert(CTB % 8 == 0, "ciphertext is a whole number of sponge words");
        const uint64_t *zw = reinterpret_cast<const uint64_t *>(sk + 768 * K + 64), *cw = reinterpret_cast<const uint64_t *>(ct);
        auto word = [&](int i) -> uint64_t { return i < 4 ? zw[i] : cw[i - 4]; };
        uint64_t a = 0;
#pragma unroll 1
        for (int blk = 0; blk < NB; blk++) { if (tid < 17) a ^= word(blk * 17 + tid); a = wk.permute(a); }
        if (tid < REM) a ^= word(NB * 17 + tid);
        if (tid == REM) a ^= 0x1FULL;
        if (tid == 16) a ^= 0x8000000000000000ULL;
        a = wk.permute(a);
        if (tid < 4) {
            uint64_t kr = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) kr |= (uint64_t)S.kr[8 * tid + i] << (8 * i);
            const uint64_t mask = 0ULL - (uint64_t)(S.fail != 0);
            const uint64_t out = kr ^ (mask & (kr ^ a));
#pragma unroll
            for (int i = 0; i < 8; i++) sss[32 * (size_t)b + 8 * tid + i] = (u8)(out >> (8 * i));
        }
    }
}

}  // namespace kosk

// ---- host side ----
struct KemState {
    u8 *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_d = nullptr; size_t cap = 0;       // grow-only staging of the host-buffer API
    u8 *d_coins64 = nullptr;                                                                   // [chunk][64] coins of crypto_kem_keypair_derand
    bool tables = false;
};
static void kem_free(KemState &ks)
{
    void *p[] = {ks.d_a, ks.d_b, ks.d_c, ks.d_d, ks.d_coins64};
    for (void *q : p) if (q) cudaFree(q);
    ks = KemState{};
}
static int kem_tables(kosk_b200_ctx *c)
{
    if (c->kem->tables) return KOSK_OK;
    uint16_t hz[128];
    for (int i = 0; i < 128; i++) { int br = 0; for (int b = 0; b < 7; b++) br |= ((i >> b) & 1) << (6 - b); hz[i] = (uint16_t)h_pow(h_pow(17, br), Q - 2); }
    CU(cudaMemcpyToSymbol(c_zeta_inv, hz, sizeof hz));
    c->kem->tables = true;
    return KOSK_OK;
}
static int kem_reserve(kosk_b200_ctx *c, size_t n)
{
    KemState &ks = *c->kem;
    if (n <= ks.cap) return KOSK_OK;
    kem_free(ks); ks.tables = false;
    const size_t sk = c->L.sk_bytes, ctb = (size_t)kem_dims(c->k).ct_bytes;
    if (cudaMalloc((void **)&ks.d_a, n * sk) != cudaSuccess || cudaMalloc((void **)&ks.d_b, n * ctb) != cudaSuccess ||
        cudaMalloc((void **)&ks.d_c, n * 32) != cudaSuccess || cudaMalloc((void **)&ks.d_d, n * 32) != cudaSuccess) { kem_free(ks); return fail(KOSK_E_NOMEM, "cudaMalloc failed for KEM staging"); }
    ks.cap = n;
    return KOSK_OK;
}
static int kem_enc_launch(kosk_b200_ctx *c, size_t n, const u8 *d_pk, const u8 *d_coins, const u8 *d_seeds, uint32_t call, u8 *d_ct, u8 *d_ss, cudaStream_t st)
{
    int rc = kem_tables(c); if (rc) return rc;
    for (size_t o = 0; o < n; o += 1u << 20) {
        const int m = (int)std::min<size_t>(1u << 20, n - o);
        const u8 *pk = d_pk + c->L.pk_bytes * o, *co = d_coins ? d_coins + 32 * o : nullptr, *se = d_seeds ? d_seeds + 32 * o : nullptr;
        u8 *ct = d_ct + (size_t)kem_dims(c->k).ct_bytes * o, *ss = d_ss + 32 * o;
        switch (c->k) {
        case 2: k_kem_enc<2><<<m, 128, 0, st>>>(pk, co, se, call, ct, ss, m); break;
        case 3: k_kem_enc<3><<<m, 128, 0, st>>>(pk, co, se, call, ct, ss, m); break;
        default: k_kem_enc<4><<<m, 128, 0, st>>>(pk, co, se, call, ct, ss, m); break;
        }
        c->launches++;
    }
    CU(cudaGetLastError());
    return KOSK_OK;
}
static int kem_dec_launch(kosk_b200_ctx *c, size_t n, const u8 *d_ct, const u8 *d_sk, u8 *d_ss, cudaStream_t st)
{
    int rc = kem_tables(c); if (rc) return rc;
    for (size_t o = 0; o < n; o += 1u << 20) {
        const int m = (int)std::min<size_t>(1u << 20, n - o);
        const u8 *ct = d_ct + (size_t)kem_dims(c->k).ct_bytes * o, *sk = d_sk + c->L.sk_bytes * o; u8 *ss = d_ss + 32 * o;
        switch (c->k) {
        case 2: k_kem_dec<2><<<m, 128, 0, st>>>(ct, sk, ss, m); break;
        case 3: k_kem_dec<3><<<m, 128, 0, st>>>(ct, sk, ss, m); break;
        default: k_kem_dec<4><<<m, 128, 0, st>>>(ct, sk, ss, m); break;
        }
        c->launches++;
    }
    CU(cudaGetLastError());
    return KOSK_OK;
}

template <int K> static void kem_launch_keygen(const ProveBufs &pb, int B, cudaStream_t st) { k_keygen<K><<<B, 128, 0, st>>>(pb); }

extern "C" {

size_t kosk_b200_ct_bytes(int k) { return (k >= 2 && k <= 4) ? (size_t)kem_dims(k).ct_bytes : 0; }

int kosk_b200_kem_enc_derand_batch_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_pk, const uint8_t *d_coins, uint8_t *d_ct, uint8_t *d_ss, void *stream)
{
    if (!c || !d_pk || !d_coins || !d_ct || !d_ss) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    return n ? kem_enc_launch(c, n, d_pk, d_coins, nullptr, 0, d_ct, d_ss, (cudaStream_t)stream) : KOSK_OK;
}
int kosk_b200_kem_dec_batch_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_ct, const uint8_t *d_sk, uint8_t *d_ss, void *stream)
{
    if (!c || !d_ct || !d_sk || !d_ss) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    return n ? kem_dec_launch(c, n, d_ct, d_sk, d_ss, (cudaStream_t)stream) : KOSK_OK;
}

static int kem_enc_host(kosk_b200_ctx *c, size_t n, const uint8_t *pk, const uint8_t *coins, const uint8_t *seed, uint32_t call, uint8_t *ct, uint8_t *ss)
{
    CU(cudaSetDevice(c->device));
    if (n == 0) return KOSK_OK;
    int rc = kem_reserve(c, n); if (rc) return rc;
    KemState &ks = *c->kem; cudaStream_t st = c->lanes[0].st;
    const size_t ctb = (size_t)kem_dims(c->k).ct_bytes;
    CU(cudaMemcpyAsync(ks.d_a, pk, n * c->L.pk_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ks.d_c, coins ? coins : seed, n * 32, cudaMemcpyHostToDevice, st));
    rc = kem_enc_launch(c, n, ks.d_a, coins ? ks.d_c : nullptr, coins ? nullptr : ks.d_c, call, ks.d_b, ks.d_d, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ct, ks.d_b, n * ctb, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ss, ks.d_d, n * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KOSK_OK;
}

int kosk_b200_kem_enc_derand_batch(kosk_b200_ctx *c, size_t n, const uint8_t *pk, const uint8_t *coins, uint8_t *ct, uint8_t *ss)
{
    if (!c || !pk || !coins || !ct || !ss) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    return kem_enc_host(c, n, pk, coins, nullptr, 0, ct, ss);
}

int kosk_b200_kem_dec_batch(kosk_b200_ctx *c, size_t n, const uint8_t *ct, const uint8_t *sk, uint8_t *ss)
{
    if (!c || !ct || !sk || !ss) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    if (n == 0) return KOSK_OK;
    int rc = kem_reserve(c, n); if (rc) return rc;
    KemState &ks = *c->kem; cudaStream_t st = c->lanes[0].st;
    const size_t ctb = (size_t)kem_dims(c->k).ct_bytes;
    CU(cudaMemcpyAsync(ks.d_a, sk, n * c->L.sk_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ks.d_b, ct, n * ctb, cudaMemcpyHostToDevice, st));
    rc = kem_dec_launch(c, n, ks.d_b, ks.d_a, ks.d_d, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ss, ks.d_d, n * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KOSK_OK;
}

// crypto_kem_keypair_derand / crypto_kem_keypair (kem.c:23-58): the keygen kernel of the KOSK path with z taken from the coins
static int kem_keypair_host(kosk_b200_ctx *c, size_t n, const uint8_t *coins /* [n][64] or nullptr = one DRBG call */, uint8_t *pk, uint8_t *sk)
{
    CU(cudaSetDevice(c->device));
    KemState &ks = *c->kem; Lane &ln = c->lanes[0];
    if (coins && !ks.d_coins64 && cudaMalloc((void **)&ks.d_coins64, (size_t)c->chunk * 64) != cudaSuccess) return fail(KOSK_E_NOMEM, "cudaMalloc failed for KEM coins");
    CU(cudaStreamSynchronize(ln.st));
    for (size_t o = 0; o < n; o += (size_t)c->chunk) {
        const int B = (int)std::min<size_t>((size_t)c->chunk, n - o);
        ProveBufs pb = ln.pb; pb.seeds = ln.d_seeds; pb.pk = ln.d_pk; pb.sk = ln.d_sk; pb.pi = ln.d_pi; pb.B = B;
        if (coins) { CU(cudaMemcpyAsync(ks.d_coins64, coins + 64 * o, 64 * (size_t)B, cudaMemcpyHostToDevice, ln.st)); pb.kem_mode = 1; pb.kem_coins = ks.d_coins64; }
        else { CU(cudaMemcpyAsync(ln.d_seeds, c->raw->seed, 32, cudaMemcpyHostToDevice, ln.st)); pb.kem_mode = 2; pb.cb_key = (int)c->raw->calls; }
        switch (c->k) { case 2: kem_launch_keygen<2>(pb, B, ln.st); break; case 3: kem_launch_keygen<3>(pb, B, ln.st); break; default: kem_launch_keygen<4>(pb, B, ln.st); }
        c->launches++;
        CU(cudaMemcpyAsync(pk + c->L.pk_bytes * o, ln.d_pk, c->L.pk_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        CU(cudaMemcpyAsync(sk + c->L.sk_bytes * o, ln.d_sk, c->L.sk_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        CU(cudaStreamSynchronize(ln.st));
    }
    CU(cudaGetLastError());
    return KOSK_OK;
}
int kosk_b200_kem_keypair_derand_batch(kosk_b200_ctx *c, size_t n, const uint8_t *coins, uint8_t *pk, uint8_t *sk)
{
    if (!c || !coins || !pk || !sk) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    return n ? kem_keypair_host(c, n, coins, pk, sk) : KOSK_OK;
}
int kosk_b200_kem_keypair(kosk_b200_ctx *c, uint8_t *pk, uint8_t *sk)
{
    if (!c || !pk || !sk) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    int rc = kem_keypair_host(c, 1, nullptr, pk, sk);
    if (!rc) c->raw->calls += 1;                 // randombytes(coins, 64), kem.c:52
    return rc;
}

// crypto_kem_enc (kem.c:114-122): coins = randombytes(32) = the next call of the context DRBG (kosk_b200_rng_reset)
int kosk_b200_kem_enc(kosk_b200_ctx *c, uint8_t *ct, uint8_t *ss, const uint8_t *pk)
{
    if (!c || !ct || !ss || !pk) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    int rc = kem_enc_host(c, 1, pk, nullptr, c->raw->seed, c->raw->calls, ct, ss);
    if (!rc) c->raw->calls += 1;
    return rc;
}
int kosk_b200_kem_dec(kosk_b200_ctx *c, uint8_t *ss, const uint8_t *ct, const uint8_t *sk)
{
    return kosk_b200_kem_dec_batch(c, 1, ct, sk, ss);
}

}  // extern "C"
