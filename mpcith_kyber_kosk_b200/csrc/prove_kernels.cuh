// prove_kernels.cuh -- sm_100a kernels of the KOSK prover (kyber_verifiable_keygen, reference kosk.cpp:72-86).
// One launch processes a whole chunk of B independent proofs; per-proof state lives in HBM as
//   Y   [B][n2][YLD]      u16  sharing inputs: 256 packed secrets | 151 tail randoms   (ss.cpp:13-20)
//   SH  [B][nslot][SLD]   u16  planes: one row of 1454 party shares per slot           (share_vec.share_y)
//   BG  [B][NP][2*BGH]    u16  beta[70]+pad | gamma[70]+pad per party                          (mpcith_vp_state)
//   TCR [B][NP][32]       u8   party commitments, VWR the same for view hashes
// Parties are the contiguous (coalesced) axis of every per-party kernel.
#pragma once
#include "keccak.cuh"
#include "gf_gemm.cuh"

namespace kosk {

constexpr int BGH = 144;        // u16 per party of one beta (or gamma) row in BG: 70 values + padding, 16B-aligned

__constant__ u16 c_zeta[128];   // 17^brv7(i) mod q: plain-residue form of kyber/ntt.c:39-56

struct ProveBufs {
    const u8 *seeds;   // [B][32]
    u16 *Y, *SH, *BG;
    u8 *TCR, *VWR;
    u16 *PW;           // [B][NA][F] powers of the FS-1 challenges alpha (mlwe_prover.cpp:144-153)
    u16 *AH;           // [B][k*k][256] matrix A-hat (NTT domain), row-major A[i][j]
    u16 *SHAT;         // [B][k][256]   s-hat
    u16 *I, *REST;     // [B][150], [B][1304]
    int8_t *YL0, *YL1; // experimental tensor path only: 7-bit limb planes of Y, [B][n2][YLD] int8
    int32_t *WS;       // [GE_WS_ELEMS] partial sums of the split-K share evaluation (latency mode, a handful of proofs)
    u8 *pk, *sk, *pi;  // outputs
    int B;
    // randombytes() call numbers (KOSK counter-mode DRBG) of the first call made by kyber_keygen, prepare_randomness,
    // prepare_range_proof and prove.  kyber_verifiable_keygen order (kosk.cpp:72-86): 0, 1, 1+3F, c_se0;
    // the struct-level API (raw_api.cuh) numbers them by the caller's own call order, as the reference's global RNG would.
    int cb_key, cb_rand, cb_eta, cb_prove;
    int tails_mask;    // k_tails: 1 = f / NTT_f sharings, 2 = eta sharings, 4 = sharings made inside prove() (s, e, z_j, A s)
    // k_keygen as crypto_kem_keypair (kyber/kem.c:23-58) instead of kyber_keygen (kosk.cpp:4-70): 0 = kyber_keygen (z = the noise seed,
    // kosk.cpp:66-68), 1 = crypto_kem_keypair_derand on kem_coins[B][64] (d | z), 2 = crypto_kem_keypair (d | z = one 64-byte DRBG call)
    int kem_mode; const u8 *kem_coins;
};
__host__ __device__ inline void set_default_calls(ProveBufs &pb, const Slots &sl)
{
    pb.cb_key = 0; pb.cb_rand = sl.c_seed0; pb.cb_eta = sl.c_eta0; pb.cb_prove = sl.c_se0; pb.tails_mask = 7;
    pb.kem_mode = 0; pb.kem_coins = nullptr;
}

__device__ __forceinline__ u16 *yrow(const ProveBufs &pb, const Slots &sl, int b, int slot) { return pb.Y + ((size_t)b * sl.n2 + slot) * YLD; }
__device__ __forceinline__ u16 *plane(const ProveBufs &pb, const Slots &sl, int b, int slot) { return pb.SH + ((size_t)b * sl.nslot + slot) * SLD + SOFF; }

// ---------------------------------------------------------------------------------------------
// 256-point Kyber NTT on canonical residues, 128 threads, data in shared memory.
// Reference: kyber/ntt.c:80-95 + poly_reduce (poly.c:261-265).
__device__ __forceinline__ void ntt256_block(u16 *p, int tid)
{
    for (int len = 128; len >= 2; len >>= 1) {
        int grp = tid / len, j = grp * 2 * len + (tid % len);
        uint32_t z = c_zeta[128 / len + grp];
        uint32_t t = gf_mul(z, p[j + len]), u = p[j];
        p[j + len] = (u16)gf_sub(u, t);
        p[j] = (u16)gf_add(u, t);
        __syncthreads();
    }
}
// pairwise product in Z_q[X]/(X^2 - zeta): kyber/ntt.c:139-146 with the signs of poly.c:290-297; plain residues
__device__ __forceinline__ void basemul_pair(uint32_t &r0, uint32_t &r1, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, int pair)
{
    uint32_t z = c_zeta[64 + (pair >> 1)];
    if (pair & 1) z = gf_sub(0, z);
    r0 = gf_add(gf_mul(gf_mul(a1, b1), z), gf_mul(a0, b0));
    r1 = gf_add(gf_mul(a0, b1), gf_mul(a1, b0));
}

// ---------------------------------------------------------------------------------------------
// Thread-level samplers with the Keccak state in registers; `blk` is scratch private to the thread (shared memory).
// SHAKE128(seed || b0 || b1) -> 256 residues by 12-bit rejection: gen_matrix / rej_uniform (indcpa.c:124-193).
__device__ __forceinline__ void xof_rej_uniform(u16 *dst, const uint64_t seed[4], uint32_t b0, uint32_t b1, uint64_t *blk /* [21] */)
{
    uint64_t a[25];
    keccak_zero(a);
    a[0] = seed[0]; a[1] = seed[1]; a[2] = seed[2]; a[3] = seed[3];
    a[4] = (uint64_t)b0 | ((uint64_t)b1 << 8) | (0x1FULL << 16);
    a[20] = 0x8000000000000000ULL;                  // rate 168 = lanes 0..20
    int ctr = 0;
    while (ctr < 256) {
        keccak_f1600(a);
#pragma unroll
        for (int l = 0; l < 21; l++) blk[l] = a[l];
        const u8 *p = reinterpret_cast<const u8 *>(blk);
        for (int g = 0; g < 56 && ctr < 256; g++) {
            const uint32_t x0 = p[3 * g], x1 = p[3 * g + 1], x2 = p[3 * g + 2];
            const uint32_t v0 = (x0 | (x1 << 8)) & 0xFFF, v1 = ((x1 >> 4) | (x2 << 4)) & 0xFFF;
            if (v0 < (uint32_t)Q) dst[ctr++] = (u16)v0;
            if (ctr < 256 && v1 < (uint32_t)Q) dst[ctr++] = (u16)v1;
        }
    }
}
// CBD_eta(SHAKE256(key || nonce)): poly_getnoise_eta1/eta2 (poly.c:225-249, cbd.c:58-107), canonical residues
template <int ETA>
__device__ __forceinline__ void prf_cbd(u16 *dst, const uint64_t key[4], u8 nonce, uint64_t *blk /* [24] */)
{
    uint64_t a[25];
    prf_begin(a, key, nonce);
#pragma unroll
    for (int l = 0; l < 17; l++) blk[l] = a[l];
    if (ETA == 3) {                                 // 192 bytes = 136 + 56
        keccak_f1600(a);
#pragma unroll
        for (int l = 0; l < 7; l++) blk[17 + l] = a[l];
    }
    const u8 *p = reinterpret_cast<const u8 *>(blk);
    if (ETA == 2) {
        for (int i = 0; i < 32; i++) {
            const uint32_t t = reinterpret_cast<const uint32_t *>(p)[i];
            const uint32_t d = (t & 0x55555555u) + ((t >> 1) & 0x55555555u);
            for (int j = 0; j < 8; j++) { const int x = (int)((d >> (4 * j)) & 3) - (int)((d >> (4 * j + 2)) & 3); dst[8 * i + j] = (u16)(x < 0 ? x + Q : x); }
        }
    } else {
        for (int i = 0; i < 64; i++) {
            const uint32_t t = (uint32_t)p[3 * i] | ((uint32_t)p[3 * i + 1] << 8) | ((uint32_t)p[3 * i + 2] << 16);
            const uint32_t d = (t & 0x249249u) + ((t >> 1) & 0x249249u) + ((t >> 2) & 0x249249u);
            for (int j = 0; j < 4; j++) { const int x = (int)((d >> (6 * j)) & 7) - (int)((d >> (6 * j + 3)) & 7); dst[4 * i + j] = (u16)(x < 0 ? x + Q : x); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Keygen: kosk.cpp:4-70 (gen_matrix indcpa.c:168-193, poly_getnoise_eta1 poly.c:225-230, cbd.c:58-107,
// polyvec_ntt, basemul_acc + tomont, tobytes poly.c:124-139).  One CTA of 128 threads per proof: the two seed hashes and
// H(pk) run warp-cooperatively on warp 0, the K*K matrix XOFs on the lanes of warp 0 and the 2K noise PRFs on warp 1.
// Also emits the secrets of the s/e sharings and of the range-proof products z_j (mlwe_prover.cpp:338-372:
// recon_secrets_2ddeg of a product sharing is the pointwise product of the secrets).
template <int K>
__global__ void __launch_bounds__(128) k_keygen(ProveBufs pb)
{
    constexpr int ETA = (K == 2) ? 3 : 2, M = 2 * ETA;
    const Slots sl = make_slots(K);
    const Layout L = make_layout(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 sA[K * K][256];
    __shared__ u16 sS[K][256], sE[K][256], sSh[K][256], sEh[K][256], sT[K][256];
    __shared__ uint64_t sSeed[8], sZ[4];
    __shared__ uint64_t sBlk[K * K + 2 * K][24];
    __shared__ __align__(8) u8 sPk[384 * K + 32];

    if (tid < 32) {
        // randombytes(buf, 64) (only bytes 0..31 are used, kosk.cpp:12-14), then sha3_512(coins || K): one state word per lane
        WarpKeccak wk; wk.init();
        uint64_t a = 0;
        if (pb.kem_mode == 1) {                                // explicit coins: d = coins[0..31], z = coins[32..63]
            if (tid < 8) for (int q = 0; q < 8; q++) a |= (uint64_t)pb.kem_coins[64 * (size_t)b + 8 * tid + q] << (8 * q);
        } else {
            const uint64_t *gs = reinterpret_cast<const uint64_t *>(pb.seeds + 32 * (size_t)b);
            if (tid < 4) a = gs[tid];
            if (tid == 4) a = (uint64_t)(uint32_t)pb.cb_key | (0x1FULL << 32);
            if (tid == 16) a = 0x8000000000000000ULL;              // SHAKE256 rate 136 = lanes 0..16
            a = wk.permute(a);
        }
        if (tid >= 4 && tid < 8) sZ[tid - 4] = a;              // bytes 32..63 of the random block: z of crypto_kem_keypair
        if (tid >= 4) a = 0;
        if (tid == 4) a = (uint64_t)K | (0x06ULL << 8);
        if (tid == 8) a = 0x8000000000000000ULL;               // SHA3-512 rate 72 = lanes 0..8
        a = wk.permute(a);
        if (tid < 8) sSeed[tid] = a;
    }
    __syncthreads();
    if (tid < K * K) {                              // A[i][j] <- SHAKE128(publicseed || j || i), 12-bit rejection
        const uint64_t ps[4] = {sSeed[0], sSeed[1], sSeed[2], sSeed[3]};
        xof_rej_uniform(sA[tid], ps, (uint32_t)(tid % K), (uint32_t)(tid / K), sBlk[tid]);
    } else if (tid >= 32 && tid < 32 + 2 * K) {     // noise: PRF(noiseseed, nonce) -> CBD_eta
        const int idx = tid - 32;
        const uint64_t ns[4] = {sSeed[4], sSeed[5], sSeed[6], sSeed[7]};
        prf_cbd<ETA>(idx < K ? sS[idx] : sE[idx - K], ns, (u8)idx, sBlk[K * K + idx]);
    }
    __syncthreads();
    // secrets of [s_i], [e_i] and of the 2*eta range-proof products z_j = prod_{m<=j+1} (x - eta_m)
    for (int c = tid; c < 256; c += 128)
        for (int i = 0; i < K; i++)
            for (int w = 0; w < 2; w++) {
                uint32_t x = w ? sE[i][c] : sS[i][c];
                yrow(pb, sl, b, (w ? sl.e0 : sl.s0) + i)[c] = (u16)x;
                uint32_t z = gf_sub(x, gf_sub(0, ETA));             // x - (-eta)
                for (int j = 0; j < M; j++) {
                    uint32_t eta_m = (j + 1 >= ETA) ? (uint32_t)(j + 1 - ETA) : (uint32_t)(Q + j + 1 - ETA);
                    z = gf_mul(z, gf_sub(x, eta_m));
                    yrow(pb, sl, b, (w ? sl.ze0 : sl.zs0) + i * M + j)[c] = (u16)z;
                }
            }
    for (int c = tid; c < 256; c += 128)
        for (int i = 0; i < K; i++) { sSh[i][c] = sS[i][c]; sEh[i][c] = sE[i][c]; }
    __syncthreads();
    for (int i = 0; i < K; i++) { ntt256_block(sSh[i], tid); ntt256_block(sEh[i], tid); }
    // t-hat = A-hat o s-hat + e-hat
    for (int i = 0; i < K; i++) {
        uint32_t a0 = 0, a1 = 0;
        for (int j = 0; j < K; j++) {
            uint32_t r0, r1;
            basemul_pair(r0, r1, sA[i * K + j][2 * tid], sA[i * K + j][2 * tid + 1], sSh[j][2 * tid], sSh[j][2 * tid + 1], tid);
            a0 = gf_add(a0, r0); a1 = gf_add(a1, r1);
        }
        sT[i][2 * tid] = (u16)gf_add(a0, sEh[i][2 * tid]); sT[i][2 * tid + 1] = (u16)gf_add(a1, sEh[i][2 * tid + 1]);
    }
    __syncthreads();
    // pack pk = tobytes(t-hat) || publicseed ; sk = tobytes(s-hat) || pk || SHA3-256(pk) || noiseseed (kosk.cpp:57-69)
    u8 *pk = pb.pk + L.pk_bytes * (size_t)b, *sk = pb.sk + L.sk_bytes * (size_t)b;
    for (int i = 0; i < K; i++) {
        uint32_t t0 = sT[i][2 * tid], t1 = sT[i][2 * tid + 1];
        u8 *d = sPk + 384 * i + 3 * tid;
        d[0] = (u8)t0; d[1] = (u8)((t0 >> 8) | (t1 << 4)); d[2] = (u8)(t1 >> 4);
        t0 = sSh[i][2 * tid]; t1 = sSh[i][2 * tid + 1];
        u8 *e = sk + 384 * i + 3 * tid;
        e[0] = (u8)t0; e[1] = (u8)((t0 >> 8) | (t1 << 4)); e[2] = (u8)(t1 >> 4);
    }
    if (tid < 32) sPk[384 * K + tid] = reinterpret_cast<const u8 *>(sSeed)[tid];
    __syncthreads();
    for (int i = tid; i < 384 * K + 32; i += 128) { pk[i] = sPk[i]; sk[384 * K + i] = sPk[i]; }
    if (tid < 32) sk[L.sk_bytes - 32 + tid] = pb.kem_mode ? reinterpret_cast<const u8 *>(sZ)[tid] : reinterpret_cast<const u8 *>(sSeed + 4)[tid];
    if (tid < 32) {                                 // sha3_256(pk), warp-cooperative: 800 / 1184 / 1568 bytes = 5 / 8 / 11 full blocks + 15 / 12 / 9 words
        WarpKeccak wk; wk.init();
        constexpr int NB = (384 * K + 32) / 136, REMW = ((384 * K + 32) % 136) / 8;
        static_assert((384 * K + 32) % 8 == 0, "pk is a whole number of sponge words");
        const uint64_t *w = reinterpret_cast<const uint64_t *>(sPk);
        uint64_t a = 0;
#pragma unroll 1
        for (int blk = 0; blk < NB; blk++) { if (tid < 17) a ^= w[blk * 17 + tid]; a = wk.permute(a); }
        if (tid < REMW) a ^= w[NB * 17 + tid];
        if (tid == REMW) a ^= 0x06ULL;
        if (tid == 16) a ^= 0x8000000000000000ULL;
        a = wk.permute(a);
        if (tid < 4) for (int i = 0; i < 8; i++) sk[L.sk_bytes - 64 + 8 * tid + i] = (u8)(a >> (8 * i));
    }
    // keep A-hat and s-hat for the online phase
    for (int i = tid; i < K * K * 256; i += 128) pb.AH[(size_t)b * K * K * 256 + i] = (&sA[0][0])[i];
    for (int i = tid; i < K * 256; i += 128) pb.SHAT[(size_t)b * K * 256 + i] = (&sSh[0][0])[i];
}

// ---------------------------------------------------------------------------------------------
// prepare_randomness, PRF part (mlwe_prover.cpp:8-14): seed_i = randombytes(32); f_i = BE16(SHAKE256(seed_i||i)[0:512]) mod q.
// One thread per (proof, i); five permutations each.
template <int K>
__global__ void __launch_bounds__(64) k_expand_f(ProveBufs pb)
{
    const Slots sl = make_slots(K);
    // rows are staged in shared memory and written out by the whole CTA: a thread storing its own row directly touches one 32-byte
    // sector per 2-byte element (rows are 832 bytes apart)
    __shared__ __align__(16) u16 srow[64][256 + 8];
    const int gid0 = blockIdx.x * 64, gid = gid0 + threadIdx.x, total = pb.B * sl.F;
    if (gid < total) {
        const int b = gid / sl.F, i = gid % sl.F;
        uint64_t sd[4], a[25];
        const uint64_t *gs = reinterpret_cast<const uint64_t *>(pb.seeds + 32 * (size_t)b);
#pragma unroll
        for (int w = 0; w < 4; w++) sd[w] = gs[w];
        drbg_begin(a, sd, pb.cb_rand + i);
        uint64_t key[4] = {a[0], a[1], a[2], a[3]};
        prf_begin(a, key, (u8)i);
        u16 *dst = srow[threadIdx.x];
#pragma unroll 1
        for (int blk = 0; blk < 4; blk++) {             // 512 bytes = 3 x 136 + 104
#pragma unroll
            for (int v = 0; v < 68; v++) {
                if (blk * 68 + v < 256) dst[blk * 68 + v] = (u16)(lane_be16(a[v >> 2], v & 3) % (uint32_t)Q);
            }
            if (blk < 3) keccak_f1600(a);
        }
    }
    __syncthreads();
    const int nrow = min(64, total - gid0);
    for (int idx = threadIdx.x; idx < nrow * 32; idx += 64) {         // 32 x 16-byte chunks per row
        const int r = idx >> 5, c = idx & 31, g = gid0 + r;
        u16 *dst = yrow(pb, sl, g / sl.F, sl.f0 + g % sl.F);
        reinterpret_cast<uint4 *>(dst)[c] = reinterpret_cast<const uint4 *>(srow[r])[c];
    }
}

// NTT_f_i = NTT(f_i) (mlwe_prover.cpp:17-26): one CTA per (proof, i)
template <int K>
__global__ void __launch_bounds__(128) k_ntt_f(ProveBufs pb)
{
    const Slots sl = make_slots(K);
    const int b = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 p[256];
    const u16 *src = yrow(pb, sl, b, sl.f0 + i);
    p[tid] = src[tid]; p[tid + 128] = src[tid + 128];
    __syncthreads();
    ntt256_block(p, tid);
    u16 *dst = yrow(pb, sl, b, sl.Tf0 + i);
    dst[tid] = p[tid]; dst[tid + 128] = p[tid + 128];
}

// The 151 tail randoms of every fresh sharing: randombytes(302) -> BE16 mod q (ss.cpp:4-11), in the reference's
// call order.  One thread per (proof, sharing); also writes the constant secrets of the
// eta sharings (mlwe_prover.cpp:42-48) and zero-fills the row padding.
template <int K>
__global__ void __launch_bounds__(64) k_tails(ProveBufs pb)
{
    const Slots sl = make_slots(K);
    const int nfresh = sl.n1 + K;                   // all challenge-independent sharings + [A s]
    // tails are staged in shared memory and written out by the whole CTA (see k_expand_f); srow[t][0..159] = Y row elements 256..415
    __shared__ __align__(16) u16 srow[64][YLD - 256];
    __shared__ int sslot[64];                       // slot of the row, -1 = not produced by this launch
    const int gid0 = blockIdx.x * 64, gid = gid0 + threadIdx.x, total = pb.B * nfresh;
    sslot[threadIdx.x] = -1;
    if (gid < total) {
        const int b = gid / nfresh, idx = gid % nfresh;
        // call numbers relative to the first call of prepare_randomness (F seeds, then f_i / NTT_f_i alternating,
        // mlwe_prover.cpp:8-38), prepare_range_proof (s / e alternating, :41-59) and prove (s_i / e_i alternating :89-101,
        // [A s]_i :292-323, z chains :338-392)
        int slot, call, group;
        if (idx < sl.n1) {
            slot = idx;
            if (slot < sl.Tf0) { call = pb.cb_rand + sl.F + 2 * (slot - sl.f0); group = 1; }
            else if (slot < sl.seta0) { call = pb.cb_rand + sl.F + 2 * (slot - sl.Tf0) + 1; group = 1; }
            else if (slot < sl.eeta0) { call = pb.cb_eta + 2 * (slot - sl.seta0); group = 2; }
            else if (slot < sl.s0) { call = pb.cb_eta + 2 * (slot - sl.eeta0) + 1; group = 2; }
            else if (slot < sl.e0) { call = pb.cb_prove + 2 * (slot - sl.s0); group = 4; }
            else if (slot < sl.zs0) { call = pb.cb_prove + 2 * (slot - sl.e0) + 1; group = 4; }
            else if (slot < sl.ze0) { call = pb.cb_prove + 3 * K + 2 * (slot - sl.zs0); group = 4; }
            else { call = pb.cb_prove + 3 * K + 2 * (slot - sl.ze0) + 1; group = 4; }
        } else { slot = sl.As0 + (idx - sl.n1); call = pb.cb_prove + 2 * K + (idx - sl.n1); group = 4; }
        if (pb.tails_mask & group) {
            sslot[threadIdx.x] = slot;
            uint64_t sd[4], a[25];
            const uint64_t *gs = reinterpret_cast<const uint64_t *>(pb.seeds + 32 * (size_t)b);
#pragma unroll
            for (int w = 0; w < 4; w++) sd[w] = gs[w];
            drbg_begin(a, sd, call);
            u16 *dst = srow[threadIdx.x];
#pragma unroll 1
            for (int blk = 0; blk < 3; blk++) {     // 302 bytes = 136 + 136 + 30
#pragma unroll
                for (int v = 0; v < 68; v++) {
                    if (blk * 68 + v <= NT) dst[blk * 68 + v] = (u16)(lane_be16(a[v >> 2], v & 3) % (uint32_t)Q);
                }
                if (blk < 2) keccak_f1600(a);
            }
            for (int c = NT + 1; c < YLD - 256; c++) dst[c] = 0;      // row padding (terms 407..415) must be zero
        }
    }
    __syncthreads();
    const int nrow = min(64, total - gid0);
    constexpr int CH = (YLD - 256) / 8;             // 20 x 16-byte chunks of tail + padding per row
    for (int idx = threadIdx.x; idx < nrow * CH; idx += 64) {
        const int r = idx / CH, c = idx % CH, slot = sslot[r];
        if (slot < 0) continue;
        u16 *dst = yrow(pb, sl, (gid0 + r) / nfresh, slot) + 256;
        reinterpret_cast<uint4 *>(dst)[c] = reinterpret_cast<const uint4 *>(srow[r])[c];
    }
    // constant secrets of the eta sharings (mlwe_prover.cpp:42-48): same constant for the s and e copies
    for (int idx = threadIdx.x; idx < nrow * 32; idx += 64) {
        const int r = idx >> 5, c = idx & 31, slot = sslot[r];
        if (slot < sl.seta0 || slot >= sl.s0) continue;
        const uint32_t ev = (uint32_t)(((slot - sl.seta0) % sl.E - sl.eta + Q) % Q), w = ev | (ev << 16);
        reinterpret_cast<uint4 *>(yrow(pb, sl, (gid0 + r) / nfresh, slot))[c] = make_uint4(w, w, w, w);
    }
}

// ---------------------------------------------------------------------------------------------
// SHA3-256 of a per-party record: commit hash (mlwe_prover.cpp:116-127, mlwe_verifier.cpp:23-35) and view
// hash (mlwe_prover.cpp:395-444, mlwe_verifier.cpp:584-632).  One thread per (proof, party).
//   prover  : record element v = plane[tab[v]][party]  -> every absorb load is coalesced across the CTA's parties
//   verifier: record element v = rec[opened index][v]   (row-major records of the 150 opened parties)
// NVALS u16 values = 2*NVALS bytes.
struct HashSrc {
    const u16 *src; long long proof_stride, item_stride, elem_stride; int off0;
    const int16_t *tab;      // nullptr = identity
    const u16 *plist; int nlist;   // nullptr: items are parties 0..NP-1; else item idx -> output party plist[b][idx]
};
// SHA3-256 of one record; element v is the u16 at cta_base + toff + soff[v].  Digest in a[0..3].
template <int NVALS>
__device__ __forceinline__ void hash_record(uint64_t (&a)[25], const char *cta_base, uint32_t toff, const uint32_t *soff)
{
    auto ld = [&](int v) -> uint64_t { return *reinterpret_cast<const u16 *>(cta_base + (toff + soff[v])); };
    keccak_zero(a);
    constexpr int NFULL = (2 * NVALS) / 136, REMV = NVALS - NFULL * 68;   // u16 values left for the last block
#pragma unroll 1
    for (int blk = 0; blk < NFULL; blk++) {
#pragma unroll
        for (int l = 0; l < 17; l++) {
            const uint32_t w0 = (uint32_t)ld(blk * 68 + 4 * l) | ((uint32_t)ld(blk * 68 + 4 * l + 1) << 16);
            const uint32_t w1 = (uint32_t)ld(blk * 68 + 4 * l + 2) | ((uint32_t)ld(blk * 68 + 4 * l + 3) << 16);
            a[l] ^= ((uint64_t)w1 << 32) | w0;
        }
        keccak_f1600(a);
    }
#pragma unroll
    for (int l = 0; l < 17; l++) {
        uint64_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (4 * l + i < REMV) w |= ld(NFULL * 68 + 4 * l + i) << (16 * i);
        if (l == REMV / 4) w |= 0x06ULL << (16 * (REMV % 4));
        if (l == 16) w |= 0x8000000000000000ULL;
        a[l] ^= w;
    }
    keccak_f1600(a);
}

template <int NVALS>
__global__ void __launch_bounds__(128)
k_hash_records(const HashSrc hs, u8 *__restrict__ out_rows, u16 *__restrict__ out_planes, int nslot, int out_plane_slot)
{
    __shared__ uint32_t soff[NVALS];      // byte offset of record element v from the CTA's (uniform) base
    for (int i = threadIdx.x; i < NVALS; i += 128) soff[i] = (uint32_t)((hs.tab ? hs.tab[i] : i) * hs.elem_stride * 2);
    __syncthreads();
    const int b = blockIdx.y, idx = blockIdx.x * 128 + threadIdx.x;
    int p = idx;
    if (hs.plist) { if (idx >= hs.nlist) return; p = hs.plist[(size_t)b * hs.nlist + idx]; if (p >= NP) return; }
    else if (idx >= NP) return;
    const char *cta_base = reinterpret_cast<const char *>(hs.src + (size_t)b * hs.proof_stride + hs.off0);
    uint64_t a[25];
    hash_record<NVALS>(a, cta_base, (uint32_t)(idx * hs.item_stride * 2), soff);
    if (out_rows) {
        uint64_t *o = reinterpret_cast<uint64_t *>(out_rows + ((size_t)b * NP + p) * 32);
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
    }
    if (out_planes) {
        u16 *o = out_planes + ((size_t)b * nslot + out_plane_slot) * SLD + SOFF + p;
#pragma unroll
        for (int i = 0; i < 16; i++) o[(size_t)i * SLD] = (u16)(a[i >> 2] >> (16 * (i & 3)));
    }
}

// Sponge kernels: one warp per proof.  With four proofs per CTA a 1024-proof launch is 256 CTAs on 148 SMs: 108 SMs host eight sponge warps and
// 40 host four, and the launch lasts as long as the crowded ones (shuffle-throughput contention).  One proof per CTA spreads them 7 / 6.
#ifndef KOSK_FS_WPC
#define KOSK_FS_WPC 1
#endif
static inline void fs_launch_dims(int B, int &ctas, int &threads) { const int wpc = B >= 256 ? KOSK_FS_WPC : 4; ctas = (B + wpc - 1) / wpc; threads = 32 * wpc; }

// FS-1: alpha = BE16(SHAKE256(SHA3-256(Tcomm_0 || ... ) || 0x01)) mod q and the power table (mlwe_prover.cpp:130-153).
// One warp per proof.
// __launch_bounds__(128, 4): without a min-blocks hint ptxas squeezes these kernels into 32 registers (full occupancy), interleaves the
// shuffles of an exchange stage with their consumers and recycles destination registers, so every shuffle waits for the previous one's
// consumer: k_fs2 then ran 27 % slower than k_fs1 on the same 343-permutation sponge (1.35 vs 1.06 ms per 1024 proofs).
template <int K>
__global__ void __launch_bounds__(128, 4) k_fs1(const u8 *__restrict__ TCR, u16 *__restrict__ PW, int B)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K;
    __shared__ u16 salpha[4][80];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + w;      // 1 .. 4 proofs per CTA (fs_launch_dims)
    if (b >= B) return;
    WarpKeccak wk; wk.init();
    uint64_t a = wk.tree_hash(TCR + (size_t)b * TREE_BYTES);
    a = wk.prf1(a);
#pragma unroll 1
    for (int blk = 0; blk < 2; blk++) {                 // 2*NA <= 156 bytes = 136 + 20
        if (lane < 17)
#pragma unroll
            for (int i = 0; i < 4; i++) { const int j = blk * 68 + 4 * lane + i; if (j < NA) salpha[w][j] = (u16)(lane_be16(a, i) % (uint32_t)Q); }
        if (blk == 0) a = wk.permute(a);
    }
    __syncwarp();
    u16 *pw = PW + (size_t)b * NA * F;
    for (int j = lane; j < NA; j += 32) {
        const uint32_t al = salpha[w][j]; uint32_t xx = 1;
        for (int kk = 0; kk < F; kk++) { pw[j * F + kk] = (u16)xx; xx = gf_mul(xx, al); }
    }
}

// FS-2: opened set I from the view hashes, with the reference's linear-probe de-duplication
// (mlwe_prover.cpp:445-474: the first free index at or after the drawn one, in draw order) and the ascending
// rest list (:480-490).  One warp per proof.
// open set from the squeezed indices: linear-probe de-duplication in draw order, then the ascending rest list
__device__ __noinline__ void fs2_open_set(const u16 *sraw, uint32_t *sused, u16 *I, u16 *rest)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < (NP + 31) / 32; i += 32) sused[i] = 0;
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < NT; i++) {
            uint32_t c = sraw[i];
            while (sused[c >> 5] & (1u << (c & 31))) c = (c + 1 == NP) ? 0 : c + 1;
            sused[c >> 5] |= 1u << (c & 31);
            I[i] = (u16)c;
        }
    }
    __syncwarp();
    int n = 0;
    for (int p0 = 0; p0 < NP; p0 += 32) {
        const int p = p0 + lane;
        const bool free_ = p < NP && !(sused[p >> 5] & (1u << (p & 31)));
        const uint32_t m = __ballot_sync(0xffffffffu, free_);
        if (free_) { const int pos = n + __popc(m & ((1u << lane) - 1)); if (pos < NR) rest[pos] = (u16)p; }
        n += __popc(m);
    }
}
__global__ void __launch_bounds__(128, 4) k_fs2(const u8 *__restrict__ VWR, u16 *__restrict__ Iout, u16 *__restrict__ REST, int B)
{
    __shared__ u16 sraw[4][NT + 2];
    __shared__ uint32_t sused[4][(NP + 31) / 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + w;      // 1 .. 4 proofs per CTA (fs_launch_dims)
    if (b >= B) return;
    WarpKeccak wk; wk.init();
    uint64_t a = wk.tree_hash(VWR + (size_t)b * TREE_BYTES);
    a = wk.prf1(a);
#pragma unroll 1
    for (int blk = 0; blk < 3; blk++) {                 // 300 bytes = 136 + 136 + 28
        if (lane < 17)
#pragma unroll
            for (int i = 0; i < 4; i++) { const int j = blk * 68 + 4 * lane + i; if (j < NT) sraw[w][j] = (u16)(lane_be16(a, i) % (uint32_t)NP); }
        if (blk < 2) a = wk.permute(a);
    }
    fs2_open_set(sraw[w], sused[w], Iout + (size_t)b * NT, REST + (size_t)b * NR);
}

// ---------------------------------------------------------------------------------------------
// Fused per-proof record hashing + Fiat-Shamir sponge (prover): one CTA per proof.  Warps 1..NHW hash the 1454 party
// records in party order (thread per party, register-resident Keccak, ALU-pipe bound) and publish how far they are in a
// shared-memory counter; warp 0 runs the strictly sequential 343-permutation tree hash (warp-cooperative Keccak, latency
// bound) over the digests as they become available, then expands the challenge.  The sponge's latency, 1.06 / 1.35 ms
// as a kernel of its own, hides behind the hashing of the same proofs.  Hashers never wait, so the CTA cannot deadlock;
// the sponge warp's spin is bounded and reports through *status.
//   MODE 1: commitments -> Tcomm rows + planes, FS-1 -> power table PW      (mlwe_prover.cpp:116-153)
//   MODE 2: view hashes -> VWR rows,            FS-2 -> I, rest list        (mlwe_prover.cpp:395-490)
template <int K, int NVALS, int MODE, int NHW>
__global__ void __launch_bounds__(32 * (NHW + 1))
k_hash_fs(const HashSrc hs, u8 *__restrict__ rows, u16 *__restrict__ out_planes, int nslot, int out_plane_slot,
          u16 *__restrict__ PW, u16 *__restrict__ Iout, u16 *__restrict__ REST, int *__restrict__ status)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K;
    constexpr int PASS = 32 * NHW, NPASS = (NP + PASS - 1) / PASS;
    __shared__ uint32_t soff[NVALS];
    __shared__ int s_done[NHW];                     // passes completed by each hasher warp
    __shared__ u16 sval[NT + 2];                    // alpha (MODE 1) or raw opened indices (MODE 2)
    __shared__ uint32_t sused[(NP + 31) / 32];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NVALS; i += blockDim.x) soff[i] = (uint32_t)((hs.tab ? hs.tab[i] : i) * hs.elem_stride * 2);
    if (threadIdx.x < NHW) s_done[threadIdx.x] = 0;
    __syncthreads();
    u8 *myrows = rows + (size_t)b * TREE_BYTES;
    if (warp > 0) {
        // ---- hashers ----
        const int h = warp - 1;
        const char *cta_base = reinterpret_cast<const char *>(hs.src + (size_t)b * hs.proof_stride + hs.off0);
        for (int j = 0; j < NPASS; j++) {
            const int p = j * PASS + 32 * h + lane;
            if (p < NP) {
                uint64_t a[25];
                hash_record<NVALS>(a, cta_base, (uint32_t)(p * hs.item_stride * 2), soff);
                uint64_t *o = reinterpret_cast<uint64_t *>(myrows + (size_t)p * 32);
                o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
                if (out_planes) {
                    u16 *q = out_planes + ((size_t)b * nslot + out_plane_slot) * SLD + SOFF + p;
#pragma unroll
                    for (int i = 0; i < 16; i++) q[(size_t)i * SLD] = (u16)(a[i >> 2] >> (16 * (i & 3)));
                }
            }
            __threadfence_block();                  // digests visible to the sponge warp before the counter moves
            __syncwarp();
            if (lane == 0) reinterpret_cast<volatile int *>(s_done)[h] = j + 1;
        }
        return;
    }
    // ---- sponge warp ----
    WarpKeccak wk; wk.init();
    const uint64_t *src = reinterpret_cast<const uint64_t *>(myrows);
    constexpr int NFULL = TREE_BYTES / 136;         // 342 full rate blocks + 16 bytes
    auto wait_party = [&](int pmax) {               // all digests of parties <= pmax written?
        const int j = pmax / PASS, hh = (pmax % PASS) / 32;
        int spins = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int w2 = 0; w2 < NHW; w2++) ok &= reinterpret_cast<volatile int *>(s_done)[w2] >= (w2 <= hh ? j + 1 : j);
            if (ok) break;
            __nanosleep(200);
            if (++spins > (1 << 24)) { if (lane == 0) atomicExch(status, 1); break; }     // ~ seconds: report instead of hanging
        }
        __threadfence_block();
    };
    uint64_t a = 0;
#pragma unroll 1
    for (int blk = 0; blk < NFULL; blk++) {
        wait_party(min(NP - 1, (136 * blk + 135) / 32));
        if (lane < 17) a ^= __ldcg(src + blk * 17 + lane);
        a = wk.permute(a);
    }
    wait_party(NP - 1);
    if (lane < 2) a ^= __ldcg(src + NFULL * 17 + lane);
    if (lane == 2) a ^= 0x06ULL;
    if (lane == 16) a ^= 0x8000000000000000ULL;
    a = wk.permute(a);
    a = wk.prf1(a);
    if (MODE == 1) {
#pragma unroll 1
        for (int blk = 0; blk < 2; blk++) {         // 2*NA <= 156 bytes = 136 + 20
            if (lane < 17)
#pragma unroll
                for (int i = 0; i < 4; i++) { const int j = blk * 68 + 4 * lane + i; if (j < NA) sval[j] = (u16)(lane_be16(a, i) % (uint32_t)Q); }
            if (blk == 0) a = wk.permute(a);
        }
        __syncwarp();
        u16 *pw = PW + (size_t)b * NA * F;
        for (int j = lane; j < NA; j += 32) {
            const uint32_t al = sval[j]; uint32_t xx = 1;
            for (int kk = 0; kk < F; kk++) { pw[j * F + kk] = (u16)xx; xx = gf_mul(xx, al); }
        }
    } else {
#pragma unroll 1
        for (int blk = 0; blk < 3; blk++) {         // 300 bytes = 136 + 136 + 28
            if (lane < 17)
#pragma unroll
                for (int i = 0; i < 4; i++) { const int j = blk * 68 + 4 * lane + i; if (j < NT) sval[j] = (u16)(lane_be16(a, i) % (uint32_t)NP); }
            if (blk < 2) a = wk.permute(a);
        }
        for (int i = lane; i < (NP + 31) / 32; i += 32) sused[i] = 0;
        __syncwarp();
        u16 *I = Iout + (size_t)b * NT;
        if (lane == 0) {
            for (int i = 0; i < NT; i++) {          // linear-probe de-duplication (mlwe_prover.cpp:459-474)
                uint32_t c = sval[i];
                while (sused[c >> 5] & (1u << (c & 31))) c = (c + 1 == NP) ? 0 : c + 1;
                sused[c >> 5] |= 1u << (c & 31);
                I[i] = (u16)c;
            }
        }
        __syncwarp();
        u16 *rest = REST + (size_t)b * NR;
        int n = 0;
        for (int p0 = 0; p0 < NP; p0 += 32) {
            const int p = p0 + lane;
            const bool free_ = p < NP && !(sused[p >> 5] & (1u << (p & 31)));
            const uint32_t m = __ballot_sync(0xffffffffu, free_);
            if (free_) { const int pos = n + __popc(m & ((1u << lane) - 1)); if (pos < NR) rest[pos] = (u16)p; }
            n += __popc(m);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// beta/gamma and r/NTT_r evaluation (mlwe_prover.cpp:159-214): per party p
//   beta[p][j]  = f[p][0]  + sum_{k>=1} alpha_j^k f[p][k]        j < 70
//   r[p][j']    = f[p][71] + sum_{k>=1} alpha_{70+j'}^k f[p][k]  j' < 2K      (c0 index quirk)
// and the same over NTT_f.  CTA = 128 parties x {f, NTT_f}; the power table sits in shared memory.
// Thread = (party pair, {f | NTT_f}, half of the challenge columns): 2 parties x NJ columns of accumulators, so one
// broadcast LDS.128 of four power-table entries feeds eight IMADs.
template <int K, int J0, int NJ>
__device__ __forceinline__ void eval_body(const ProveBufs &pb, const Slots &sl, const int32_t (*spw)[((MK + 2 * K + 3) & ~3)], int b, int half, int p0)
{
    constexpr int F = MK + 2 * K + 1;
    const bool two = p0 + 1 < NP;
    const u16 *src = plane(pb, sl, b, half ? sl.Tf0 : sl.f0) + p0;
    int32_t acc0[NJ], acc1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) { acc0[j] = 0; acc1[j] = 0; }
    constexpr int U = 4;                     // software-pipelined share loads
    int32_t v0[U], v1[U];
#pragma unroll
    for (int u = 0; u < U; u++) { v0[u] = src[(size_t)u * SLD]; v1[u] = two ? src[(size_t)u * SLD + 1] : 0; }
    int32_t c71a = 0, c71b = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < F; k0 += U) {
        int32_t a0[U], a1[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            a0[u] = gf_center(v0[u]); a1[u] = gf_center(v1[u]);
            if (k0 + U + u < F) { v0[u] = src[(size_t)(k0 + U + u) * SLD]; v1[u] = two ? src[(size_t)(k0 + U + u) * SLD + 1] : 0; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + u < F) {
                if (k0 + u == MK + 1) { c71a = a0[u]; c71b = a1[u]; }
#pragma unroll
                for (int j4 = 0; j4 < NJ / 4; j4++) {
                    const int4 w = *reinterpret_cast<const int4 *>(&spw[k0 + u][J0 + 4 * j4]);
                    acc0[4 * j4] += a0[u] * w.x; acc0[4 * j4 + 1] += a0[u] * w.y; acc0[4 * j4 + 2] += a0[u] * w.z; acc0[4 * j4 + 3] += a0[u] * w.w;
                    acc1[4 * j4] += a1[u] * w.x; acc1[4 * j4 + 1] += a1[u] * w.y; acc1[4 * j4 + 2] += a1[u] * w.z; acc1[4 * j4 + 3] += a1[u] * w.w;
                }
            }
        }
    }
#pragma unroll
    for (int pp = 0; pp < 2; pp++) {
        if (pp == 1 && !two) break;
        const int p = p0 + pp;
        const int32_t *acc = pp ? acc1 : acc0;
        const int32_t c71 = pp ? c71b : c71a;
        // beta | gamma row of this party: 16-byte stores (BGH = 144 u16 per half-row, 16B-aligned)
        uint4 *bg = reinterpret_cast<uint4 *>(pb.BG + ((size_t)b * NP + p) * (2 * BGH) + half * BGH + J0);
#pragma unroll
        for (int q = 0; q < (NJ + 7) / 8; q++) {
            if (J0 + 8 * q >= BGH) break;
            uint32_t wv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j0 = 8 * q + 2 * i, j1 = j0 + 1;
                const uint32_t lo = (j0 < NJ && J0 + j0 < MK) ? gf_canon(acc[j0 < NJ ? j0 : 0]) : 0;
                const uint32_t hi = (j1 < NJ && J0 + j1 < MK) ? gf_canon(acc[j1 < NJ ? j1 : 0]) : 0;
                wv[i] = lo | (hi << 16);
            }
            bg[q] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int jj = J0 + j;
            if (jj < K) plane(pb, sl, b, (half ? sl.G0 : sl.B0) + jj)[p] = (u16)gf_canon(acc[j]);
            if (jj >= MK && jj < MK + 2 * K) plane(pb, sl, b, (half ? sl.TR0 : sl.R0) + (jj - MK))[p] = (u16)gf_canon(acc[j] + c71);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256, 2) k_eval(ProveBufs pb)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K, NAP = (NA + 3) & ~3, JSPLIT = 40;
    const Slots sl = make_slots(K);
    __shared__ __align__(16) int32_t spw[F][NAP];
    const int b = blockIdx.y, tid = threadIdx.x;
    // power table of this proof: [NA][F] u16, read as coalesced 32-bit words and transposed/centered into shared memory
    const uint32_t *pw2 = reinterpret_cast<const uint32_t *>(pb.PW + (size_t)b * NA * F);       // NA*F is even for K = 2, 3, 4
    for (int i = tid; i < F * NAP; i += 256) (&spw[0][0])[i] = 0;
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < NA * F / 2; i += 256) {
        const uint32_t v = pw2[i];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = 2 * i + h, j = e / F, kk = e % F;
            if (!(kk == 0 && j >= MK)) spw[kk][j] = gf_center(h ? v >> 16 : v & 0xFFFF);
        }
    }
    __syncthreads();
    const int w = tid >> 5, lane = tid & 31;
    const int half = (w >> 2) & 1, jh = (w >> 1) & 1;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < (NP + 127) / 128; tile += gridDim.x) {      // several party tiles per CTA amortise the table load
        const int p0 = tile * 128 + 2 * ((w & 1) * 32 + lane);
        if (p0 >= NP) continue;
        if (jh == 0) eval_body<K, 0, JSPLIT>(pb, sl, spw, b, half, p0);
        else eval_body<K, JSPLIT, NAP - JSPLIT>(pb, sl, spw, b, half, p0);
    }
}

// ---------------------------------------------------------------------------------------------
// Online opening, in the clear (mlwe_prover.cpp:222-318).  By linearity of the sharing the opened
// masked secrets are s + r with r = f_71 + sum_k alpha^k f_k over the clear f vectors, so no
// reconstruction mat-vec is needed.  Emits the Y rows of [NTT(s+r)], [NTT(e+r)], [A(s+r)] (all
// re-using the tail of [s+r]/[e+r]) and the secrets of [A s].  One CTA per proof.
template <int K>
__global__ void __launch_bounds__(128) k_open(ProveBufs pb)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K;
    const Slots sl = make_slots(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 sSR[K][256], sER[K][256];
    __shared__ u16 spw[2 * K][F];
    for (int i = tid; i < 2 * K * F; i += 128) spw[i / F][i % F] = pb.PW[(size_t)b * NA * F + (MK + i / F) * F + (i % F)];
    __syncthreads();
    for (int c = tid; c < 256; c += 128) {
        uint32_t acc[2 * K];
        const uint32_t f71 = yrow(pb, sl, b, sl.f0 + MK + 1)[c];
#pragma unroll
        for (int j = 0; j < 2 * K; j++) acc[j] = f71;
        for (int kk = 1; kk < F; kk++) {
            const uint32_t v = yrow(pb, sl, b, sl.f0 + kk)[c];
#pragma unroll
            for (int j = 0; j < 2 * K; j++) acc[j] = gf_add(acc[j], gf_mul(spw[j][kk], v));
        }
#pragma unroll
        for (int i = 0; i < K; i++) {
            sSR[i][c] = (u16)gf_add(yrow(pb, sl, b, sl.s0 + i)[c], acc[i]);
            sER[i][c] = (u16)gf_add(yrow(pb, sl, b, sl.e0 + i)[c], acc[K + i]);
        }
    }
    __syncthreads();
    for (int i = 0; i < K; i++) { ntt256_block(sSR[i], tid); ntt256_block(sER[i], tid); }
    const u16 *AH = pb.AH + (size_t)b * K * K * 256, *SHAT = pb.SHAT + (size_t)b * K * 256;
    for (int i = 0; i < K; i++) {
        uint32_t s0 = 0, s1 = 0, q0 = 0, q1 = 0;
        for (int j = 0; j < K; j++) {
            const uint32_t a0 = AH[(i * K + j) * 256 + 2 * tid], a1 = AH[(i * K + j) * 256 + 2 * tid + 1];
            uint32_t r0, r1;
            basemul_pair(r0, r1, a0, a1, SHAT[j * 256 + 2 * tid], SHAT[j * 256 + 2 * tid + 1], tid);
            s0 = gf_add(s0, r0); s1 = gf_add(s1, r1);
            basemul_pair(r0, r1, a0, a1, sSR[j][2 * tid], sSR[j][2 * tid + 1], tid);
            q0 = gf_add(q0, r0); q1 = gf_add(q1, r1);
        }
        u16 *yAs = yrow(pb, sl, b, sl.As0 + i), *yAsr = yrow(pb, sl, b, sl.Asr0 + i);
        yAs[2 * tid] = (u16)s0; yAs[2 * tid + 1] = (u16)s1;
        yAsr[2 * tid] = (u16)q0; yAsr[2 * tid + 1] = (u16)q1;
        u16 *yTsr = yrow(pb, sl, b, sl.Tsr0 + i), *yTer = yrow(pb, sl, b, sl.Ter0 + i);
        yTsr[tid] = sSR[i][tid]; yTsr[tid + 128] = sSR[i][tid + 128];
        yTer[tid] = sER[i][tid]; yTer[tid + 128] = sER[i][tid + 128];
        // tails: shares of [s+r] / [e+r] held by parties 0..150
        for (int q = tid; q <= NT; q += 128) {
            const u16 srq = (u16)gf_add(plane(pb, sl, b, sl.s0 + i)[q], plane(pb, sl, b, sl.R0 + i)[q]);
            const u16 erq = (u16)gf_add(plane(pb, sl, b, sl.e0 + i)[q], plane(pb, sl, b, sl.R0 + K + i)[q]);
            yTsr[256 + q] = srq; yAsr[256 + q] = srq; yTer[256 + q] = erq;
        }
        for (int c = D1 + tid; c < YLD; c += 128) { yTsr[c] = 0; yTer[c] = 0; yAsr[c] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------
// Per-party derived planes hashed into the view: [s+r], [e+r] (mlwe_prover.cpp:227-245) and the
// range-proof zero sharings u_j = z_j^(2d) - z_j^(d) (:338-381).  One thread per (proof, party).
template <int K>
__global__ void __launch_bounds__(128) k_derive(ProveBufs pb)
{
    constexpr int ETA = (K == 2) ? 3 : 2, E = 2 * ETA + 1, M = 2 * ETA;
    const Slots sl = make_slots(K);
    const int b = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
    if (p >= NP) return;
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const uint32_t x = plane(pb, sl, b, (w ? sl.e0 : sl.s0) + i)[p];
            plane(pb, sl, b, (w ? sl.ER0 : sl.SR0) + i)[p] = (u16)gf_add(x, plane(pb, sl, b, sl.R0 + w * K + i)[p]);
            uint32_t sub[E];
#pragma unroll
            for (int m = 0; m < E; m++) sub[m] = gf_sub(x, plane(pb, sl, b, (w ? sl.eeta0 : sl.seta0) + i * E + m)[p]);
            uint32_t prev = sub[0];
#pragma unroll
            for (int j = 0; j < M; j++) {
                const uint32_t z2 = gf_mul(prev, sub[j + 1]);
                const uint32_t zd = plane(pb, sl, b, (w ? sl.ze0 : sl.zs0) + i * M + j)[p];
                plane(pb, sl, b, (w ? sl.UE0 : sl.US0) + i * M + j)[p] = (u16)gf_sub(z2, zd);
                prev = zd;
            }
        }
}

// ---------------------------------------------------------------------------------------------
// Proof assembly (mlwe_prover.cpp:480-537): gather the opened-set and rest-set fields into the packed byte layout of struct
// mpcith_proof.  grid = (ASM_OPENED_CTAS + rest tiles, B), 256 threads.
//   Opened set (150 parties in the order of I, ~220 planes each): an element-wise gather costs one 32-byte sector per 2 useful bytes and
//   is bound by DRAM latency (round 1: 1.53 GB read for 0.68 GB of proof, DRAM at 43 %).  Instead a CTA streams whole plane rows (2.9 KB,
//   coalesced 16-byte loads), picks the 150 opened parties of each row through shared memory into a compact [plane][opened] tile and
//   emits its fields from that tile with coalesced stores.  Three CTAs per proof: the f planes, the NTT_f planes, everything else.
//   Rest set (1304 parties in ascending order): tiles of 64 consecutive rest parties; their plane reads are near-contiguous.
#ifndef KOSK_ASM_RB
#define KOSK_ASM_RB 4
#endif
#ifndef KOSK_ASM_MINB
#define KOSK_ASM_MINB 5     // five CTAs per SM (48 registers): 0.40 -> 0.38 ms; six (40 registers) lose the register prefetch (0.55 ms)
#endif
constexpr int ASM_OPENED_CTAS = 3, ASM_RB = KOSK_ASM_RB, ASM_OS = 154, ASM_REST_ROWS = 64;
template <int K> struct AsmPlanes {              // local plane numbering of the "everything else" CTA
    static constexpr int ETA = (K == 2) ? 3 : 2, E = 2 * ETA + 1, M = 2 * ETA;
    static constexpr int S = 0, Ee = K, TSR = 2 * K, TER = 3 * K, ASR = 4 * K, AS = 5 * K, TR = 6 * K, SETA = 8 * K, EETA = SETA + K * E, ZS = EETA + K * E, ZE = ZS + K * M,
                         N = ZE + K * M;
    __device__ static int slot(const Slots &sl, int l)
    {
        if (l < Ee) return sl.s0 + l;
        if (l < TSR) return sl.e0 + (l - Ee);
        if (l < TER) return sl.Tsr0 + (l - TSR);
        if (l < ASR) return sl.Ter0 + (l - TER);
        if (l < AS) return sl.Asr0 + (l - ASR);
        if (l < TR) return sl.As0 + (l - AS);
        if (l < SETA) return sl.TR0 + (l - TR);
        if (l < EETA) return sl.seta0 + (l - SETA);
        if (l < ZS) return sl.eeta0 + (l - EETA);
        if (l < ZE) return sl.zs0 + (l - ZS);
        return sl.ze0 + (l - ZE);
    }
};

template <int K>
__global__ void __launch_bounds__(256, KOSK_ASM_MINB) k_assemble(ProveBufs pb)
{
    constexpr int ETA = (K == 2) ? 3 : 2, E = 2 * ETA + 1, M = 2 * ETA, F = MK + 2 * K + 1;
    using AP = AsmPlanes<K>;
    constexpr int NPLMAX = AP::N > F ? AP::N : F;
    const Slots sl = make_slots(K);
    const Layout L = make_layout(K);
    const int b = blockIdx.y, tid = threadIdx.x;
    u8 *pi = pb.pi + L.proof_bytes * (size_t)b;
    auto out16 = [&](size_t off, size_t idx) -> u16 * { return reinterpret_cast<u16 *>(pi + off) + idx; };
    __shared__ u16 sI[NT + 2];
    __shared__ __align__(16) u16 srow[ASM_RB][SLD];
    __shared__ u16 sO[NPLMAX][ASM_OS];
    if ((int)blockIdx.x < ASM_OPENED_CTAS) {
        const int grp = blockIdx.x, npl = grp < 2 ? F : AP::N;
        const u16 *I = pb.I + (size_t)b * NT;
        for (int i = tid; i < NT; i += 256) sI[i] = I[i];
        // ---- compact: O[l][r] = plane(l)[I[r]] ----
        const u16 *planes = pb.SH + (size_t)b * sl.nslot * SLD;
        // the rows of group g + 1 are loaded into registers while group g is picked: with load -> barrier -> pick -> barrier in sequence every
        // group exposed a DRAM round trip (ncu: 49 % of the stall samples on long_scoreboard, 10 % on the barriers)
        constexpr int NLD = (ASM_RB * (SLD / 8) + 255) / 256;
        uint4 pre[NLD];
        auto issue = [&](int base) {
            const int nrow = min(ASM_RB, npl - base);
#pragma unroll
            for (int u = 0; u < NLD; u++) {
                const int idx = tid + 256 * u;
                if (idx < nrow * (SLD / 8)) {
                    const int r = idx / (SLD / 8), c = idx % (SLD / 8), l = base + r;
                    const int slot = grp == 0 ? sl.f0 + l : grp == 1 ? sl.Tf0 + l : AP::slot(sl, l);
                    pre[u] = __ldcs(reinterpret_cast<const uint4 *>(planes + (size_t)slot * SLD) + c);
                }
            }
        };
        issue(0);
        for (int base = 0; base < npl; base += ASM_RB) {
            const int nrow = min(ASM_RB, npl - base);
            __syncthreads();                                  // previous rows consumed (and sI visible)
#pragma unroll
            for (int u = 0; u < NLD; u++) {
                const int idx = tid + 256 * u;
                if (idx < nrow * (SLD / 8)) reinterpret_cast<uint4 *>(srow[idx / (SLD / 8)])[idx % (SLD / 8)] = pre[u];
            }
            __syncthreads();
            if (base + ASM_RB < npl) issue(base + ASM_RB);
            for (int idx = tid; idx < nrow * NT; idx += 256) {
                const int r = idx / NT, i = idx % NT;
                sO[base + r][i] = srow[r][SOFF + sI[i]];
            }
        }
        __syncthreads();
        // ---- emit ----
        if (grp < 2) {                                        // f_shares / NTT_f_shares [T][F]: two elements per 32-bit store
            uint32_t *dst = reinterpret_cast<uint32_t *>(pi + (grp == 0 ? L.o_f : L.o_Tf));
            for (int w = tid; w < NT * F / 2; w += 256) {
                const int e0 = 2 * w, r0 = e0 / F, j0 = e0 % F, e1 = e0 + 1, r1 = e1 / F, j1 = e1 % F;
                dst[w] = (uint32_t)sO[j0][r0] | ((uint32_t)sO[j1][r1] << 16);
            }
        } else {
            for (int i = tid; i < NT; i += 256) *out16(L.o_I, i) = sI[i];
            for (int idx = tid; idx < NT * K; idx += 256) {
                const int r = idx / K, j = idx % K;
                const uint32_t As = sO[AP::AS + j][r];
                *out16(L.o_s, idx) = sO[AP::S + j][r];
                *out16(L.o_e, idx) = sO[AP::Ee + j][r];
                *out16(L.o_NTTs, idx) = (u16)gf_sub(sO[AP::TSR + j][r], sO[AP::TR + j][r]);
                *out16(L.o_NTTe, idx) = (u16)gf_sub(sO[AP::TER + j][r], sO[AP::TR + K + j][r]);
                *out16(L.o_NTTAr, idx) = (u16)gf_sub(sO[AP::ASR + j][r], As);
                *out16(L.o_NTTAs, idx) = (u16)As;
            }
            for (int idx = tid; idx < NT * K * E; idx += 256) {
                const int r = idx / (K * E), jm = idx % (K * E), j = jm / E;
                *out16(L.o_ssub, idx) = (u16)gf_sub(sO[AP::S + j][r], sO[AP::SETA + jm][r]);
                *out16(L.o_esub, idx) = (u16)gf_sub(sO[AP::Ee + j][r], sO[AP::EETA + jm][r]);
            }
            for (int idx = tid; idx < NT * K * M; idx += 256) {
                const int r = idx / (K * M), jm = idx % (K * M);
                *out16(L.o_zs, idx) = sO[AP::ZS + jm][r];
                *out16(L.o_ze, idx) = sO[AP::ZE + jm][r];
            }
        }
    } else {
        // The loops below are latency-bound unless several loads are in flight per thread: all sources are read through the non-coherent
        // path (nothing this kernel writes is read back), so the compiler may hoist the loads of an unrolled body above its stores.
        constexpr int ROWS = ASM_REST_ROWS;
        const u16 *rest = pb.REST + (size_t)b * NR;
        const u16 *__restrict__ planes = pb.SH + (size_t)b * sl.nslot * SLD + SOFF;
        auto P = [&](int slot, int p) -> uint32_t { return __ldg(planes + (size_t)slot * SLD + p); };
        const int r0 = (blockIdx.x - ASM_OPENED_CTAS) * ROWS, nr = min(ROWS, NR - r0);
        if (tid < nr) sI[tid] = rest[r0 + tid];
        __syncthreads();
        {   // beta / gamma [R][70]: rows of 70 u16 = 35 words, source rows 16-byte aligned, destination 4-byte aligned
            uint32_t *__restrict__ db = reinterpret_cast<uint32_t *>(pi + L.o_beta) + (size_t)r0 * (MK / 2);
            uint32_t *__restrict__ dg = reinterpret_cast<uint32_t *>(pi + L.o_gamma) + (size_t)r0 * (MK / 2);
            const uint32_t *__restrict__ bgb = reinterpret_cast<const uint32_t *>(pb.BG + (size_t)b * NP * (2 * BGH));
            constexpr int U = 5;
            for (int base = tid; base < nr * (MK / 2); base += 256 * U) {
                uint32_t vb[U], vg[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int idx = base + 256 * u;
                    if (idx < nr * (MK / 2)) { const int r = idx / (MK / 2), w = idx % (MK / 2); const uint32_t *bg = bgb + (size_t)sI[r] * BGH; vb[u] = __ldg(bg + w); vg[u] = __ldg(bg + BGH / 2 + w); }
                }
#pragma unroll
                for (int u = 0; u < U; u++) { const int idx = base + 256 * u; if (idx < nr * (MK / 2)) { db[idx] = vb[u]; dg[idx] = vg[u]; } }
            }
        }
        {   // 32-byte digests as 8 x u32 (proofs are only 4-byte aligned)
            uint32_t *__restrict__ dt = reinterpret_cast<uint32_t *>(pi + L.o_Tcomm) + (size_t)r0 * 8, *__restrict__ dc = reinterpret_cast<uint32_t *>(pi + L.o_comm) + (size_t)r0 * 8;
            const uint32_t *__restrict__ tcr = reinterpret_cast<const uint32_t *>(pb.TCR + (size_t)b * NP * 32), *__restrict__ vwr = reinterpret_cast<const uint32_t *>(pb.VWR + (size_t)b * NP * 32);
            uint32_t vt[2], vc[2];
#pragma unroll
            for (int u = 0; u < 2; u++) { const int idx = tid + 256 * u; if (idx < nr * 8) { const int p = sI[idx / 8]; vt[u] = __ldg(tcr + (size_t)p * 8 + idx % 8); vc[u] = __ldg(vwr + (size_t)p * 8 + idx % 8); } }
#pragma unroll
            for (int u = 0; u < 2; u++) { const int idx = tid + 256 * u; if (idx < nr * 8) { dt[idx] = vt[u]; dc[idx] = vc[u]; } }
        }
        {   // sr, er, t [R][K]; s_eta, e_eta [R][K][E]; u_s, u_e [R][K][M]: all loads first, then the stores
            constexpr int NKE = (ROWS * K * E + 255) / 256, NKM = (ROWS * K * M + 255) / 256;
            uint32_t vk[5], ve[NKE][2], vm[NKM][2];
            const bool hk = tid < nr * K;
            if (hk) { const int r = tid / K, j = tid % K, p = sI[r]; vk[0] = P(sl.SR0 + j, p); vk[1] = P(sl.ER0 + j, p); vk[2] = P(sl.As0 + j, p); vk[3] = P(sl.Ter0 + j, p); vk[4] = P(sl.TR0 + K + j, p); }
#pragma unroll
            for (int u = 0; u < NKE; u++) { const int idx = tid + 256 * u; if (idx < nr * K * E) { const int p = sI[idx / (K * E)], jm = idx % (K * E); ve[u][0] = P(sl.seta0 + jm, p); ve[u][1] = P(sl.eeta0 + jm, p); } }
#pragma unroll
            for (int u = 0; u < NKM; u++) { const int idx = tid + 256 * u; if (idx < nr * K * M) { const int p = sI[idx / (K * M)], jm = idx % (K * M); vm[u][0] = P(sl.US0 + jm, p); vm[u][1] = P(sl.UE0 + jm, p); } }
            if (hk) {
                const size_t o = (size_t)r0 * K + tid;
                *out16(L.o_sr, o) = (u16)vk[0]; *out16(L.o_er, o) = (u16)vk[1];
                *out16(L.o_t, o) = (u16)gf_add(vk[2], gf_sub(vk[3], vk[4]));
            }
#pragma unroll
            for (int u = 0; u < NKE; u++) { const int idx = tid + 256 * u; if (idx < nr * K * E) { const size_t o = (size_t)r0 * K * E + idx; *out16(L.o_seta, o) = (u16)ve[u][0]; *out16(L.o_eeta, o) = (u16)ve[u][1]; } }
#pragma unroll
            for (int u = 0; u < NKM; u++) { const int idx = tid + 256 * u; if (idx < nr * K * M) { const size_t o = (size_t)r0 * K * M + idx; *out16(L.o_us, o) = (u16)vm[u][0]; *out16(L.o_ue, o) = (u16)vm[u][1]; } }
        }
    }
}

}  // namespace kosk
