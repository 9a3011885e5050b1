// gf_gemm.cuh -- dense GF(3329) contraction on the SM integer pipe (IMAD), the hot loop of the KOSK path.
//   C[m][n] = sum_k A[m][k] * Bt[n][k]  mod 3329
// Replaces the reference's table mat-vecs: share_secrets_ddeg / recompute_share_secrets_ddeg (ss.cpp:23-32,
// :88-97; Bt = S, 1303 x 407), recon_secrets_ddeg (ss.cpp:44-50; Bt = R1), recon_secrets_2ddeg (ss.cpp:63-69;
// Bt = R2) and the verifier's NTL interpolate+eval (mlwe_verifier.cpp:188-224 etc.; Bt = per-proof Lagrange matrix).
// Operands are held as centered residues (|v| <= 1664): 813 * 1664^2 > 2^31, so the accumulator is reduced
// once every 32 k-steps (512 terms * 1664^2 + 1664 < 2^31); a 407-term row never needs the mid reduction.
// CTA tile (16*TM) rows x 128 columns x 16 terms, 256 threads, TM x 8 register tile per thread, split 4+4 so
// that every shared-memory read is a conflict-free (or broadcast) LDS.128; register-staged double buffering.
#pragma once
#include "kosk_common.cuh"

namespace kosk {

typedef uint16_t u16;
typedef uint8_t u8;

constexpr int GE_BN = 128, GE_BK = 16;
constexpr int GE_NPAD = ((NX + GE_BN - 1) / GE_BN) * GE_BN;   // 1408 rows of the share table incl. zero padding

struct GemmArgs {
    const u16 *A;        // canonical residues (< q), rows padded with zeros to ksteps*16 terms, 16B-aligned rows
    const int16_t *Bt;   // centered residues, [n padded to 128][ldb], zero padded
    u16 *C;              // canonical residues out
    long long lda, ldb, ldc;
    long long a_batch, b_batch, c_batch;   // strides (elements) of blockIdx.z
    int mtotal, ksteps, nvalid, c_off;
    // row m -> storage row (m / rpp) * slots + slot_lo + m % rpp  (planes of one proof are `slots` rows apart)
    int rpp, slot_lo, a_slots, c_slots;
    int tail;            // 1: also copy A[row][tail_off .. +150] to C[row][c_off-151 ..] (parties 0..150, ss.cpp:7-11,:77-80)
    int tail_off;        // element offset of the 151 tail values in an A row (256 unless A was advanced past the secrets)
    // Constant-secret rows (the eta sharings, mlwe_prover.cpp:41-59): all 256 secrets equal c, so only the 151 tail terms go
    // through the contraction (A and Bt advanced by 256, ksteps = 10) and c * U[n], U[n] = sum_{j<256} Bt[n][j], is added here.
    const int16_t *addvec;   // centered U, or nullptr
    const u16 *scale_src;    // c of row r = scale_src[r * lda] (canonical), same row mapping as A
};

template <int TM, int NREG>
__global__ void __maxnreg__(NREG) k_gf_gemm(const GemmArgs g)
{
    constexpr int BM = 16 * TM;
    __shared__ __align__(16) int32_t As[2][GE_BK][BM];
    __shared__ __align__(16) int32_t Bs[2][GE_BK][GE_BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * GE_BN;
    const u16 *Ab = g.A + (size_t)blockIdx.z * g.a_batch;
    const int16_t *Bb = g.Bt + (size_t)blockIdx.z * g.b_batch;
    u16 *Cb = g.C + (size_t)blockIdx.z * g.c_batch;
    // loader mapping: thread -> (row, 8-term half)
    const int lrb = tid & 127, lhb = tid >> 7;
    const int lra = tid & (BM - 1), lha = (tid / BM) & 1;
    const bool a_thr = tid < 2 * BM;
    const int am = m0 + lra;
    const bool a_ok = a_thr && am < g.mtotal;
    const u16 *a_src = Ab;
    if (a_ok) a_src = Ab + ((size_t)(am / g.rpp) * g.a_slots + g.slot_lo + am % g.rpp) * g.lda + lha * 8;
    const int16_t *b_src = Bb + (size_t)(n0 + lrb) * g.ldb + lhb * 8;

    int32_t acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0;

    uint4 ra = make_uint4(0, 0, 0, 0), rb;
    if (a_ok) ra = *reinterpret_cast<const uint4 *>(a_src);
    rb = *reinterpret_cast<const uint4 *>(b_src);
    auto stage = [&](int buf) {
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (a_thr) {
                const int32_t a0 = (int32_t)(wa[i] & 0xFFFF), a1 = (int32_t)(wa[i] >> 16);
                As[buf][lha * 8 + 2 * i][lra] = a0 > Q / 2 ? a0 - Q : a0;
                As[buf][lha * 8 + 2 * i + 1][lra] = a1 > Q / 2 ? a1 - Q : a1;
            }
            Bs[buf][lhb * 8 + 2 * i][lrb] = (int32_t)(int16_t)(wb[i] & 0xFFFF);
            Bs[buf][lhb * 8 + 2 * i + 1][lrb] = (int32_t)(int16_t)(wb[i] >> 16);
        }
    };
    stage(0);
    __syncthreads();
#pragma unroll 1
    for (int kt = 0; kt < g.ksteps; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < g.ksteps) {
            if (a_ok) ra = *reinterpret_cast<const uint4 *>(a_src + (kt + 1) * GE_BK);
            rb = *reinterpret_cast<const uint4 *>(b_src + (kt + 1) * GE_BK);
        }
#pragma unroll
        for (int kk = 0; kk < GE_BK; kk++) {
            int32_t av[TM], bv[8];
#pragma unroll
            for (int h = 0; h < TM / 4; h++) {
                const int4 a4 = *reinterpret_cast<const int4 *>(&As[buf][kk][h * (BM / 2) + ty * 4]);
                av[4 * h] = a4.x; av[4 * h + 1] = a4.y; av[4 * h + 2] = a4.z; av[4 * h + 3] = a4.w;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int4 b4 = *reinterpret_cast<const int4 *>(&Bs[buf][kk][h * 64 + tx * 4]);
                bv[4 * h] = b4.x; bv[4 * h + 1] = b4.y; bv[4 * h + 2] = b4.z; bv[4 * h + 3] = b4.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] += av[i] * bv[j];
        }
        if ((kt & 31) == 31) {          // keep |acc| < 2^31 for rows longer than 512 terms
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = acc[i][j] % Q;
        }
        if (kt + 1 < g.ksteps) { stage(buf ^ 1); __syncthreads(); }
    }
    // epilogue: canonical residues, 8-byte stores along the contiguous (party / coefficient) axis
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
        if (m >= g.mtotal) continue;
        u16 *dst = Cb + ((size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp) * g.ldc + g.c_off;
        if (g.addvec) {
            const int32_t cst = gf_center(g.scale_src[((size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp) * g.lda]);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][h * 4 + j] += cst * (int32_t)g.addvec[n0 + h * 64 + tx * 4 + j];
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int x = n0 + h * 64 + tx * 4;
            u16 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (u16)gf_canon(acc[i][h * 4 + j]);
            if (x + 3 < g.nvalid) {
                *reinterpret_cast<uint2 *>(dst + x) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (x + j < g.nvalid) dst[x + j] = v[j];
            }
        }
    }
    if (g.tail && blockIdx.x == 0) {
        for (int idx = tid; idx < BM * (NT + 1); idx += 256) {
            const int r = idx / (NT + 1), c = idx % (NT + 1), m = m0 + r;
            if (m >= g.mtotal) continue;
            const size_t ar = (size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp, cr = (size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp;
            Cb[cr * g.ldc + g.c_off - (NT + 1) + c] = Ab[ar * g.lda + g.tail_off + c];
        }
    }
}

// host-side launcher; returns the number of kernels launched
template <int TM, int NREG = 128>
static inline int gf_gemm_launch(const GemmArgs &g, int npad, int nbatch, cudaStream_t st)
{
    dim3 grid(npad / GE_BN, (g.mtotal + 16 * TM - 1) / (16 * TM), nbatch);
    k_gf_gemm<TM, NREG><<<grid, 256, 0, st>>>(g);
    return 1;
}

}  // namespace kosk
