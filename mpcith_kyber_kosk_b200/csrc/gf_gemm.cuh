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
#include <algorithm>

namespace kosk {

typedef uint16_t u16;
typedef uint8_t u8;

constexpr int GE_BN = 128, GE_BK = 16;
constexpr int GE_NPAD = ((NX + GE_BN - 1) / GE_BN) * GE_BN;   // 1408 rows of the share table incl. zero padding
constexpr int GE_NCOLS7 = ((NX + 111) / 112) * 112;           // 1344: columns covered by 112-wide tiles

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
    int half_last;           // 1: only the first 8 terms of the last 16-term step are non-zero (407 = 25*16 + 7)
    const u16 *colscale;     // optional per-column factor (canonical), applied after the reduction: C[m][n] = acc * colscale[n] mod q
    long long colscale_batch;
    int colscale_by_group;   // 1: the factor row is chosen by the row's group m / rpp (one proof) instead of blockIdx.z
    // Latency mode (a handful of rows): split-K over blockIdx.z.  Every slice writes its raw int32 accumulators to
    // ws[slice][row][ws_ld]; k_gf_gemm_finish sums the slices and runs the epilogue.  ws = nullptr disables it.
    int32_t *ws; long long ws_elems; int ws_ld;
};
constexpr long long GE_WS_ELEMS = 6ll << 20;      // 24 MB of int32 partial sums per scratch set

// TN = columns per thread: 8 -> 128-column CTA tile (4 + 4 split), 7 -> 112-column tile (4 + 2 + 1 split).  1303 columns
// are 12 x 112 = 1344 (3 % padding) instead of 11 x 128 = 1408 (8 %); all shared loads stay conflict-free / broadcast.
template <int TM, int NREG, int TN, bool SPLITK = false>
__global__ void __maxnreg__(NREG) k_gf_gemm(const GemmArgs g)
{
    static_assert(TN == 8 || TN == 7, "TN");
    constexpr int BM = 16 * TM, BN = 16 * TN;
    __shared__ __align__(16) int32_t As[2][GE_BK][BM];
    __shared__ __align__(16) int32_t Bs[2][GE_BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int zb = SPLITK ? 0 : blockIdx.z;
    const int kt_lo = SPLITK ? (int)((long long)blockIdx.z * g.ksteps / gridDim.z) : 0;
    const int kt_hi = SPLITK ? (int)((long long)(blockIdx.z + 1) * g.ksteps / gridDim.z) : g.ksteps;
    const u16 *Ab = g.A + (size_t)zb * g.a_batch;
    const int16_t *Bb = g.Bt + (size_t)zb * g.b_batch;
    u16 *Cb = g.C + (size_t)zb * g.c_batch;
    // loader mapping: thread -> (row, 8-term half)
    const bool b_thr = tid < 2 * BN;
    const int lrb = b_thr ? tid % BN : 0, lhb = b_thr ? tid / BN : 0;
    const int lra = tid & (BM - 1), lha = (tid / BM) & 1;
    const bool a_thr = tid < 2 * BM;
    const int am = m0 + lra;
    const bool a_ok = a_thr && am < g.mtotal;
    const u16 *a_src = Ab;
    if (a_ok) a_src = Ab + ((size_t)(am / g.rpp) * g.a_slots + g.slot_lo + am % g.rpp) * g.lda + lha * 8;
    const int16_t *b_src = Bb + (size_t)(n0 + lrb) * g.ldb + lhb * 8;
    if (SPLITK) { a_src += (size_t)kt_lo * GE_BK; b_src += (size_t)kt_lo * GE_BK; }
    // column c of this thread's TN-wide strip sits at tile column col_of(c)
    auto col_of = [&](int c) { return c < 4 ? tx * 4 + c : (TN == 8 ? 64 + tx * 4 + (c - 4) : (c < 6 ? 64 + tx * 2 + (c - 4) : 96 + tx)); };

    int32_t acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0;

    uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (a_ok) ra = *reinterpret_cast<const uint4 *>(a_src);
    if (b_thr) rb = *reinterpret_cast<const uint4 *>(b_src);
    auto stage = [&](int buf) {
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (a_thr) {
                int32_t a0 = (int32_t)(wa[i] & 0xFFFF), a1 = (int32_t)(wa[i] >> 16);
                if (a0 >= Q) a0 %= Q;                       // rows need not be canonical (any u16 acts as its residue, utils/gf3329.c:282-284)
                if (a1 >= Q) a1 %= Q;
                As[buf][lha * 8 + 2 * i][lra] = a0 > Q / 2 ? a0 - Q : a0;
                As[buf][lha * 8 + 2 * i + 1][lra] = a1 > Q / 2 ? a1 - Q : a1;
            }
            if (b_thr) {
                Bs[buf][lhb * 8 + 2 * i][lrb] = (int32_t)(int16_t)(wb[i] & 0xFFFF);
                Bs[buf][lhb * 8 + 2 * i + 1][lrb] = (int32_t)(int16_t)(wb[i] >> 16);
            }
        }
    };
    stage(0);
    __syncthreads();
#pragma unroll 1
    for (int kt = kt_lo; kt < kt_hi; kt++) {
        const int buf = (kt - kt_lo) & 1;
        if (kt + 1 < kt_hi) {
            if (a_ok) ra = *reinterpret_cast<const uint4 *>(a_src + (kt + 1 - kt_lo) * GE_BK);
            if (b_thr) rb = *reinterpret_cast<const uint4 *>(b_src + (kt + 1 - kt_lo) * GE_BK);
        }
        auto mac = [&](int kk) {
            int32_t av[TM], bv[8];
#pragma unroll
            for (int h = 0; h < TM / 4; h++) {
                const int4 a4 = *reinterpret_cast<const int4 *>(&As[buf][kk][h * (BM / 2) + ty * 4]);
                av[4 * h] = a4.x; av[4 * h + 1] = a4.y; av[4 * h + 2] = a4.z; av[4 * h + 3] = a4.w;
            }
            const int4 b4 = *reinterpret_cast<const int4 *>(&Bs[buf][kk][tx * 4]);
            bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
            if (TN == 8) {
                const int4 c4 = *reinterpret_cast<const int4 *>(&Bs[buf][kk][64 + tx * 4]);
                bv[4] = c4.x; bv[5] = c4.y; bv[6] = c4.z; bv[7] = c4.w;
            } else {
                const int2 c2 = *reinterpret_cast<const int2 *>(&Bs[buf][kk][64 + tx * 2]);
                bv[4] = c2.x; bv[5] = c2.y; bv[6] = Bs[buf][kk][96 + tx]; bv[7] = 0;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] += av[i] * bv[j];
        };
#pragma unroll
        for (int kk = 0; kk < GE_BK / 2; kk++) mac(kk);
        if (kt + 1 < g.ksteps || !g.half_last) {      // a 407-term row needs only 8 terms of its 26th step (terms 408..415 are padding)
#pragma unroll
            for (int kk = GE_BK / 2; kk < GE_BK; kk++) mac(kk);
        }
        if ((kt & 31) == 31) {          // keep |acc| < 2^31 for rows longer than 512 terms
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = acc[i][j] % Q;
        }
        if (kt + 1 < kt_hi) { stage(buf ^ 1); __syncthreads(); }
    }
    if (SPLITK) {            // raw partial sums of this k-slice; the epilogue runs in k_gf_gemm_finish
#pragma unroll
        for (int i = 0; i < TM; i++) {
            const int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
            if (m >= g.mtotal) continue;
            int32_t *w = g.ws + ((size_t)blockIdx.z * g.mtotal + m) * g.ws_ld + n0;
#pragma unroll
            for (int j = 0; j < TN; j++) w[col_of(j)] = acc[i][j];
        }
        return;
    }
    // epilogue: canonical residues, vector stores along the contiguous (party / coefficient) axis
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
        if (m >= g.mtotal) continue;
        u16 *dst = Cb + ((size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp) * g.ldc + g.c_off;
        if (g.addvec) {
            const int32_t cst = gf_center(g.scale_src[((size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp) * g.lda] % (uint32_t)Q);
#pragma unroll
            for (int j = 0; j < TN; j++) acc[i][j] += cst * (int32_t)g.addvec[n0 + col_of(j)];
        }
        u16 v[TN];
#pragma unroll
        for (int j = 0; j < TN; j++) v[j] = (u16)gf_canon(acc[i][j]);
        if (g.colscale) {
            const u16 *cs = g.colscale + (size_t)(g.colscale_by_group ? m / g.rpp : blockIdx.z) * g.colscale_batch + n0;
#pragma unroll
            for (int j = 0; j < TN; j++) v[j] = (u16)gf_mul(v[j], cs[col_of(j)]);
        }
        auto put4 = [&](int x, const u16 *w) {
            if (x + 3 < g.nvalid) *reinterpret_cast<uint2 *>(dst + x) = make_uint2((uint32_t)w[0] | ((uint32_t)w[1] << 16), (uint32_t)w[2] | ((uint32_t)w[3] << 16));
            else { for (int j = 0; j < 4; j++) if (x + j < g.nvalid) dst[x + j] = w[j]; }
        };
        put4(n0 + tx * 4, v);
        if (TN == 8) put4(n0 + 64 + tx * 4, v + 4);
        else {
            const int x = n0 + 64 + tx * 2;
            if (x + 1 < g.nvalid) *reinterpret_cast<uint32_t *>(dst + x) = (uint32_t)v[4] | ((uint32_t)v[5] << 16);
            else if (x < g.nvalid) dst[x] = v[4];
            if (n0 + 96 + tx < g.nvalid) dst[n0 + 96 + tx] = v[6];
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

// second pass of the split-K form: sum the k-slices and run the epilogue of k_gf_gemm (one thread per output element)
__global__ void __launch_bounds__(256) k_gf_gemm_finish(const GemmArgs g, int nslices)
{
    const int m = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    const size_t ar = (size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp, cr = (size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp;
    if (n < g.nvalid) {
        long long sum = 0;
        for (int z = 0; z < nslices; z++) sum += g.ws[((size_t)z * g.mtotal + m) * g.ws_ld + n];
        if (g.addvec) sum += (long long)gf_center(g.scale_src[ar * g.lda] % (uint32_t)Q) * (long long)g.addvec[n];
        long long r = sum % Q; if (r < 0) r += Q;
        uint32_t v = (uint32_t)r;
        if (g.colscale) v = gf_mul(v, g.colscale[(size_t)(g.colscale_by_group ? m / g.rpp : 0) * g.colscale_batch + n]);
        g.C[cr * g.ldc + g.c_off + n] = (u16)v;
    }
    if (g.tail && n <= NT) g.C[cr * g.ldc + g.c_off - (NT + 1) + n] = g.A[ar * g.lda + g.tail_off + n];
}

// host-side launcher; returns the number of kernels launched
// npad = number of Bt rows available (zero padded to a multiple of the column tile)
template <int TM, int NREG = 128, int TN = 8>
static inline int gf_gemm_launch(const GemmArgs &g, int npad, int nbatch, cudaStream_t st)
{
    dim3 grid(npad / (16 * TN), (g.mtotal + 16 * TM - 1) / (16 * TM), nbatch);
    k_gf_gemm<TM, NREG, TN><<<grid, 256, 0, st>>>(g);
    return 1;
}

// Tile height by occupancy: when 128-row tiles would not give every SM a CTA (a handful of proofs in flight), 64-row tiles
// halve the latency of a tile; throughput-sized launches keep the 128-row tile.
template <int TN>
static inline int gf_gemm_launch_auto(const GemmArgs &g, int npad, int nbatch, cudaStream_t st)
{
    const long long ctas = (long long)(npad / (16 * TN)) * ((g.mtotal + 127) / 128) * nbatch;
    if (ctas >= 148) return gf_gemm_launch<8, 128, TN>(g, npad, nbatch, st);
    const int ctas64 = (npad / (16 * TN)) * ((g.mtotal + 63) / 64);
    int nz = std::min(std::min(8, g.ksteps / 4), 148 / std::max(1, ctas64));
    if (nbatch == 1 && g.ws && nz >= 2 && (long long)nz * g.mtotal * npad <= g.ws_elems) {
        // a handful of rows: 64-row tiles, k split over blockIdx.z so that the launch covers the SMs, then the epilogue pass
        GemmArgs h = g; h.ws_ld = npad;
        dim3 grid(npad / (16 * TN), (g.mtotal + 63) / 64, nz);
        k_gf_gemm<4, 128, TN, true><<<grid, 256, 0, st>>>(h);
        k_gf_gemm_finish<<<dim3((std::max(g.nvalid, NT + 1) + 255) / 256, g.mtotal), 256, 0, st>>>(h, nz);
        return 2;
    }
    return gf_gemm_launch<4, 128, TN>(g, npad, nbatch, st);
}

}  // namespace kosk
