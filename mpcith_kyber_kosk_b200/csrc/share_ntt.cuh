// share_ntt.cuh -- packed-Shamir share evaluation as a blocked cyclic convolution over GF(3329).
// Same function as the table mat-vec of share_secrets_ddeg / recompute_share_secrets_ddeg (reference ss.cpp:23-32, :88-97):
//   share[151 + x] = sum_{j<407} S[x][j] y[j],   S[x][j] = l_j(x + 407) over the nodes 0..406   
// but S is never formed.  In barycentric form S[x][j] = P(x) * w_j / (x + 407 - j) with w_j = 1 / prod_{m != j} (j - m) and
// P(x) = prod_{m<407} (x + 407 - m), so the sum is a convolution of u_j = w_j y_j with the fixed sequence c[m] = 1 / m
// (m = 1..1709), scaled by P(x).  GF(3329) has 256-th roots of unity (17 generates them: the Kyber NTT's root), so the
// convolution is cut into 128-wide input blocks i (4 of them) and output blocks o (11): block (o, i) is a length-256 cyclic
// convolution with the kernel segment K_{o-i}[d] = c[128 (o - i) + 407 + d], d in [-127, 127], and
//   out block o = INTT( sum_i NTT(K_{o-i}) . NTT(u block i) ) [0..127].
// Per sharing: 4 forward and 11 inverse 256-point NTTs plus 11 x 4 x 256 pointwise MACs, about 115 k IMADs instead of the
// 530 k MACs of the dense table (every value is a field element, so the result is the same canonical residue).
// A 256-point NTT is done as 16 x 16 (n = 16 a + b, k = k1 + 16 k2): a 16-point DFT down the columns, the twiddle
// w^(b k1), a 16-point DFT along the rows; the 16-point DFTs are radix-2 networks in registers whose twiddles are compile-time
// constants, i.e. IMAD immediates (SnDft below).  One warp per sharing; its two half-warps run two NTTs at a time (two input
// blocks, then two output blocks), lane = column (then row) of the 16 x 16 matrix, transposed once per NTT through shared memory.
// Arithmetic: int32 values kept unreduced wherever the interval bounds allow it (tools/exp/fft16_plan.py); reductions are Barrett
// (a - mulhi(a, floor(2^32 / q)) q) and multiplications by constants Shoup's (s w - mulhi(s, w') q with w' = round(w 2^32 / q)), both
// on the FMA pipe; no Montgomery factors in any table.
// Two kernels: k_conv_ntt<NIN, NOUT, ...>, generic over the block counts (the verifier's other Toeplitz products, and the sharing with
// KOSK_B200_SHARE_NTT=1), and k_share_ntt2, the sharing's own kernel (unequal blocks, IDP.2A pointwise stage; the default).
#pragma once
#include "gf_gemm.cuh"
#include <vector>

namespace kosk {

constexpr int SN_NIN = 4, SN_NOUT = 11;                                  // the sharing: 407 inputs in 4 blocks, 1303 outputs in 11
#ifndef KOSK_SN_WARPS
#define KOSK_SN_WARPS 4
#endif
#ifndef KOSK_SN_MINB
#define KOSK_SN_MINB 4
#endif
constexpr int SN_WARPS = KOSK_SN_WARPS;                                  // rows in flight per CTA
// transpose buffers: 16 rows of SN_TS int16 per half-warp.  An even stride of 18 (9 words) puts the 16 rows a half-warp reads at 16 distinct banks (9 c mod 32)
// for every column, and the second half-warp's buffer (288 int16 = 16 banks further) on the other 16; with 17 rows 0 and 15 collided for odd columns.
constexpr int SN_TS = 18;
constexpr int SN_LD = 24;               // int16 per k1 row of a spectrum in shared memory: 16 values (k2) + pad, so that the 128-bit row reads of a quarter-warp hit distinct banks

// compile-time helpers for the constant tables (on sm_100a an IMAD takes no constant-bank operand: a value folded from a constexpr table in an
// unrolled loop is a 32-bit immediate of the IMAD itself, a __constant__ array element a separate LDC)
constexpr uint32_t sn_cpow(uint32_t b, uint32_t e) { uint32_t r = 1; b %= Q; while (e) { if (e & 1) r = r * b % Q; b = b * b % Q; e >>= 1; } return r; }
constexpr int32_t sn_ccenter(uint32_t v) { return (int32_t)(v % Q) > Q / 2 ? (int32_t)(v % Q) - Q : (int32_t)(v % Q); }

// Modular arithmetic of the kernel (all on the FMA pipe, no shifts or sign extensions on the ALU pipe):
//   sn_barrett(a)      a mod q          in (-105, 2q)   for |a| < 2^28:  a - mulhi(a, floor(2^32 / q)) q
//   sn_shoup(s, w, w') s w mod q        in (-105, q + 105) for |s| < 2^28, a constant w in [-1664, 1664] and its companion
//                                       w' = round(w 2^32 / q):  s w - mulhi(s, w') q  (32-bit wrapping; the true value is small)
// (for |a|, |s| up to 2^31 the results stay within (-q/4, 2q + q/2) and (-q/4, 5q/4): what the DFT networks below rely on)
constexpr int32_t SN_BARRETT_M = (int32_t)((1ull << 32) / Q);
__device__ __forceinline__ int32_t sn_barrett(int32_t a) { return a - __mulhi(a, SN_BARRETT_M) * Q; }
__device__ __forceinline__ int32_t sn_shoup(int32_t s, int32_t w, int32_t wp)
{
    return (int32_t)((uint32_t)s * (uint32_t)w - (uint32_t)__mulhi(s, wp) * (uint32_t)Q);
}
__host__ __device__ inline int2 sn_pair(uint32_t v)                  // residue -> (centered w, w' = round(w 2^32 / q))
{
    const int32_t w = (int32_t)(v % Q) > Q / 2 ? (int32_t)(v % Q) - Q : (int32_t)(v % Q);
    const long long num = (long long)w * 4294967296LL;
    return make_int2(w, (int32_t)((num >= 0 ? num + Q / 2 : num - Q / 2) / Q));
}

// 16-point DFTs as radix-2 decimation-in-time networks in registers (natural order in and out) instead of 16 x 16 mat-vecs:
// 17 twiddle multiplications per transform.  A multiplication is either LAZY (one IMAD, the product of a small value with a centered constant
// stays an unreduced int32) or a SHOUP multiplication (IMAD.HI + 2 IMAD = 4 issue slots, result in (-q/4, 5q/4) for any |s| < 2^31); which one
// is a 17-bit compile-time mask, chosen by exhaustive search over interval bounds (tools/exp/fft16_plan.py) so that no intermediate exceeds
// 2^31.  Multiplication ids: size-4 transform at offset o (x I): o (0..3); size-8 transform at offset o, k = 1..3: 4 + 3 o + k - 1; size 16,
// k = 1..7: 9 + k.  SN_FFT_SMALL (inputs below 1.25 q: 32 slots) keeps only the operands of later lazy multiplications small; SN_FFT_BIG
// (inputs up to 2^25.5, the unreduced pointwise sums: 68 slots) reduces in every multiplication, which replaces 16 Barrett reductions.
constexpr uint32_t SN_FFT_SMALL = 0x386u, SN_FFT_BIG = 0x1ffffu;
struct SnTw16 { int32_t w[16], wp[16]; };
constexpr SnTw16 sn_make_tw16(bool inv)
{
    SnTw16 t{};
    const uint32_t om16 = sn_cpow(17, 16), w = inv ? sn_cpow(om16, Q - 2) : om16;
    for (int j = 0; j < 16; j++) {
        t.w[j] = sn_ccenter(sn_cpow(w, (uint32_t)j));
        const long long num = (long long)t.w[j] * 4294967296LL;
        t.wp[j] = (int32_t)((num >= 0 ? num + Q / 2 : num - Q / 2) / Q);
    }
    return t;
}
__device__ constexpr SnTw16 c_sn_tw16f = sn_make_tw16(false), c_sn_tw16i = sn_make_tw16(true);

template <int N, int OFF, int STRIDE, bool INV, uint32_t MASK>
struct SnDft {
    static __device__ __forceinline__ void run(const int32_t (&x)[16], int32_t (&out)[N])
    {
        int32_t E[N / 2], O[N / 2];
        SnDft<N / 2, OFF, 2 * STRIDE, INV, MASK>::run(x, E);
        SnDft<N / 2, OFF + STRIDE, 2 * STRIDE, INV, MASK>::run(x, O);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            int32_t t = O[k];
            if (k > 0) {
                const int id = N == 4 ? OFF : N == 8 ? 4 + 3 * OFF + k - 1 : 9 + k, j = k * (16 / N);
                const int32_t w = INV ? c_sn_tw16i.w[j] : c_sn_tw16f.w[j], wp = INV ? c_sn_tw16i.wp[j] : c_sn_tw16f.wp[j];
                t = ((MASK >> id) & 1u) ? sn_shoup(O[k], w, wp) : O[k] * w;
            }
            out[k] = E[k] + t; out[k + N / 2] = E[k] - t;
        }
    }
};
template <int OFF, int STRIDE, bool INV, uint32_t MASK>
struct SnDft<1, OFF, STRIDE, INV, MASK> {
    static __device__ __forceinline__ void run(const int32_t (&x)[16], int32_t (&out)[1]) { out[0] = x[OFF]; }
};
template <bool INV, uint32_t MASK>
__device__ __forceinline__ void sn_dft16(const int32_t (&x)[16], int32_t (&out)[16]) { SnDft<16, 0, 1, INV, MASK>::run(x, out); }

// One Toeplitz product per row:  C[row][c_off + x] = post[x] * sum_j c[x - j + OFF] * (pre[j] * A[row][j]),  c[m] = 1/m (0 for m = 0),
// x < nout <= 128 NOUT, j < nin <= 128 NIN.  OFF lives in the kernel-segment table `khat` ([NIN + NOUT - 1] segments for o - i).
//   sharing (ss.cpp:23-32, :88-97)                 OFF = 407,  pre = w_j,      post = P(x)        (4, 11)
//   recon_secrets_ddeg / _2ddeg (ss.cpp:37-73)      OFF = -256, pre = w_j,      post = P'(i)       (4, 2) / (7, 2)
//   verifier interpolation (mlwe_verifier.cpp:188-224 etc.), rows already scaled by the per-proof weights, columns = parties
//                                                   OFF = -256, pre = none,     post = per-proof P(t) (5, 4) / (8, 2)
// Work ticket of k_share_ntt2: a device counter private to one stream.  Every processed row draws exactly one ticket, so a launch over
// mtotal rows advances the counter by mtotal and the host knows its value at the start of the next launch without resetting it.
struct SnTicket { unsigned *ctr = nullptr; unsigned base = 0; };
struct ConvArgs {
    const u16 *A; u16 *C; long long lda, ldc;
    int mtotal, rpp, slot_lo, a_slots, c_slots, c_off;   // row m -> storage row (m / rpp) * slots + slot_lo + m % rpp
    int tail, tail_off;                                  // sharing: copy the 151 tail values to parties 0..150
    const int2 *tw;          // [2][16][16]  (w, w') of 17^(+-b k1) (symmetric in b, k1)
    const int16_t *khat;     // [NIN + NOUT - 1][16 k1][SN_LD]  NTT(K_delta)[k1 + 16 k2] / 256, centered, at [k1][k2], delta = o - i
    const uint32_t *kpk;     // k_share_ntt2 only: [S2_NOUT][2][256] packed limb words of the (o, i) segment spectra
    unsigned *ctr; unsigned ctr_base;   // k_share_ntt2 only: work ticket (SnTicket); nullptr = rows strided statically over the warps
    const int2 *pre;         // [128 NIN] (w, w') of the input factors, or nullptr
    const int2 *post;        // (w, w') of the output factors
    long long post_group;    // 0: one table; else the table of row m starts at post + (m / rpp) * post_group
};

// NINV / NOUTV = valid inputs / outputs, PRE = input factors present, PGROUP = per-row-group output factors: compile-time, so that
// the sharing's kernel carries none of the other products' branches
template <int NIN, int NOUT, int NINV, int NOUTV, bool PRE, bool PGROUP>
__global__ void __launch_bounds__(32 * SN_WARPS, (NIN <= 5 ? KOSK_SN_MINB : KOSK_SN_MINB - 1)) k_conv_ntt(const ConvArgs g)
{
    constexpr int NK = NIN + NOUT - 1, NINP = (NIN + 1) & ~1;
    __shared__ __align__(16) int16_t s_kh[NK * 16 * SN_LD];
    __shared__ __align__(8) int2 s_tw[2 * 256];
    __shared__ __align__(16) int16_t s_uh[SN_WARPS][NINP][16 * SN_LD];
    // transpose buffers, one per half-warp (SN_TS)
    __shared__ __align__(4) int16_t s_t[SN_WARPS][2][288];      // 4-byte aligned: the halfword loads of a row (36 c bytes in) pair up into LDS.32
    for (int i = threadIdx.x; i < NK * 16 * SN_LD; i += blockDim.x) s_kh[i] = g.khat[i];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tw[i] = g.tw[i];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, c = lane & 15;
    int16_t (*uh)[16 * SN_LD] = s_uh[wid];
    int16_t *T = s_t[wid][hw];
    auto lo16 = [](uint32_t w) -> int32_t { return (int32_t)(int16_t)(w & 0xFFFFu); };
    auto hi16 = [](uint32_t w) -> int32_t { return (int32_t)w >> 16; };
    // rows are drawn from a work ticket and the next row is prefetched, as in k_share_ntt2 below
    const int nw = gridDim.x * SN_WARPS;
    int m = blockIdx.x * SN_WARPS + wid;
    while (m < g.mtotal) {
        int mn = m + nw;
        if (g.ctr) {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(g.ctr, 1u) - g.ctr_base;
            mn = nw + (int)__shfl_sync(0xffffffffu, t, 0);
        }
        const u16 *yrow = g.A + ((size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp) * g.lda;
        u16 *dst = g.C + ((size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp) * g.ldc + g.c_off;
        const int2 *post = g.post + (PGROUP ? (size_t)(m / g.rpp) * g.post_group : 0);
        if (mn < g.mtotal && lane * 64 < NINV) {
            const u16 *nrow = g.A + ((size_t)(mn / g.rpp) * g.a_slots + g.slot_lo + mn % g.rpp) * g.lda + lane * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow));
        }
        if (g.tail) for (int t = lane; t <= NT; t += 32) dst[t - (NT + 1)] = yrow[g.tail_off + t];
        // ---- forward: u_j = pre_j A_j, NTT of the zero-padded 128-wide input blocks (two per pass) ----
#pragma unroll 1
        for (int it = 0; it < NINP / 2; it++) {
            const int blk = 2 * it + hw;
            int32_t x[16];                                                 // rows 8..15 of a block are zero padding: they fold away in the transform
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const int j = 128 * blk + 16 * a + c;
                int32_t v = 0;
                if (a < 8 && j < NINV) {
                    v = (int32_t)yrow[j];
                    if (PRE) { const int2 p = __ldg(g.pre + j); v = sn_shoup(v, p.x, p.y); }   // without a factor: any u16, the transform's bounds hold up to 65535
                }
                x[a] = v;
            }
            int32_t y[16];
            sn_dft16<false, SN_FFT_SMALL>(x, y);
#pragma unroll
            for (int k = 0; k < 16; k++) { const int2 t = s_tw[k * 16 + c]; T[k * SN_TS + c] = (int16_t)sn_shoup(y[k], t.x, t.y); }
            __syncwarp();
            int32_t in[16];
#pragma unroll
            for (int b = 0; b < 16; b++) in[b] = T[c * SN_TS + b];
            int32_t X[16];
            sn_dft16<false, SN_FFT_SMALL>(in, X);
#pragma unroll
            for (int k = 0; k < 16; k++) X[k] = sn_barrett(X[k]);
            {   // row k1 = c of the spectrum: 16 values (k2) as two 128-bit stores
                uint32_t w[8];
#pragma unroll
                for (int k = 0; k < 8; k++) w[k] = ((uint32_t)X[2 * k] & 0xFFFFu) | ((uint32_t)X[2 * k + 1] << 16);
                uint4 *p = reinterpret_cast<uint4 *>(&uh[blk][c * SN_LD]);
                p[0] = make_uint4(w[0], w[1], w[2], w[3]); p[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            __syncwarp();
        }
        // ---- inverse: output blocks two per pass ----
#pragma unroll 1
        for (int it = 0; it < (NOUT + 1) / 2; it++) {
            const int o = 2 * it + hw;
            const bool live = o < NOUT;
            int32_t O[16];
            {
                int32_t acc[16];
#pragma unroll
                for (int k = 0; k < 16; k++) acc[k] = 0;
                if (live) {
#pragma unroll
                    for (int i = 0; i < NIN; i++) {
                        const uint4 *pu = reinterpret_cast<const uint4 *>(&uh[i][c * SN_LD]);
                        const uint4 *pk = reinterpret_cast<const uint4 *>(&s_kh[((o - i + NIN - 1) * 16 + c) * SN_LD]);
#pragma unroll
                        for (int h2 = 0; h2 < 2; h2++) {
                            const uint4 u4 = pu[h2], k4 = pk[h2];
                            const uint32_t uw[4] = {u4.x, u4.y, u4.z, u4.w}, kw[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                acc[8 * h2 + 2 * e] += lo16(uw[e]) * lo16(kw[e]);
                                acc[8 * h2 + 2 * e + 1] += hi16(uw[e]) * hi16(kw[e]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; k++) O[k] = acc[k];                 // at most 8 x 6700 x 1664 < 2^26.5: reduced inside the transform (SN_FFT_BIG)
            }
            int32_t v[16];
            sn_dft16<true, SN_FFT_BIG>(O, v);
#pragma unroll
            for (int b = 0; b < 16; b++) { const int2 t = s_tw[256 + b * 16 + c]; T[b * SN_TS + c] = (int16_t)sn_shoup(v[b], t.x, t.y); }
            __syncwarp();
            int32_t in[16], out[16];
#pragma unroll
            for (int k = 0; k < 16; k++) in[k] = T[c * SN_TS + k];
            sn_dft16<true, SN_FFT_SMALL>(in, out);                          // only out[0..7] are used: the rest of the last level is dead code
            const int na = !live ? 0 : min(8, (NOUTV - (128 * o + c) + 15) >> 4);
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const int xo = 128 * o + 16 * a + c;
                if (a < na) {
                    const int2 pf = PGROUP ? post[xo] : __ldg(post + xo);
                    uint32_t r = (uint32_t)sn_shoup(out[a], pf.x, pf.y);     // in (-q/4, 5q/4): canonical with two unsigned minima
                    r = min(r, r + Q);
                    r = min(r, r - Q);
                    dst[xo] = (u16)r;
                }
            }
            __syncwarp();
        }
        m = mn;
    }
}

template <int NIN, int NOUT, int NINV, int NOUTV, bool PRE, bool PGROUP>
static inline int conv_ntt_launch(const ConvArgs &g0, cudaStream_t st, SnTicket *tk = nullptr)
{
    ConvArgs g = g0;
    const int ctas = std::min((g.mtotal + SN_WARPS - 1) / SN_WARPS, 148 * 8);
    g.ctr = nullptr; g.ctr_base = 0;
    if (tk && tk->ctr && ctas > 0) { g.ctr = tk->ctr; g.ctr_base = tk->base; tk->base += (unsigned)g.mtotal; }
    if (ctas > 0) k_conv_ntt<NIN, NOUT, NINV, NOUTV, PRE, PGROUP><<<ctas, 32 * SN_WARPS, 0, st>>>(g);
    if (g.ctr && cudaPeekAtLastError() != cudaSuccess) tk->base -= (unsigned)g.mtotal;      // the launch never ran: no ticket was drawn
    return 1;
}

// ---------------------------------------------------------------------------------------------
// The sharing itself (the (4, 11) product above, 96 % of all rows) with unequal blocks and a packed pointwise stage.
//   * Blocks: 126 inputs x 131 outputs per length-256 cyclic convolution (126 + 131 - 1 = 256) instead of 128 x 128: 4 input blocks
//     (504 >= 407) and TEN output blocks (1310 >= 1303) instead of eleven, i.e. five two-block passes instead of six.  The ninth output row
//     (x' = 128..130) of a block is free: w16^(8 k) = (-1)^k, so rows 0 and 8 share one even / odd sum.  The kernel segment of block (o, i)
//     is K[d] = c[131 o - 126 i + 407 + d], d in [-125, 130]: it depends on o and i separately, 40 spectra instead of 14.
//   * Pointwise stage: the spectra of input blocks 2p and 2p + 1 sit in one 32-bit word (two int16), and the segment spectra of (o, 2p),
//     (o, 2p + 1) are split into signed limbs k = 64 k1 + k0 (k0 in [-32, 32), |k1| <= 26) and packed as the four bytes (k0, k0', k1, k1') of
//     one word, so that   acc0 += u . (k0, k0')  and  acc1 += u . (k1, k1')   are one IDP.2A.LO and one IDP.2A.HI on the FMA-heavy pipe (full
//     rate, tools/exp/idp_bench.cu) with both operands used exactly as loaded: 64 IDP per pass replace 64 IMAD + 128 ALU-pipe unpacks
//     (PRMT / SHF) of the int16 spectra; the limbs are recombined inside the Barrett step's input (acc0 + 64 acc1 < 2^26).
// Shared-memory rows are 16 words without padding; the 16-byte chunk index is XOR-swizzled with bits 1..2 of the row so that the 128-bit
// row reads of a quarter-warp hit distinct banks.
constexpr int S2_BI = 126, S2_BO = 131, S2_NIN = 4, S2_NOUT = 10;
#ifndef KOSK_S2_WARPS
#define KOSK_S2_WARPS 7
#endif
#ifndef KOSK_S2_MINB
#define KOSK_S2_MINB 4
#endif
#ifndef KOSK_S2_DUAL
#define KOSK_S2_DUAL 1     // two output blocks per half-warp and inverse iteration (0.92 -> 0.89 ms per 1024 proofs)
#endif
constexpr int S2_WARPS = KOSK_S2_WARPS;
static_assert(S2_BI + S2_BO - 1 == 256 && S2_NIN * S2_BI >= D1 && S2_NOUT * S2_BO >= NX && S2_NIN == 4 && S2_NOUT % 2 == 0, "share_ntt2 blocking");
__host__ __device__ constexpr int s2_word(int k1, int k2) { return k1 * 16 + ((((k2 >> 2) ^ (k1 >> 1)) & 3) << 2) + (k2 & 3); }   // word of bin (k1, k2) in a swizzled [16][16] tile

template <int NINV, int NOUTV>
__global__ void __launch_bounds__(32 * S2_WARPS, KOSK_S2_MINB) k_share_ntt2(const ConvArgs g)
{
    __shared__ __align__(16) uint32_t s_kp[S2_NOUT * 2 * 256];             // [o][pair][k1][k2 swizzled]: bytes (k0, k0', k1, k1') of segments (o, 2 pair), (o, 2 pair + 1)
    __shared__ __align__(8) int2 s_tw[2 * 256];
    __shared__ __align__(16) uint32_t s_up[S2_WARPS][2 * 256];             // [pair][k1][k2 swizzled]: spectra of input blocks 2 pair (low half) and 2 pair + 1
    __shared__ __align__(4) int16_t s_t[S2_WARPS][2][288];
    for (int i = threadIdx.x; i < S2_NOUT * 2 * 256; i += blockDim.x) s_kp[i] = g.kpk[i];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tw[i] = g.tw[i];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, c = lane & 15;
    uint32_t *up = s_up[wid];
    int16_t *T = s_t[wid][hw];
    // the buffer is 4-byte aligned and a row starts at 36 c bytes, so the 16 halfword loads of a row pair up into 8 LDS.32
    auto t_store = [&](const int32_t (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 16; j++) T[j * SN_TS + c] = (int16_t)v[j];
    };
    auto t_load = [&](int32_t (&in)[16]) {
#pragma unroll
        for (int b2 = 0; b2 < 16; b2++) in[b2] = T[c * SN_TS + b2];
    };
    const int sw = (c >> 1) & 3;                                            // chunk ch of row c lives at chunk ch ^ sw
    // Rows are handed out dynamically (one ticket per row): with a static stride the warps of the single wave drifted apart and the SMs
    // idled 18 % of the warp slots at the tail (ncu: 23.1 of 28 warps active on average).
    const int nw = gridDim.x * S2_WARPS;
    int m = blockIdx.x * S2_WARPS + wid;
    while (m < g.mtotal) {
        int mn = m + nw;
        if (g.ctr) {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(g.ctr, 1u) - g.ctr_base;
            mn = nw + (int)__shfl_sync(0xffffffffu, t, 0);
        }
        const u16 *yrow = g.A + ((size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp) * g.lda;
        u16 *dst = g.C + ((size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp) * g.ldc + g.c_off;
        {   // the first touch of an input row is a DRAM miss that nothing hides (ncu: 16 % of the stall samples sat on the first uses of
            // yrow): ask for this warp's NEXT row now, a whole sharing ahead
            if (mn < g.mtotal && lane * 64 < (int)g.lda) {
                const u16 *nrow = g.A + ((size_t)(mn / g.rpp) * g.a_slots + g.slot_lo + mn % g.rpp) * g.lda + lane * 64;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow));
            }
        }
        u16 tl[(NT + 32) / 32];                                             // tail values: loaded here, stored after the forward passes
        if (g.tail)
#pragma unroll
            for (int i = 0; i < (NT + 32) / 32; i++) { const int t = lane + 32 * i; tl[i] = t <= NT ? yrow[g.tail_off + t] : (u16)0; }
        // ---- forward: u_j = w_j y_j, NTT of the zero-padded 126-wide input blocks (two per pass) ----
#pragma unroll 1
        for (int it = 0; it < S2_NIN / 2; it++) {
            const int blk = 2 * it + hw;
            int32_t x[16];
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const int jl = 16 * a + c, j = S2_BI * blk + jl;
                int32_t v = 0;
                if (a < 8 && (a < 7 || jl < S2_BI) && j < NINV) { const int2 p = __ldg(g.pre + j); v = sn_shoup((int32_t)yrow[j], p.x, p.y); }
                x[a] = v;
            }
            int32_t y[16];
            sn_dft16<false, SN_FFT_SMALL>(x, y);                           // rows 8..15 of the block are zero padding: x[8..15] = 0 folds away
#pragma unroll
            for (int k = 0; k < 16; k++) { const int2 t = s_tw[k * 16 + c]; y[k] = sn_shoup(y[k], t.x, t.y); }
            t_store(y);
            __syncwarp();
            int32_t in[16], X[16];
            t_load(in);
            sn_dft16<false, SN_FFT_SMALL>(in, X);
            int16_t *urow = reinterpret_cast<int16_t *>(up + it * 256 + c * 16) + hw;
#pragma unroll
            for (int k = 0; k < 16; k++) urow[2 * ((((k >> 2) ^ sw) << 2) + (k & 3))] = (int16_t)sn_barrett(X[k]);
            __syncwarp();
        }
        if (g.tail)
#pragma unroll
            for (int i = 0; i < (NT + 32) / 32; i++) { const int t = lane + 32 * i; if (t <= NT) dst[t - (NT + 1)] = tl[i]; }
        // ---- inverse ----
        // second DFT stage, output factors and stores of block o from the transposed values in T
        auto stage2_emit = [&](int o) {
            int32_t in[16], out[16];
            t_load(in);
            sn_dft16<true, SN_FFT_SMALL>(in, out);                          // only out[0..8] are used: the rest of the last level is dead code
            const int xb = S2_BO * o + c;                                   // first output of this lane; row a adds 16 a
            const int2 *postp = g.post + xb;
            u16 *dstp = dst + xb;
            // rows a < na are valid: row 8 only holds x' = 128..130, and the last block ends at NOUTV
            const int na = min(c < S2_BO - 128 ? 9 : 8, (NOUTV - xb + 15) >> 4);
#pragma unroll
            for (int a = 0; a < 9; a++)
                if (a < na) {
                    const int2 pf = __ldg(postp + 16 * a);
                    uint32_t r = (uint32_t)sn_shoup(out[a], pf.x, pf.y);    // in (-q/4, 5q/4): canonical with two unsigned minima, no predicates
                    r = min(r, r + Q);
                    r = min(r, r - Q);
                    dstp[16 * a] = (u16)r;
                }
        };
#if KOSK_S2_DUAL
        // Two output blocks per half-warp and iteration (blocks o and o + 2) share the loads of the input spectra and of the twiddle pairs: the kernel
        // is bound by the L1 / shared-memory data pipe (ncu: 92 %), and those two are 64 of the 128 shared-memory wavefronts of a single-block pass.
#pragma unroll 1
        for (int it = 0; it < S2_NOUT / 4; it++) {
            const int oa = 4 * it + hw, ob = oa + 2;
            int32_t Oa[16], Ob[16];
            {
                const uint4 *pu = reinterpret_cast<const uint4 *>(up + c * 16);
                const uint4 *pka = reinterpret_cast<const uint4 *>(s_kp + oa * 512 + c * 16), *pkb = reinterpret_cast<const uint4 *>(s_kp + ob * 512 + c * 16);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint4 u0 = pu[ch ^ sw], u1 = pu[64 + (ch ^ sw)];
                    const uint32_t uw0[4] = {u0.x, u0.y, u0.z, u0.w}, uw1[4] = {u1.x, u1.y, u1.z, u1.w};
                    {
                        const uint4 k0 = pka[ch ^ sw], k1 = pka[64 + (ch ^ sw)];
                        const uint32_t kw0[4] = {k0.x, k0.y, k0.z, k0.w}, kw1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            Oa[4 * ch + e] = __dp2a_hi((int)uw1[e], (int)kw1[e], __dp2a_hi((int)uw0[e], (int)kw0[e], 0)) * 64 + __dp2a_lo((int)uw1[e], (int)kw1[e], __dp2a_lo((int)uw0[e], (int)kw0[e], 0));
                    }
                    {
                        const uint4 k0 = pkb[ch ^ sw], k1 = pkb[64 + (ch ^ sw)];
                        const uint32_t kw0[4] = {k0.x, k0.y, k0.z, k0.w}, kw1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            Ob[4 * ch + e] = __dp2a_hi((int)uw1[e], (int)kw1[e], __dp2a_hi((int)uw0[e], (int)kw0[e], 0)) * 64 + __dp2a_lo((int)uw1[e], (int)kw1[e], __dp2a_lo((int)uw0[e], (int)kw0[e], 0));
                    }
                }
            }
            int32_t va[16], vb[16];
            sn_dft16<true, SN_FFT_BIG>(Oa, va);
            sn_dft16<true, SN_FFT_BIG>(Ob, vb);
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const int2 t = s_tw[256 + b * 16 + c];
                va[b] = sn_shoup(va[b], t.x, t.y);
                vb[b] = sn_shoup(vb[b], t.x, t.y);                          // waits in registers for the transposition buffer
            }
            t_store(va);
            __syncwarp();
            stage2_emit(oa);
            __syncwarp();
            t_store(vb);
            __syncwarp();
            stage2_emit(ob);
            __syncwarp();
        }
        constexpr int O_SINGLE = 4 * (S2_NOUT / 4);
#else
        constexpr int O_SINGLE = 0;
#endif
        // one output block per half-warp and pass
#pragma unroll 1
        for (int o = O_SINGLE + hw; o < S2_NOUT; o += 2) {
            int32_t O[16];
            {
                const uint4 *pu = reinterpret_cast<const uint4 *>(up + c * 16);
                const uint4 *pk = reinterpret_cast<const uint4 *>(s_kp + o * 512 + c * 16);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint4 u0 = pu[ch ^ sw], u1 = pu[64 + (ch ^ sw)], k0 = pk[ch ^ sw], k1 = pk[64 + (ch ^ sw)];
                    const uint32_t uw0[4] = {u0.x, u0.y, u0.z, u0.w}, uw1[4] = {u1.x, u1.y, u1.z, u1.w}, kw0[4] = {k0.x, k0.y, k0.z, k0.w}, kw1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int32_t lo = __dp2a_lo((int)uw1[e], (int)kw1[e], __dp2a_lo((int)uw0[e], (int)kw0[e], 0));
                        const int32_t hi = __dp2a_hi((int)uw1[e], (int)kw1[e], __dp2a_hi((int)uw0[e], (int)kw0[e], 0));
                        O[4 * ch + e] = hi * 64 + lo;                       // < 2^25.5, reduced inside the transform (SN_FFT_BIG)
                    }
                }
            }
            int32_t v[16];
            sn_dft16<true, SN_FFT_BIG>(O, v);
#pragma unroll
            for (int b = 0; b < 16; b++) { const int2 t = s_tw[256 + b * 16 + c]; v[b] = sn_shoup(v[b], t.x, t.y); }
            t_store(v);
            __syncwarp();
            stage2_emit(o);
            __syncwarp();
        }
        m = mn;
    }
}
static inline int share_ntt2_launch(const ConvArgs &g0, cudaStream_t st, SnTicket *tk = nullptr)
{
    ConvArgs g = g0;
    const int ctas = std::min((g.mtotal + S2_WARPS - 1) / S2_WARPS, 148 * KOSK_S2_MINB);
    g.ctr = nullptr; g.ctr_base = 0;
    if (tk && tk->ctr && ctas > 0) { g.ctr = tk->ctr; g.ctr_base = tk->base; tk->base += (unsigned)g.mtotal; }
    if (ctas > 0) k_share_ntt2<D1, NX><<<ctas, 32 * S2_WARPS, 0, st>>>(g);
    if (g.ctr && cudaPeekAtLastError() != cudaSuccess) tk->base -= (unsigned)g.mtotal;      // the launch never ran: no ticket was drawn
    return 1;
}

// device tables of all Toeplitz products of the KOSK path (one allocation per context)
struct ShareNttTables {
    const int2 *tw;
    const int16_t *kh_share;      // OFF = 407, (4, 11)
    const uint32_t *kp_share;     // OFF = 407, 126 x 131 blocks, packed limbs (k_share_ntt2)
    const int16_t *kh_m256;       // OFF = -256: segments delta in [-7, 3]; a product with (NIN, NOUT) starts at delta = -(NIN - 1)
    const int2 *wj;               // [512]   (w, w') of w_j over 407 consecutive nodes (0 beyond)
    const int2 *wj2;              // [896]   the same over 813 consecutive nodes
    const int2 *px;               // [1408]  P(x) of the sharing
    const int2 *pr1, *pr2;        // [256]   prod_m (i - 256 - m) for m < 407 / m < 813 (recon_secrets_ddeg / _2ddeg)
};
constexpr int SN_M256_DMIN = -7, SN_M256_DMAX = 3;
__host__ __device__ inline const int16_t *sn_kh_m256(const ShareNttTables &t, int nin) { return t.kh_m256 + (size_t)(-(nin - 1) - SN_M256_DMIN) * 16 * SN_LD; }
// the sharing of the rows described by a GemmArgs (row mapping, A / C, tail) as a Toeplitz product
static inline ConvArgs share_conv_args(const GemmArgs &g, const ShareNttTables &t)
{
    ConvArgs a{};
    a.A = g.A; a.C = g.C; a.lda = g.lda; a.ldc = g.ldc; a.mtotal = g.mtotal; a.rpp = g.rpp; a.slot_lo = g.slot_lo; a.a_slots = g.a_slots; a.c_slots = g.c_slots;
    a.c_off = g.c_off; a.tail = g.tail; a.tail_off = g.tail_off;
    a.tw = t.tw; a.khat = t.kh_share; a.kpk = t.kp_share; a.pre = t.wj; a.post = t.px; a.post_group = 0;
    return a;
}
// variant 2 (default) = k_share_ntt2, variant 1 = the generic equal-block kernel
static inline int share_ntt_launch(const ConvArgs &g, cudaStream_t st, int variant = 2, SnTicket *tk = nullptr)
{
    return variant >= 2 ? share_ntt2_launch(g, st, tk) : conv_ntt_launch<SN_NIN, SN_NOUT, D1, NX, true, false>(g, st, tk);
}

// ---- host: table construction (plain modular arithmetic, once per context) ----
struct ShareNttHost {
    std::vector<int32_t> w16f, w16i;
    std::vector<int2> tw, wj, wj2, px, pr1, pr2;
    std::vector<int16_t> kh_share, kh_m256;
    std::vector<uint32_t> kp_share;
};
static inline ShareNttHost share_ntt_tables()
{
    auto pw = [](uint32_t b, uint32_t e) { uint32_t r = 1; b %= Q; while (e) { if (e & 1) r = r * b % Q; b = b * b % Q; e >>= 1; } return r; };
    auto inv = [&](uint32_t a) { return pw(a % Q, Q - 2); };
    const uint32_t om = 17, iom = inv(17), om16 = pw(om, 16), iom16 = inv(om16);
    ShareNttHost h;
    h.w16f.resize(256); h.w16i.resize(256); h.tw.resize(512);
    for (int a = 0; a < 16; a++)
        for (int k = 0; k < 16; k++) {
            h.w16f[a * 16 + k] = gf_center(pw(om16, a * k)); h.w16i[a * 16 + k] = gf_center(pw(iom16, a * k));
            h.tw[a * 16 + k] = sn_pair(pw(om, a * k)); h.tw[256 + a * 16 + k] = sn_pair(pw(iom, a * k));
        }
    // barycentric weights over n consecutive nodes: w_j = 1 / prod_{m != j} (j - m)
    auto weights = [&](int n, int padded) {
        std::vector<int2> w(padded, make_int2(0, 0));
        for (int j = 0; j < n; j++) {
            uint32_t d = 1;
            for (int m = 0; m < n; m++) if (m != j) d = d * (uint32_t)(((j - m) % Q + Q) % Q) % Q;
            w[j] = sn_pair(inv(d));
        }
        return w;
    };
    h.wj = weights(D1, 512); h.wj2 = weights(D2, 896);
    // P at the targets: the sharing evaluates at x + 407 over the nodes 0..406; the reconstructions at i over the nodes 256..256+n-1
    h.px.assign(1408, make_int2(0, 0)); h.pr1.assign(256, make_int2(0, 0)); h.pr2.assign(256, make_int2(0, 0));
    for (int x = 0; x < NX; x++) {
        uint32_t p = 1;
        for (int m = 0; m < D1; m++) p = p * (uint32_t)((x + D1 - m) % Q) % Q;
        h.px[x] = sn_pair(p);
    }
    for (int i = 0; i < NL; i++) {
        uint32_t p1 = 1, p2 = 1;
        for (int m = 0; m < D2; m++) { const uint32_t f = (uint32_t)(((i - 256 - m) % Q + Q) % Q); p2 = p2 * f % Q; if (m < D1) p1 = p1 * f % Q; }
        h.pr1[i] = sn_pair(p1); h.pr2[i] = sn_pair(p2);
    }
    // kernel segments K_delta[d] = c[128 delta + OFF + d], d in [-127, 127], c[m] = 1/m (0 when m = 0 mod q), and their 256-point
    // NTTs, index k = k1 + 16 k2 stored at [k1][k2], scaled by 1/256
    std::vector<uint32_t> opw(256);
    for (int i = 0; i < 256; i++) opw[i] = pw(om, i);
    const uint32_t i256 = inv(256);
    auto segments = [&](int off, int dmin, int dmax) {
        std::vector<int16_t> out((size_t)(dmax - dmin + 1) * 16 * SN_LD, 0);
        for (int dl = dmin; dl <= dmax; dl++) {
            uint32_t K[256];
            for (int t = 0; t < 256; t++) {
                const int d = t < 128 ? t : t - 256;
                const int mm = ((128 * dl + off + d) % Q + Q) % Q;
                K[t] = (t == 128 || mm == 0) ? 0 : inv((uint32_t)mm);
            }
            for (int k = 0; k < 256; k++) {
                uint32_t s = 0;
                for (int t = 0; t < 256; t++) s = (s + K[t] * opw[(t * k) & 255]) % Q;
                out[((size_t)(dl - dmin) * 16 + (k & 15)) * SN_LD + (k >> 4)] = (int16_t)gf_center(s * i256 % Q);
            }
        }
        return out;
    };
    h.kh_share = segments(D1, -(SN_NIN - 1), SN_NOUT - 1);
    h.kh_m256 = segments(-256, SN_M256_DMIN, SN_M256_DMAX);
    // k_share_ntt2: segment (o, i) K[d] = c[131 o - 126 i + 407 + d], d in [-125, 130] at t = d mod 256; spectrum / 256, centered, as limbs k = 64 k1 + k0
    h.kp_share.assign((size_t)S2_NOUT * 2 * 256, 0);
    for (int o = 0; o < S2_NOUT; o++)
        for (int i = 0; i < S2_NIN; i++) {
            uint32_t K[256];
            for (int t = 0; t < 256; t++) {
                const int d = t < S2_BO ? t : t - 256;
                const int mm = ((S2_BO * o - S2_BI * i + D1 + d) % Q + Q) % Q;
                K[t] = mm == 0 ? 0 : inv((uint32_t)mm);
            }
            for (int k = 0; k < 256; k++) {
                uint32_t s = 0;
                for (int t = 0; t < 256; t++) s = (s + K[t] * opw[(t * k) & 255]) % Q;
                const int v = gf_center(s * i256 % Q), k0 = ((v + 32) & 63) - 32, k1 = (v - k0) / 64;
                uint32_t &w = h.kp_share[((size_t)o * 2 + i / 2) * 256 + s2_word(k & 15, k >> 4)];
                w |= ((uint32_t)(uint8_t)(int8_t)k0) << (8 * (i & 1)) | ((uint32_t)(uint8_t)(int8_t)k1) << (16 + 8 * (i & 1));
            }
        }
    return h;
}

}  // namespace kosk
