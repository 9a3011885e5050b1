// verify_kernels.cuh -- sm_100a kernels of the KOSK verifier (kyber_kosk_verify, reference kosk.cpp:88-117 and
// verify(), mlwe_verifier.cpp:4-686).  One launch handles a chunk of B independent proofs.  Every failed check
// ORs a bit into flags[b]; the proof is accepted iff flags[b] == 0 (the reference returns false at the first
// failing check, so only the conjunction matters).  The reference's NTL interpolate/eval over the rest-party
// nodes (mlwe_verifier.cpp:188-224, :326-352, :397-443, :510-543) becomes one Lagrange matrix per proof,
// applied with the same integer-pipe GEMM as the prover's share evaluation.
#pragma once
#include "prove_kernels.cuh"
#include "share_ntt.cuh"
#include <cuda_runtime.h>

namespace kosk {

constexpr int VR2LD = 816;        // row stride of 813-term rows (16B-aligned rows, 51 k-steps)
constexpr int LM1_ROWS = 512;     // 407 targets padded to the GEMM's column tiles (stride of PT1)
// Interpolation operands indexed by PARTY (not by rest-set position): the first 407 (813) rest parties lie among parties
// 0..556 (0..962) because at most 150 parties are opened, so one fixed Cauchy matrix U[t][p] = 1 / (t - (p + 256)) serves every proof.
constexpr int KP1 = 560, KP2 = 976;          // party columns of the two operands (multiples of the 16-term k-step)
constexpr int U1_ROWS = 448;                 // 407 targets in four 112-column tiles
constexpr int OPLD = 160;         // opened-party values: beta[70] gamma[70] r[2k] NTT_r[2k]

enum VFlag { VF_I = 1, VF_BG = 2, VF_SR = 4, VF_NTT = 8, VF_ASR = 16, VF_T = 32, VF_TREL = 64, VF_ETA = 128,
             VF_SUBETA = 256, VF_UZ = 512, VF_U2D = 1024, VF_FS2 = 2048, VF_STRICT = 4096 };

struct VerifyTables { const int16_t *St, *R1, *R2; const u16 *inv; const int16_t *SU; const u16 *fact; /* [2][FACT_N]: i!, 1/i! */
                      const int16_t *U1, *U2; /* Cauchy operands [U1_ROWS][KP1], [256][KP2], centered */
                      const ShareNttTables *sn; /* share evaluation as an NTT convolution (share_ntt.cuh); nullptr = dense table GEMM */
                      int sn_variant = 2; /* share_ntt_launch variant */
                      SnTicket *tk_main = nullptr, *tk_side = nullptr; /* work tickets of the main / side stream (share_ntt.cuh) */ };

struct VerifyBufs {
    int *flags = nullptr;
    u16 *I = nullptr, *REST = nullptr, *I2 = nullptr, *REST2 = nullptr; int16_t *POS = nullptr;
    u16 *AH = nullptr, *TPK = nullptr, *PW = nullptr;
    u8 *TCR = nullptr, *VWR = nullptr;
    u16 *CR = nullptr, *VR = nullptr, *OPV = nullptr;
    u16 *ABG = nullptr, *BS = nullptr, *A1 = nullptr, *YV = nullptr, *A2 = nullptr, *UZ = nullptr, *VSH = nullptr, *U2 = nullptr, *UR = nullptr;
    u16 *W1 = nullptr, *W2 = nullptr, *PT1 = nullptr, *PT2 = nullptr;   // barycentric weights per node / P(t) per target
    int32_t *WS = nullptr;                                              // split-K partial sums (latency mode of the GEMMs)
    int k = 0, chunk = 0;
    int strict = 0;   // hardened decoding: off by default = the reference's accept set
    int raw_inst = 0; // verify() on a caller-supplied mlwe_inst (raw_api.cuh): AH / TPK are preloaded, pk is not parsed
    int pt_mont = 0;  // P(t) tables hold (w, w') int2 pairs for the NTT-convolution path (share_ntt.cuh, sn_shoup); plain u16 residues for the dense GEMM path
};

struct VDims {
    int k, eta, E, M, F, NA, nc, nv, crld, vrld, n1rows, nyrows, n2rows;
};
KOSK_HD VDims make_vdims(int k)
{
    VDims d; d.k = k; d.eta = (k == 2) ? 3 : 2; d.E = 2 * d.eta + 1; d.M = 2 * d.eta; d.F = MK + 2 * k + 1; d.NA = MK + 2 * k;
    d.nc = 2 * (k + d.F); d.nv = 16 + d.nc + 4 * k + 4 * k * d.M;
    d.crld = (d.nc + 3) & ~3; d.vrld = (d.nv + 3) & ~3;
    d.n1rows = k * (3 + 2 * d.E); d.nyrows = d.n1rows + 3 * k; d.n2rows = 2 * k * d.M;
    return d;
}

static inline void verify_free(VerifyBufs &v)
{
    void *p[] = {v.flags, v.I, v.REST, v.I2, v.REST2, v.POS, v.AH, v.TPK, v.PW, v.TCR, v.VWR, v.CR, v.VR, v.OPV,
                 v.ABG, v.BS, v.A1, v.YV, v.A2, v.UZ, v.VSH, v.U2, v.UR, v.W1, v.W2, v.PT1, v.PT2, v.WS};
    for (void *q : p) if (q) cudaFree(q);
    v = VerifyBufs{};
}
static inline int verify_alloc(VerifyBufs &v, int k, int chunk)
{
    const VDims d = make_vdims(k); const size_t B = (size_t)chunk;
    v.k = k; v.chunk = chunk;
#define VA(ptr, bytes, zero) do { if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) return -1; if (zero) cudaMemset((ptr), 0, (bytes)); } while (0)
    VA(v.flags, B * 4, 1); VA(v.I, B * NT * 2, 0); VA(v.REST, B * NR * 2, 0); VA(v.I2, B * NT * 2, 0); VA(v.REST2, B * NR * 2, 0); VA(v.POS, B * NP * 2, 0);
    VA(v.AH, B * k * k * 256 * 2, 0); VA(v.TPK, B * k * 256 * 2, 0); VA(v.PW, B * d.NA * d.F * 2, 0);
    VA(v.TCR, B * TREE_BYTES, 0); VA(v.VWR, B * TREE_BYTES, 0);
    VA(v.CR, B * NT * d.crld * 2, 1); VA(v.VR, B * NT * d.vrld * 2, 1); VA(v.OPV, B * NT * OPLD * 2, 1);
    VA(v.ABG, B * 2 * MK * YLD * 2, 1); VA(v.BS, B * 2 * MK * 256 * 2, 0);
    VA(v.A1, B * d.n1rows * KP1 * 2, 1); VA(v.YV, B * d.nyrows * YLD * 2, 1);
    VA(v.A2, B * d.n2rows * KP2 * 2, 1); VA(v.UZ, B * d.n2rows * 256 * 2, 0);
    VA(v.VSH, B * d.nyrows * SLD * 2, 1); VA(v.U2, B * d.n2rows * VR2LD * 2, 1); VA(v.UR, B * d.n2rows * 256 * 2, 0);
    VA(v.WS, (size_t)GE_WS_ELEMS * 4, 0);
    VA(v.W1, B * YLD * 2, 1); VA(v.W2, B * VR2LD * 2, 1); VA(v.PT1, B * LM1_ROWS * 8, 1); VA(v.PT2, B * 256 * 8, 1);      // sized for int2 entries
#undef VA
    return 0;
}

__device__ __forceinline__ u16 pi16(const u8 *pi, size_t off, size_t idx) { return reinterpret_cast<const u16 *>(pi + off)[idx]; }

// V1 (mlwe_verifier.cpp:9-19) + kosk.cpp:94-112: opened/rest sets, A-hat from the pk seed, t from pk, and
// the rest parties' commitment / view digests copied into the per-party digest rows (:36-38, :645-647).
template <int K>
__global__ void __launch_bounds__(128) kv_setup(VerifyBufs vb, const u8 *__restrict__ pis, const u8 *__restrict__ pks)
{
    const Layout L = make_layout(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    const u8 *pi = pis + L.proof_bytes * (size_t)b, *pk = pks + L.pk_bytes * (size_t)b;
    __shared__ int cnt[NP];
    __shared__ int bad, base[129];
    __shared__ u16 sI[NT];
    for (int p = tid; p < NP; p += 128) cnt[p] = 0;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int i = tid; i < NT; i += 128) {
        u16 v = pi16(pi, L.o_I, i); sI[i] = v;
        if (v >= NP) atomicOr(&bad, 1); else if (atomicAdd(&cnt[v], 1) != 0) atomicOr(&bad, 1);
    }
    __syncthreads();
    if (bad) {   // malformed I is UB in the reference (mlwe_verifier.cpp:12); rejected here, with a sane set for the remaining kernels
        for (int p = tid; p < NP; p += 128) cnt[p] = p < NT ? 1 : 0;
        for (int i = tid; i < NT; i += 128) sI[i] = (u16)i;
        if (tid == 0) atomicOr(&vb.flags[b], VF_I);
    }
    __syncthreads();
    for (int i = tid; i < NT; i += 128) { vb.I[(size_t)b * NT + i] = sI[i]; vb.POS[(size_t)b * NP + sI[i]] = (int16_t)(-1 - i); }
    constexpr int PER = 12;   // 128 * 12 >= 1454
    int mine = 0;
    for (int p = tid * PER; p < min(NP, tid * PER + PER); p++) mine += cnt[p] == 0;
    base[tid + 1] = mine;
    if (tid == 0) base[0] = 0;
    __syncthreads();
    if (tid == 0) for (int i = 1; i <= 128; i++) base[i] += base[i - 1];
    __syncthreads();
    { int j = base[tid];
      for (int p = tid * PER; p < min(NP, tid * PER + PER); p++) if (cnt[p] == 0) { vb.REST[(size_t)b * NR + j] = (u16)p; vb.POS[(size_t)b * NP + p] = (int16_t)j; j++; } }
    // t from pk: polyvec_frombytes (kyber/poly.c:151-158), raw 12-bit values
    if (!vb.raw_inst) for (int i = 0; i < K; i++) {
        const u8 *a = pk + 384 * i + 3 * tid;
        vb.TPK[((size_t)b * K + i) * 256 + 2 * tid] = (u16)((a[0] | ((u16)a[1] << 8)) & 0xFFF);
        vb.TPK[((size_t)b * K + i) * 256 + 2 * tid + 1] = (u16)(((a[1] >> 4) | ((u16)a[2] << 4)) & 0xFFF);
    }
    if (tid < K * K && !vb.raw_inst) {      // gen_matrix (indcpa.c:168-193): A[i][j] <- SHAKE128(seed || j || i)
        __shared__ uint64_t sBlk[K * K][21];
        uint64_t sd[4];
        for (int w = 0; w < 4; w++) { uint64_t v = 0; for (int q = 0; q < 8; q++) v |= (uint64_t)pk[384 * K + 8 * w + q] << (8 * q); sd[w] = v; }
        xof_rej_uniform(vb.AH + ((size_t)b * K * K + tid) * 256, sd, (uint32_t)(tid % K), (uint32_t)(tid / K), sBlk[tid]);
    }
}

// Second half of the setup, spread over several CTAs per proof (latency of a single verification): the rest parties' commitment /
// view digests copied into the per-party digest rows (mlwe_verifier.cpp:36-38, :645-647) and the commit records of the opened
// parties (:23-33), row-major.  grid (SETUP_COPY_CTAS, B).
constexpr int SETUP_COPY_CTAS = 8;
template <int K>
__global__ void __launch_bounds__(128) kv_setup_copy(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const VDims d = make_vdims(K);
    const int b = blockIdx.y, t0 = blockIdx.x * 128 + threadIdx.x, stride = gridDim.x * 128;
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    for (int idx = t0; idx < NR * 8; idx += stride) {
        const int j = idx / 8, w = idx % 8, p = vb.REST[(size_t)b * NR + j];
        reinterpret_cast<uint32_t *>(vb.TCR + ((size_t)b * NP + p) * 32)[w] = reinterpret_cast<const uint32_t *>(pi + L.o_Tcomm)[(size_t)j * 8 + w];
        reinterpret_cast<uint32_t *>(vb.VWR + ((size_t)b * NP + p) * 32)[w] = reinterpret_cast<const uint32_t *>(pi + L.o_comm)[(size_t)j * 8 + w];
    }
    for (int idx = t0; idx < NT * d.nc; idx += stride) {
        const int i = idx / d.nc, v = idx % d.nc;
        u16 x;
        if (v < K) x = pi16(pi, L.o_s, i * K + v);
        else if (v < 2 * K) x = pi16(pi, L.o_e, i * K + v - K);
        else if (v < 2 * K + d.F) x = pi16(pi, L.o_f, i * d.F + v - 2 * K);
        else x = pi16(pi, L.o_Tf, i * d.F + v - 2 * K - d.F);
        vb.CR[((size_t)b * NT + i) * d.crld + v] = x;
    }
}

// V4 + V8 (mlwe_verifier.cpp:67-89, :148-170): beta/gamma/r/NTT_r of the opened parties.
// Exact reference semantics: the sum starts from the raw (possibly non-canonical) share f[c0] and every step is
// gf3329_add(acc, gf3329_mul(pow, f)); with a canonical first term this equals plain mod-q arithmetic, which is what the
// fast path computes (shares held as centered residues in registers, one IMAD per term, one reduction per output).
template <int K>
__global__ void __launch_bounds__(320) kv_eval_opened(VerifyBufs vb)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K, FP = (F + 3) & ~3;
    const VDims d = make_vdims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ __align__(16) int32_t spw[NA][FP];          // centered powers, row j = challenge j
    for (int i = tid; i < NA * FP; i += 320) {
        const int j = i / FP, kk = i % FP;
        spw[j][kk] = kk < F ? gf_center(vb.PW[(size_t)b * NA * F + j * F + kk]) : 0;
    }
    __syncthreads();
    if (tid >= 2 * NT) return;
    const int i = tid >> 1, half = tid & 1;
    const u16 *f = vb.CR + ((size_t)b * NT + i) * d.crld + 2 * K + half * F;
    u16 *out = vb.OPV + ((size_t)b * NT + i) * OPLD;
    int32_t fr[FP];
#pragma unroll
    for (int kk = 0; kk < FP; kk++) fr[kk] = kk < F ? gf_center(f[kk] % (uint32_t)Q) : 0;
    const bool canon0 = f[0] < Q, canon71 = f[MK + 1] < Q;
    const int32_t c0 = fr[0], c71 = fr[MK + 1];
    auto put = [&](int j, u16 res) { if (j < MK) out[half * MK + j] = res; else out[2 * MK + half * 2 * K + (j - MK)] = res; };
    static_assert(NA % 2 == 0 && MK % 2 == 0, "challenge pairs must not straddle the beta / r boundary");
#pragma unroll 1
    for (int j = 0; j < NA; j += 2) {                           // two challenges per iteration: two independent IMAD chains per thread
        const bool fast = j < MK ? canon0 : canon71;
        if (fast) {
            int32_t acc0 = 0, acc1 = 0;                          // 78 * 1664^2 < 2^31; the k = 0 term is replaced by the c0 share below
#pragma unroll
            for (int k4 = 0; k4 < FP / 4; k4++) {
                const int4 w0 = *reinterpret_cast<const int4 *>(&spw[j][4 * k4]), w1 = *reinterpret_cast<const int4 *>(&spw[j + 1][4 * k4]);
                acc0 += fr[4 * k4] * w0.x + fr[4 * k4 + 1] * w0.y + fr[4 * k4 + 2] * w0.z + fr[4 * k4 + 3] * w0.w;
                acc1 += fr[4 * k4] * w1.x + fr[4 * k4 + 1] * w1.y + fr[4 * k4 + 2] * w1.z + fr[4 * k4 + 3] * w1.w;
            }
            const int32_t cc = j < MK ? c0 : c71;
            put(j, (u16)gf_canon(acc0 + cc - fr[0] * spw[j][0]));
            put(j + 1, (u16)gf_canon(acc1 + cc - fr[0] * spw[j + 1][0]));
        } else {                                                 // non-canonical first term: the reference's u16 chain, verbatim
            for (int jj = j; jj < j + 2; jj++) {
                u16 acc = f[jj < MK ? 0 : MK + 1];
                for (int kk = 1; kk < F; kk++) acc = ref_add(acc, (u16)(((uint32_t)gf_canon(spw[jj][kk]) * f[kk]) % (uint32_t)Q));
                put(jj, acc);
            }
        }
    }
}

// Batch form of kv_eval_opened on the dot-product instruction: thread = two records (one opened party: f and NTT_f), held as packed int16 pairs
// (f[k], f[k+1]); the power table as signed limbs w = 64 w1 + w0 packed (w0[k], w0[k+1], w1[k], w1[k+1]) per word, so that one broadcast LDS.128
// (eight powers of one challenge) feeds sixteen IDP.2A -- four times less shared-memory traffic per multiply-add than kv_eval_opened, which reads one
// LDS.128 per four IMADs and is bound by the shared-memory pipe (32 % of the IMAD peak).  The k = 0 entries of the table are zero: the constant term
// is added explicitly (f[0] for beta / gamma, f[71] for the r columns: the c0 quirk).
template <int K>
__global__ void __launch_bounds__(160, 4) kv_eval_opened_idp(VerifyBufs vb)
{
    constexpr int F = MK + 2 * K + 1, NA = MK + 2 * K, FW = (F + 1) / 2, FWP = (FW + 3) & ~3;      // words per challenge row, padded to LDS.128
    const VDims d = make_vdims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ __align__(16) uint32_t spk[NA][FWP];
    for (int i = tid; i < NA * FWP; i += 160) {
        const int j = i / FWP, wd = i % FWP;
        uint32_t word = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int kk = 2 * wd + h;
            if (kk >= 1 && kk < F) {
                const int v = gf_center(vb.PW[(size_t)b * NA * F + j * F + kk]), w0 = ((v + 32) & 63) - 32, w1 = (v - w0) / 64;
                word |= ((uint32_t)(uint8_t)(int8_t)w0) << (8 * h) | ((uint32_t)(uint8_t)(int8_t)w1) << (16 + 8 * h);
            }
        }
        spk[j][wd] = word;
    }
    __syncthreads();
    if (tid >= NT) return;
    const int i = tid;                                          // opened party: rows f (half 0) and NTT_f (half 1)
    const u16 *f0 = vb.CR + ((size_t)b * NT + i) * d.crld + 2 * K, *f1 = f0 + F;
    u16 *out = vb.OPV + ((size_t)b * NT + i) * OPLD;
    uint32_t p0[FWP], p1[FWP];                                  // packed centered pairs (f[2w], f[2w+1])
#pragma unroll
    for (int wd = 0; wd < FWP; wd++) {
        uint32_t a = 0, c = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int kk = 2 * wd + h;
            if (kk < F) {
                a |= ((uint32_t)gf_center(f0[kk] % (uint32_t)Q) & 0xFFFFu) << (16 * h);
                c |= ((uint32_t)gf_center(f1[kk] % (uint32_t)Q) & 0xFFFFu) << (16 * h);
            }
        }
        p0[wd] = a; p1[wd] = c;
    }
    const u16 r00 = f0[0], r071 = f0[MK + 1], r10 = f1[0], r171 = f1[MK + 1];
    const bool can00 = r00 < Q, can071 = r071 < Q, can10 = r10 < Q, can171 = r171 < Q;
    const int32_t c00 = gf_center(r00 % (uint32_t)Q), c071 = gf_center(r071 % (uint32_t)Q), c10 = gf_center(r10 % (uint32_t)Q), c171 = gf_center(r171 % (uint32_t)Q);
    auto put = [&](int half, int j, u16 res) { if (j < MK) out[half * MK + j] = res; else out[2 * MK + half * 2 * K + (j - MK)] = res; };
    auto slow = [&](const u16 *f, int j) -> u16 {              // non-canonical first term: the reference's u16 chain, verbatim
        u16 acc = f[j < MK ? 0 : MK + 1];
        for (int kk = 1; kk < F; kk++) {
            const uint32_t word = spk[j][kk >> 1];
            const int w = 64 * (int)(int8_t)(word >> (16 + 8 * (kk & 1))) + (int)(int8_t)(word >> (8 * (kk & 1)));
            acc = ref_add(acc, (u16)(((uint32_t)gf_canon(w) * f[kk]) % (uint32_t)Q));
        }
        return acc;
    };
#pragma unroll 1
    for (int j = 0; j < NA; j++) {
        int32_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
#pragma unroll
        for (int w4 = 0; w4 < FWP / 4; w4++) {
            const uint4 w = *reinterpret_cast<const uint4 *>(&spk[j][4 * w4]);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                lo0 = __dp2a_lo((int)p0[4 * w4 + e], (int)ww[e], lo0); hi0 = __dp2a_hi((int)p0[4 * w4 + e], (int)ww[e], hi0);
                lo1 = __dp2a_lo((int)p1[4 * w4 + e], (int)ww[e], lo1); hi1 = __dp2a_hi((int)p1[4 * w4 + e], (int)ww[e], hi1);
            }
        }
        const bool beta = j < MK;
        put(0, j, (beta ? can00 : can071) ? (u16)gf_canon(lo0 + 64 * hi0 + (beta ? c00 : c071)) : slow(f0, j));
        put(1, j, (beta ? can10 : can171) ? (u16)gf_canon(lo1 + 64 * hi1 + (beta ? c10 : c171)) : slow(f1, j));
    }
}

// Row gathers for the table / interpolation contractions (values reduced mod q exactly where the reference
// reduces them: gf3329_mul in recon_* and the NTL ZZ_p assignment in the interpolation inputs).
//   ABG[j*2+w][p]  p<407: beta/gamma share of party p      (mlwe_verifier.cpp:97-108)
//   A1 [row][j]    j<407: sr, er, t, s_eta, e_eta shares of rest party j, times the barycentric weight w_j  (:178-186, :320-324, :390-395)
//   A2 [row][j]    j<813: u_s, u_e shares of rest party j, times w_j  (:503-508)
template <int K>
__global__ void __launch_bounds__(128) kv_gather(VerifyBufs vb, const u8 *__restrict__ pis, int row0)
{
    const Layout L = make_layout(K);
    const VDims d = make_vdims(K);
    const int b = blockIdx.y, row = blockIdx.x + row0, tid = threadIdx.x;
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    if (row < 2 * MK) {
        const int j = row >> 1, w = row & 1;
        u16 *dst = vb.ABG + ((size_t)b * 2 * MK + row) * YLD;
        for (int p = tid; p < D1; p += 128) {
            const int pos = vb.POS[(size_t)b * NP + p];
            const u16 v = pos >= 0 ? pi16(pi, w ? L.o_gamma : L.o_beta, (size_t)pos * MK + j) : vb.OPV[((size_t)b * NT + (-1 - pos)) * OPLD + w * MK + j];
            dst[p] = (u16)(v % (uint32_t)Q);
        }
    } else if (row < 2 * MK + d.n1rows) {
        const int r = row - 2 * MK;
        size_t off; int mul, add;
        if (r < K) { off = L.o_sr; mul = K; add = r; }
        else if (r < 2 * K) { off = L.o_er; mul = K; add = r - K; }
        else if (r < 3 * K) { off = L.o_t; mul = K; add = r - 2 * K; }
        else if (r < 3 * K + K * d.E) { off = L.o_seta; mul = K * d.E; add = r - 3 * K; }
        else { off = L.o_eeta; mul = K * d.E; add = r - 3 * K - K * d.E; }
        // operand row indexed by party: share * barycentric weight at the first 407 rest parties, 0 at opened / later parties
        u16 *dst = vb.A1 + ((size_t)b * d.n1rows + r) * KP1;
        const u16 *w1 = vb.W1 + (size_t)b * YLD;
        const int16_t *pos = vb.POS + (size_t)b * NP;
        for (int p = tid; p < KP1; p += 128) {
            const int j = pos[p];
            dst[p] = (j >= 0 && j < D1) ? (u16)gf_mul(pi16(pi, off, (size_t)j * mul + add) % (uint32_t)Q, w1[j]) : (u16)0;
        }
    } else {
        const int r = row - 2 * MK - d.n1rows;              // r = w*K*M + i*M + m
        const int w = r / (K * d.M), im = r % (K * d.M);
        u16 *dst = vb.A2 + ((size_t)b * d.n2rows + r) * KP2;
        const u16 *w2 = vb.W2 + (size_t)b * VR2LD;
        const int16_t *pos = vb.POS + (size_t)b * NP;
        for (int p = tid; p < KP2; p += 128) {
            const int j = pos[p];
            dst[p] = (j >= 0 && j < D2) ? (u16)gf_mul(pi16(pi, w ? L.o_ue : L.o_us, (size_t)j * K * d.M + im) % (uint32_t)Q, w2[j]) : (u16)0;
        }
    }
}

// ABG rows (beta / gamma shares of parties 0..406, one row per challenge column) through a shared-memory transpose: the proof holds
// beta / gamma party-major ([rest party][70], 140-byte rows), the reconstruction wants them column-major.  A per-element gather costs
// one 32-byte sector per value (kv_gather's first branch: 0.25 ms per 1024 proofs); here a CTA reads the rows of 64 consecutive
// parties as words and writes 128-byte segments of the 140 output rows.  grid (ceil(407 / 64), B).
constexpr int VG_TP = 64;
template <int K>
__global__ void __launch_bounds__(256) kv_gather_bg(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const int b = blockIdx.y, p0 = blockIdx.x * VG_TP, tid = threadIdx.x, np = min(VG_TP, D1 - p0);
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    __shared__ uint32_t st[VG_TP][MK + 1];           // [party][35 words of beta | 35 words of gamma], odd stride
    const uint32_t *__restrict__ beta = reinterpret_cast<const uint32_t *>(pi + L.o_beta), *__restrict__ gamma = reinterpret_cast<const uint32_t *>(pi + L.o_gamma);
    for (int idx = tid; idx < np * MK; idx += 256) {
        const int r = idx / MK, w = idx % MK, h = w >= MK / 2, ww = h ? w - MK / 2 : w;
        const int pos = vb.POS[(size_t)b * NP + p0 + r];
        uint32_t v;
        if (pos >= 0) v = __ldg((h ? gamma : beta) + (size_t)pos * (MK / 2) + ww);
        else v = __ldg(reinterpret_cast<const uint32_t *>(vb.OPV + ((size_t)b * NT + (-1 - pos)) * OPLD + h * MK) + ww);
        st[r][w] = v;
    }
    __syncthreads();
    u16 *dst = vb.ABG + (size_t)b * 2 * MK * YLD + p0;
    for (int idx = tid; idx < 2 * MK * VG_TP; idx += 256) {
        const int row = idx / VG_TP, r = idx % VG_TP;      // row = j * 2 + w
        if (r >= np) continue;
        const int j = row >> 1, h = row & 1;
        const uint32_t wv = st[r][h * (MK / 2) + (j >> 1)];
        const uint32_t v = (j & 1) ? wv >> 16 : wv & 0xFFFFu;
        dst[(size_t)row * YLD + r] = (u16)(v % (uint32_t)Q);
    }
}

// Targets 256..406 of the d-degree interpolation are themselves nodes when party t - 256 is a rest party: the interpolant passes
// through the share, so the value is the share itself (the Cauchy operand has a zero there and P(t) = 0).  grid (n1rows, B).
template <int K>
__global__ void __launch_bounds__(160) kv_node_targets(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const VDims d = make_vdims(K);
    const int b = blockIdx.y, r = blockIdx.x, p = threadIdx.x;
    if (p > NT) return;
    const int j = vb.POS[(size_t)b * NP + p];
    if (j < 0) return;                                      // opened party: interpolated like any other target
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    size_t off; int mul, add;
    if (r < K) { off = L.o_sr; mul = K; add = r; }
    else if (r < 2 * K) { off = L.o_er; mul = K; add = r - K; }
    else if (r < 3 * K) { off = L.o_t; mul = K; add = r - 2 * K; }
    else if (r < 3 * K + K * d.E) { off = L.o_seta; mul = K * d.E; add = r - 3 * K; }
    else { off = L.o_eeta; mul = K * d.E; add = r - 3 * K - K * d.E; }
    vb.YV[((size_t)b * d.nyrows + r) * YLD + 256 + p] = (u16)(pi16(pi, off, (size_t)j * mul + add) % (uint32_t)Q);
}

// Lagrange interpolation over the rest-party nodes x_k = rest[k] + 256 in barycentric form,
//   l_k(t) = P(t) * w_k / (t - x_k):
// the GEMM operand is the fixed Cauchy matrix 1 / (t - x) (VerifyTables.U1 / U2, indexed by party); the weights w_k scale the
// gathered input rows (kv_gather) and P(t) scales the output columns (GemmArgs.colscale).  This kernel computes w_k and P(t) per proof.
//   d-degree:  t = 0..406 over x_0..x_406;  2d-degree: t = 0..255 over x_0..x_812.
// The nodes are the integers of [256, hi] minus at most 150 "holes" (opened parties), so the O(n^2) products of the
// barycentric weights and of P(t) = prod_m (t - x_m) collapse to factorials times a product over the holes:
//   prod_{m != k} (x_k - x_m) = (x_k-256)! (hi-x_k)! (-1)^(hi-x_k) / prod_h (x_k - h)
//   prod_m (t - x_m)          = (-1)^(hi-255) (hi-t)! / (255-t)! / prod_h (t - h)                      (t < 256)
//                             = (t-256)! (hi-t)! (-1)^(hi-t) / prod_{h != t} (t - h)                  (t a hole in [256, hi])
// A target that is itself a node (t = x_h, only for t in 256..406) has P(t) = 0; kv_node_targets writes the share there.
constexpr int FACT_N = 1712;      // factorials up to hi <= 1709
__global__ void __launch_bounds__(256) kv_lagrange(VerifyBufs vb, const u16 *__restrict__ inv_g, const u16 *__restrict__ fact_g)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 x[D2], inv[Q], fact[FACT_N], ifact[FACT_N], hole[NT];
    __shared__ int nh;
    for (int i = tid; i < Q; i += 256) inv[i] = inv_g[i];
    for (int i = tid; i < FACT_N; i += 256) { fact[i] = fact_g[i]; ifact[i] = fact_g[FACT_N + i]; }
    for (int i = tid; i < D2; i += 256) x[i] = (u16)(vb.REST[(size_t)b * NR + i] + 256);
    __syncthreads();
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass ? D2 : D1, nt = pass ? 256 : D1;
        const int hi = x[n - 1];
        if (tid == 0) {                                   // holes of [256, hi]: opened parties below the last node
            int c = 0;
            for (int i = 0; i < NT; i++) { const int h = vb.I[(size_t)b * NT + i] + 256; if (h < hi) hole[c++] = (u16)h; }
            nh = c;
        }
        __syncthreads();
        for (int k = tid; k < n; k += 256) {
            const uint32_t xk = x[k];
            uint32_t num = 1;
            for (int i = 0; i < nh; i++) num = gf_mul(num, gf_sub(xk, hole[i]));
            uint32_t v = gf_mul(gf_mul(num, ifact[xk - 256]), ifact[hi - xk]);
            if ((hi - xk) & 1) v = gf_sub(0, v);
            (pass ? vb.W2 + (size_t)b * VR2LD : vb.W1 + (size_t)b * YLD)[k] = (u16)v;
        }
        for (int t = tid; t < nt; t += 256) {
            uint32_t den = 1; bool node = false;
            if (t >= 256) node = vb.POS[(size_t)b * NP + (t - 256)] >= 0;       // only in pass 0: party t - 256 is a rest party = a node
            uint32_t v = 0;
            if (!node) {
                for (int i = 0; i < nh; i++) { const uint32_t dd = gf_sub((uint32_t)t, hole[i]); if (dd) den = gf_mul(den, dd); }
                if (t < 256) { v = gf_mul(fact[hi - t], ifact[255 - t]); if ((hi - 255) & 1) v = gf_sub(0, v); }
                else { v = gf_mul(fact[t - 256], fact[hi - t]); if ((hi - t) & 1) v = gf_sub(0, v); }
                v = gf_mul(v, inv[den]);
            }
            if (vb.pt_mont) (pass ? reinterpret_cast<int2 *>(vb.PT2) + (size_t)b * 256 : reinterpret_cast<int2 *>(vb.PT1) + (size_t)b * LM1_ROWS)[t] = sn_pair(v);
            else (pass ? vb.PT2 + (size_t)b * 256 : vb.PT1 + (size_t)b * LM1_ROWS)[t] = (u16)v;
        }
        __syncthreads();
    }
}

// V5+V6 (mlwe_verifier.cpp:97-124): NTT(recon(beta_j)) == recon(gamma_j).  BS rows are j*2 + {beta,gamma}.
__global__ void __launch_bounds__(128) kv_check_bg(VerifyBufs vb)
{
    const int b = blockIdx.y, j = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 p[256];
    __shared__ int bad;
    const u16 *bs = vb.BS + ((size_t)b * 2 * MK + 2 * j) * 256, *gs = bs + 256;
    p[tid] = bs[tid]; p[tid + 128] = bs[tid + 128];
    if (tid == 0) bad = 0;
    __syncthreads();
    ntt256_block(p, tid);
    if (p[tid] != gs[tid] || p[tid + 128] != gs[tid + 128]) bad = 1;
    __syncthreads();
    if (tid == 0 && bad) atomicOr(&vb.flags[b], VF_BG);
}

// Checks on the interpolated secrets and the rows of the re-sharing GEMM (mlwe_verifier.cpp:257-270, :287-301,
// :354-363, :415-441, :528-543).  YV rows: [0,K) s+r, [K,2K) e+r, [2K,3K) t, then s_eta[K][E], e_eta[K][E];
// appended here: NTT(s+r), NTT(e+r), A o NTT(s+r), all with the tail of the row they derive from.
template <int K>
__global__ void __launch_bounds__(128) kv_open(VerifyBufs vb)
{
    const VDims d = make_vdims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ u16 sSR[K][256], sER[K][256];
    __shared__ int bad;
    if (tid == 0) bad = 0;
    __syncthreads();
    u16 *yv = vb.YV + (size_t)b * d.nyrows * YLD;
    int f = 0;
    for (int c = tid; c < 256; c += 128) {
        for (int i = 0; i < K; i++) {
            sSR[i][c] = yv[(size_t)i * YLD + c]; sER[i][c] = yv[(size_t)(K + i) * YLD + c];
            if (yv[(size_t)(2 * K + i) * YLD + c] != vb.TPK[((size_t)b * K + i) * 256 + c]) f |= VF_T;
            for (int m = 0; m < d.E; m++) {
                const u16 cur = (u16)(m >= d.eta ? m - d.eta : m + Q - d.eta);     // gf3329_sub(j, KYBER_ETA1), :418
                if (yv[(size_t)(3 * K + i * d.E + m) * YLD + c] != cur) f |= VF_ETA;
                if (yv[(size_t)(3 * K + K * d.E + i * d.E + m) * YLD + c] != cur) f |= VF_ETA;
            }
        }
        for (int r = 0; r < d.n2rows; r++) if (vb.UZ[((size_t)b * d.n2rows + r) * 256 + c] != 0) f |= VF_UZ;
    }
    if (f) atomicOr(&bad, f);
    __syncthreads();
    for (int i = 0; i < K; i++) { ntt256_block(sSR[i], tid); ntt256_block(sER[i], tid); }
    const u16 *AH = vb.AH + (size_t)b * K * K * 256;
    for (int i = 0; i < K; i++) {
        uint32_t q0 = 0, q1 = 0;
        for (int j = 0; j < K; j++) {
            uint32_t r0, r1;
            basemul_pair(r0, r1, AH[(i * K + j) * 256 + 2 * tid], AH[(i * K + j) * 256 + 2 * tid + 1], sSR[j][2 * tid], sSR[j][2 * tid + 1], tid);
            q0 = gf_add(q0, r0); q1 = gf_add(q1, r1);
        }
        u16 *yTsr = yv + (size_t)(d.n1rows + i) * YLD, *yTer = yv + (size_t)(d.n1rows + K + i) * YLD, *yAsr = yv + (size_t)(d.n1rows + 2 * K + i) * YLD;
        yAsr[2 * tid] = (u16)q0; yAsr[2 * tid + 1] = (u16)q1;
        yTsr[tid] = sSR[i][tid]; yTsr[tid + 128] = sSR[i][tid + 128];
        yTer[tid] = sER[i][tid]; yTer[tid + 128] = sER[i][tid + 128];
        for (int q = 256 + tid; q < D1; q += 128) {
            const u16 srq = yv[(size_t)i * YLD + q], erq = yv[(size_t)(K + i) * YLD + q];
            yTsr[q] = srq; yAsr[q] = srq; yTer[q] = erq;
        }
    }
    if (tid == 0 && bad) atomicOr(&vb.flags[b], bad);
}

// Per-party checks against the regenerated sharings (planes VSH, same row order as YV) and construction of the
// inputs of the last two steps: the u^(2d) rows for recon_secrets_2ddeg (mlwe_verifier.cpp:469-556) and the view
// records of the opened parties (:584-630).  kv_check_parties: the rest set, one thread per (proof, party);
// kv_check_opened (below): the opened set, one thread per record element / check.
template <int K>
__global__ void __launch_bounds__(128) kv_check_parties(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const VDims d = make_vdims(K);
    constexpr int ETA = (K == 2) ? 3 : 2, E = 2 * ETA + 1, M = 2 * ETA, F = MK + 2 * K + 1;
    const int b = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
    if (p >= NP) return;
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    const u16 *vsh = vb.VSH + (size_t)b * d.nyrows * SLD + SOFF + p;
    // every source of this kernel is read-only here: loads through the non-coherent path may be hoisted above the U2 stores
    auto V = [&](int row) -> u16 { return __ldg(vsh + (size_t)row * SLD); };
    auto pi16 = [](const u8 *q, size_t off, size_t idx) -> u16 { return __ldg(reinterpret_cast<const u16 *>(q + off) + idx); };
    const int pos = vb.POS[(size_t)b * NP + p];
    if (pos < 0) return;                                       // opened party: kv_check_opened
    int f = 0;
    {                                                          // rest party, index pos in the proof's [R] arrays
        for (int i = 0; i < K; i++) {
            if (V(i) != pi16(pi, L.o_sr, (size_t)pos * K + i)) f |= VF_SR;              // :232-246
            if (V(K + i) != pi16(pi, L.o_er, (size_t)pos * K + i)) f |= VF_SR;
            if (vb.strict) {   // the reference only consumes the first 407 rest parties of these fields (mlwe_verifier.cpp:321-323, :390-394)
                if (V(2 * K + i) != pi16(pi, L.o_t, (size_t)pos * K + i)) f |= VF_STRICT;
                for (int m = 0; m < E; m++) {
                    if (V(3 * K + i * E + m) != pi16(pi, L.o_seta, ((size_t)pos * K + i) * E + m)) f |= VF_STRICT;
                    if (V(3 * K + K * E + i * E + m) != pi16(pi, L.o_eeta, ((size_t)pos * K + i) * E + m)) f |= VF_STRICT;
                }
            }
        }
        if (p < D2)                                            // :547-552 rest shares of u^(2d), used raw in gf3329_mul
            for (int r = 0; r < d.n2rows; r++) {
                const int w = r / (K * M), im = r % (K * M);
                vb.U2[((size_t)b * d.n2rows + r) * VR2LD + p] = (u16)(pi16(pi, w ? L.o_ue : L.o_us, (size_t)pos * K * M + im) % (uint32_t)Q);
            }
    }
    if (f) atomicOr(&vb.flags[b], f);
}

// The opened parties' share of V9-V15 (mlwe_verifier.cpp:249-312, :365-376, :447-493) and their view records (:584-632): 64 threads per opened party
// (four parties per CTA), one thread per record element / check.  As a branch of the thread-per-party kernel above this was a chain of some 250 dependent loads in one lane
// of nearly every warp (0.30 ms per 1024 proofs for very little work).
template <int K>
__global__ void __launch_bounds__(256) kv_check_opened(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const VDims d = make_vdims(K);
    constexpr int ETA = (K == 2) ? 3 : 2, E = 2 * ETA + 1, M = 2 * ETA;
    const int o = blockIdx.x * 4 + (threadIdx.x >> 6), b = blockIdx.y, tid = threadIdx.x & 63;      // four opened parties per CTA, 64 threads each
    if (o >= NT) return;
    const int p = vb.I[(size_t)b * NT + o];                    // kv_setup left a valid, duplicate-free set here even for a malformed proof
    const u8 *pi = pis + L.proof_bytes * (size_t)b;
    const u16 *vsh = vb.VSH + (size_t)b * d.nyrows * SLD + SOFF + p;
    auto V = [&](int row) -> u16 { return __ldg(vsh + (size_t)row * SLD); };
    auto pi16 = [](const u8 *q, size_t off, size_t idx) -> u16 { return __ldg(reinterpret_cast<const u16 *>(q + off) + idx); };
    const u16 *opv = vb.OPV + ((size_t)b * NT + o) * OPLD;
    u16 *vr = vb.VR + ((size_t)b * NT + o) * d.vrld;
    const u16 *tc = reinterpret_cast<const u16 *>(vb.TCR + ((size_t)b * NP + p) * 32);
    const u16 *cr = vb.CR + ((size_t)b * NT + o) * d.crld;
    int f = 0;
    // view record: commitment digest | s e f NTT_f record | beta[0..K) gamma[0..K) [s+r] [e+r] | per i: z_s z_e u_s u_e
    for (int idx = tid; idx < 16 + d.nc; idx += 64) vr[idx] = idx < 16 ? __ldg(tc + idx) : __ldg(cr + idx - 16);
    if (tid < 4 * K) {
        const int kind = tid / K, i = tid % K;
        vr[16 + d.nc + tid] = kind == 0 ? __ldg(opv + i) : kind == 1 ? __ldg(opv + MK + i) : kind == 2 ? V(i) : V(K + i);
    }
    if (tid < K) {
        const int i = tid;
        const size_t oi = (size_t)o * K + i;
        const u16 As = pi16(pi, L.o_NTTAs, oi), Ar = pi16(pi, L.o_NTTAr, oi), Te = pi16(pi, L.o_NTTe, oi);
        if (pi16(pi, L.o_NTTs, oi) != ref_sub(V(d.n1rows + i), __ldg(opv + 2 * MK + 2 * K + i))) f |= VF_NTT;          // :273-284
        if (Te != ref_sub(V(d.n1rows + K + i), __ldg(opv + 2 * MK + 2 * K + K + i))) f |= VF_NTT;
        if (V(d.n1rows + 2 * K + i) != ref_add(As, Ar)) f |= VF_ASR;                                           // :304-312
        if (V(2 * K + i) != ref_add(As, Te)) f |= VF_TREL;                                                     // :365-376
    }
    if (tid < K * E) {                                                                                        // :447-466
        const int i = tid / E, m = tid % E;
        const size_t oi = (size_t)o * K + i;
        if (pi16(pi, L.o_ssub, oi * E + m) != ref_sub(pi16(pi, L.o_s, oi), V(3 * K + i * E + m))) f |= VF_SUBETA;
        if (pi16(pi, L.o_esub, oi * E + m) != ref_sub(pi16(pi, L.o_e, oi), V(3 * K + K * E + i * E + m))) f |= VF_SUBETA;
    }
    if (tid < K * M) {                                                                                        // :471-493
        const int i = tid / M, m = tid % M;
        const size_t oi = (size_t)o * K + i;
        const u16 as_ = m == 0 ? pi16(pi, L.o_ssub, oi * E) : pi16(pi, L.o_zs, oi * M + m - 1);
        const u16 ae_ = m == 0 ? pi16(pi, L.o_esub, oi * E) : pi16(pi, L.o_ze, oi * M + m - 1);
        const u16 z2s = (u16)(((uint32_t)as_ * pi16(pi, L.o_ssub, oi * E + m + 1)) % (uint32_t)Q);
        const u16 z2e = (u16)(((uint32_t)ae_ * pi16(pi, L.o_esub, oi * E + m + 1)) % (uint32_t)Q);
        const u16 zs = pi16(pi, L.o_zs, oi * M + m), ze = pi16(pi, L.o_ze, oi * M + m);
        const u16 us = ref_sub(z2s, zs), ue = ref_sub(z2e, ze);
        if (p < D2) {
            vb.U2[((size_t)b * d.n2rows + i * M + m) * VR2LD + p] = (u16)(us % (uint32_t)Q);
            vb.U2[((size_t)b * d.n2rows + K * M + i * M + m) * VR2LD + p] = (u16)(ue % (uint32_t)Q);
        }
        u16 *vz = vr + 16 + d.nc + 4 * K + i * 4 * M;
        vz[m] = zs; vz[M + m] = ze; vz[2 * M + m] = us; vz[3 * M + m] = ue;
    }
    if (f) atomicOr(&vb.flags[b], f);
}

// Final: recon_secrets_2ddeg outputs must vanish (:555-569), recomputed I must equal pi->I (:678-683).
template <int K>
__global__ void __launch_bounds__(128) kv_final(VerifyBufs vb, u8 *__restrict__ ok)
{
    const VDims d = make_vdims(K);
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int bad;
    if (tid == 0) bad = 0;
    __syncthreads();
    int f = 0;
    for (int idx = tid; idx < d.n2rows * 256; idx += 128) if (vb.UR[(size_t)b * d.n2rows * 256 + idx] != 0) f |= VF_U2D;
    for (int i = tid; i < NT; i += 128) if (vb.I2[(size_t)b * NT + i] != vb.I[(size_t)b * NT + i]) f |= VF_FS2;
    if (f) atomicOr(&bad, f);
    __syncthreads();
    if (tid == 0) {
        const int fl = vb.flags[b] | bad;
        vb.flags[b] = fl;
        ok[b] = fl == 0 ? 1 : 0;
    }
}

// Hardened decoding: every u16 field element of the proof must be a canonical residue (< q).  The reference silently
// reduces (NTL assignment, gf3329_mul) or never reads some of them.
template <int K>
__global__ void __launch_bounds__(256) kv_strict_scan(VerifyBufs vb, const u8 *__restrict__ pis)
{
    const Layout L = make_layout(K);
    const int b = blockIdx.y;
    const u16 *w = reinterpret_cast<const u16 *>(pis + L.proof_bytes * (size_t)b);
    const size_t n = L.proof_bytes / 2, i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const size_t off = 2 * i;
    const bool digest = (off >= L.o_Tcomm && off < L.o_I) || off >= L.o_comm;       // byte strings, not field elements
    const bool idx = off >= L.o_I && off < L.o_s;                                     // opened-set indices (checked in kv_setup)
    if (!digest && !idx && w[i] >= Q) atomicOr(&vb.flags[b], VF_STRICT);
}

template <int K> __global__ void kv_clear(VerifyBufs vb, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) vb.flags[b] = 0;
}

// launch sequence; returns the number of kernels launched or -1
// Side stream: everything that does not depend on the FS-1 challenges (Lagrange weights, the interpolations, the checks on the
// interpolated secrets, the regeneration of the sharings) runs on `side` next to commit hashes -> FS-1 sponge (one latency-bound warp per
// proof) -> opened-party evaluation -> beta/gamma reconstruction on the main stream; the two join before the per-party checks.
struct VerifySide { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
template <int K>
static int verify_chunk_t(VerifyBufs &vb, const VerifyTables &vt, int B, const u8 *d_pi, const u8 *d_pk, u8 *d_ok, cudaStream_t st, const VerifySide &sd)
{
    const VDims d = make_vdims(K);
    int nl = 0;
    const int ptiles = (NP + 127) / 128;
    vb.pt_mont = vt.sn != nullptr;
    kv_clear<K><<<(B + 127) / 128, 128, 0, st>>>(vb, B); nl++;
    kv_setup<K><<<B, 128, 0, st>>>(vb, d_pi, d_pk); nl++;
    kv_setup_copy<K><<<dim3(SETUP_COPY_CTAS, B), 128, 0, st>>>(vb, d_pi); nl++;
    if (vb.strict) { kv_strict_scan<K><<<dim3((unsigned)((make_layout(K).proof_bytes / 2 + 255) / 256), B), 256, 0, st>>>(vb, d_pi); nl++; }
    // the NTT-convolution path has no shared scratch between the two chains (the dense GEMMs share the split-K workspace)
    const bool fork = sd.st != nullptr && vt.sn != nullptr;
    cudaStream_t s2 = fork ? sd.st : st;
    if (fork) { cudaEventRecord(sd.fork, st); cudaStreamWaitEvent(s2, sd.fork, 0); }
    GemmArgs g{};
    ConvArgs cv{};
    // ---- main chain: V2 commitments of the opened parties, FS-1, opened-party evaluation, beta/gamma reconstruction ----
    {
        constexpr int NC = 2 * (K + MK + 2 * K + 1);
        HashSrc hs{vb.CR, (long long)NT * d.crld, d.crld, 1, 0, nullptr, vb.I, NT};
        k_hash_records<NC><<<dim3(2, B), 128, 0, st>>>(hs, vb.TCR, nullptr, 0, 0); nl++;
    }
    { int fc, ft; fs_launch_dims(B, fc, ft); k_fs1<K><<<fc, ft, 0, st>>>(vb.TCR, vb.PW, B); nl++; }
    if (B >= 64) { kv_eval_opened_idp<K><<<B, 160, 0, st>>>(vb); nl++; }      // batches: two records per thread on IDP.2A; few proofs: one record per thread (latency)
    else { kv_eval_opened<K><<<B, 320, 0, st>>>(vb); nl++; }
    // ---- side chain (independent of the challenges) ----
    kv_lagrange<<<B, 256, 0, s2>>>(vb, vt.inv, vt.fact); nl++;
    kv_gather<K><<<dim3(d.n1rows + d.n2rows, B), 128, 0, s2>>>(vb, d_pi, 2 * MK); nl++;
    // interpolation-apply: rows of all proofs against the fixed Cauchy operand 1 / (t - (p + 256)), columns scaled by each proof's P(t)
    if (vt.sn) {
        cv = ConvArgs{}; cv.A = vb.A1; cv.C = vb.YV; cv.lda = KP1; cv.ldc = YLD; cv.mtotal = B * d.n1rows; cv.rpp = d.n1rows; cv.a_slots = d.n1rows; cv.c_slots = d.nyrows;
        cv.tw = vt.sn->tw; cv.khat = sn_kh_m256(*vt.sn, 5); cv.post = reinterpret_cast<const int2 *>(vb.PT1); cv.post_group = LM1_ROWS;
        nl += conv_ntt_launch<5, 4, KP1, D1, false, true>(cv, s2, fork ? vt.tk_side : vt.tk_main);
        kv_node_targets<K><<<dim3(d.n1rows, B), 160, 0, s2>>>(vb, d_pi); nl++;
        cv = ConvArgs{}; cv.A = vb.A2; cv.C = vb.UZ; cv.lda = KP2; cv.ldc = 256; cv.mtotal = B * d.n2rows; cv.rpp = d.n2rows; cv.a_slots = d.n2rows; cv.c_slots = d.n2rows;
        cv.tw = vt.sn->tw; cv.khat = sn_kh_m256(*vt.sn, 8); cv.post = reinterpret_cast<const int2 *>(vb.PT2); cv.post_group = 256;
        nl += conv_ntt_launch<8, 2, KP2, 256, false, true>(cv, s2, fork ? vt.tk_side : vt.tk_main);
    } else {
        g = GemmArgs{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = vb.A1; g.Bt = vt.U1; g.C = vb.YV; g.lda = KP1; g.ldb = KP1; g.ldc = YLD;
        g.mtotal = B * d.n1rows; g.ksteps = KP1 / GE_BK; g.nvalid = D1; g.rpp = d.n1rows; g.a_slots = d.n1rows; g.c_slots = d.nyrows;
        g.colscale = vb.PT1; g.colscale_batch = LM1_ROWS; g.colscale_by_group = 1;
        nl += gf_gemm_launch_auto<7>(g, U1_ROWS, 1, s2);
        kv_node_targets<K><<<dim3(d.n1rows, B), 160, 0, s2>>>(vb, d_pi); nl++;
        g = GemmArgs{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = vb.A2; g.Bt = vt.U2; g.C = vb.UZ; g.lda = KP2; g.ldb = KP2; g.ldc = 256;
        g.mtotal = B * d.n2rows; g.ksteps = KP2 / GE_BK; g.nvalid = 256; g.rpp = d.n2rows; g.a_slots = d.n2rows; g.c_slots = d.n2rows;
        g.colscale = vb.PT2; g.colscale_batch = 256; g.colscale_by_group = 1;
        nl += gf_gemm_launch_auto<8>(g, 256, 1, s2);
    }
    kv_open<K><<<B, 128, 0, s2>>>(vb); nl++;
    // regenerate every sharing at all 1454 parties: YV x S.  Row groups per proof: [0,3K) s+r, e+r, t | [3K, n1rows) the eta
    // sharings, whose 256 secrets were just checked to be one constant (short path: tail terms only) | [n1rows, nyrows)
    if (vt.sn) {
        g = GemmArgs{}; g.A = vb.YV; g.C = vb.VSH; g.lda = YLD; g.ldc = SLD; g.rpp = d.nyrows; g.slot_lo = 0; g.a_slots = d.nyrows; g.c_slots = d.nyrows;
        g.mtotal = B * d.nyrows; g.c_off = SOFF + NT + 1; g.tail = 1; g.tail_off = NL;
        nl += share_ntt_launch(share_conv_args(g, *vt.sn), s2, vt.sn_variant, fork ? vt.tk_side : vt.tk_main);
    } else {
        const int grp_lo[3] = {0, 3 * K, d.n1rows}, grp_hi[3] = {3 * K, d.n1rows, d.nyrows};
        for (int gi = 0; gi < 3; gi++) {
            g = GemmArgs{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = vb.YV; g.Bt = vt.St; g.C = vb.VSH; g.lda = YLD; g.ldb = YLD; g.ldc = SLD;
            g.rpp = grp_hi[gi] - grp_lo[gi]; g.slot_lo = grp_lo[gi]; g.a_slots = d.nyrows; g.c_slots = d.nyrows;
            g.mtotal = B * g.rpp; g.ksteps = YLD / GE_BK; g.nvalid = NX; g.c_off = SOFF + NT + 1; g.tail = 1; g.tail_off = NL;
            if (gi == 1) { g.A = vb.YV + NL; g.Bt = vt.St + NL; g.ksteps = (YLD - NL) / GE_BK; g.tail_off = 0; g.addvec = vt.SU; g.scale_src = vb.YV; }
            g.half_last = 1;
            nl += gf_gemm_launch_auto<7>(g, GE_NCOLS7, 1, s2);
        }
    }
    if (fork) cudaEventRecord(sd.join, s2);
    // ---- main chain, continued: beta/gamma of parties 0..406 (rest: from the proof, opened: just evaluated), reconstruction, NTT check ----
    kv_gather_bg<K><<<dim3((D1 + VG_TP - 1) / VG_TP, B), 256, 0, st>>>(vb, d_pi); nl++;
    // beta/gamma reconstruction: ABG x R1 (recon_secrets_ddeg, ss.cpp:37-54)
    if (vt.sn) {
        cv = ConvArgs{}; cv.A = vb.ABG; cv.C = vb.BS; cv.lda = YLD; cv.ldc = 256; cv.mtotal = B * 2 * MK; cv.rpp = cv.mtotal;
        cv.tw = vt.sn->tw; cv.khat = sn_kh_m256(*vt.sn, 4); cv.pre = vt.sn->wj; cv.post = vt.sn->pr1;
        nl += conv_ntt_launch<4, 2, D1, 256, true, false>(cv, st, vt.tk_main);
    } else {
        g = GemmArgs{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = vb.ABG; g.Bt = vt.R1; g.C = vb.BS; g.lda = YLD; g.ldb = YLD; g.ldc = 256;
        g.mtotal = B * 2 * MK; g.ksteps = YLD / GE_BK; g.nvalid = 256; g.rpp = g.mtotal; g.c_off = 0; g.half_last = 1;
        nl += gf_gemm_launch_auto<8>(g, 256, 1, st);
    }
    kv_check_bg<<<dim3(MK, B), 128, 0, st>>>(vb); nl++;
    if (fork) cudaStreamWaitEvent(st, sd.join, 0);
    kv_check_parties<K><<<dim3(ptiles, B), 128, 0, st>>>(vb, d_pi); nl++;
    kv_check_opened<K><<<dim3((NT + 3) / 4, B), 256, 0, st>>>(vb, d_pi); nl++;
    if (vt.sn) {      // recon_secrets_2ddeg over parties 0..812 (ss.cpp:56-73)
        cv = ConvArgs{}; cv.A = vb.U2; cv.C = vb.UR; cv.lda = VR2LD; cv.ldc = 256; cv.mtotal = B * d.n2rows; cv.rpp = cv.mtotal;
        cv.tw = vt.sn->tw; cv.khat = sn_kh_m256(*vt.sn, 7); cv.pre = vt.sn->wj2; cv.post = vt.sn->pr2;
        nl += conv_ntt_launch<7, 2, D2, 256, true, false>(cv, st, vt.tk_main);
    } else {
        g = GemmArgs{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = vb.U2; g.Bt = vt.R2; g.C = vb.UR; g.lda = VR2LD; g.ldb = VR2LD; g.ldc = 256;
        g.mtotal = B * d.n2rows; g.ksteps = VR2LD / GE_BK; g.nvalid = 256; g.rpp = g.mtotal;
        nl += gf_gemm_launch_auto<8>(g, 256, 1, st);
    }
    {   // V16: view hashes of the opened parties, FS-2, compare
        constexpr int ETA = (K == 2) ? 3 : 2, NV = 16 + 2 * (K + MK + 2 * K + 1) + 4 * K + 8 * ETA * K;
        HashSrc hs{vb.VR, (long long)NT * d.vrld, d.vrld, 1, 0, nullptr, vb.I, NT};
        k_hash_records<NV><<<dim3(2, B), 128, 0, st>>>(hs, vb.VWR, nullptr, 0, 0); nl++;
    }
    { int fc, ft; fs_launch_dims(B, fc, ft); k_fs2<<<fc, ft, 0, st>>>(vb.VWR, vb.I2, vb.REST2, B); nl++; }
    kv_final<K><<<B, 128, 0, st>>>(vb, d_ok); nl++;
    return cudaGetLastError() == cudaSuccess ? nl : -1;
}

static inline int verify_chunk(int k, VerifyBufs &vb, const VerifyTables &vt, int B, const u8 *d_pi, const u8 *d_pk, u8 *d_ok, cudaStream_t st, const VerifySide &sd)
{
    switch (k) {
    case 2: return verify_chunk_t<2>(vb, vt, B, d_pi, d_pk, d_ok, st, sd);
    case 3: return verify_chunk_t<3>(vb, vt, B, d_pi, d_pk, d_ok, st, sd);
    default: return verify_chunk_t<4>(vb, vt, B, d_pi, d_pk, d_ok, st, sd);
    }
}

}  // namespace kosk
