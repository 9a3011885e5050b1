// gf_gemm_imma.cuh -- EXPERIMENTAL, opt-in (KOSK_B200_TENSOR=1): the GF(3329) share-evaluation contraction on the int8
// tensor-core path.  The plan of record (north_star) is the INT32 pipe (gf_gemm.cuh); this variant exists to measure what
// a limb-split integer tensor formulation buys and is bit-identical by construction:
//   centered residues a, b in [-1664, 1664] are split into signed 7-bit limbs  a = 128*a1 + a0  (|a0| <= 64, |a1| <= 13),
//   a*b = a0*b0 + 128*(a0*b1 + a1*b0) + 16384*a1*b1, each partial sum an exact int32 (407 * 64*64 < 2^21), and
//   P00 + 128*Pc + 16384*P11 < 2^31, reduced mod q once per output.
// Instruction: mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 (legacy warp-level IMMA; compiles for sm_100a).
// CTA tile 128 rows x 64 columns x 32 terms, 8 warps (4 x 2), warp tile 32 x 32, three accumulator sets per warp.
#pragma once
#include "gf_gemm.cuh"

namespace kosk {

constexpr int IM_BM = 128, IM_BN = 64, IM_PITCH = 48;     // 32 data bytes + 16 pad per smem row: conflict-free ldmatrix
constexpr int IM_STAGES = 3;
constexpr int IM_STAGE_BYTES = 2 * IM_BM * IM_PITCH + 2 * IM_BN * IM_PITCH;     // A limbs + B limbs of one 32-term step
constexpr int IM_SMEM = IM_STAGES * IM_STAGE_BYTES;

struct ImmaTables {
    const int8_t *B0, *B1;     // limb planes of the centered table, [npad][ldb] int8
    const int8_t *A0, *A1;     // limb planes of the A rows (k_limb_split), same row mapping and stride (bytes) as g.A
};

// u16 canonical residues -> two int8 limb planes (a = 128*a1 + a0 centered) for rows [slot_lo, slot_lo + rows) of every
// proof (row stride lda elements / bytes, `slots` rows per proof).  One 16-byte chunk (8 residues) per thread.
__global__ void __launch_bounds__(256) k_limb_split(const u16 *__restrict__ Y, int8_t *__restrict__ L0, int8_t *__restrict__ L1,
                                                    int rows, int slot_lo, int slots, int lda, size_t nchunks)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nchunks) return;
    const int cpr = lda / 8;                                   // chunks per row
    const size_t r = i / cpr; const int q = (int)(i % cpr);
    const size_t row = (r / rows) * slots + slot_lo + r % rows;
    const uint4 v = *reinterpret_cast<const uint4 *>(Y + row * lda + q * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo[2] = {0, 0}, hi[2] = {0, 0};
#pragma unroll
    for (int e = 0; e < 8; e++) {
        int32_t a = (int32_t)((e & 1) ? w[e >> 1] >> 16 : w[e >> 1] & 0xFFFF);
        if (a >= Q) a %= Q;                                     // rows need not be canonical
        a = a > Q / 2 ? a - Q : a;
        const int32_t a0 = ((a + 64) & 127) - 64, a1 = (a - a0) >> 7;
        lo[e >> 2] |= (uint32_t)(a0 & 0xFF) << (8 * (e & 3));
        hi[e >> 2] |= (uint32_t)(a1 & 0xFF) << (8 * (e & 3));
    }
    *reinterpret_cast<uint2 *>(L0 + row * lda + q * 8) = make_uint2(lo[0], lo[1]);
    *reinterpret_cast<uint2 *>(L1 + row * lda + q * 8) = make_uint2(hi[0], hi[1]);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void imma_16832(int32_t (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;          // src-size 0 -> 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(a), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(256, 2) k_gf_gemm_imma(const GemmArgs g, const ImmaTables tb)
{
    extern __shared__ __align__(16) unsigned char im_smem[];
    auto sA = [&](int stg, int limb) { return im_smem + stg * IM_STAGE_BYTES + limb * IM_BM * IM_PITCH; };
    auto sB = [&](int stg, int limb) { return im_smem + stg * IM_STAGE_BYTES + 2 * IM_BM * IM_PITCH + limb * IM_BN * IM_PITCH; };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
    const int m0 = blockIdx.y * IM_BM, n0 = blockIdx.x * IM_BN;
    const u16 *Ab = g.A + (size_t)blockIdx.z * g.a_batch;
    u16 *Cb = g.C + (size_t)blockIdx.z * g.c_batch;
    // A loader: per 32-term step 128 rows x 2 chunks x 2 limbs = 512 chunks of 16 bytes: thread -> (limb, row, half), twice
    const int8_t *a_src[2]; bool a_ok[2]; int a_dst[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int c = tid + 256 * h, limb = c >> 8, row = (c & 255) >> 1, half = c & 1, m = m0 + row;
        a_ok[h] = m < g.mtotal;
        const size_t srow = a_ok[h] ? (size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp : 0;
        a_src[h] = (limb ? tb.A1 : tb.A0) + srow * g.lda + half * 16;
        a_dst[h] = limb * IM_BM * IM_PITCH + row * IM_PITCH + half * 16;
    }
    // B loader: 64 rows x 2 chunks x 2 limbs = 256 chunks: one per thread
    const int bl = tid >> 7, bn = (tid & 127) >> 1, bh = tid & 1;
    const int8_t *b_src = (bl ? tb.B1 : tb.B0) + (size_t)(n0 + bn) * g.ldb + bh * 16;
    const int b_dst = 2 * IM_BM * IM_PITCH + bl * IM_BN * IM_PITCH + bn * IM_PITCH + bh * 16;
    auto issue = [&](int kt, int stg) {
        unsigned char *base = im_smem + stg * IM_STAGE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; h++) cp_async16(base + a_dst[h], a_src[h] + kt * 32, a_ok[h]);
        cp_async16(base + b_dst, b_src + kt * 32, true);
    };

    int32_t p00[2][4][4], pc[2][4][4], p11[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) { p00[i][j][e] = 0; pc[i][j][e] = 0; p11[i][j][e] = 0; }

    const int nk = g.ksteps / 2;        // ksteps counts 16-term steps; this kernel consumes 32 terms per step
#pragma unroll
    for (int s = 0; s < IM_STAGES - 1; s++) { if (s < nk) issue(s, s); asm volatile("cp.async.commit_group;\n"); }
#pragma unroll 1
    for (int kt = 0; kt < nk; kt++) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(IM_STAGES - 2));
        __syncthreads();
        { const int nx = kt + IM_STAGES - 1; if (nx < nk) issue(nx, nx % IM_STAGES); asm volatile("cp.async.commit_group;\n"); }
        const int stg = kt % IM_STAGES;
        uint32_t fa[2][2][4], fb[2][2][4];     // [limb][m-tile][4] ; [limb][n-tile pair][4] (two n-tiles per ldmatrix.x4)
#pragma unroll
        for (int l = 0; l < 2; l++) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int row = wm * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kb = (lane >> 4) * 16;
                ldmatrix_x4(fa[l][mt], sA(stg, l) + row * IM_PITCH + kb);
            }
#pragma unroll
            for (int np = 0; np < 2; np++) {
                const int n = wn * 32 + (np * 2 + (lane >> 4)) * 8 + (lane & 7), kb = ((lane >> 3) & 1) * 16;
                ldmatrix_x4(fb[l][np], sB(stg, l) + n * IM_PITCH + kb);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int np = nt >> 1, o = (nt & 1) * 2;
                imma_16832(p00[mt][nt], fa[0][mt], fb[0][np][o], fb[0][np][o + 1]);
                imma_16832(pc[mt][nt], fa[0][mt], fb[1][np][o], fb[1][np][o + 1]);
                imma_16832(pc[mt][nt], fa[1][mt], fb[0][np][o], fb[0][np][o + 1]);
                imma_16832(p11[mt][nt], fa[1][mt], fb[1][np][o], fb[1][np][o + 1]);
            }
    }
    // epilogue: combine limbs, add the constant-secret term if any, reduce, store 2 x u16 per 32-bit store
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hr = 0; hr < 2; hr++) {
            const int m = m0 + wm * 32 + mt * 16 + gid + hr * 8;
            if (m >= g.mtotal) continue;
            const size_t srow = (size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp;
            u16 *dst = Cb + ((size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp) * g.ldc + g.c_off;
            const int32_t cst = g.addvec ? gf_center(g.scale_src[srow * g.lda]) : 0;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int x = n0 + wn * 32 + nt * 8 + tig * 2;
                uint32_t v[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int ci = hr * 2 + e;
                    int32_t t = p00[mt][nt][ci] + 128 * pc[mt][nt][ci] + 16384 * p11[mt][nt][ci];
                    if (g.addvec) t = t % Q + cst * (int32_t)g.addvec[x + e];
                    v[e] = gf_canon(t);
                }
                if (x + 1 < g.nvalid) *reinterpret_cast<uint32_t *>(dst + x) = v[0] | (v[1] << 16);
                else if (x < g.nvalid) dst[x] = (u16)v[0];
            }
        }
    if (g.tail && blockIdx.x == 0) {
        for (int idx = tid; idx < IM_BM * (NT + 1); idx += 256) {
            const int r = idx / (NT + 1), c = idx % (NT + 1), m = m0 + r;
            if (m >= g.mtotal) continue;
            const size_t ar = (size_t)(m / g.rpp) * g.a_slots + g.slot_lo + m % g.rpp, cr = (size_t)(m / g.rpp) * g.c_slots + g.slot_lo + m % g.rpp;
            Cb[cr * g.ldc + g.c_off - (NT + 1) + c] = Ab[ar * g.lda + g.tail_off + c];
        }
    }
}

static inline int gf_gemm_imma_launch(const GemmArgs &g, const ImmaTables &tb, int ncols, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_gf_gemm_imma, cudaFuncAttributeMaxDynamicSharedMemorySize, IM_SMEM); attr_set = true; }
    dim3 grid((ncols + IM_BN - 1) / IM_BN, (g.mtotal + IM_BM - 1) / IM_BM, 1);
    k_gf_gemm_imma<<<grid, 256, IM_SMEM, st>>>(g, tb);
    return 1;
}

}  // namespace kosk
