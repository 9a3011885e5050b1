// wire_kernels.cuh -- device side of the compact wire format (kosk_common.cuh, WireLayout): struct mpcith_proof bytes
// (reference mlwe_prover.hpp:57-75, 16 bits per field element) <-> 12 bits per field element, so that the PCIe link carries
// 519 / 532 / 579 KB per proof instead of 664 / 681 / 744 KB.  Pure streaming kernels (HBM-bound): one thread per group of 8
// elements (16 B <-> 12 B) or per 16-byte chunk of the digest arrays; proofs are only 4-byte aligned, so all accesses are 32-bit.
#pragma once
#include "kosk_common.cuh"

namespace kosk {

__device__ __forceinline__ void wire_pack8(const uint32_t (&in)[4], uint32_t (&w)[3])
{
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { e[2 * i] = in[i] & 0xFFFu; e[2 * i + 1] = (in[i] >> 16) & 0xFFFu; }
    w[0] = e[0] | (e[1] << 12) | (e[2] << 24);
    w[1] = (e[2] >> 8) | (e[3] << 4) | (e[4] << 16) | (e[5] << 28);
    w[2] = (e[5] >> 4) | (e[6] << 8) | (e[7] << 20);
}
__device__ __forceinline__ void wire_unpack8(const uint32_t (&w)[3], uint32_t (&out)[4])
{
    const uint32_t e0 = w[0] & 0xFFFu, e1 = (w[0] >> 12) & 0xFFFu, e2 = (w[0] >> 24) | ((w[1] & 0xFu) << 8), e3 = (w[1] >> 4) & 0xFFFu;
    const uint32_t e4 = (w[1] >> 16) & 0xFFFu, e5 = (w[1] >> 28) | ((w[2] & 0xFFu) << 4), e6 = (w[2] >> 8) & 0xFFFu, e7 = w[2] >> 20;
    out[0] = e0 | (e1 << 16); out[1] = e2 | (e3 << 16); out[2] = e4 | (e5 << 16); out[3] = e6 | (e7 << 16);
}

// grid = (item tiles, B); items of a proof: groups of run A, groups of run B, 16-byte chunks of Tcomm, of comm
template <bool PACK>
__global__ void __launch_bounds__(256) k_wire(const u8 *__restrict__ src, u8 *__restrict__ dst, const WireLayout W)
{
    const uint32_t GA = (W.nA + 7) / 8, GB = (W.nB + 7) / 8, HC = NR * 32 / 16;
    const uint32_t item = blockIdx.x * 256 + threadIdx.x;
    const size_t b = blockIdx.y;
    const u8 *pi_c = (PACK ? src : dst) + b * W.proof_bytes;     // only used for address arithmetic below
    const u8 *wr_c = (PACK ? dst : src) + b * W.wire_bytes;
    if (item < GA + GB) {
        const bool inA = item < GA;
        const uint32_t g = inA ? item : item - GA, n = inA ? W.nA : W.nB;
        const size_t po = (inA ? W.o_A : W.o_B) + 16 * (size_t)g, wo = (inA ? W.w_A : W.w_B) + 12 * (size_t)g;
        const uint32_t rem = min(8u, n - 8 * g);                   // elements of this group (even)
        if (PACK) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(pi_c + po);
            uint32_t in[4], w[3];
#pragma unroll
            for (int i = 0; i < 4; i++) in[i] = (2u * i < rem) ? p[i] : 0u;
            wire_pack8(in, w);
            u8 *q = const_cast<u8 *>(wr_c) + wo;
            if (g + 1 < (inA ? GA : GB)) {
#pragma unroll
                for (int i = 0; i < 3; i++) reinterpret_cast<uint32_t *>(q)[i] = w[i];
            } else {                                               // last group of the run: its valid bytes, then zeros up to the next segment
                const uint32_t vb = rem / 2 * 3, end = (uint32_t)((inA ? W.w_Tcomm : W.w_comm) - wo);
                for (uint32_t i = 0; i < end; i++) q[i] = i < vb ? (u8)(w[i >> 2] >> (8 * (i & 3))) : (u8)0;
            }
        } else {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(wr_c + wo);
            uint32_t w[3], out[4];
#pragma unroll
            for (int i = 0; i < 3; i++) w[i] = q[i];              // a partial last group reads into the segment padding: in bounds
            wire_unpack8(w, out);
            uint32_t *p = reinterpret_cast<uint32_t *>(const_cast<u8 *>(pi_c) + po);
#pragma unroll
            for (int i = 0; i < 4; i++) if (2u * i < rem) p[i] = out[i];
        }
    } else if (item < GA + GB + 2 * HC) {
        const uint32_t h = item - GA - GB;
        const bool first = h < HC;
        const size_t po = (first ? W.o_Tcomm : W.o_comm) + 16 * (size_t)(first ? h : h - HC), wo = (first ? W.w_Tcomm : W.w_comm) + 16 * (size_t)(first ? h : h - HC);
        const uint32_t *s = reinterpret_cast<const uint32_t *>(PACK ? pi_c + po : wr_c + wo);
        uint32_t *d = reinterpret_cast<uint32_t *>(const_cast<u8 *>(PACK ? wr_c + wo : pi_c + po));
        const uint32_t a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3];
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
    }
}

static inline int wire_launch(bool pack, const u8 *src, u8 *dst, int k, int B, cudaStream_t st)
{
    if (B <= 0) return 0;
    const WireLayout W = make_wire_layout(k);
    const uint32_t items = (W.nA + 7) / 8 + (W.nB + 7) / 8 + 2 * (NR * 32 / 16);
    const dim3 grid((items + 255) / 256, B);
    if (pack) k_wire<true><<<grid, 256, 0, st>>>(src, dst, W);
    else k_wire<false><<<grid, 256, 0, st>>>(src, dst, W);
    return 1;
}

}  // namespace kosk
