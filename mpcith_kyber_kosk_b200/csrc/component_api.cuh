// component_api.cuh -- component entry points for the reconstruction and interpolation products of the verifier, so that the
// kernels behind them can be checked element-wise against the oracle (ko_recon_ddeg / ko_recon_2ddeg, ok_lagrange_matrix):
//   recon_rows   recon_secrets_ddeg / _2ddeg (ss.cpp:37-73): k_conv_ntt<4,2> / <7,2>, or the dense table GEMM with KOSK_B200_SHARE_NTT=0
//   interp_rows  the verifier's NTL interpolate + eval over the rest-party nodes (mlwe_verifier.cpp:188-224, :326-352, :397-443,
//                :510-543): kv_lagrange (weights, P(t)) + scatter to party columns + k_conv_ntt<5,4> / <8,2> (or the dense GEMM on the
//                Cauchy operands) + the node targets -- the same kernels and operands verify_chunk_t launches.
// Included by kosk_b200.cu after the context is defined.  Host buffers, synchronous, lane 0.
#pragma once

namespace kosk {

// A[row][p] = share of rest party p (position pos[p] in the rest list, < nn) times its barycentric weight, 0 elsewhere; the same operand
// rows kv_gather builds from a proof
__global__ void __launch_bounds__(128) kc_scatter(const u16 *__restrict__ shares, int nn, const int16_t *__restrict__ pos, const u16 *__restrict__ w, u16 *__restrict__ A, int kp)
{
    const int row = blockIdx.x;
    for (int p = threadIdx.x; p < kp; p += 128) {
        const int j = p < NP ? pos[p] : -1;
        A[(size_t)row * kp + p] = (j >= 0 && j < nn) ? (u16)gf_mul(shares[(size_t)row * nn + j] % (uint32_t)Q, w[j]) : (u16)0;
    }
}
// targets 256..406 that are themselves nodes take the share (kv_node_targets)
__global__ void __launch_bounds__(160) kc_node_targets(const u16 *__restrict__ shares, const int16_t *__restrict__ pos, u16 *__restrict__ out)
{
    const int row = blockIdx.x, p = threadIdx.x;
    if (p > NT) return;
    const int j = pos[p];
    if (j >= 0) out[(size_t)row * YLD + 256 + p] = (u16)(shares[(size_t)row * D1 + j] % (uint32_t)Q);
}
// opened set -> I / REST / POS of "proof" 0 of a VerifyBufs (what kv_setup derives from pi->I)
__global__ void kc_sets(VerifyBufs vb, const u16 *__restrict__ opened)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int p = 0; p < NP; p++) vb.POS[p] = 0;
    for (int i = 0; i < NT; i++) { vb.I[i] = opened[i]; vb.POS[opened[i]] = (int16_t)(-1 - i); }
    int j = 0;
    for (int p = 0; p < NP; p++) if (vb.POS[p] == 0) { vb.REST[j] = (u16)p; vb.POS[p] = (int16_t)j; j++; }
}

}  // namespace kosk

extern "C" {

int kosk_b200_recon_rows(kosk_b200_ctx *c, int degree2, size_t n, const uint16_t *shares, uint16_t *secrets)
{
    if (!c || !shares || !secrets || n == 0 || n > (1u << 20)) return fail(KOSK_E_ARG, "bad argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    const int nin = degree2 ? D2 : D1, ld = degree2 ? VR2LD : YLD;
    DevBuf ba, bc;
    std::vector<u16> ha(n * ld, 0);
    for (size_t r = 0; r < n; r++) memcpy(&ha[r * ld], shares + r * nin, (size_t)nin * 2);
    CU(ba.alloc(ha.size() * 2 + 16)); CU(bc.alloc(n * 256 * 2));
    CU(cudaMemcpy(ba.p, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    cudaStream_t st = c->lanes[0].st;
    if (c->use_ntt) {
        ConvArgs cv{}; cv.A = ba.as<u16>(); cv.C = bc.as<u16>(); cv.lda = ld; cv.ldc = 256; cv.mtotal = (int)n; cv.rpp = (int)n;
        cv.tw = c->sn.tw; cv.khat = sn_kh_m256(c->sn, degree2 ? 7 : 4); cv.pre = degree2 ? c->sn.wj2 : c->sn.wj;
        cv.post = degree2 ? c->sn.pr2 : c->sn.pr1;
        c->launches += degree2 ? conv_ntt_launch<7, 2, D2, 256, true, false>(cv, st) : conv_ntt_launch<4, 2, D1, 256, true, false>(cv, st);
    } else {
        GemmArgs g{}; g.ws = c->lanes[0].vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = ba.as<u16>(); g.Bt = degree2 ? c->d_R2 : c->d_R1; g.C = bc.as<u16>();
        g.lda = ld; g.ldb = ld; g.ldc = 256; g.mtotal = (int)n; g.ksteps = ld / GE_BK; g.nvalid = 256; g.rpp = (int)n; g.half_last = degree2 ? 0 : 1;
        c->launches += gf_gemm_launch_auto<8>(g, 256, 1, st);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(secrets, bc.p, n * 256 * 2, cudaMemcpyDeviceToHost));
    return KOSK_OK;
}

int kosk_b200_interp_rows(kosk_b200_ctx *c, int degree2, const uint16_t *opened, size_t n, const uint16_t *shares, uint16_t *out)
{
    if (!c || !opened || !shares || !out || n == 0 || n > 4096) return fail(KOSK_E_ARG, "bad argument (1 <= n <= 4096)");
    {   // the opened set must be 150 distinct parties
        std::vector<char> seen(NP, 0);
        for (int i = 0; i < NT; i++) { if (opened[i] >= NP || seen[opened[i]]) return fail(KOSK_E_ARG, "opened set: 150 distinct parties < 1454"); seen[opened[i]] = 1; }
    }
    LOCK(c);
    CU(cudaSetDevice(c->device));
    Lane &ln = c->lanes[0];
    cudaStream_t st = ln.st;
    VerifyBufs &vb = ln.vb;
    const int nn = degree2 ? D2 : D1, kp = degree2 ? KP2 : KP1, nt = degree2 ? 256 : D1, ldo = degree2 ? 256 : YLD;
    DevBuf bo, bs, bA, bC;
    CU(bo.alloc(NT * 2)); CU(bs.alloc(n * nn * 2)); CU(bA.alloc(n * kp * 2)); CU(bC.alloc(n * ldo * 2));
    CU(cudaMemcpyAsync(bo.p, opened, NT * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(bs.p, shares, n * nn * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(bC.p, 0, n * ldo * 2, st));
    vb.pt_mont = c->use_ntt ? 1 : 0;
    kc_sets<<<1, 32, 0, st>>>(vb, bo.as<u16>());
    kv_lagrange<<<1, 256, 0, st>>>(vb, c->d_inv, c->d_fact);
    kc_scatter<<<(unsigned)n, 128, 0, st>>>(bs.as<u16>(), nn, vb.POS, degree2 ? vb.W2 : vb.W1, bA.as<u16>(), kp);
    c->launches += 3;
    if (c->use_ntt) {
        ConvArgs cv{}; cv.A = bA.as<u16>(); cv.C = bC.as<u16>(); cv.lda = kp; cv.ldc = ldo; cv.mtotal = (int)n; cv.rpp = (int)n; cv.a_slots = (int)n; cv.c_slots = (int)n;
        cv.tw = c->sn.tw; cv.khat = sn_kh_m256(c->sn, degree2 ? 8 : 5); cv.post = reinterpret_cast<const int2 *>(degree2 ? vb.PT2 : vb.PT1); cv.post_group = degree2 ? 256 : LM1_ROWS;
        c->launches += degree2 ? conv_ntt_launch<8, 2, KP2, 256, false, true>(cv, st) : conv_ntt_launch<5, 4, KP1, D1, false, true>(cv, st);
    } else {
        GemmArgs g{}; g.ws = vb.WS; g.ws_elems = GE_WS_ELEMS; g.A = bA.as<u16>(); g.Bt = degree2 ? c->d_U2 : c->d_U1; g.C = bC.as<u16>(); g.lda = kp; g.ldb = kp; g.ldc = ldo;
        g.mtotal = (int)n; g.ksteps = kp / GE_BK; g.nvalid = nt; g.rpp = (int)n; g.a_slots = (int)n; g.c_slots = (int)n;
        g.colscale = degree2 ? vb.PT2 : vb.PT1; g.colscale_batch = degree2 ? 256 : LM1_ROWS; g.colscale_by_group = 1;
        c->launches += degree2 ? gf_gemm_launch_auto<8>(g, 256, 1, st) : gf_gemm_launch_auto<7>(g, U1_ROWS, 1, st);
    }
    if (!degree2) { kc_node_targets<<<(unsigned)n, 160, 0, st>>>(bs.as<u16>(), vb.POS, bC.as<u16>()); c->launches++; }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    std::vector<u16> ho(n * ldo);
    CU(cudaMemcpy(ho.data(), bC.p, ho.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < n; r++) memcpy(out + r * nt, &ho[r * ldo], (size_t)nt * 2);
    return KOSK_OK;
}

}  // extern "C"
