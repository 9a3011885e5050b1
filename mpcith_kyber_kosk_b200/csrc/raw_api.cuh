// raw_api.cuh -- struct-level API of the reference: prepare_randomness / prepare_range_proof
// (mlwe_prover.cpp:4-59), kyber_keygen with the raw MLWE instance (kosk.cpp:4-70), prove (mlwe_prover.cpp:81-538) and
// verify (mlwe_verifier.cpp:4-686) on the reference's own structs, as main.cpp:16-59 uses them.  The structs cross the
// C ABI as byte images with the reference's layout (x86-64, no padding):
//   mlwe_inst          int16  A[K][K][256] | t[K][256] | s[K][256] | e[K][256]                    (mlwe_prover.hpp:34-37)
//   mpcith_randomness  u16 f[F][256] | u16 NTT_f[F][256] | share_vec f_shares[F] | share_vec NTT_f_shares[F]   (:39-44)
//   mpcith_range_proof share_vec s_eta_shares[K][E] | share_vec e_eta_shares[K][E]                 (:46-49)
//   share_vec          size_t len | u16 share_x[1454] | u16 share_y[1454]  = 5824 bytes            (ss.hpp:33-37)
//   mpcith_proof       the proof bytes themselves (encode_mpcith_proof is a memcpy, mlwe_prover.cpp:540-543)
// The kernels below only convert between these images and the plane / Y-row layout of the batch pipeline (B = 1);
// all sharing, hashing and checking runs in the same kernels as kyber_verifiable_keygen / kyber_kosk_verify.
// Included by kosk_b200.cu after the context definition.
#pragma once

namespace kosk {

constexpr size_t SHARE_VEC_BYTES = 8 + 2 * 2 * (size_t)NP;     // 5824

struct RawSizes { size_t inst, rand, eta, rand_f, rand_Tf, rand_fsh, rand_Tfsh; };
KOSK_HD RawSizes raw_sizes(int k)
{
    const Slots sl = make_slots(k);
    RawSizes r;
    r.inst = (size_t)(k * k + 3 * k) * 256 * 2;
    r.rand_f = 0; r.rand_Tf = (size_t)sl.F * 512; r.rand_fsh = 2 * (size_t)sl.F * 512; r.rand_Tfsh = r.rand_fsh + (size_t)sl.F * SHARE_VEC_BYTES;
    r.rand = r.rand_Tfsh + (size_t)sl.F * SHARE_VEC_BYTES;
    r.eta = 2 * (size_t)k * sl.E * SHARE_VEC_BYTES;
    return r;
}

// planes [slot_lo, slot_lo + n) of proof 0 -> n consecutive share_vec images (share_x[p] = p + 256, ss.cpp:29)
__global__ void __launch_bounds__(256) k_export_share_vecs(ProveBufs pb, int nslot, int slot_lo, u8 *out)
{
    const u16 *src = pb.SH + (size_t)(slot_lo + blockIdx.x) * SLD + SOFF;
    u8 *sv = out + (size_t)blockIdx.x * SHARE_VEC_BYTES;
    if (threadIdx.x < 4) reinterpret_cast<u16 *>(sv)[threadIdx.x] = 0;      // len: never set by the reference (ss.hpp:34)
    u16 *sx = reinterpret_cast<u16 *>(sv + 8), *sy = sx + NP;
    for (int p = threadIdx.x; p < NP; p += 256) { sx[p] = (u16)(p + NL); sy[p] = src[p]; }
}
// secrets (first 256 elements) of Y rows [slot_lo, slot_lo + n) of proof 0 -> out[n][256]
__global__ void __launch_bounds__(256) k_export_secrets(ProveBufs pb, int n2, int slot_lo, u16 *out)
{
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = pb.Y[(size_t)(slot_lo + blockIdx.x) * YLD + threadIdx.x];
}
// mlwe_inst image from the state k_keygen left for proof 0: A-hat, t-hat (from pk), s, e (kosk.cpp:22-54)
template <int K>
__global__ void __launch_bounds__(256) k_export_inst(ProveBufs pb, int16_t *inst)
{
    const Slots sl = make_slots(K);
    const int c = threadIdx.x;
    for (int ij = 0; ij < K * K; ij++) inst[ij * 256 + c] = (int16_t)pb.AH[ij * 256 + c];          // gen_matrix output, [0, q)
    int16_t *t = inst + K * K * 256, *s = t + K * 256, *e = s + K * 256;
    for (int i = 0; i < K; i++) {
        const u8 *a = pb.pk + 384 * i + 3 * (c >> 1);                                              // poly_tobytes layout
        const uint32_t v = (c & 1) ? (((uint32_t)a[1] >> 4) | ((uint32_t)a[2] << 4)) & 0xFFF : ((uint32_t)a[0] | ((uint32_t)a[1] << 8)) & 0xFFF;
        t[i * 256 + c] = (int16_t)gf_center(v);                                                    // polyvec_reduce: centered representative
        s[i * 256 + c] = (int16_t)gf_center(yrow(pb, sl, 0, sl.s0 + i)[c]);                        // CBD output in [-eta, eta]
        e[i * 256 + c] = (int16_t)gf_center(yrow(pb, sl, 0, sl.e0 + i)[c]);
    }
}

// mpcith_randomness + mpcith_range_proof images -> planes and Y rows of proof 0 (what PH_OFFLINE would have left there).
// grid = 2F + 2KE blocks (one per sharing).
template <int K>
__global__ void __launch_bounds__(256) k_import_pre(ProveBufs pb, const u8 *rand, const u8 *eta)
{
    const Slots sl = make_slots(K);
    const RawSizes rs = raw_sizes(K);
    const int j = blockIdx.x, tid = threadIdx.x;
    int slot; const u8 *sv; const u16 *sec = nullptr;
    if (j < sl.F) { slot = sl.f0 + j; sv = rand + rs.rand_fsh + (size_t)j * SHARE_VEC_BYTES; sec = reinterpret_cast<const u16 *>(rand + rs.rand_f) + j * 256; }
    else if (j < 2 * sl.F) { slot = sl.Tf0 + (j - sl.F); sv = rand + rs.rand_Tfsh + (size_t)(j - sl.F) * SHARE_VEC_BYTES; sec = reinterpret_cast<const u16 *>(rand + rs.rand_Tf) + (j - sl.F) * 256; }
    else { slot = sl.seta0 + (j - 2 * sl.F); sv = eta + (size_t)(j - 2 * sl.F) * SHARE_VEC_BYTES; }
    const u16 *sy = reinterpret_cast<const u16 *>(sv + 8) + NP;
    u16 *pl = plane(pb, sl, 0, slot), *y = yrow(pb, sl, 0, slot);
    for (int p = tid; p < NP; p += 256) pl[p] = sy[p];
    if (sec) y[tid] = sec[tid];
    else { const int m = (slot - sl.seta0) % sl.E; y[tid] = (u16)((m - sl.eta + Q) % Q); }           // mlwe_prover.cpp:42-48
    for (int c = tid; c < YLD - 256; c += 256) y[256 + c] = c <= NT ? sy[c] : 0;                      // parties 0..150 hold the tail verbatim
}

// mlwe_inst image -> what k_keygen leaves for the online phase: A-hat, s-hat, the secrets of [s], [e] and of the z_j products
template <int K>
__global__ void __launch_bounds__(128) k_import_inst(ProveBufs pb, const int16_t *inst)
{
    constexpr int ETA = (K == 2) ? 3 : 2, M = 2 * ETA;
    const Slots sl = make_slots(K);
    const int tid = threadIdx.x;
    __shared__ u16 sSh[K][256];
    auto enc = [](int16_t v) -> uint32_t { int32_t r = (int32_t)v % Q; return (uint32_t)(r < 0 ? r + Q : r); };      // encode_to_gf3329 (gf3329.c:308-310) for in-range input
    const int16_t *A = inst, *s = inst + (K * K + K) * 256, *e = s + K * 256;
    for (int i = tid; i < K * K * 256; i += 128) pb.AH[i] = (u16)enc(A[i]);
    for (int c = tid; c < 256; c += 128)
        for (int i = 0; i < K; i++)
            for (int w = 0; w < 2; w++) {
                const uint32_t x = enc(w ? e[i * 256 + c] : s[i * 256 + c]);
                yrow(pb, sl, 0, (w ? sl.e0 : sl.s0) + i)[c] = (u16)x;
                if (!w) sSh[i][c] = (u16)x;
                uint32_t z = gf_sub(x, gf_sub(0, ETA));
                for (int j = 0; j < M; j++) {
                    const uint32_t eta_m = (j + 1 >= ETA) ? (uint32_t)(j + 1 - ETA) : (uint32_t)(Q + j + 1 - ETA);
                    z = gf_mul(z, gf_sub(x, eta_m));
                    yrow(pb, sl, 0, (w ? sl.ze0 : sl.zs0) + i * M + j)[c] = (u16)z;
                }
            }
    __syncthreads();
    for (int i = 0; i < K; i++) ntt256_block(sSh[i], tid);
    for (int i = tid; i < K * 256; i += 128) pb.SHAT[i] = (&sSh[0][0])[i];
}

}  // namespace kosk

// ---- host side ----
struct RawState {
    uint8_t seed[32] = {0};
    uint32_t calls = 0;
    u8 *d_rand = nullptr, *d_eta = nullptr; int16_t *d_inst = nullptr;
};

static int raw_prepare(kosk_b200_ctx *c, RawState &rs)
{
    const RawSizes sz = raw_sizes(c->k);
    if (!rs.d_rand) {
        if (cudaMalloc((void **)&rs.d_rand, sz.rand) != cudaSuccess || cudaMalloc((void **)&rs.d_eta, sz.eta) != cudaSuccess ||
            cudaMalloc((void **)&rs.d_inst, sz.inst) != cudaSuccess) return fail(KOSK_E_NOMEM, "cudaMalloc failed for the struct staging buffers");
    }
    return KOSK_OK;
}
static void raw_free(RawState &rs)
{
    void *p[] = {rs.d_rand, rs.d_eta, rs.d_inst};
    for (void *q : p) if (q) cudaFree(q);
    rs.d_rand = rs.d_eta = nullptr; rs.d_inst = nullptr;
}

template <int K> static void launch_export_inst(const ProveBufs &pb, int16_t *d, cudaStream_t st) { k_export_inst<K><<<1, 256, 0, st>>>(pb, d); }
template <int K> static void launch_import_pre(const ProveBufs &pb, const u8 *r, const u8 *e, int nblk, cudaStream_t st) { k_import_pre<K><<<nblk, 256, 0, st>>>(pb, r, e); }
template <int K> static void launch_import_inst(const ProveBufs &pb, const int16_t *d, cudaStream_t st) { k_import_inst<K><<<1, 128, 0, st>>>(pb, d); }
template <int K> static void launch_expand(const ProveBufs &pb, const Slots &sl, cudaStream_t st)
{
    k_expand_f<K><<<(sl.F + 63) / 64, 64, 0, st>>>(pb);
    k_ntt_f<K><<<dim3(sl.F, 1), 128, 0, st>>>(pb);
}
template <int K> static void launch_tails(const ProveBufs &pb, const Slots &sl, cudaStream_t st) { k_tails<K><<<(sl.n1 + K + 63) / 64, 64, 0, st>>>(pb); }
template <int K> static void launch_keygen(const ProveBufs &pb, cudaStream_t st) { k_keygen<K><<<1, 128, 0, st>>>(pb); }
#define RAW_DISPATCH(fn, ...) do { switch (c->k) { case 2: fn<2>(__VA_ARGS__); break; case 3: fn<3>(__VA_ARGS__); break; default: fn<4>(__VA_ARGS__); } } while (0)

// one-proof view of lane 0's scratch with the seed uploaded and the DRBG call bases of this call
static int raw_begin(kosk_b200_ctx *c, Lane *&ln, ProveBufs &pb)
{
    CU(cudaSetDevice(c->device));
    // the DRBG addresses randombytes() calls with LE32(call number): refuse to come near the wrap instead of replaying randomness
    if (c->raw->calls > 0x7FF00000u) return fail(KOSK_E_ARG, "DRBG call counter exhausted: re-seed with kosk_b200_rng_reset");
    int rc = raw_prepare(c, *c->raw); if (rc) return rc;
    ln = &c->lanes[0];
    CU(cudaStreamSynchronize(ln->st));
    CU(cudaMemcpyAsync(ln->d_seeds, c->raw->seed, 32, cudaMemcpyHostToDevice, ln->st));
    pb = ln->pb; pb.seeds = ln->d_seeds; pb.pk = ln->d_pk; pb.sk = ln->d_sk; pb.pi = ln->d_pi; pb.B = 1;
    return KOSK_OK;
}

extern "C" {

size_t kosk_b200_inst_bytes(int k) { return (k >= 2 && k <= 4) ? raw_sizes(k).inst : 0; }
size_t kosk_b200_randomness_bytes(int k) { return (k >= 2 && k <= 4) ? raw_sizes(k).rand : 0; }
size_t kosk_b200_range_proof_bytes(int k) { return (k >= 2 && k <= 4) ? raw_sizes(k).eta : 0; }

int kosk_b200_rng_reset(kosk_b200_ctx *c, const uint8_t seed[32])
{
    if (!c || !seed) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    memcpy(c->raw->seed, seed, 32); c->raw->calls = 0;
    return KOSK_OK;
}
uint32_t kosk_b200_rng_calls(const kosk_b200_ctx *c) { return c ? c->raw->calls : 0; }

// kyber_verifiable_keygen (kosk.cpp:72-86) drawing from the context DRBG at its current call number, like every other
// function of this file (the seeded form kosk_b200_verifiable_keygen starts a fresh DRBG per proof instead)
int kosk_b200_verifiable_keygen_rng(kosk_b200_ctx *c, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    if (!c || !pk || !sk || !pi) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    Lane *ln; ProveBufs pb; int rc = raw_begin(c, ln, pb); if (rc) return rc;
    RawState &rs = *c->raw;
    const int base = (int)rs.calls;
    pb.cb_key += base; pb.cb_rand += base; pb.cb_eta += base; pb.cb_prove += base;
    rc = prove_chunk_k(c, *ln, pb, 1, ln->d_seeds, ln->d_pk, ln->d_sk, ln->d_pi, PH_OFFLINE | PH_ONLINE);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pk, ln->d_pk, c->L.pk_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaMemcpyAsync(sk, ln->d_sk, c->L.sk_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaMemcpyAsync(pi, ln->d_pi, c->L.proof_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaStreamSynchronize(ln->st));
    rs.calls += (uint32_t)c->sl.ncalls;
    return KOSK_OK;
}

int kosk_b200_prepare_randomness(kosk_b200_ctx *c, void *rand_image)
{
    if (!c || !rand_image) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    Lane *ln; ProveBufs pb; int rc = raw_begin(c, ln, pb); if (rc) return rc;
    const Slots &sl = c->sl; const RawSizes sz = raw_sizes(c->k); RawState &rs = *c->raw;
    pb.cb_rand = (int)rs.calls; pb.tails_mask = 1;
    RAW_DISPATCH(launch_expand, pb, sl, ln->st);
    RAW_DISPATCH(launch_tails, pb, sl, ln->st);
    launch_share_eval(c, pb.Y, pb.SH, sl.f0, 2 * sl.F, sl.n2, sl.nslot, 1, ln->st, false, nullptr, nullptr, pb.WS);
    k_export_secrets<<<2 * sl.F, 256, 0, ln->st>>>(pb, sl.n2, sl.f0, reinterpret_cast<u16 *>(rs.d_rand));
    k_export_share_vecs<<<2 * sl.F, 256, 0, ln->st>>>(pb, sl.nslot, sl.f0, rs.d_rand + sz.rand_fsh);
    c->launches += 5;
    CU(cudaMemcpyAsync(rand_image, rs.d_rand, sz.rand, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaStreamSynchronize(ln->st));
    rs.calls += 3 * sl.F;                      // F seeds + 2F sharings (mlwe_prover.cpp:8-38)
    return KOSK_OK;
}

int kosk_b200_prepare_range_proof(kosk_b200_ctx *c, void *eta_image)
{
    if (!c || !eta_image) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    Lane *ln; ProveBufs pb; int rc = raw_begin(c, ln, pb); if (rc) return rc;
    const Slots &sl = c->sl; const RawSizes sz = raw_sizes(c->k); RawState &rs = *c->raw;
    pb.cb_eta = (int)rs.calls; pb.tails_mask = 2;
    RAW_DISPATCH(launch_tails, pb, sl, ln->st);
    launch_share_eval(c, pb.Y, pb.SH, sl.seta0, 2 * c->k * sl.E, sl.n2, sl.nslot, 1, ln->st, true, nullptr, nullptr, pb.WS);
    k_export_share_vecs<<<2 * c->k * sl.E, 256, 0, ln->st>>>(pb, sl.nslot, sl.seta0, rs.d_eta);
    c->launches += 2;
    CU(cudaMemcpyAsync(eta_image, rs.d_eta, sz.eta, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaStreamSynchronize(ln->st));
    rs.calls += 2 * c->k * sl.E;               // mlwe_prover.cpp:41-59
    return KOSK_OK;
}

int kosk_b200_keygen(kosk_b200_ctx *c, uint8_t *pk, uint8_t *sk, void *inst_image)
{
    if (!c || !pk || !sk) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    Lane *ln; ProveBufs pb; int rc = raw_begin(c, ln, pb); if (rc) return rc;
    const RawSizes sz = raw_sizes(c->k); RawState &rs = *c->raw;
    pb.cb_key = (int)rs.calls;
    RAW_DISPATCH(launch_keygen, pb, ln->st); c->launches++;
    if (inst_image) {
        RAW_DISPATCH(launch_export_inst, pb, rs.d_inst, ln->st); c->launches++;
        CU(cudaMemcpyAsync(inst_image, rs.d_inst, sz.inst, cudaMemcpyDeviceToHost, ln->st));
    }
    CU(cudaMemcpyAsync(pk, ln->d_pk, c->L.pk_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaMemcpyAsync(sk, ln->d_sk, c->L.sk_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaStreamSynchronize(ln->st));
    rs.calls += 1;                             // randombytes(buf, 64), kosk.cpp:11
    return KOSK_OK;
}

int kosk_b200_prove(kosk_b200_ctx *c, uint8_t *pi, const void *inst_image, const void *rand_image, const void *eta_image)
{
    if (!c || !pi || !inst_image || !rand_image || !eta_image) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    Lane *ln; ProveBufs pb; int rc = raw_begin(c, ln, pb); if (rc) return rc;
    const Slots &sl = c->sl; const RawSizes sz = raw_sizes(c->k); RawState &rs = *c->raw;
    CU(cudaMemcpyAsync(rs.d_rand, rand_image, sz.rand, cudaMemcpyHostToDevice, ln->st));
    CU(cudaMemcpyAsync(rs.d_eta, eta_image, sz.eta, cudaMemcpyHostToDevice, ln->st));
    CU(cudaMemcpyAsync(rs.d_inst, inst_image, sz.inst, cudaMemcpyHostToDevice, ln->st));
    pb.cb_prove = (int)rs.calls; pb.tails_mask = 4;
    RAW_DISPATCH(launch_import_pre, pb, rs.d_rand, rs.d_eta, 2 * sl.F + 2 * c->k * sl.E, ln->st);
    RAW_DISPATCH(launch_import_inst, pb, rs.d_inst, ln->st);
    RAW_DISPATCH(launch_tails, pb, sl, ln->st);
    c->launches += 3;
    rc = prove_chunk_k(c, *ln, pb, 1, ln->d_seeds, ln->d_pk, ln->d_sk, ln->d_pi, PH_ONLINE | PH_NOKEYGEN);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pi, ln->d_pi, c->L.proof_bytes, cudaMemcpyDeviceToHost, ln->st));
    CU(cudaStreamSynchronize(ln->st));
    rs.calls += 3 * c->k + 2 * c->k * sl.M;    // 2K (s, e) + K ([A s]) + 2K*2eta (z chains), SURVEY Appendix C
    return KOSK_OK;
}

int kosk_b200_verify(kosk_b200_ctx *c, const uint8_t *pi, const void *inst_image)
{
    if (!c || !pi || !inst_image) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    Lane &ln = c->lanes[0];
    const int K = c->k;
    // A and t as the verifier consumes them: encode_to_gf3329 of the int16 coefficients (mlwe_verifier.cpp:289, :358)
    const int16_t *inst = static_cast<const int16_t *>(inst_image);
    std::vector<u16> at((size_t)(K * K + K) * 256);
    // encode_to_gf3329 (gf3329.c:308-310) only adds q to negative values.  A enters products, where the reference's u16 % q arithmetic makes
    // the reduction of that u16 neutral; t is COMPARED with recomputed canonical shares (mlwe_verifier.cpp:358-376), so it must stay
    // unreduced: an instance whose t has non-canonical coefficients is rejected, as in the reference.
    const size_t nA = (size_t)K * K * 256;
    for (size_t i = 0; i < at.size(); i++) {
        const int v = inst[i];
        const u16 enc = (u16)(v < 0 ? v + Q : v);
        at[i] = i < nA ? (u16)(enc % Q) : enc;
    }
    CU(cudaStreamSynchronize(ln.st));
    CU(cudaMemcpyAsync(ln.vb.AH, at.data(), (size_t)K * K * 512, cudaMemcpyHostToDevice, ln.st));
    CU(cudaMemcpyAsync(ln.vb.TPK, at.data() + (size_t)K * K * 256, (size_t)K * 512, cudaMemcpyHostToDevice, ln.st));
    CU(cudaMemcpyAsync(ln.d_pi, pi, c->L.proof_bytes, cudaMemcpyHostToDevice, ln.st));
    ln.vb.raw_inst = 1;
    int rc = verify_chunk_lane(c, ln, 1, ln.d_pi, ln.d_pk, ln.d_ok);
    ln.vb.raw_inst = 0;
    if (rc) return rc;
    uint8_t ok = 0;
    CU(cudaMemcpyAsync(&ok, ln.d_ok, 1, cudaMemcpyDeviceToHost, ln.st));
    CU(cudaStreamSynchronize(ln.st));
    return (int)ok;
}

}  // extern "C"
