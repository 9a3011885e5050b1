// kosk_b200.cu -- context, host orchestration and the C ABI (include/kosk_b200.h) of the B200 KOSK core.
// Host side is C++ because the reference's host side is C++ (kosk.cpp); nothing here runs field or hash
// arithmetic on the CPU except the one-time construction of the Lagrange tables at context creation.
#include "../../include/kosk_b200.h"
#include "prove_kernels.cuh"
#include "verify_kernels.cuh"
#include "gf_gemm_imma.cuh"
#include "share_ntt.cuh"
#include "wire_kernels.cuh"
#include "wire_host.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <memory>
#include <atomic>
#include <chrono>
#include <thread>
#include <algorithm>
#include <utility>
#include <map>

using namespace kosk;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(KOSK_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// inside kosk_b200_create_ex after the context exists: release it (and every table allocated so far) before reporting the error
#define CUC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { ctx_free(c); return fail(KOSK_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
    } while (0)

// ---- one-time host table construction (the reference's missing utils/precomputed_kyber.c) ----
static uint32_t h_pow(uint32_t b, uint32_t e) { uint32_t r = 1; b %= Q; while (e) { if (e & 1) r = r * b % Q; b = b * b % Q; e >>= 1; } return r; }
// out[t*n + j] = l_j^{nodes}(targets[t]) mod q (barycentric form); nodes distinct mod q, no target equals a node
static void h_lagrange(std::vector<uint16_t> &out, const std::vector<int> &nodes, const std::vector<int> &targets)
{
    const int n = (int)nodes.size(), nt = (int)targets.size();
    std::vector<uint32_t> w(n);
    for (int k = 0; k < n; k++) {
        uint32_t d = 1;
        for (int m = 0; m < n; m++) if (m != k) d = d * ((uint32_t)(nodes[k] - nodes[m] + 2 * Q) % Q) % Q;
        w[k] = h_pow(d, Q - 2);
    }
    out.assign((size_t)n * nt, 0);
    for (int t = 0; t < nt; t++) {
        uint32_t full = 1;
        for (int m = 0; m < n; m++) full = full * ((uint32_t)(targets[t] - nodes[m] + 2 * Q) % Q) % Q;
        for (int k = 0; k < n; k++)
            out[(size_t)t * n + k] = (uint16_t)(full * w[k] % Q * h_pow((uint32_t)(targets[t] - nodes[k] + 2 * Q) % Q, Q - 2) % Q);
    }
}

enum { PH_OFFLINE = 1, PH_ONLINE = 2, PH_NOKEYGEN = 4 /* raw prove(): the instance comes from the caller */ };
struct RawState;
static RawState *raw_new();
static void raw_delete(RawState *);
struct KemState;
static KemState *kem_new();
static void kem_delete(KemState *);
enum { KOSK_PH_KEYGEN = 0, KOSK_PH_EXPAND, KOSK_PH_SHARE1, KOSK_PH_COMMIT, KOSK_PH_FS1, KOSK_PH_EVAL, KOSK_PH_OPEN, KOSK_PH_SHARE2,
       KOSK_PH_VIEW, KOSK_PH_FS2, KOSK_PH_ASSEMBLE, KOSK_PH_VERIFY, KOSK_NPHASE };

// One pipeline lane: a stream plus the per-chunk scratch and staging buffers of one sub-batch.  Consecutive sub-batches
// alternate over the lanes; an event chain keeps their kernels back to back (co-running them was measured to be a loss)
// while the host copies of one sub-batch overlap the kernels of the next.
struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;           // everything enqueued on the lane so far (join with the caller's stream)
    cudaEvent_t computed = nullptr;       // the kernels of the lane's latest sub-batch (the next sub-batch's kernels wait on it)
    cudaEvent_t pre_tail = nullptr;       // ... up to its view hashes: FS-2 and the assembly may overlap the next sub-batch (tail overlap)
    ProveBufs pb{};
    u8 *d_seeds = nullptr, *d_pk = nullptr, *d_sk = nullptr, *d_pi = nullptr, *d_ok = nullptr;   // staging of the host-buffer API
    VerifyBufs vb{};
    VerifySide vside{};                     // second stream of the verifier (challenge-independent chain), KOSK_B200_VERIFY_SIDE=0 disables it
    std::vector<cudaEvent_t> ev; int ev_used = 0;
    std::vector<std::pair<int, int>> ev_phase;     // (phase id, event index of its start); end = next event
    // compact wire format (wire_kernels.cuh / wire_host.h): device and pinned host staging of the packed proofs of one sub-batch,
    // one event per D2H slice (the unpack workers wait on it), and the event after the lane's latest H2D from h_wire
    u8 *d_wire = nullptr, *h_wire = nullptr;
    std::vector<cudaEvent_t> wev;
    std::unique_ptr<std::atomic<int>[]> wctr;      // per slice: unpack jobs still reading that slice of h_wire
    cudaEvent_t h2d_done = nullptr; bool h2d_pending = false;
    volatile int wire_flag = 0;                    // set by a worker whose event wait failed
};

struct kosk_b200_ctx {
    int k = 0, device = 0, chunk = 0, use_tensor = 0, fuse_fs = 0, fuse_max = 128;
    int *d_status = nullptr;               // device word set by a kernel that gave up waiting (never expected)
    size_t next_lane = 0;
    cudaEvent_t last_computed = nullptr;   // compute-done event of the most recently enqueued sub-batch (any lane)
    cudaEvent_t last_gate = nullptr;       // event the next prove sub-batch waits for: last_computed, or the previous sub-batch's pre_tail
    int overlap_fs1 = 0;                   // KOSK_B200_OVERLAP_FS1=1: eta / z_j sharings on the lane's side stream next to commit hashes + FS-1 (measured: a loss, off)
    int overlap_tail = 1;                  // KOSK_B200_OVERLAP_TAIL: let a sub-batch start while the previous one runs FS-2 + assembly
    int use_ntt = 2;                       // KOSK_B200_SHARE_NTT: share evaluation as a blocked NTT convolution (share_ntt.cuh; 2 = k_share_ntt2, 1 = the generic equal-block kernel) instead of the dense table GEMM (0)
    uint8_t *d_sn = nullptr;               // its tables (ShareNttTables), one allocation
    ShareNttTables sn{};
    std::map<cudaStream_t, SnTicket> sn_tickets;   // work tickets of k_share_ntt2, one device counter per stream that launches it (created on first use)
    Slots sl; Layout L;
    uint64_t launches = 0;
    // constant tables
    int16_t *d_St = nullptr;               // [GE_NPAD][YLD] centered share table S (zero padded)
    int8_t *tmpL0 = nullptr, *tmpL1 = nullptr; size_t tmp_rows = 0;
    int8_t *d_St0 = nullptr, *d_St1 = nullptr;   // experimental tensor path: 7-bit limb planes of S, [GE_NPAD][YLD] int8
    int16_t *d_SU = nullptr;               // [GE_NPAD] centered U[x] = sum_{j<256} S[x][j]: share of the all-ones secret vector
    int16_t *d_R1 = nullptr, *d_R2 = nullptr; // verifier: centered recon tables [256][YLD], [256][VR2LD]
    int16_t *d_U1 = nullptr, *d_U2 = nullptr; // verifier: Cauchy interpolation operands 1 / (t - (p + 256)), [U1_ROWS][KP1], [256][KP2]
    u16 *d_inv = nullptr;                  // [3329] inverses
    u16 *d_fact = nullptr;                 // [2][FACT_N] factorials and inverse factorials mod q (verifier's Lagrange weights)
    int16_t *d_tab_commit = nullptr, *d_tab_view = nullptr;
    std::vector<Lane> lanes;
    cudaEvent_t ev_start = nullptr;
    // optional phase timing with CUDA events on the launching stream (bench.py roofline)
    bool prof = false;
    double phase_ms[KOSK_NPHASE] = {0}; uint64_t phase_calls[KOSK_NPHASE] = {0};
    RawState *raw = nullptr;               // struct-level API (raw_api.cuh): DRBG state and staging buffers
    KemState *kem = nullptr;               // Kyber KEM encaps / decaps (kem_kernels.cuh)
    // Host-buffer batch calls: wire_mode = percentage of the proofs that cross the link in the compact wire format (packed / unpacked on the
    // device, expanded / packed by the worker pool on the host); the others cross it as struct mpcith_proof bytes straight between the
    // caller's buffer and the device (round-1 behaviour = 0).  The link favours 100, a host short of memory bandwidth something lower.
    int wire_mode = 100, wire_threads = 0, wire_slice = 16, wire_acc = 0;
    WirePool *wpool = nullptr;
    int live_pools = 0;                    // preprocessing pools created from this context and not yet destroyed
    std::recursive_mutex mu;               // one caller at a time: the lanes' scratch and the DRBG state are per context
};
#define LOCK(c) std::lock_guard<std::recursive_mutex> lock_((c)->mu)

static void prof_mark(kosk_b200_ctx *c, Lane &ln, int phase, cudaStream_t on = nullptr)
{
    if (!c->prof) return;
    if (ln.ev_used >= (int)ln.ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ln.ev.push_back(e); }
    cudaEventRecord(ln.ev[ln.ev_used], on ? on : ln.st);
    ln.ev_phase.push_back({phase, ln.ev_used});
    ln.ev_used++;
}
static void prof_collect(kosk_b200_ctx *c)
{
    for (Lane &ln : c->lanes) {
        for (size_t i = 0; i + 1 < ln.ev_phase.size(); i++) {
            const int ph = ln.ev_phase[i].first;
            if (ph < 0) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ln.ev[ln.ev_phase[i].second], ln.ev[ln.ev_phase[i + 1].second]) == cudaSuccess) { c->phase_ms[ph] += ms; c->phase_calls[ph]++; }
        }
        ln.ev_phase.clear(); ln.ev_used = 0;
    }
}

static void free_prove_bufs(ProveBufs &pb)
{
    void *lp[] = {pb.Y, pb.SH, pb.BG, pb.TCR, pb.VWR, pb.PW, pb.AH, pb.SHAT, pb.I, pb.REST, pb.YL0, pb.YL1, pb.WS};
    for (void *p : lp) if (p) cudaFree(p);
    pb = ProveBufs{};
}
static void scrub_prove_bufs(ProveBufs &pb, const Slots &sl, int k, size_t B)
{
    if (pb.Y) cudaMemset(pb.Y, 0, B * sl.n2 * YLD * 2);
    if (pb.SH) cudaMemset(pb.SH, 0, B * sl.nslot * SLD * 2);
    if (pb.SHAT) cudaMemset(pb.SHAT, 0, B * k * 256 * 2);
}
static int alloc_prove_bufs(ProveBufs &pb, const Slots &sl, int k, size_t B, bool tensor)
{
#define PA(ptr, bytes, zero) do { if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) { free_prove_bufs(pb); return -1; } if (zero) cudaMemset((ptr), 0, (bytes)); } while (0)
    PA(pb.Y, B * sl.n2 * YLD * 2, 1);          // zero: row padding (terms 407..415) must stay 0
    PA(pb.SH, B * sl.nslot * SLD * 2, 1);
    PA(pb.BG, B * NP * 2 * BGH * 2, 0);
    PA(pb.TCR, B * TREE_BYTES, 0); PA(pb.VWR, B * TREE_BYTES, 0);
    PA(pb.PW, B * (MK + 2 * k) * sl.F * 2, 0);
    PA(pb.AH, B * k * k * 256 * 2, 0); PA(pb.SHAT, B * k * 256 * 2, 0);
    PA(pb.I, B * NT * 2, 0); PA(pb.REST, B * NR * 2, 0);
    if (tensor) { PA(pb.YL0, B * sl.n2 * YLD, 1); PA(pb.YL1, B * sl.n2 * YLD, 1); }
    PA(pb.WS, (size_t)GE_WS_ELEMS * 4, 0);
#undef PA
    set_default_calls(pb, sl);
    return 0;
}

static void ctx_free(kosk_b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    wire_pool_destroy(c->wpool); c->wpool = nullptr;
    void *ptrs[] = {c->d_sn, c->d_U1, c->d_U2, c->d_status, c->d_fact, c->tmpL0, c->tmpL1, c->d_St, c->d_St0, c->d_St1, c->d_SU, c->d_R1, c->d_R2, c->d_inv, c->d_tab_commit, c->d_tab_view};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &kv : c->sn_tickets) if (kv.second.ctr) cudaFree(kv.second.ctr);
    c->sn_tickets.clear();
    const size_t Bc = (size_t)c->chunk;
    for (Lane &ln : c->lanes) {
        // scratch that held seeds, s, e and secret keys is cleared before it goes back to the allocator
        if (ln.d_seeds) cudaMemset(ln.d_seeds, 0, Bc * 32);
        if (ln.d_sk) cudaMemset(ln.d_sk, 0, Bc * c->L.sk_bytes);
        scrub_prove_bufs(ln.pb, c->sl, c->k, Bc);
        free_prove_bufs(ln.pb);
        void *lp[] = {ln.d_seeds, ln.d_pk, ln.d_sk, ln.d_pi, ln.d_ok, ln.d_wire};
        for (void *p : lp) if (p) cudaFree(p);
        if (ln.h_wire) cudaFreeHost(ln.h_wire);
        for (cudaEvent_t e : ln.wev) cudaEventDestroy(e);
        if (ln.h2d_done) cudaEventDestroy(ln.h2d_done);
        verify_free(ln.vb);
        for (cudaEvent_t e : ln.ev) cudaEventDestroy(e);
        if (ln.done) cudaEventDestroy(ln.done);
        if (ln.computed) cudaEventDestroy(ln.computed);
        if (ln.pre_tail) cudaEventDestroy(ln.pre_tail);
        if (ln.vside.fork) cudaEventDestroy(ln.vside.fork);
        if (ln.vside.join) cudaEventDestroy(ln.vside.join);
        if (ln.vside.st) cudaStreamDestroy(ln.vside.st);
        if (ln.st) cudaStreamDestroy(ln.st);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    raw_delete(c->raw);
    kem_delete(c->kem);
    delete c;
}

// device status word: set by a kernel that gave up waiting (never expected); read after the streams are idle
static int check_status(kosk_b200_ctx *c)
{
    int st = 0;
    CU(cudaMemcpy(&st, c->d_status, 4, cudaMemcpyDeviceToHost));
    if (st) return fail(KOSK_E_CUDA, "a kernel gave up waiting for its producer warps (internal error)");
    return KOSK_OK;
}

extern "C" {

size_t kosk_b200_pk_bytes(int k) { return (k >= 2 && k <= 4) ? make_layout(k).pk_bytes : 0; }
size_t kosk_b200_sk_bytes(int k) { return (k >= 2 && k <= 4) ? make_layout(k).sk_bytes : 0; }
size_t kosk_b200_proof_bytes(int k) { return (k >= 2 && k <= 4) ? make_layout(k).proof_bytes : 0; }
const char *kosk_b200_last_error(void) { return g_err.c_str(); }
const char *kosk_b200_version(void) { return "kosk_b200 0.1 (sm_100a)"; }

int kosk_b200_create(kosk_b200_ctx **out, int k, int device, int max_chunk)
{
    const char *e = getenv("KOSK_B200_LANES");
    const char *t = getenv("KOSK_B200_TENSOR");
    return kosk_b200_create_ex(out, k, device, max_chunk, e ? atoi(e) : 0, (t && atoi(t)) ? KOSK_F_TENSOR : 0);
}

int kosk_b200_create_ex(kosk_b200_ctx **out, int k, int device, int max_chunk, int nlanes, int flags)
{
    if (!out || k < 2 || k > 4) return fail(KOSK_E_ARG, "kyber_k must be 2, 3 or 4");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KOSK_E_CUDA, "no CUDA device: the KOSK core has no CPU path");
    if (device < 0 || device >= ndev) return fail(KOSK_E_ARG, "bad device index");
    CU(cudaSetDevice(device));
    kosk_b200_ctx *c = new kosk_b200_ctx;
    c->k = k; c->device = device; c->sl = make_slots(k); c->L = make_layout(k);
    c->raw = raw_new(); c->kem = kem_new();
    c->chunk = max_chunk > 0 ? max_chunk : 1024;
    if (c->chunk > 16384) c->chunk = 16384;
    if (nlanes <= 0) nlanes = 2;
    if (nlanes > 8) nlanes = 8;
    c->use_tensor = (flags & KOSK_F_TENSOR) ? 1 : 0;
    { const char *e = getenv("KOSK_B200_FUSE_FS"); if (e) c->fuse_fs = atoi(e); }
    { const char *e = getenv("KOSK_B200_FUSE_MAX"); if (e) c->fuse_max = atoi(e); }
    { const char *e = getenv("KOSK_B200_OVERLAP_TAIL"); if (e) c->overlap_tail = atoi(e); }
    { const char *e = getenv("KOSK_B200_OVERLAP_FS1"); if (e) c->overlap_fs1 = atoi(e); }
    { const char *e = getenv("KOSK_B200_SHARE_NTT"); if (e) c->use_ntt = atoi(e); }
    if (c->use_tensor) c->use_ntt = 0;
    { const char *e = getenv("KOSK_B200_WIRE"); if (e) c->wire_mode = std::min(100, std::max(0, atoi(e))); }
    { const char *e = getenv("KOSK_B200_WIRE_SLICE"); if (e && atoi(e) > 0) c->wire_slice = atoi(e); }
    const Slots &sl = c->sl; const Layout &L = c->L;
    const size_t B = (size_t)c->chunk;
#define ALLOC(ptr, bytes) do { if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) { ctx_free(c); return fail(KOSK_E_NOMEM, "cudaMalloc failed for " #ptr); } } while (0)
    // ---- tables ----
    {
        uint16_t hz[128];
        for (int i = 0; i < 128; i++) { int br = 0; for (int b = 0; b < 7; b++) br |= ((i >> b) & 1) << (6 - b); hz[i] = (uint16_t)h_pow(17, br); }
        CUC(cudaMemcpyToSymbol(c_zeta, hz, sizeof hz));
        std::vector<int> nodes(D1), targets(NX);
        for (int j = 0; j < D1; j++) nodes[j] = j;
        for (int x = 0; x < NX; x++) targets[x] = x + D1;
        std::vector<uint16_t> S; h_lagrange(S, nodes, targets);
        std::vector<int16_t> St((size_t)GE_NPAD * YLD, 0);
        for (int x = 0; x < NX; x++) for (int j = 0; j < D1; j++) St[(size_t)x * YLD + j] = (int16_t)gf_center(S[(size_t)x * D1 + j]);
        ALLOC(c->d_St, St.size() * 2); CUC(cudaMemcpy(c->d_St, St.data(), St.size() * 2, cudaMemcpyHostToDevice));
        {   // limb planes for the opt-in tensor path: s = 128*s1 + s0, s0 in [-64, 63]
            std::vector<int8_t> L0(St.size()), L1(St.size());
            for (size_t i = 0; i < St.size(); i++) { const int v = St[i], v0 = ((v + 64) & 127) - 64; L0[i] = (int8_t)v0; L1[i] = (int8_t)((v - v0) >> 7); }
            ALLOC(c->d_St0, L0.size()); CUC(cudaMemcpy(c->d_St0, L0.data(), L0.size(), cudaMemcpyHostToDevice));
            ALLOC(c->d_St1, L1.size()); CUC(cudaMemcpy(c->d_St1, L1.data(), L1.size(), cudaMemcpyHostToDevice));
        }
        std::vector<int16_t> SU(GE_NPAD, 0);
        for (int x = 0; x < NX; x++) { uint32_t u = 0; for (int j = 0; j < NL; j++) u = (u + S[(size_t)x * D1 + j]) % Q; SU[x] = (int16_t)gf_center(u); }
        ALLOC(c->d_SU, SU.size() * 2); CUC(cudaMemcpy(c->d_SU, SU.data(), SU.size() * 2, cudaMemcpyHostToDevice));
        // verifier recon tables R1 (256 x 407 over nodes 256..662) and R2 (256 x 813 over nodes 256..1068)
        std::vector<int> n1(D1), n2(D2), tg(NL);
        for (int j = 0; j < D1; j++) n1[j] = 256 + j;
        for (int j = 0; j < D2; j++) n2[j] = 256 + j;
        for (int i = 0; i < NL; i++) tg[i] = i;
        std::vector<uint16_t> R1, R2; h_lagrange(R1, n1, tg); h_lagrange(R2, n2, tg);
        std::vector<int16_t> R1p((size_t)NL * YLD, 0), R2p((size_t)NL * VR2LD, 0);
        for (int i = 0; i < NL; i++) { for (int j = 0; j < D1; j++) R1p[(size_t)i * YLD + j] = (int16_t)gf_center(R1[(size_t)i * D1 + j]);
                                       for (int j = 0; j < D2; j++) R2p[(size_t)i * VR2LD + j] = (int16_t)gf_center(R2[(size_t)i * D2 + j]); }
        ALLOC(c->d_R1, R1p.size() * 2); CUC(cudaMemcpy(c->d_R1, R1p.data(), R1p.size() * 2, cudaMemcpyHostToDevice));
        ALLOC(c->d_R2, R2p.size() * 2); CUC(cudaMemcpy(c->d_R2, R2p.data(), R2p.size() * 2, cudaMemcpyHostToDevice));
        std::vector<uint16_t> inv(Q, 0); for (int a = 1; a < Q; a++) inv[a] = (uint16_t)h_pow(a, Q - 2);
        ALLOC(c->d_inv, Q * 2); CUC(cudaMemcpy(c->d_inv, inv.data(), Q * 2, cudaMemcpyHostToDevice));
        {   // Cauchy operands of the verifier's interpolation (mlwe_verifier.cpp:188-224 etc.): U[t][p] = 1 / (t - (p + 256)), 0 where t is the node
            std::vector<int16_t> U1((size_t)U1_ROWS * KP1, 0), U2((size_t)NL * KP2, 0);
            for (int t = 0; t < D1; t++) for (int p = 0; p < KP1; p++) { const int dd = ((t - p - 256) % Q + Q) % Q; U1[(size_t)t * KP1 + p] = (int16_t)(dd ? gf_center(inv[dd]) : 0); }
            for (int t = 0; t < NL; t++) for (int p = 0; p < KP2; p++) { const int dd = ((t - p - 256) % Q + Q) % Q; U2[(size_t)t * KP2 + p] = (int16_t)gf_center(inv[dd]); }
            ALLOC(c->d_U1, U1.size() * 2); CUC(cudaMemcpy(c->d_U1, U1.data(), U1.size() * 2, cudaMemcpyHostToDevice));
            ALLOC(c->d_U2, U2.size() * 2); CUC(cudaMemcpy(c->d_U2, U2.data(), U2.size() * 2, cudaMemcpyHostToDevice));
        }
        {   // tables of the NTT-convolution share evaluation (share_ntt.cuh)
            const ShareNttHost sh = share_ntt_tables();
            for (int j = 0; j < 16; j++)       // the kernels' compile-time twiddles w16^(+-j) against the host construction
                if (sh.w16f[16 + j] != sn_make_tw16(false).w[j] || sh.w16i[16 + j] != sn_make_tw16(true).w[j]) { ctx_free(c); return fail(KOSK_E_UNSUPPORTED, "share_ntt: DFT table mismatch (internal error)"); }
            // one byte blob, every part 16-byte aligned
            std::vector<uint8_t> all;
            size_t offs[9]; int np = 0;
            auto put = [&](const void *p, size_t bytes) { offs[np++] = all.size(); const uint8_t *q = static_cast<const uint8_t *>(p); all.insert(all.end(), q, q + bytes); while (all.size() % 16) all.push_back(0); };
            put(sh.tw.data(), sh.tw.size() * sizeof(int2)); put(sh.kh_share.data(), sh.kh_share.size() * 2); put(sh.kh_m256.data(), sh.kh_m256.size() * 2);
            put(sh.wj.data(), sh.wj.size() * sizeof(int2)); put(sh.wj2.data(), sh.wj2.size() * sizeof(int2)); put(sh.px.data(), sh.px.size() * sizeof(int2));
            put(sh.pr1.data(), sh.pr1.size() * sizeof(int2)); put(sh.pr2.data(), sh.pr2.size() * sizeof(int2)); put(sh.kp_share.data(), sh.kp_share.size() * 4);
            ALLOC(c->d_sn, all.size()); CUC(cudaMemcpy(c->d_sn, all.data(), all.size(), cudaMemcpyHostToDevice));
            const uint8_t *base = reinterpret_cast<const uint8_t *>(c->d_sn);
            c->sn.tw = reinterpret_cast<const int2 *>(base + offs[0]); c->sn.kh_share = reinterpret_cast<const int16_t *>(base + offs[1]);
            c->sn.kh_m256 = reinterpret_cast<const int16_t *>(base + offs[2]); c->sn.wj = reinterpret_cast<const int2 *>(base + offs[3]);
            c->sn.wj2 = reinterpret_cast<const int2 *>(base + offs[4]); c->sn.px = reinterpret_cast<const int2 *>(base + offs[5]);
            c->sn.pr1 = reinterpret_cast<const int2 *>(base + offs[6]); c->sn.pr2 = reinterpret_cast<const int2 *>(base + offs[7]);
            c->sn.kp_share = reinterpret_cast<const uint32_t *>(base + offs[8]);
        }
        std::vector<uint16_t> fc(2 * FACT_N);
        { uint32_t f = 1; for (int i = 0; i < FACT_N; i++) { if (i) f = f * i % Q; fc[i] = (uint16_t)f; fc[FACT_N + i] = inv[f]; } }
        ALLOC(c->d_fact, fc.size() * 2); CUC(cudaMemcpy(c->d_fact, fc.data(), fc.size() * 2, cudaMemcpyHostToDevice));
        // hashed-record slot tables: commitment (mlwe_prover.cpp:117-126) and view (:398-443)
        std::vector<int16_t> tc, tv;
        for (int j = 0; j < k; j++) tc.push_back((int16_t)(sl.s0 + j));
        for (int j = 0; j < k; j++) tc.push_back((int16_t)(sl.e0 + j));
        for (int j = 0; j < sl.F; j++) tc.push_back((int16_t)(sl.f0 + j));
        for (int j = 0; j < sl.F; j++) tc.push_back((int16_t)(sl.Tf0 + j));
        for (int j = 0; j < 16; j++) tv.push_back((int16_t)(sl.TC0 + j));
        tv.insert(tv.end(), tc.begin(), tc.end());
        for (int j = 0; j < k; j++) tv.push_back((int16_t)(sl.B0 + j));
        for (int j = 0; j < k; j++) tv.push_back((int16_t)(sl.G0 + j));
        for (int j = 0; j < k; j++) tv.push_back((int16_t)(sl.SR0 + j));
        for (int j = 0; j < k; j++) tv.push_back((int16_t)(sl.ER0 + j));
        for (int j = 0; j < k; j++) {
            for (int m = 0; m < sl.M; m++) tv.push_back((int16_t)(sl.zs0 + j * sl.M + m));
            for (int m = 0; m < sl.M; m++) tv.push_back((int16_t)(sl.ze0 + j * sl.M + m));
            for (int m = 0; m < sl.M; m++) tv.push_back((int16_t)(sl.US0 + j * sl.M + m));
            for (int m = 0; m < sl.M; m++) tv.push_back((int16_t)(sl.UE0 + j * sl.M + m));
        }
        ALLOC(c->d_tab_commit, tc.size() * 2); CUC(cudaMemcpy(c->d_tab_commit, tc.data(), tc.size() * 2, cudaMemcpyHostToDevice));
        ALLOC(c->d_tab_view, tv.size() * 2); CUC(cudaMemcpy(c->d_tab_view, tv.data(), tv.size() * 2, cudaMemcpyHostToDevice));
    }
    ALLOC(c->d_status, 4); CUC(cudaMemset(c->d_status, 0, 4));
    // ---- per-lane scratch ----
    c->lanes.resize(nlanes);
    for (Lane &ln : c->lanes) {
        if (alloc_prove_bufs(ln.pb, sl, k, B, c->use_tensor != 0) != 0) { ctx_free(c); return fail(KOSK_E_NOMEM, "cudaMalloc failed for prover scratch"); }
        ALLOC(ln.d_seeds, B * 32); ALLOC(ln.d_pk, B * L.pk_bytes); ALLOC(ln.d_sk, B * L.sk_bytes); ALLOC(ln.d_pi, B * L.proof_bytes); ALLOC(ln.d_ok, B);
        if (verify_alloc(ln.vb, k, c->chunk) != 0) { ctx_free(c); return fail(KOSK_E_NOMEM, "cudaMalloc failed for verifier scratch"); }
        CUC(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
        CUC(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&ln.computed, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&ln.pre_tail, cudaEventDisableTiming));
        { const char *e = getenv("KOSK_B200_VERIFY_SIDE");
          if (!e || atoi(e)) {
              CUC(cudaStreamCreateWithFlags(&ln.vside.st, cudaStreamNonBlocking));
              CUC(cudaEventCreateWithFlags(&ln.vside.fork, cudaEventDisableTiming));
              CUC(cudaEventCreateWithFlags(&ln.vside.join, cudaEventDisableTiming));
          } }
    }
    CUC(cudaEventCreate(&c->ev_start));
    CUC(cudaDeviceSynchronize());
    *out = c;
    return KOSK_OK;
}

void kosk_b200_destroy(kosk_b200_ctx *c)
{
    if (c && c->live_pools > 0) { fail(KOSK_E_ARG, "kosk_b200_destroy: the context still has preprocessing pools; destroy them first (context kept)"); return; }
    ctx_free(c);
}
uint64_t kosk_b200_kernel_launches(const kosk_b200_ctx *c) { return c ? c->launches : 0; }
int kosk_b200_sync(kosk_b200_ctx *c)
{
    if (!c) return KOSK_E_ARG;
    CU(cudaSetDevice(c->device));
    LOCK(c);
    for (Lane &ln : c->lanes) CU(cudaStreamSynchronize(ln.st));
    if (c->wpool) wire_pool_wait_all(c->wpool);       // packed proofs still being expanded into the caller's buffers
    for (Lane &ln : c->lanes) if (ln.wire_flag) { ln.wire_flag = 0; return fail(KOSK_E_CUDA, "a D2H slice of the compact wire path failed"); }
    return check_status(c);
}
int kosk_b200_lanes(const kosk_b200_ctx *c) { return c ? (int)c->lanes.size() : 0; }

}  // extern "C"

// work ticket of the share evaluation for launches on `st` (nullptr if the counter cannot be allocated: static striding)
static SnTicket *sn_ticket(kosk_b200_ctx *c, cudaStream_t st)
{
    auto it = c->sn_tickets.find(st);
    if (it != c->sn_tickets.end()) return &it->second;
    SnTicket tk;
    if (cudaMalloc(&tk.ctr, sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(tk.ctr, 0, sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); cudaFree(tk.ctr); return nullptr; }
    return &(c->sn_tickets[st] = tk);
}

// ---- launch sequence for one chunk of B proofs on one lane ----
static void launch_share_eval(kosk_b200_ctx *c, const u16 *Y, u16 *SH, int slot_lo, int rows, int y_slots, int sh_slots, int B, cudaStream_t st,
                              bool const_secret = false, int8_t *YL0 = nullptr, int8_t *YL1 = nullptr, int32_t *ws = nullptr)
{
    if (rows <= 0) return;
    GemmArgs g{};
    g.ws = ws; g.ws_elems = GE_WS_ELEMS;
    g.A = Y; g.Bt = c->d_St; g.C = SH; g.lda = YLD; g.ldb = YLD; g.ldc = SLD;
    g.mtotal = B * rows; g.ksteps = YLD / GE_BK; g.nvalid = NX; g.c_off = SOFF + NT + 1;
    g.rpp = rows; g.slot_lo = slot_lo; g.a_slots = y_slots; g.c_slots = sh_slots; g.tail = 1; g.tail_off = NL;
    if (c->use_ntt && !(c->use_tensor && YL0)) {      // one warp per sharing; constant-secret rows need no special case here
        c->launches += share_ntt_launch(share_conv_args(g, c->sn), st, c->use_ntt, sn_ticket(c, st));
        return;
    }
    const int koff = const_secret ? NL : 0;
    if (const_secret) {      // eta sharings: contraction over the 151 tail terms only, constant part added in the epilogue
        g.A = Y + NL; g.Bt = c->d_St + NL; g.ksteps = (YLD - NL) / GE_BK; g.tail_off = 0;
        g.addvec = c->d_SU; g.scale_src = Y;
    }
    if (c->use_tensor && YL0) {      // experimental int8 tensor-core path: split the rows into limb planes, then IMMA
        const size_t nchunks = (size_t)B * rows * (YLD / 8);
        k_limb_split<<<(unsigned)((nchunks + 255) / 256), 256, 0, st>>>(Y, YL0, YL1, rows, slot_lo, y_slots, YLD, nchunks);
        ImmaTables tb{c->d_St0 + koff, c->d_St1 + koff, YL0 + koff, YL1 + koff};
        c->launches += 1 + gf_gemm_imma_launch(g, tb, NX, st);
        return;
    }
    g.half_last = 1;         // terms 407..415 of Y rows and of the S table are zero padding
    // latency mode: a handful of proofs does not fill the 148 SMs with 128-row tiles; 64-row tiles halve the time of a tile
    c->launches += gf_gemm_launch_auto<7>(g, GE_NCOLS7, 1, st);
}
// first share evaluation of a prove chunk over slots [lo, hi): the eta-constant sharings [seta0, s0) take the short path
static void launch_share_eval_prove(kosk_b200_ctx *c, const ProveBufs &pb, int lo, int hi, int B, cudaStream_t st)
{
    const Slots &sl = c->sl;
    if (c->use_ntt && !c->use_tensor) {      // the NTT-convolution kernel has no short path to select: one launch over all slots
        launch_share_eval(c, pb.Y, pb.SH, lo, hi - lo, sl.n2, sl.nslot, B, st, false, nullptr, nullptr, pb.WS);
        return;
    }
    const int a_lo = lo, a_hi = std::min(hi, sl.seta0), b_lo = std::max(lo, sl.seta0), b_hi = std::min(hi, sl.s0), c_lo = std::max(lo, sl.s0), c_hi = hi;
    launch_share_eval(c, pb.Y, pb.SH, a_lo, a_hi - a_lo, sl.n2, sl.nslot, B, st, false, pb.YL0, pb.YL1, pb.WS);
    launch_share_eval(c, pb.Y, pb.SH, b_lo, b_hi - b_lo, sl.n2, sl.nslot, B, st, true, pb.YL0, pb.YL1, pb.WS);
    launch_share_eval(c, pb.Y, pb.SH, c_lo, c_hi - c_lo, sl.n2, sl.nslot, B, st, false, pb.YL0, pb.YL1, pb.WS);
}

template <int K>
static int prove_chunk(kosk_b200_ctx *c, Lane &ln, const ProveBufs &bufs, int B, const u8 *d_seeds, u8 *d_pk, u8 *d_sk, u8 *d_pi, int phases)
{
    // phases: PH_OFFLINE = key-independent preprocessing (prepare_randomness + prepare_range_proof, mlwe_prover.cpp:4-59),
    //         PH_ONLINE  = keygen + prove() + encode; both = kyber_verifiable_keygen
    const Slots &sl = c->sl;
    cudaStream_t st = ln.st;
    ProveBufs pb = bufs;
    pb.seeds = d_seeds; pb.pk = d_pk; pb.sk = d_sk; pb.pi = d_pi; pb.B = B;
    constexpr int NCOMMIT = 2 * (K + MK + 2 * K + 1), ETA = (K == 2) ? 3 : 2;
    constexpr int NVIEW = 16 + NCOMMIT + 4 * K + 8 * ETA * K;
    const int ptiles = (NP + 127) / 128;
    const bool off = phases & PH_OFFLINE, on = phases & PH_ONLINE;
    cudaEvent_t tail_gate = nullptr;
    // Lanes pipeline copies against compute, not compute against compute: the kernels of consecutive sub-batches run one
    // after the other (co-running them was measured to slow the latency-bound FS sponges 3x), while the D2H copy of a
    // finished sub-batch overlaps the kernels of the next one on the other lane.
    // Tail overlap: the previous sub-batch's FS-2 sponge (one latency-bound warp per proof, ~20 % of the issue slots) and its
    // assembly (HBM-bound gather) leave the integer pipes to this sub-batch's keygen / expansion / share evaluation.
    if (c->last_gate && c->last_gate != ln.computed && c->last_gate != ln.pre_tail) CU(cudaStreamWaitEvent(st, c->last_gate, 0));
    // Latency mode: the key-independent expansion (PRF, NTT, sharing tails) does not depend on keygen; with a handful of proofs in flight the
    // two run side by side on the lane's two streams (each is a short chain of small kernels) and join before the share evaluation.
    const bool kside = on && off && !(phases & PH_NOKEYGEN) && B <= 16 && ln.vside.st != nullptr;
    cudaStream_t se = kside ? ln.vside.st : st;
    if (kside) { CU(cudaEventRecord(ln.vside.fork, st)); CU(cudaStreamWaitEvent(se, ln.vside.fork, 0)); }
    if (on && !(phases & PH_NOKEYGEN)) {
        prof_mark(c, ln, KOSK_PH_KEYGEN);
        k_keygen<K><<<B, 128, 0, st>>>(pb); c->launches++;
    }
    if (off) {
        prof_mark(c, ln, KOSK_PH_EXPAND);
        k_expand_f<K><<<(B * sl.F + 63) / 64, 64, 0, se>>>(pb);
        k_ntt_f<K><<<dim3(sl.F, B), 128, 0, se>>>(pb);
        k_tails<K><<<(B * (sl.n1 + K) + 63) / 64, 64, 0, se>>>(pb);
        c->launches += 3;
    }
    if (kside) { CU(cudaEventRecord(ln.vside.join, se)); CU(cudaStreamWaitEvent(st, ln.vside.join, 0)); }
    // first share evaluation: slots [0, s0) (f, NTT_f, eta constants) are key-independent, [s0, n1) (s, e, z_j) are not
    const int lo = off ? 0 : sl.s0, hi = on ? sl.n1 : sl.s0;
    prof_mark(c, ln, KOSK_PH_SHARE1);
    // The commitments hash only the s, e, f and NTT_f sharings (mlwe_prover.cpp:117-126): when a whole kyber_verifiable_keygen runs here, the
    // eta-constant and z_j sharings (52 of 206 per proof) are evaluated on the lane's side stream, next to commit hashes -> FS-1 sponge
    // (one latency-bound warp per proof) -> eval -> open, and joined before k_derive reads them.  Opt-in (KOSK_B200_OVERLAP_FS1=1): measured
    // 107.7 k vs 110.6 k proofs/s -- next to the convolution kernel the sponge slows more than the 0.57 ms of share evaluation it hides.
    const bool side = c->overlap_fs1 && off && on && c->use_ntt && !c->use_tensor && ln.vside.st != nullptr;
    if (side) {
        CU(cudaEventRecord(ln.vside.fork, st)); CU(cudaStreamWaitEvent(ln.vside.st, ln.vside.fork, 0));
        launch_share_eval(c, pb.Y, pb.SH, sl.f0, sl.seta0 - sl.f0, sl.n2, sl.nslot, B, st, false, nullptr, nullptr, pb.WS);
        launch_share_eval(c, pb.Y, pb.SH, sl.s0, sl.zs0 - sl.s0, sl.n2, sl.nslot, B, st, false, nullptr, nullptr, pb.WS);
        launch_share_eval(c, pb.Y, pb.SH, sl.seta0, sl.s0 - sl.seta0, sl.n2, sl.nslot, B, ln.vside.st, false, nullptr, nullptr, pb.WS);
        launch_share_eval(c, pb.Y, pb.SH, sl.zs0, sl.n1 - sl.zs0, sl.n2, sl.nslot, B, ln.vside.st, false, nullptr, nullptr, pb.WS);
        CU(cudaEventRecord(ln.vside.join, ln.vside.st));
    } else launch_share_eval_prove(c, pb, lo, hi, B, st);
    if (!on) { prof_mark(c, ln, -1); CU(cudaEventRecord(ln.computed, st)); c->last_computed = c->last_gate = ln.computed; CU(cudaGetLastError()); return KOSK_OK; }
    prof_mark(c, ln, KOSK_PH_COMMIT);
    HashSrc hc{pb.SH, (long long)sl.nslot * SLD, 1, SLD, SOFF, c->d_tab_commit, nullptr, 0};
    // Fused form (opt-in, KOSK_B200_FUSE_FS=1, sub-batches <= fuse_max): commit hashes and the FS-1 sponge of a proof in one CTA, so
    // the sponge starts while the parties are still being hashed.  Off by default: its sponge warp shares an SM with the hasher warps
    // and runs 2.9 us per permutation against 2.6 us for the standalone sponge kernel, which costs more than the 25 us of hashing
    // it hides (single proof: 2.53 ms fused, 2.37 ms unfused); for large sub-batches it is hash-throughput bound and gains nothing.
    const bool fuse = c->fuse_fs && B <= c->fuse_max;
    if (fuse) {
        k_hash_fs<K, NCOMMIT, 1, 2><<<B, 96, 0, st>>>(hc, pb.TCR, pb.SH, sl.nslot, sl.TC0, pb.PW, nullptr, nullptr, c->d_status);
        prof_mark(c, ln, KOSK_PH_FS1);
    } else {
        k_hash_records<NCOMMIT><<<dim3(ptiles, B), 128, 0, st>>>(hc, pb.TCR, pb.SH, sl.nslot, sl.TC0);
        prof_mark(c, ln, KOSK_PH_FS1);
        { int fc, ft; fs_launch_dims(B, fc, ft); k_fs1<K><<<fc, ft, 0, st>>>(pb.TCR, pb.PW, B); }
    }
    prof_mark(c, ln, KOSK_PH_EVAL);
    k_eval<K><<<dim3(B < 16 ? (NP + 127) / 128 : 3, B), 256, 0, st>>>(pb);     // few proofs: one party tile per CTA (latency); many: 4 tiles per CTA amortise the table load
    prof_mark(c, ln, KOSK_PH_OPEN);
    k_open<K><<<B, 128, 0, st>>>(pb);
    prof_mark(c, ln, KOSK_PH_SHARE2);
    launch_share_eval(c, pb.Y, pb.SH, sl.n1, 4 * K, sl.n2, sl.nslot, B, st, false, pb.YL0, pb.YL1, pb.WS);
    prof_mark(c, ln, KOSK_PH_VIEW);
    if (side) CU(cudaStreamWaitEvent(st, ln.vside.join, 0));
    k_derive<K><<<dim3(ptiles, B), 128, 0, st>>>(pb);
    HashSrc hv{pb.SH, (long long)sl.nslot * SLD, 1, SLD, SOFF, c->d_tab_view, nullptr, 0};
    if (fuse) {
        k_hash_fs<K, NVIEW, 2, 2><<<B, 96, 0, st>>>(hv, pb.VWR, nullptr, 0, 0, nullptr, pb.I, pb.REST, c->d_status);
        prof_mark(c, ln, KOSK_PH_FS2);
    } else {
        k_hash_records<NVIEW><<<dim3(ptiles, B), 128, 0, st>>>(hv, pb.VWR, nullptr, 0, 0);
        if (c->overlap_tail && c->lanes.size() > 1) { CU(cudaEventRecord(ln.pre_tail, st)); tail_gate = ln.pre_tail; }
        prof_mark(c, ln, KOSK_PH_FS2);
        { int fc, ft; fs_launch_dims(B, fc, ft); k_fs2<<<fc, ft, 0, st>>>(pb.VWR, pb.I, pb.REST, B); }
    }
    prof_mark(c, ln, KOSK_PH_ASSEMBLE);
    k_assemble<K><<<dim3(ASM_OPENED_CTAS + (NR + ASM_REST_ROWS - 1) / ASM_REST_ROWS, B), 256, 0, st>>>(pb);
    prof_mark(c, ln, -1);
    CU(cudaEventRecord(ln.computed, st)); c->last_computed = ln.computed;
    c->last_gate = tail_gate ? tail_gate : ln.computed;
    c->launches += fuse ? 6 : 8;
    CU(cudaGetLastError());
    return KOSK_OK;
}

static int prove_chunk_k(kosk_b200_ctx *c, Lane &ln, const ProveBufs &bufs, int B, const u8 *s, u8 *pk, u8 *sk, u8 *pi, int phases)
{
    switch (c->k) {
    case 2: return prove_chunk<2>(c, ln, bufs, B, s, pk, sk, pi, phases);
    case 3: return prove_chunk<3>(c, ln, bufs, B, s, pk, sk, pi, phases);
    default: return prove_chunk<4>(c, ln, bufs, B, s, pk, sk, pi, phases);
    }
}

static int verify_chunk_lane(kosk_b200_ctx *c, Lane &ln, int B, const u8 *d_pi, const u8 *d_pk, u8 *d_ok)
{
    VerifyTables vt{c->d_St, c->d_R1, c->d_R2, c->d_inv, c->d_SU, c->d_fact, c->d_U1, c->d_U2, c->use_ntt ? &c->sn : nullptr, c->use_ntt, sn_ticket(c, ln.st), ln.vside.st ? sn_ticket(c, ln.vside.st) : nullptr};
    if (c->last_computed && c->last_computed != ln.computed) CU(cudaStreamWaitEvent(ln.st, c->last_computed, 0));
    prof_mark(c, ln, KOSK_PH_VERIFY);
    int nl = verify_chunk(c->k, ln.vb, vt, B, d_pi, d_pk, d_ok, ln.st, ln.vside);
    prof_mark(c, ln, -1);
    CU(cudaEventRecord(ln.computed, ln.st)); c->last_computed = c->last_gate = ln.computed;
    if (nl < 0) return fail(KOSK_E_CUDA, "verify launch failed");
    c->launches += nl;
    CU(cudaGetLastError());
    return KOSK_OK;
}

// sub-batch size: at most `chunk` proofs per wave; consecutive sub-batches alternate over the lanes
static size_t sub_batch(const kosk_b200_ctx *c, size_t n)
{
    return std::max<size_t>(1, std::min<size_t>((size_t)c->chunk, n));
}
// lanes start after everything already enqueued on the caller's stream ...
static int lanes_fork(kosk_b200_ctx *c, cudaStream_t caller)
{
    CU(cudaEventRecord(c->ev_start, caller));
    for (Lane &ln : c->lanes) CU(cudaStreamWaitEvent(ln.st, c->ev_start, 0));
    return KOSK_OK;
}
// ... and the caller's stream continues after all lanes are done
static int lanes_join(kosk_b200_ctx *c, cudaStream_t caller)
{
    for (Lane &ln : c->lanes) { CU(cudaEventRecord(ln.done, ln.st)); CU(cudaStreamWaitEvent(caller, ln.done, 0)); }
    return KOSK_OK;
}

// device allocation released on every return path of the component entry points
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// ---- generic component kernels ----
__global__ void __launch_bounds__(128) k_ntt_rows(u16 *a)
{
    __shared__ u16 p[256];
    u16 *row = a + (size_t)blockIdx.x * 256;
    const int tid = threadIdx.x;
    p[tid] = row[tid]; p[tid + 128] = row[tid + 128];
    __syncthreads();
    ntt256_block(p, tid);
    row[tid] = p[tid]; row[tid + 128] = p[tid + 128];
}

__global__ void __launch_bounds__(64) k_sha3_rows(const u8 *in, u8 *out, size_t n, size_t len)
{
    const size_t r = (size_t)blockIdx.x * 64 + threadIdx.x;
    if (r >= n) return;
    const u8 *src = in + r * len;
    uint64_t a[25]; keccak_zero(a);
    size_t off = 0;
    while (len - off >= 136) {
#pragma unroll
        for (int l = 0; l < 17; l++) { uint64_t w = 0; for (int i = 0; i < 8; i++) w |= (uint64_t)src[off + 8 * l + i] << (8 * i); a[l] ^= w; }
        keccak_f1600(a);
        off += 136;
    }
    uint64_t last[17];
    for (int l = 0; l < 17; l++) last[l] = 0;
    const int rem = (int)(len - off);
    for (int i = 0; i < rem; i++) last[i >> 3] |= (uint64_t)src[off + i] << (8 * (i & 7));
    last[rem >> 3] |= 0x06ULL << (8 * (rem & 7));
    last[16] |= 0x8000000000000000ULL;
#pragma unroll
    for (int l = 0; l < 17; l++) a[l] ^= last[l];
    keccak_f1600(a);
    uint64_t *o = reinterpret_cast<uint64_t *>(out + r * 32);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
}

extern "C" {

int kosk_b200_prove_batch_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_seeds, uint8_t *d_pk, uint8_t *d_sk, uint8_t *d_pi, void *stream)
{
    if (!c || !d_seeds || !d_pk || !d_sk || !d_pi) return fail(KOSK_E_ARG, "null argument");
    if (((uintptr_t)d_seeds & 7) || ((uintptr_t)d_pi & 3)) return fail(KOSK_E_ARG, "misaligned device buffer (d_seeds: 8 bytes, d_pi: 4 bytes)");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    if (n == 0) return KOSK_OK;
    const Layout &L = c->L;
    int rc = lanes_fork(c, (cudaStream_t)stream); if (rc) return rc;
    const size_t sub = sub_batch(c, n);
    for (size_t o = 0; o < n && !rc; o += sub) {
        const int B = (int)std::min<size_t>(sub, n - o);
        Lane &ln = c->lanes[c->next_lane++ % c->lanes.size()];
        rc = prove_chunk_k(c, ln, ln.pb, B, d_seeds + 32 * o, d_pk + L.pk_bytes * o, d_sk + L.sk_bytes * o, d_pi + L.proof_bytes * o, PH_OFFLINE | PH_ONLINE);
    }
    // the caller's stream is ordered after whatever was enqueued on the lanes, also when a later sub-batch failed to launch
    const int rj = lanes_join(c, (cudaStream_t)stream);
    return rc ? rc : rj;
}

int kosk_b200_prove_batch(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    int rc = kosk_b200_prove_batch_async(c, n, seeds, pk, sk, pi);
    const int rs = c ? kosk_b200_sync(c) : KOSK_OK;      // also after a mid-batch error: nothing may still be writing the caller's buffers
    return rc ? rc : rs;
}

// ---- compact wire format on the host-buffer paths (wire_kernels.cuh, wire_host.h) ----
// the gate thread polls: a blocking-sync event makes the copy engine raise an interrupt per slice (measured: +1 ms per 64 slices of link
// idle time), a spinning cudaEventSynchronize burns a core that the unpack workers need
static int g_wire_poll_us = 20;      // KOSK_B200_WIRE_POLL_US: 0 = block in cudaEventSynchronize (events are then created with cudaEventBlockingSync)
static int wire_wait_event(void *gate)
{
    if (g_wire_poll_us <= 0) return cudaEventSynchronize((cudaEvent_t)gate) == cudaSuccess ? 0 : 1;
    for (;;) {
        const cudaError_t e = cudaEventQuery((cudaEvent_t)gate);
        if (e == cudaSuccess) return 0;
        if (e != cudaErrorNotReady) return 1;
        std::this_thread::sleep_for(std::chrono::microseconds(g_wire_poll_us));
    }
}
static int wire_default_threads()
{
    if (const char *e = getenv("KOSK_B200_WIRE_THREADS")) { const int t = atoi(e); if (t > 0) return t; }
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::min(16u, std::max(2u, hc / 2));
}
// worker pool on first use; for every lane the packed device buffer and (host_too) the pinned staging buffer with its slice events.
// All lanes at once: pinning half a gigabyte takes a quarter of a second, which must not land in the middle of a pipelined run.
static int wire_ensure(kosk_b200_ctx *c, bool host_too)
{
    const WireLayout W = make_wire_layout(c->k);
    if (host_too && !c->wpool) c->wpool = wire_pool_create(c->wire_threads > 0 ? c->wire_threads : wire_default_threads(), wire_wait_event);
    const size_t bytes = (size_t)c->chunk * W.wire_bytes;
    for (Lane &ln : c->lanes) {
        if (!ln.d_wire) {
            if (cudaMalloc((void **)&ln.d_wire, bytes) != cudaSuccess) { ln.d_wire = nullptr; return fail(KOSK_E_NOMEM, "cudaMalloc failed for the packed proofs"); }
        }
        if (host_too && !ln.h_wire) {
            if (cudaHostAlloc((void **)&ln.h_wire, bytes, cudaHostAllocDefault) != cudaSuccess) { ln.h_wire = nullptr; return fail(KOSK_E_NOMEM, "cudaHostAlloc failed for the wire staging buffer"); }
            const int nsl = (c->chunk + c->wire_slice - 1) / c->wire_slice;
            ln.wev.resize(nsl);
            if (const char *e = getenv("KOSK_B200_WIRE_POLL_US")) g_wire_poll_us = atoi(e);
            for (cudaEvent_t &e : ln.wev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (g_wire_poll_us <= 0 ? cudaEventBlockingSync : 0)));
            ln.wctr.reset(new std::atomic<int>[nsl]);
            for (int i = 0; i < nsl; i++) ln.wctr[i].store(0);
            CU(cudaEventCreateWithFlags(&ln.h2d_done, cudaEventDisableTiming));
        }
    }
    return KOSK_OK;
}

// packed = the caller keeps the compact bytes (out = wire[n][wire_bytes]); else out = pi[n][proof_bytes] in the reference layout
static int prove_async_impl(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *out, bool packed)
{
    if (!c || !seeds || !pk || !sk || !out) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    const Layout &L = c->L;
    const WireLayout W = make_wire_layout(c->k);
    // a handful of proofs is latency-bound: the pack kernel and the hand-over to a worker cost more than the bytes saved on the link
    const bool wire = packed || (c->wire_mode != 0 && n >= 8);
    const size_t sub = sub_batch(c, n);
    for (size_t o = 0; o < n; o += sub) {
        const int B = (int)std::min<size_t>(sub, n - o);
        const size_t li = c->next_lane++ % c->lanes.size();
        Lane &ln = c->lanes[li];
        if (wire) {
            int rc = wire_ensure(c, !packed); if (rc) return rc;
            static const int lanewait = getenv("KOSK_B200_WIRE_LANEWAIT") ? atoi(getenv("KOSK_B200_WIRE_LANEWAIT")) : 0;      // measurement switch
            if (lanewait && !packed) for (size_t si = 0; si < ln.wev.size(); si++) wire_pool_wait_counter(c->wpool, &ln.wctr[si]);
        }
        wire_pool_trace(c->wpool, 10, li);
        CU(cudaMemcpyAsync(ln.d_seeds, seeds + 32 * o, 32 * (size_t)B, cudaMemcpyHostToDevice, ln.st));
        int rc = prove_chunk_k(c, ln, ln.pb, B, ln.d_seeds, ln.d_pk, ln.d_sk, ln.d_pi, PH_OFFLINE | PH_ONLINE);
        if (rc) return rc;
        wire_pool_trace(c->wpool, 11, li);
        if (wire) { c->launches += wire_launch(true, ln.d_pi, ln.d_wire, c->k, B, ln.st); CU(cudaGetLastError()); }
        CU(cudaMemcpyAsync(pk + L.pk_bytes * o, ln.d_pk, L.pk_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        CU(cudaMemcpyAsync(sk + L.sk_bytes * o, ln.d_sk, L.sk_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        if (!wire) {
            CU(cudaMemcpyAsync(out + L.proof_bytes * o, ln.d_pi, L.proof_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        } else if (packed) {
            CU(cudaMemcpyAsync(out + W.wire_bytes * o, ln.d_wire, W.wire_bytes * (size_t)B, cudaMemcpyDeviceToHost, ln.st));
        } else {
            // the proofs cross the link in slices.  A packed slice lands in the staging buffer and the workers expand it into the caller's
            // buffer as soon as its event fires, while the following slices are still on the link; a raw slice (error-diffused share of
            // 100 - wire_mode percent) goes straight from d_pi to the caller's buffer and costs the host nothing.
            for (int s0 = 0; s0 < B; s0 += c->wire_slice) {
                const int ns = std::min(c->wire_slice, B - s0);
                c->wire_acc += c->wire_mode;
                if (c->wire_acc < 100) {
                    CU(cudaMemcpyAsync(out + L.proof_bytes * (o + s0), ln.d_pi + L.proof_bytes * (size_t)s0, L.proof_bytes * (size_t)ns, cudaMemcpyDeviceToHost, ln.st));
                    continue;
                }
                c->wire_acc -= 100;
                const int si = s0 / c->wire_slice;
                cudaEvent_t ev = ln.wev[si];
                // the kernels of this sub-batch are already enqueued; only the copy into a slice of the staging buffer waits until the
                // workers have expanded what the lane's previous sub-batch left there
                wire_pool_wait_counter(c->wpool, &ln.wctr[si]);
                CU(cudaMemcpyAsync(ln.h_wire + W.wire_bytes * (size_t)s0, ln.d_wire + W.wire_bytes * (size_t)s0, W.wire_bytes * (size_t)ns, cudaMemcpyDeviceToHost, ln.st));
                CU(cudaEventRecord(ev, ln.st));
                wire_pool_submit(c->wpool, ev, 0, c->k, (size_t)ns, 1, ln.h_wire + W.wire_bytes * (size_t)s0, out + L.proof_bytes * (o + s0), &ln.wctr[si], &ln.wire_flag);
            }
            wire_pool_trace(c->wpool, 12, li);
        }
    }
    return KOSK_OK;
}

int kosk_b200_prove_batch_async(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    return prove_async_impl(c, n, seeds, pk, sk, pi, false);
}
int kosk_b200_prove_batch_packed_async(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *wire)
{
    return prove_async_impl(c, n, seeds, pk, sk, wire, true);
}
int kosk_b200_prove_batch_packed(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *wire)
{
    int rc = prove_async_impl(c, n, seeds, pk, sk, wire, true);
    const int rs = c ? kosk_b200_sync(c) : KOSK_OK;
    return rc ? rc : rs;
}

int kosk_b200_verifiable_keygen(kosk_b200_ctx *c, const uint8_t seed[32], uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    return kosk_b200_prove_batch(c, 1, seed, pk, sk, pi);
}

int kosk_b200_verify_batch_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_pi, const uint8_t *d_pk, uint8_t *d_ok, void *stream)
{
    if (!c || !d_pi || !d_pk || !d_ok) return fail(KOSK_E_ARG, "null argument");
    if ((uintptr_t)d_pi & 3) return fail(KOSK_E_ARG, "misaligned device buffer (d_pi: 4 bytes)");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    if (n == 0) return KOSK_OK;
    const Layout &L = c->L;
    int rc = lanes_fork(c, (cudaStream_t)stream); if (rc) return rc;
    const size_t sub = sub_batch(c, n);
    size_t i = 0;
    for (size_t o = 0; o < n && !rc; o += sub, i++) {
        const int B = (int)std::min<size_t>(sub, n - o);
        rc = verify_chunk_lane(c, c->lanes[i % c->lanes.size()], B, d_pi + L.proof_bytes * o, d_pk + L.pk_bytes * o, d_ok + o);
    }
    const int rj = lanes_join(c, (cudaStream_t)stream);
    return rc ? rc : rj;
}

// packed: in = wire[n][wire_bytes]; else in = pi[n][proof_bytes] (reference layout), packed on the host when the wire mode is on
static int verify_batch_impl(kosk_b200_ctx *c, size_t n, const uint8_t *in, const uint8_t *pk, uint8_t *ok, bool packed, bool async)
{
    if (!c || !in || !pk || !ok) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    const Layout &L = c->L;
    const WireLayout W = make_wire_layout(c->k);
    // a single proof is latency-bound: packing it on the host costs more than the 145 KB it saves on the link
    const bool wire = packed || (c->wire_mode != 0 && n >= 8);
    const size_t sub = sub_batch(c, n);
    int rc = KOSK_OK;
    for (size_t o = 0; o < n && !rc; o += sub) {
        const int B = (int)std::min<size_t>(sub, n - o);
        Lane &ln = c->lanes[c->next_lane++ % c->lanes.size()];
        // proofs [0, np) of the sub-batch travel packed, [np, B) as struct bytes straight from the caller's buffer (that copy is enqueued
        // first, so it is on the link while the workers pack)
        int np = !wire ? 0 : packed ? B : (int)((size_t)B * c->wire_mode / 100);
        if (wire) { rc = wire_ensure(c, !packed); if (rc) break; }
        // the public keys first: a small H2D queued behind the device unpack kernel would hold the copy engine's queue (and the next call's big copy
        // behind it) until that kernel gets SMs next to the other lane's verify kernels (packed end to end: 12.1 ms per 1024 proofs for a 9.6 ms copy)
        CU(cudaMemcpyAsync(ln.d_pk, pk + L.pk_bytes * o, L.pk_bytes * (size_t)B, cudaMemcpyHostToDevice, ln.st));
        if (!packed && np < B)
            CU(cudaMemcpyAsync(ln.d_pi + L.proof_bytes * (size_t)np, in + L.proof_bytes * (o + np), L.proof_bytes * (size_t)(B - np), cudaMemcpyHostToDevice, ln.st));
        if (np > 0 && !packed) {
            if (ln.h2d_pending) { CU(cudaEventSynchronize(ln.h2d_done)); ln.h2d_pending = false; }      // the previous copy out of the staging buffer
            // a proof with a u16 >= 4096 has no wire image: such a sub-batch crosses the link in the reference layout (same verdicts)
            if (wire_pool_run(c->wpool, 1, c->k, (size_t)np, in + L.proof_bytes * o, ln.h_wire) != 0) {
                CU(cudaMemcpyAsync(ln.d_pi, in + L.proof_bytes * o, L.proof_bytes * (size_t)np, cudaMemcpyHostToDevice, ln.st));
                np = 0;
            }
        }
        if (np > 0) {
            CU(cudaMemcpyAsync(ln.d_wire, packed ? in + W.wire_bytes * o : ln.h_wire, W.wire_bytes * (size_t)np, cudaMemcpyHostToDevice, ln.st));
            if (!packed) { CU(cudaEventRecord(ln.h2d_done, ln.st)); ln.h2d_pending = true; }
            c->launches += wire_launch(false, ln.d_wire, ln.d_pi, c->k, np, ln.st);
        }
        rc = verify_chunk_lane(c, ln, B, ln.d_pi, ln.d_pk, ln.d_ok);
        if (rc) break;
        CU(cudaMemcpyAsync(ok + o, ln.d_ok, (size_t)B, cudaMemcpyDeviceToHost, ln.st));
    }
    if (async && !rc) return KOSK_OK;                      // the caller waits with kosk_b200_sync
    for (Lane &ln : c->lanes) { cudaError_t e = cudaStreamSynchronize(ln.st); if (e != cudaSuccess && !rc) rc = fail(KOSK_E_CUDA, cudaGetErrorString(e)); ln.h2d_pending = false; }
    if (rc) return rc;
    return check_status(c);
}

int kosk_b200_verify_batch(kosk_b200_ctx *c, size_t n, const uint8_t *pi, const uint8_t *pk, uint8_t *ok)
{
    return verify_batch_impl(c, n, pi, pk, ok, false, false);
}
int kosk_b200_verify_batch_async(kosk_b200_ctx *c, size_t n, const uint8_t *pi, const uint8_t *pk, uint8_t *ok)
{
    return verify_batch_impl(c, n, pi, pk, ok, false, true);
}
int kosk_b200_verify_batch_packed_async(kosk_b200_ctx *c, size_t n, const uint8_t *wire, const uint8_t *pk, uint8_t *ok)
{
    return verify_batch_impl(c, n, wire, pk, ok, true, true);
}
int kosk_b200_verify_batch_packed(kosk_b200_ctx *c, size_t n, const uint8_t *wire, const uint8_t *pk, uint8_t *ok)
{
    return verify_batch_impl(c, n, wire, pk, ok, true, false);
}

int kosk_b200_kosk_verify(kosk_b200_ctx *c, const uint8_t *pi, const uint8_t *pk)
{
    uint8_t ok = 0;
    int rc = kosk_b200_verify_batch(c, 1, pi, pk, &ok);
    return rc ? rc : (int)ok;
}

size_t kosk_b200_wire_bytes(int k) { return (k >= 2 && k <= 4) ? make_wire_layout(k).wire_bytes : 0; }

int kosk_b200_set_wire(kosk_b200_ctx *c, int mode, int threads)
{
    if (!c || mode < 0 || mode > 100) return fail(KOSK_E_ARG, "bad argument");
    LOCK(c);
    int rc = kosk_b200_sync(c); if (rc) return rc;
    c->wire_mode = mode;
    if (threads > 0 && threads != c->wire_threads) { c->wire_threads = threads; wire_pool_destroy(c->wpool); c->wpool = nullptr; }
    return KOSK_OK;
}
int kosk_b200_wire_info(const kosk_b200_ctx *c, int *mode, int *threads, const char **simd)
{
    if (!c) return fail(KOSK_E_ARG, "null argument");
    if (mode) *mode = c->wire_mode;
    if (threads) *threads = c->wpool ? wire_pool_threads(c->wpool) : (c->wire_threads > 0 ? c->wire_threads : wire_default_threads());
    if (simd) *simd = wire_simd_name();
    return KOSK_OK;
}

// instrumentation of the wire pipeline: out[0] = ns the gate thread spent waiting for D2H slices, [1] = slices, [2] = ns summed over
// the workers spent converting, [3] = jobs (one proof each); reset != 0 clears the counters
int kosk_b200_wire_stats(kosk_b200_ctx *c, uint64_t *out, int reset)
{
    if (!c || !out) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    out[0] = out[1] = out[2] = out[3] = 0;
    if (c->wpool) wire_pool_stats(c->wpool, out, reset);
    return KOSK_OK;
}

int kosk_b200_wire_pack_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_pi, uint8_t *d_wire, void *stream)
{
    if (!c || !d_pi || !d_wire || ((uintptr_t)d_pi & 3) || ((uintptr_t)d_wire & 15)) return fail(KOSK_E_ARG, "null or misaligned argument (d_pi: 4 bytes, d_wire: 16 bytes)");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    for (size_t o = 0; o < n; o += 32768) c->launches += wire_launch(true, d_pi + c->L.proof_bytes * o, d_wire + make_wire_layout(c->k).wire_bytes * o, c->k, (int)std::min<size_t>(32768, n - o), (cudaStream_t)stream);
    CU(cudaGetLastError());
    return KOSK_OK;
}
int kosk_b200_wire_unpack_device(kosk_b200_ctx *c, size_t n, const uint8_t *d_wire, uint8_t *d_pi, void *stream)
{
    if (!c || !d_pi || !d_wire || ((uintptr_t)d_pi & 3) || ((uintptr_t)d_wire & 15)) return fail(KOSK_E_ARG, "null or misaligned argument (d_pi: 4 bytes, d_wire: 16 bytes)");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    for (size_t o = 0; o < n; o += 32768) c->launches += wire_launch(false, d_wire + make_wire_layout(c->k).wire_bytes * o, d_pi + c->L.proof_bytes * o, c->k, (int)std::min<size_t>(32768, n - o), (cudaStream_t)stream);
    CU(cudaGetLastError());
    return KOSK_OK;
}

// host codec, no context and no device: wire[n][wire_bytes] <-> pi[n][proof_bytes]; threads <= 1 runs on the calling thread
static int wire_host_run(int kind, int k, size_t n, const uint8_t *src, uint8_t *dst, int threads)
{
    if (k < 2 || k > 4 || !src || !dst) return fail(KOSK_E_ARG, "bad argument");
    const WireLayout W = make_wire_layout(k);
    if (threads <= 1 || n < 2) {
        int bad = 0;
        for (size_t i = 0; i < n; i++) {
            if (kind == 0) wire_unpack_proof(k, src + i * W.wire_bytes, dst + i * W.proof_bytes);
            else bad |= wire_pack_proof(k, src + i * W.proof_bytes, dst + i * W.wire_bytes);
        }
        return bad ? fail(KOSK_E_ARG, "a proof holds a u16 >= 4096 and has no wire image") : KOSK_OK;
    }
    WirePool *p = wire_pool_create(threads, nullptr);
    const int bad = wire_pool_run(p, kind, k, n, src, dst);
    wire_pool_destroy(p);
    return bad ? fail(KOSK_E_ARG, "a proof holds a u16 >= 4096 and has no wire image") : KOSK_OK;
}
int kosk_b200_wire_pack(int k, size_t n, const uint8_t *pi, uint8_t *wire, int threads) { return wire_host_run(1, k, n, pi, wire, threads); }
int kosk_b200_wire_unpack(int k, size_t n, const uint8_t *wire, uint8_t *pi, int threads) { return wire_host_run(0, k, n, wire, pi, threads); }
const char *kosk_b200_wire_simd(void) { return wire_simd_name(); }

// ---- offline / online split ----
struct kosk_b200_pool {
    kosk_b200_ctx *ctx; size_t n;
    ProveBufs pb{};
    u8 *d_seeds = nullptr, *d_pk = nullptr, *d_sk = nullptr, *d_pi = nullptr;
};

void kosk_b200_pool_destroy(kosk_b200_pool *p)
{
    if (!p) return;
    LOCK(p->ctx);
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->lanes[0].st);
    if (p->d_seeds) cudaMemset(p->d_seeds, 0, 32 * p->n);
    if (p->d_sk) cudaMemset(p->d_sk, 0, p->ctx->L.sk_bytes * p->n);
    scrub_prove_bufs(p->pb, p->ctx->sl, p->ctx->k, p->n);
    free_prove_bufs(p->pb);
    p->ctx->live_pools--;
    void *lp[] = {p->d_seeds, p->d_pk, p->d_sk, p->d_pi};
    for (void *q : lp) if (q) cudaFree(q);
    delete p;
}

int kosk_b200_pool_create(kosk_b200_ctx *c, size_t n, const uint8_t *seeds, kosk_b200_pool **out)
{
    if (!c || !seeds || !out || n == 0 || n > 16384) return fail(KOSK_E_ARG, "bad argument (1 <= n <= 16384)");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    kosk_b200_pool *p = new kosk_b200_pool; p->ctx = c; p->n = n; c->live_pools++;
    if (alloc_prove_bufs(p->pb, c->sl, c->k, n, c->use_tensor != 0) != 0 || cudaMalloc((void **)&p->d_seeds, 32 * n) != cudaSuccess) {
        kosk_b200_pool_destroy(p); return fail(KOSK_E_NOMEM, "cudaMalloc failed for the preprocessing pool");
    }
    Lane &ln = c->lanes[0];
    if (cudaMemcpyAsync(p->d_seeds, seeds, 32 * n, cudaMemcpyHostToDevice, ln.st) != cudaSuccess) { kosk_b200_pool_destroy(p); return fail(KOSK_E_CUDA, "seed copy failed"); }
    int rc = prove_chunk_k(c, ln, p->pb, (int)n, p->d_seeds, nullptr, nullptr, nullptr, PH_OFFLINE);
    if (rc) { kosk_b200_pool_destroy(p); return rc; }
    if (cudaStreamSynchronize(ln.st) != cudaSuccess) { kosk_b200_pool_destroy(p); return fail(KOSK_E_CUDA, "preprocessing failed"); }
    *out = p;
    return KOSK_OK;
}

int kosk_b200_pool_prove(kosk_b200_pool *p, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    if (!p || !pk || !sk || !pi) return fail(KOSK_E_ARG, "null argument");
    kosk_b200_ctx *c = p->ctx;
    LOCK(c);
    CU(cudaSetDevice(c->device));
    const Layout &L = c->L; const size_t n = p->n;
    if (!p->d_pi) {
        if (cudaMalloc((void **)&p->d_pk, L.pk_bytes * n) != cudaSuccess || cudaMalloc((void **)&p->d_sk, L.sk_bytes * n) != cudaSuccess ||
            cudaMalloc((void **)&p->d_pi, L.proof_bytes * n) != cudaSuccess) return fail(KOSK_E_NOMEM, "cudaMalloc failed for pool outputs");
    }
    Lane &ln = c->lanes[0];
    int rc = prove_chunk_k(c, ln, p->pb, (int)n, p->d_seeds, p->d_pk, p->d_sk, p->d_pi, PH_ONLINE);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pk, p->d_pk, L.pk_bytes * n, cudaMemcpyDeviceToHost, ln.st));
    CU(cudaMemcpyAsync(sk, p->d_sk, L.sk_bytes * n, cudaMemcpyDeviceToHost, ln.st));
    CU(cudaMemcpyAsync(pi, p->d_pi, L.proof_bytes * n, cudaMemcpyDeviceToHost, ln.st));
    CU(cudaStreamSynchronize(ln.st));
    return KOSK_OK;
}

// ---- pool serialiser: the working form of the reference's unused prepare_randomness (de)serialiser (mlwe_prover.cpp:61-79, which drops the
// range-proof half).  Image = header | seeds[n][32] | Y[n][n2][YLD] u16 (every sharing's 256 secrets and 151 tail randoms, the prove-time
// ones included: they are functions of the seed) | planes[n][s0][SLD] u16 (the evaluated f / NTT_f / eta sharings).  It holds the seeds:
// treat it like a secret key.
struct PoolHeader { char magic[8]; uint32_t version, k, n, n2, s0, yld, sld, reserved; };
static const char POOL_MAGIC[8] = {'K', 'O', 'S', 'K', 'P', 'O', 'O', 'L'};
static size_t pool_image_bytes(const kosk_b200_ctx *c, size_t n)
{
    return sizeof(PoolHeader) + n * (32 + (size_t)c->sl.n2 * YLD * 2 + (size_t)c->sl.s0 * SLD * 2);
}
size_t kosk_b200_pool_bytes(const kosk_b200_pool *p) { return p ? pool_image_bytes(p->ctx, p->n) : 0; }

int kosk_b200_pool_export(kosk_b200_pool *p, void *image, size_t bytes)
{
    if (!p || !image) return fail(KOSK_E_ARG, "null argument");
    kosk_b200_ctx *c = p->ctx;
    LOCK(c);
    if (bytes < pool_image_bytes(c, p->n)) return fail(KOSK_E_ARG, "image buffer too small (kosk_b200_pool_bytes)");
    CU(cudaSetDevice(c->device));
    const Slots &sl = c->sl; const size_t n = p->n;
    PoolHeader h{}; memcpy(h.magic, POOL_MAGIC, 8); h.version = 1; h.k = (uint32_t)c->k; h.n = (uint32_t)n; h.n2 = (uint32_t)sl.n2; h.s0 = (uint32_t)sl.s0; h.yld = YLD; h.sld = SLD;
    u8 *o = static_cast<u8 *>(image);
    memcpy(o, &h, sizeof h); o += sizeof h;
    CU(cudaStreamSynchronize(c->lanes[0].st));
    CU(cudaMemcpy(o, p->d_seeds, 32 * n, cudaMemcpyDeviceToHost)); o += 32 * n;
    CU(cudaMemcpy(o, p->pb.Y, n * sl.n2 * YLD * 2, cudaMemcpyDeviceToHost)); o += n * sl.n2 * YLD * 2;
    CU(cudaMemcpy2D(o, (size_t)sl.s0 * SLD * 2, p->pb.SH, (size_t)sl.nslot * SLD * 2, (size_t)sl.s0 * SLD * 2, n, cudaMemcpyDeviceToHost));
    return KOSK_OK;
}

int kosk_b200_pool_import(kosk_b200_ctx *c, const void *image, size_t bytes, kosk_b200_pool **out)
{
    if (!c || !image || !out || bytes < sizeof(PoolHeader)) return fail(KOSK_E_ARG, "bad argument");
    PoolHeader h; memcpy(&h, image, sizeof h);
    const Slots &sl = c->sl;
    if (memcmp(h.magic, POOL_MAGIC, 8) || h.version != 1) return fail(KOSK_E_ARG, "not a pool image");
    if (h.k != (uint32_t)c->k || h.n2 != (uint32_t)sl.n2 || h.s0 != (uint32_t)sl.s0 || h.yld != YLD || h.sld != SLD) return fail(KOSK_E_ARG, "pool image of another KYBER_K / layout");
    const size_t n = h.n;
    if (n == 0 || n > 16384 || bytes < pool_image_bytes(c, n)) return fail(KOSK_E_ARG, "truncated pool image");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    kosk_b200_pool *p = new kosk_b200_pool; p->ctx = c; p->n = n; c->live_pools++;
    if (alloc_prove_bufs(p->pb, sl, c->k, n, c->use_tensor != 0) != 0 || cudaMalloc((void **)&p->d_seeds, 32 * n) != cudaSuccess) {
        kosk_b200_pool_destroy(p); return fail(KOSK_E_NOMEM, "cudaMalloc failed for the preprocessing pool");
    }
    const u8 *o = static_cast<const u8 *>(image) + sizeof h;
    cudaError_t e = cudaMemcpy(p->d_seeds, o, 32 * n, cudaMemcpyHostToDevice); o += 32 * n;
    if (e == cudaSuccess) e = cudaMemcpy(p->pb.Y, o, n * sl.n2 * YLD * 2, cudaMemcpyHostToDevice);
    o += n * sl.n2 * YLD * 2;
    if (e == cudaSuccess) e = cudaMemcpy2D(p->pb.SH, (size_t)sl.nslot * SLD * 2, o, (size_t)sl.s0 * SLD * 2, (size_t)sl.s0 * SLD * 2, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { kosk_b200_pool_destroy(p); return fail(KOSK_E_CUDA, cudaGetErrorString(e)); }
    *out = p;
    return KOSK_OK;
}

int kosk_b200_share_eval_device(kosk_b200_ctx *c, size_t n, const uint16_t *d_y, uint16_t *d_planes, void *stream)
{
    if (!c || !d_y || !d_planes) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    if (c->use_tensor && c->tmp_rows < n) {       // grow-only scratch for the limb planes of the caller's rows
        if (c->tmpL0) { cudaFree(c->tmpL0); cudaFree(c->tmpL1); c->tmpL0 = c->tmpL1 = nullptr; }
        if (cudaMalloc((void **)&c->tmpL0, n * YLD) != cudaSuccess || cudaMalloc((void **)&c->tmpL1, n * YLD) != cudaSuccess) return fail(KOSK_E_NOMEM, "cudaMalloc failed");
        c->tmp_rows = n;
    }
    for (size_t o = 0; o < n; o += (1u << 22)) {
        const int m = (int)std::min<size_t>(1u << 22, n - o);
        launch_share_eval(c, d_y + o * YLD, d_planes + o * SLD, 0, m, m, m, 1, (cudaStream_t)stream, false,
                          c->use_tensor ? c->tmpL0 + o * YLD : nullptr, c->use_tensor ? c->tmpL1 + o * YLD : nullptr);
    }
    CU(cudaGetLastError());
    return KOSK_OK;
}

int kosk_b200_share_eval(kosk_b200_ctx *c, size_t n, const uint16_t *y, uint16_t *shares)
{
    if (!c || !y || !shares) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    DevBuf by, bp;
    std::vector<u16> hy(n * YLD, 0), hp(n * SLD);
    for (size_t r = 0; r < n; r++) memcpy(&hy[r * YLD], y + r * D1, D1 * 2);
    CU(by.alloc(hy.size() * 2 + 16)); CU(bp.alloc(hp.size() * 2 + 16));
    u16 *dy = by.as<u16>(), *dp = bp.as<u16>();
    CU(cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice));
    int rc = kosk_b200_share_eval_device(c, n, dy, dp, c->lanes[0].st);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->lanes[0].st));
    CU(cudaMemcpy(hp.data(), dp, hp.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < n; r++) memcpy(shares + r * NP, &hp[r * SLD + SOFF], NP * 2);
    return KOSK_OK;
}

int kosk_b200_sha3_256_rows(kosk_b200_ctx *c, size_t n, size_t len, const uint8_t *in, uint8_t *out)
{
    if (!c || !in || !out || n == 0) return fail(KOSK_E_ARG, "bad argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    DevBuf bi, bo;
    CU(bi.alloc(n * len + 16)); CU(bo.alloc(n * 32));
    u8 *di = bi.as<u8>(), *dout = bo.as<u8>();
    CU(cudaMemcpy(di, in, n * len, cudaMemcpyHostToDevice));
    k_sha3_rows<<<(unsigned)((n + 63) / 64), 64, 0, c->lanes[0].st>>>(di, dout, n, len); c->launches++;
    CU(cudaStreamSynchronize(c->lanes[0].st));
    CU(cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost));
    return KOSK_OK;
}

int kosk_b200_ntt_rows(kosk_b200_ctx *c, size_t n, uint16_t *a)
{
    if (!c || !a || n == 0) return fail(KOSK_E_ARG, "bad argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    DevBuf bd;
    CU(bd.alloc(n * 512));
    u16 *d = bd.as<u16>();
    CU(cudaMemcpy(d, a, n * 512, cudaMemcpyHostToDevice));
    k_ntt_rows<<<(unsigned)n, 128, 0, c->lanes[0].st>>>(d); c->launches++;
    CU(cudaStreamSynchronize(c->lanes[0].st));
    CU(cudaMemcpy(a, d, n * 512, cudaMemcpyDeviceToHost));
    return KOSK_OK;
}

int kosk_b200_debug_fetch(kosk_b200_ctx *c, const char *what, void *out, size_t bytes)
{
    if (!c || !what || !out) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    const Slots &sl = c->sl;
    const size_t B = (size_t)c->chunk;
    const void *src = nullptr; size_t avail = 0;
    std::string w(what);
    if (w == "alpha_pow") { src = c->lanes[0].pb.PW; avail = B * (MK + 2 * c->k) * sl.F * 2; }
    else if (w == "I") { src = c->lanes[0].pb.I; avail = B * NT * 2; }
    else if (w == "rest") { src = c->lanes[0].pb.REST; avail = B * NR * 2; }
    else if (w == "tcomm") { src = c->lanes[0].pb.TCR; avail = B * TREE_BYTES; }
    else if (w == "views") { src = c->lanes[0].pb.VWR; avail = B * TREE_BYTES; }
    else if (w == "Y") { src = c->lanes[0].pb.Y; avail = B * sl.n2 * YLD * 2; }
    else if (w == "planes") { src = c->lanes[0].pb.SH; avail = B * sl.nslot * SLD * 2; }
    else if (w == "bg") { src = c->lanes[0].pb.BG; avail = B * NP * 2 * BGH * 2; }
    else if (w == "vflags") { src = c->lanes[0].vb.flags; avail = B * 4; }
    else return fail(KOSK_E_ARG, "unknown buffer name");
    CU(cudaMemcpy(out, src, std::min(bytes, avail), cudaMemcpyDeviceToHost));
    return KOSK_OK;
}


// debug: timeline of the phase marks of the last profiled batch: triples (lane, phase, ms since the batch's fork event)
int kosk_b200_debug_trace(kosk_b200_ctx *c, double *out, int max_triples)
{
    if (!c || !out) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    int n = 0;
    for (size_t li = 0; li < c->lanes.size(); li++) {
        Lane &ln = c->lanes[li];
        for (size_t i = 0; i < ln.ev_phase.size() && n < max_triples; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->ev_start, ln.ev[ln.ev_phase[i].second]) != cudaSuccess) ms = -1;
            out[3 * n] = (double)li; out[3 * n + 1] = ln.ev_phase[i].first; out[3 * n + 2] = ms; n++;
        }
    }
    return n;
}

int kosk_b200_set_strict(kosk_b200_ctx *c, int on)
{
    if (!c) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    for (Lane &ln : c->lanes) ln.vb.strict = on != 0;
    return KOSK_OK;
}

int kosk_b200_set_profiling(kosk_b200_ctx *c, int on)
{
    if (!c) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    c->prof = on != 0;
    return KOSK_OK;
}

int kosk_b200_phase_times(kosk_b200_ctx *c, double *ms, uint64_t *calls, int n, int reset)
{
    if (!c || !ms || !calls) return fail(KOSK_E_ARG, "null argument");
    LOCK(c);
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    prof_collect(c);
    for (int i = 0; i < n && i < KOSK_NPHASE; i++) { ms[i] = c->phase_ms[i]; calls[i] = c->phase_calls[i]; }
    if (reset) for (int i = 0; i < KOSK_NPHASE; i++) { c->phase_ms[i] = 0; c->phase_calls[i] = 0; }
    return KOSK_OK;
}

}  // extern "C"

#include "component_api.cuh"
#include "raw_api.cuh"
static RawState *raw_new() { return new RawState; }
static void raw_delete(RawState *r) { if (r) { raw_free(*r); delete r; } }
#include "kem_kernels.cuh"
static KemState *kem_new() { return new KemState; }
static void kem_delete(KemState *k) { if (k) { kem_free(*k); delete k; } }

// ---- integer-pipe issue-rate microbenchmarks (roofline denominators; MEASURED_PEAKS.json has no integer entry) ----
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    const uint32_t m = seed | 1u, x = seed * 0x85EBCA6Bu + 3u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) r[i] = r[i] * m + x;                                   // IMAD (fma pipe)
                else if (MODE == 1) r[i] = r[i] ^ (~r[(i + 1) & 15] & m);                  // LOP3 (alu pipe)
                else if (MODE == 2) r[i] = __funnelshift_l(r[i], r[(i + 5) & 15], 7);  // SHF (alu pipe)
                else r[i] = (uint32_t)__mulhi((int)r[i], (int)m) + x;                  // IMAD.HI (fma pipe: is the high half full rate?)
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= r[i];
    if (acc == 0x12345678u) out[blockIdx.x] = acc;      // keeps the chain live
}

// warp-level int8 MMA issue rate (legacy mma.sync path): 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) k_imma_peak(int32_t *out, int iters)
{
    int32_t d[8][4];
    uint32_t a[4] = {0x01020304u + threadIdx.x, 0x05060708u, 0x090a0b0cu, 0x0d0e0f10u}, b0 = 0x01010101u * (threadIdx.x & 7), b1 = 0x02020202u;
#pragma unroll
    for (int t = 0; t < 8; t++) for (int e = 0; e < 4; e++) d[t][e] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 8; t++) imma_16832(d[t], a, b0 + t, b1);
    }
    int32_t acc = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) for (int e = 0; e < 4; e++) acc ^= d[t][e];
    if (acc == 0x12345678) out[blockIdx.x] = acc;
}

extern "C" int kosk_b200_int_peak(kosk_b200_ctx *c, double *ops_per_s /* [5]: IMAD, LOP3, SHF thread-ops/s, IMMA int8 MAC/s, IMAD.HI thread-ops/s */)
{
    if (!c || !ops_per_s) return fail(KOSK_E_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    uint32_t *d = nullptr; CU(cudaMalloc(&d, blocks * 4));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, c->lanes[0].st));
            if (mode == 0) k_int_peak<0><<<blocks, 256, 0, c->lanes[0].st>>>(d, iters, 12345u + rep);
            else if (mode == 1) k_int_peak<1><<<blocks, 256, 0, c->lanes[0].st>>>(d, iters, 12345u + rep);
            else if (mode == 2) k_int_peak<2><<<blocks, 256, 0, c->lanes[0].st>>>(d, iters, 12345u + rep);
            else k_int_peak<3><<<blocks, 256, 0, c->lanes[0].st>>>(d, iters, 12345u + rep);
            CU(cudaEventRecord(e1, c->lanes[0].st));
            CU(cudaEventSynchronize(e1));
            float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        ops_per_s[mode == 3 ? 4 : mode] = (double)blocks * 256.0 * iters * 64.0 / (best * 1e-3);
    }
    {
        float best = 1e30f; const int it2 = 8192;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, c->lanes[0].st));
            k_imma_peak<<<blocks, 256, 0, c->lanes[0].st>>>((int32_t *)d, it2);
            CU(cudaEventRecord(e1, c->lanes[0].st));
            CU(cudaEventSynchronize(e1));
            float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        ops_per_s[3] = (double)blocks * 8.0 * it2 * 8.0 * (16.0 * 8 * 32) / (best * 1e-3);     // warps x iters x tiles x MACs per m16n8k32
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return KOSK_OK;
}
