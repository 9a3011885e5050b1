// kosk_common.cuh -- shared constants, proof layout, plane ("slot") map and GF(3329) helpers.
//
// Domain vocabulary follows the reference: parties (N=1454), opened set I (T=150), rest set (R=1304),
// sharings (one d-degree packed-Shamir sharing = one row of 1454 party shares), secrets (256 per sharing).
// Reference parameters: params.hpp:12-37, ss.hpp:56-57, kyber/params.h:29-41.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define KOSK_HD __host__ __device__ __forceinline__
#else
#define KOSK_HD inline
#endif

namespace kosk {

typedef uint16_t u16;
typedef uint8_t u8;

constexpr int Q = 3329;
constexpr int NP = 1454;          // MPCITH_N parties
constexpr int NT = 150;           // MPCITH_T opened parties
constexpr int NR = NP - NT;       // rest parties
constexpr int NL = 256;           // packed secrets per sharing (KYBER_N)
constexpr int MK = 70;            // MPCITH_K
constexpr int D1 = 407;           // d-degree sharing: 256 secrets + 151 random evaluations
constexpr int D2 = 813;           // 2d-degree
constexpr int NX = NP - (NT + 1); // 1303 parties whose share is an evaluation of the table S
constexpr int YLD = 416;          // row stride (u16) of the sharing-input matrix Y (407 padded, 16B-aligned rows)
constexpr int SLD = 1456;         // row stride (u16) of a plane: [pad][party 0..1453][pad]; party 151 is 16B-aligned
constexpr int SOFF = 1;           // plane element of party p is row[SOFF + p]
constexpr int TREE_BYTES = NP * 32;

// ---- proof byte layout: struct mpcith_proof, reference mlwe_prover.hpp:57-75 (no padding) ----
struct Layout {
    int k, eta, F, E, M;
    size_t pk_bytes, sk_bytes, proof_bytes;
    size_t o_f, o_Tf, o_beta, o_gamma, o_Tcomm, o_I, o_s, o_e, o_t, o_NTTs, o_NTTe, o_NTTAr, o_NTTAs,
           o_sr, o_er, o_seta, o_eeta, o_ssub, o_esub, o_zs, o_ze, o_us, o_ue, o_comm;
};

KOSK_HD Layout make_layout(int k)
{
    Layout L;
    L.k = k; L.eta = (k == 2) ? 3 : 2; L.F = MK + 2 * k + 1; L.E = 2 * L.eta + 1; L.M = 2 * L.eta;
    L.pk_bytes = 384 * (size_t)k + 32; L.sk_bytes = 768 * (size_t)k + 96;
    size_t o = 0, TF = 2 * (size_t)NT * L.F, RK = 2 * (size_t)NR * MK, Tk = 2 * (size_t)NT * k, Rk = 2 * (size_t)NR * k;
    L.o_f = o; o += TF; L.o_Tf = o; o += TF; L.o_beta = o; o += RK; L.o_gamma = o; o += RK;
    L.o_Tcomm = o; o += (size_t)NR * 32; L.o_I = o; o += 2 * NT;
    L.o_s = o; o += Tk; L.o_e = o; o += Tk; L.o_t = o; o += Rk;
    L.o_NTTs = o; o += Tk; L.o_NTTe = o; o += Tk; L.o_NTTAr = o; o += Tk; L.o_NTTAs = o; o += Tk;
    L.o_sr = o; o += Rk; L.o_er = o; o += Rk;
    L.o_seta = o; o += Rk * L.E; L.o_eeta = o; o += Rk * L.E;
    L.o_ssub = o; o += Tk * L.E; L.o_esub = o; o += Tk * L.E;
    L.o_zs = o; o += Tk * L.M; L.o_ze = o; o += Tk * L.M;
    L.o_us = o; o += Rk * L.M; L.o_ue = o; o += Rk * L.M;
    L.o_comm = o; o += (size_t)NR * 32;
    L.proof_bytes = o;
    return L;
}

// ---- compact wire format of a proof: the same fields in the same order, field elements as 12 bits ----
// struct mpcith_proof is u16 field elements (and the u16 party indices I) around two byte arrays (Tcomm, comm).  On the wire the
// two u16 runs A = [f .. gamma] and B = [I .. u_e] are packed two elements per three bytes, little-endian like kyber/poly.c:124-139
// (b0 = a0, b1 = a0 >> 8 | a1 << 4, b2 = a1 >> 4), the digests are copied, every segment starts 16-byte aligned:
//   wire = pack12(A) | pad | Tcomm[NR][32] | pack12(B) | pad | comm[NR][32]          519 136 / 531 616 / 578 992 bytes for K = 2 / 3 / 4
// Every u16 < 4096 is representable (all canonical proofs are); decoding is the exact inverse.
struct WireLayout {
    uint32_t nA, nB;                       // u16 elements of the two runs (both even for K = 2, 3, 4)
    size_t o_A, o_Tcomm, o_B, o_comm, proof_bytes;    // byte offsets in struct mpcith_proof
    size_t w_A, w_Tcomm, w_B, w_comm, wire_bytes;     // byte offsets on the wire
};
KOSK_HD WireLayout make_wire_layout(int k)
{
    const Layout L = make_layout(k);
    WireLayout W;
    W.o_A = 0; W.o_Tcomm = L.o_Tcomm; W.o_B = L.o_I; W.o_comm = L.o_comm; W.proof_bytes = L.proof_bytes;
    W.nA = (uint32_t)(L.o_Tcomm / 2); W.nB = (uint32_t)((L.o_comm - L.o_I) / 2);
    const size_t HB = (size_t)NR * 32;
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    W.w_A = 0; W.w_Tcomm = al(((size_t)W.nA + 1) / 2 * 3); W.w_B = W.w_Tcomm + HB;
    W.w_comm = al(W.w_B + ((size_t)W.nB + 1) / 2 * 3); W.wire_bytes = al(W.w_comm + HB);
    return W;
}

// ---- plane map: every per-party u16 quantity of one proof lives in plane[slot][party] ----
// Slots < n2 are sharings produced by the share-evaluation kernel from a Y row (256 secrets | 151 tail);
// the first n1 of them have no dependence on the Fiat-Shamir challenge and go through one launch.
struct Slots {
    int k, eta, F, E, M;
    int f0, Tf0, seta0, eeta0, s0, e0, zs0, ze0, n1;      // challenge-independent sharings
    int Tsr0, Ter0, Asr0, As0, n2;                        // sharings after FS-1
    int R0, TR0, SR0, ER0, US0, UE0, B0, G0, TC0, nslot;  // derived planes
    // randombytes() call schedule of one kyber_verifiable_keygen
    int c_seed0, c_f0, c_eta0, c_se0, c_As0, c_z0, ncalls;
};

KOSK_HD Slots make_slots(int k)
{
    Slots s;
    s.k = k; s.eta = (k == 2) ? 3 : 2; s.F = MK + 2 * k + 1; s.E = 2 * s.eta + 1; s.M = 2 * s.eta;
    int o = 0;
    s.f0 = o; o += s.F; s.Tf0 = o; o += s.F; s.seta0 = o; o += k * s.E; s.eeta0 = o; o += k * s.E;
    s.s0 = o; o += k; s.e0 = o; o += k; s.zs0 = o; o += k * s.M; s.ze0 = o; o += k * s.M; s.n1 = o;
    s.Tsr0 = o; o += k; s.Ter0 = o; o += k; s.Asr0 = o; o += k; s.As0 = o; o += k; s.n2 = o;
    s.R0 = o; o += 2 * k; s.TR0 = o; o += 2 * k; s.SR0 = o; o += k; s.ER0 = o; o += k;
    s.US0 = o; o += k * s.M; s.UE0 = o; o += k * s.M; s.B0 = o; o += k; s.G0 = o; o += k; s.TC0 = o; o += 16;
    s.nslot = o;
    s.c_seed0 = 1; s.c_f0 = 1 + s.F; s.c_eta0 = 1 + 3 * s.F; s.c_se0 = s.c_eta0 + 2 * k * s.E;
    s.c_As0 = s.c_se0 + 2 * k; s.c_z0 = s.c_As0 + k; s.ncalls = s.c_z0 + 2 * k * s.M;
    return s;
}

// ---- GF(3329) helpers on canonical residues ----
KOSK_HD uint32_t gf_add(uint32_t a, uint32_t b) { uint32_t r = a + b; return r >= (uint32_t)Q ? r - Q : r; }
KOSK_HD uint32_t gf_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + Q - b; }
KOSK_HD uint32_t gf_mul(uint32_t a, uint32_t b) { return (a * b) % (uint32_t)Q; }
KOSK_HD int32_t gf_center(uint32_t a) { return (int32_t)a > Q / 2 ? (int32_t)a - Q : (int32_t)a; }   // [0,q) -> [-1664,1664]
KOSK_HD uint32_t gf_canon(int32_t a) { int32_t r = a % Q; return (uint32_t)(r < 0 ? r + Q : r); }    // any int32 -> [0,q)
// reference u16 semantics for possibly non-canonical operands (utils/gf3329.c:274-280), used by the verifier
KOSK_HD uint16_t ref_add(uint16_t a, uint16_t b) { int s = (int)a + (int)b; return (uint16_t)(s < Q ? s : s - Q); }
KOSK_HD uint16_t ref_sub(uint16_t a, uint16_t b) { return (uint16_t)(a < b ? (int)a + Q - (int)b : (int)a - (int)b); }
KOSK_HD uint16_t be16_mod(uint8_t hi, uint8_t lo, int m) { return (uint16_t)((((uint32_t)hi << 8) | lo) % (uint32_t)m); }

}  // namespace kosk
