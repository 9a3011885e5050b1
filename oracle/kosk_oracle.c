/* oracle/kosk_oracle.c -- TEST INFRASTRUCTURE (CPU oracle); see kosk_oracle.h.
 * A restatement, not a copy: KYBER_K is a run-time argument, all field values are kept as canonical
 * residues (the reference's Montgomery/Barrett forms are value-transparent), the verifier's
 * NTL interpolate/eval (third-party NTL 11.5.1, not vendored) is restated as one Lagrange matrix per
 * proof.  Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "kosk_oracle.h"
#include "ok_keccak.h"
#include "ok_rng.h"
#include "ok_tables.h"
#include <stdlib.h>
#include <stdio.h>

#define Q KO_Q
#define N KO_N
#define T KO_T
#define R KO_R
#define D1 KO_D1
#define D2 KO_D2

typedef uint16_t u16;

/* ---- utils/gf3329.c:274-284,308-323: exact u16 semantics (also for non-canonical inputs) ---- */
static u16 gf_add(u16 a, u16 b) { return (u16)(a + b < Q ? a + b : a + b - Q); }
static u16 gf_sub(u16 a, u16 b) { return (u16)(a < b ? a + Q - b : a - b); }
static u16 gf_mul(u16 a, u16 b) { return (u16)((uint32_t)a * b % Q); }

void ko_sha3_256(uint8_t h[32], const uint8_t *in, size_t n) { ok_sha3_256(h, in, n); }
void ko_sha3_512(uint8_t h[64], const uint8_t *in, size_t n) { ok_sha3_512(h, in, n); }
void ko_shake256(uint8_t *o, size_t on, const uint8_t *in, size_t n) { ok_shake256(o, on, in, n); }
void ko_shake128(uint8_t *o, size_t on, const uint8_t *in, size_t n) { ok_shake128(o, on, in, n); }
void ko_keccak_f1600(uint64_t a[25]) { ok_keccak_f1600(a); }
void ko_randombytes_at(const uint8_t seed[32], uint32_t c, uint8_t *out, size_t n)
{
    uint8_t in[36]; memcpy(in, seed, 32);
    in[32] = (uint8_t)c; in[33] = (uint8_t)(c >> 8); in[34] = (uint8_t)(c >> 16); in[35] = (uint8_t)(c >> 24);
    ok_shake256(out, n, in, 36);
}

/* kyber/symmetric-shake.c:43-51 : SHAKE256(key || nonce) */
static void prf(uint8_t *out, size_t outlen, const uint8_t key[32], uint8_t nonce)
{
    uint8_t ext[33]; memcpy(ext, key, 32); ext[32] = nonce; ok_shake256(out, outlen, ext, 33);
}

/* ---- layout: mlwe_prover.hpp:57-75 (no padding) ---- */
int kosk_oracle_layout(int k, ko_layout *L)
{
    if (k < 2 || k > 4) return -1;
    L->k = k; L->eta = (k == 2) ? 3 : 2;                  /* kyber/params.h:29-41 */
    L->F = KO_MK + 2 * k + 1; L->E = 2 * L->eta + 1; L->M = 2 * L->eta;
    L->pk_bytes = 384 * (size_t)k + 32; L->sk_bytes = 384 * (size_t)k + L->pk_bytes + 64;
    size_t o = 0;
#define FIELD(name, bytes) do { L->name = o; o += (bytes); } while (0)
    FIELD(o_f, 2 * (size_t)T * L->F);      FIELD(o_Tf, 2 * (size_t)T * L->F);
    FIELD(o_beta, 2 * (size_t)R * KO_MK);  FIELD(o_gamma, 2 * (size_t)R * KO_MK);
    FIELD(o_Tcomm, (size_t)R * 32);        FIELD(o_I, 2 * (size_t)T);
    FIELD(o_s, 2 * (size_t)T * k);         FIELD(o_e, 2 * (size_t)T * k);       FIELD(o_t, 2 * (size_t)R * k);
    FIELD(o_NTTs, 2 * (size_t)T * k);      FIELD(o_NTTe, 2 * (size_t)T * k);
    FIELD(o_NTTAr, 2 * (size_t)T * k);     FIELD(o_NTTAs, 2 * (size_t)T * k);
    FIELD(o_sr, 2 * (size_t)R * k);        FIELD(o_er, 2 * (size_t)R * k);
    FIELD(o_seta, 2 * (size_t)R * k * L->E); FIELD(o_eeta, 2 * (size_t)R * k * L->E);
    FIELD(o_ssub, 2 * (size_t)T * k * L->E); FIELD(o_esub, 2 * (size_t)T * k * L->E);
    FIELD(o_zs, 2 * (size_t)T * k * L->M);   FIELD(o_ze, 2 * (size_t)T * k * L->M);
    FIELD(o_us, 2 * (size_t)R * k * L->M);   FIELD(o_ue, 2 * (size_t)R * k * L->M);
    FIELD(o_comm, (size_t)R * 32);
#undef FIELD
    L->proof_bytes = o;
    return 0;
}

/* ---- Kyber NTT in plain residues.  kyber/ntt.c:39-56 stores zetas[i] = 17^brv7(i) * 2^16 (centered);
 * fqmul (ntt.c:68-70) multiplies by 2^-16, so in plain residues the twiddle is 17^brv7(i). ---- */
static u16 zeta_plain[128];
static void zetas_init(void)
{
    if (zeta_plain[0]) return;
    for (int i = 0; i < 128; i++) {
        int br = 0; for (int b = 0; b < 7; b++) br |= ((i >> b) & 1) << (6 - b);
        uint32_t z = 1; for (int e = 0; e < br; e++) z = z * 17 % Q;
        zeta_plain[i] = (u16)z;
    }
}
/* kyber/ntt.c:80-95 followed by poly_reduce (poly.c:261-265); canonical in, canonical out */
void ko_ntt(u16 r[256])
{
    zetas_init();
    int k = 1;
    for (int len = 128; len >= 2; len >>= 1)
        for (int start = 0; start < 256; start += 2 * len) {
            u16 z = zeta_plain[k++];
            for (int j = start; j < start + len; j++) {
                u16 t = gf_mul(z, r[j + len]);
                r[j + len] = gf_sub(r[j], t);
                r[j] = gf_add(r[j], t);
            }
        }
}
/* polyvec_basemul_acc_montgomery (polyvec.c:202-214; basemul ntt.c:139-146, zeta signs poly.c:290-297)
 * followed by poly_tomont (poly.c:307-313): net scaling 2^-16 * 2^16 = 1, i.e. the plain product. */
void ko_basemul_acc(int k, u16 r[256], const u16 *a, const u16 *b)
{
    zetas_init();
    for (int c = 0; c < 256; c++) r[c] = 0;
    for (int v = 0; v < k; v++) {
        const u16 *x = a + 256 * v, *y = b + 256 * v;
        for (int i = 0; i < 128; i++) {
            u16 z = zeta_plain[64 + i / 2]; if (i & 1) z = gf_sub(0, z);
            u16 r0 = gf_add(gf_mul(gf_mul(x[2 * i + 1], y[2 * i + 1]), z), gf_mul(x[2 * i], y[2 * i]));
            u16 r1 = gf_add(gf_mul(x[2 * i], y[2 * i + 1]), gf_mul(x[2 * i + 1], y[2 * i]));
            r[2 * i] = gf_add(r[2 * i], r0); r[2 * i + 1] = gf_add(r[2 * i + 1], r1);
        }
    }
}
/* indcpa.c:124-193 : SHAKE128(seed || j || i) with 12-bit rejection sampling; A[i][j] not transposed */
void ko_gen_matrix(int k, u16 *A, const uint8_t seed[32])
{
    for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) {
        uint8_t ext[34]; memcpy(ext, seed, 32); ext[32] = (uint8_t)j; ext[33] = (uint8_t)i;
        ok_sponge s; ok_sponge_init(&s, 168); ok_sponge_absorb(&s, ext, 34); ok_sponge_finalize(&s, 0x1F);
        u16 *p = A + 256 * (i * k + j); int ctr = 0;
        while (ctr < 256) {                 /* blocks are multiples of 3 bytes, so no carry-over (indcpa.c:161-163) */
            uint8_t b[3]; ok_sponge_squeeze(&s, b, 3);
            u16 v0 = (u16)((b[0] | ((u16)b[1] << 8)) & 0xFFF), v1 = (u16)(((b[1] >> 4) | ((u16)b[2] << 4)) & 0xFFF);
            if (v0 < Q) p[ctr++] = v0;
            if (ctr < 256 && v1 < Q) p[ctr++] = v1;
        }
    }
}
/* cbd.c:58-107; output canonical (encode_to_gf3329 of the signed value, gf3329.c:308-310) */
static void cbd_eta(int eta, u16 r[256], const uint8_t *buf)
{
    if (eta == 2) {
        for (int i = 0; i < 32; i++) {
            uint32_t t = buf[4 * i] | (uint32_t)buf[4 * i + 1] << 8 | (uint32_t)buf[4 * i + 2] << 16 | (uint32_t)buf[4 * i + 3] << 24;
            uint32_t d = (t & 0x55555555) + ((t >> 1) & 0x55555555);
            for (int j = 0; j < 8; j++) { int a = (d >> (4 * j)) & 3, b = (d >> (4 * j + 2)) & 3; r[8 * i + j] = (u16)((a - b + Q) % Q); }
        }
    } else {
        for (int i = 0; i < 64; i++) {
            uint32_t t = buf[3 * i] | (uint32_t)buf[3 * i + 1] << 8 | (uint32_t)buf[3 * i + 2] << 16;
            uint32_t d = (t & 0x249249) + ((t >> 1) & 0x249249) + ((t >> 2) & 0x249249);
            for (int j = 0; j < 4; j++) { int a = (d >> (6 * j)) & 7, b = (d >> (6 * j + 3)) & 7; r[4 * i + j] = (u16)((a - b + Q) % Q); }
        }
    }
}
/* poly.c:124-139 on canonical coefficients */
static void poly_tobytes(uint8_t *r, const u16 a[256])
{
    for (int i = 0; i < 128; i++) {
        u16 t0 = a[2 * i], t1 = a[2 * i + 1];
        r[3 * i] = (uint8_t)t0; r[3 * i + 1] = (uint8_t)((t0 >> 8) | (t1 << 4)); r[3 * i + 2] = (uint8_t)(t1 >> 4);
    }
}

/* ---- ss.cpp ---- */
void ko_share_ddeg(u16 sh[N], const u16 y[D1])             /* ss.cpp:76-99 (and :13-33) */
{
    const u16 *S = ok_table_share_ddeg();
    for (int i = 0; i <= T; i++) sh[i] = y[i + 256];
    for (int x = 0; x < N - (T + 1); x++) {
        u16 acc = 0;
        for (int j = 0; j < D1; j++) acc = gf_add(acc, gf_mul(y[j], S[x * D1 + j]));
        sh[T + 1 + x] = acc;
    }
}
static void share_fresh(u16 sh[N], const u16 secret[256])   /* ss.cpp:3-34 */
{
    uint8_t rb[(T + 1) * 2]; u16 y[D1];
    randombytes(rb, sizeof rb);
    memcpy(y, secret, 512);
    for (int i = 0; i <= T; i++) y[256 + i] = (u16)(((rb[2 * i] << 8) | rb[2 * i + 1]) % Q);
    ko_share_ddeg(sh, y);
}
void ko_recon_ddeg(u16 sec[256], const u16 sh[D1])          /* ss.cpp:37-54 */
{
    const u16 *R1 = ok_table_recon_ddeg();
    for (int i = 0; i < 256; i++) { u16 a = 0; for (int j = 0; j < D1; j++) a = gf_add(a, gf_mul(sh[j], R1[i * D1 + j])); sec[i] = a; }
}
void ko_recon_2ddeg(u16 sec[256], const u16 sh[D2])         /* ss.cpp:56-73 */
{
    const u16 *R2 = ok_table_recon_2ddeg();
    for (int i = 0; i < 256; i++) { u16 a = 0; for (int j = 0; j < D2; j++) a = gf_add(a, gf_mul(sh[j], R2[i * D2 + j])); sec[i] = a; }
}
static void vadd(u16 *r, const u16 *a, const u16 *b) { for (int i = 0; i < N; i++) r[i] = gf_add(a[i], b[i]); } /* ss.cpp:101-111 */
static void vsub(u16 *r, const u16 *a, const u16 *b) { for (int i = 0; i < N; i++) r[i] = gf_sub(a[i], b[i]); } /* ss.cpp:114-124 */
static void vmul(u16 *r, const u16 *a, const u16 *b) { for (int i = 0; i < N; i++) r[i] = gf_mul(a[i], b[i]); } /* ss.cpp:126-136 */

static __thread ko_trace g_trace;
const ko_trace *kosk_oracle_last_trace(void) { return &g_trace; }

/* FS challenge helpers */
static void power_table(int F, u16 *pw, const u16 *alpha, int n)   /* mlwe_prover.cpp:144-153 ; pw[n][F] */
{
    for (int i = 0; i < n; i++) { pw[i * F] = 1; pw[i * F + 1] = alpha[i]; for (int j = 2; j < F; j++) pw[i * F + j] = gf_mul(pw[i * F + j - 1], alpha[i]); }
}
static void derive_alpha(int nalpha, u16 *alpha, uint8_t digest[32], const uint8_t *tcomm_all)  /* mlwe_prover.cpp:130-142 */
{
    uint8_t ab[2 * 78];
    ok_sha3_256(digest, tcomm_all, (size_t)N * 32);
    prf(ab, 2 * (size_t)nalpha, digest, 1);
    for (int i = 0; i < nalpha; i++) alpha[i] = (u16)(((ab[2 * i] << 8) | ab[2 * i + 1]) % Q);
}
static void derive_I(u16 I[T], uint8_t ch[32], const uint8_t *views_all)   /* mlwe_prover.cpp:445-474 */
{
    uint8_t ib[2 * T];
    ok_sha3_256(ch, views_all, (size_t)N * 32);
    prf(ib, sizeof ib, ch, 1);
    for (int i = 0; i < T; i++) I[i] = (u16)(((ib[2 * i] << 8) | ib[2 * i + 1]) % N);
    for (int i = 1; i < T; i++) {
        u16 inc = 0; int dup;
        do { dup = 0; for (int j = 0; j < i; j++) if ((I[i] + inc) % N == I[j]) { dup = 1; inc = (u16)(inc + 1); break; } } while (dup);
        I[i] = (u16)((I[i] + inc) % N);
    }
}
/* beta/gamma/r evaluation for one party: mlwe_prover.cpp:159-203 (c0 = f[0] for j<70, f[71] for j>=70) */
static u16 eval_comb(int F, const u16 *pw_row, const u16 *fsh, int c0)
{
    u16 acc = fsh[c0];
    for (int k = 1; k < F; k++) acc = gf_add(acc, gf_mul(pw_row[k], fsh[k]));
    return acc;
}
static void put16(uint8_t *p, u16 v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static u16 get16(const uint8_t *p) { return (u16)(p[0] | (p[1] << 8)); }

/* ================= prover: kosk.cpp:72-86 ================= */
void kosk_oracle_verifiable_keygen(int k, const uint8_t seed[32], int rng_mode, uint8_t *pk, uint8_t *sk, uint8_t *pi)
{
    ko_layout L; if (kosk_oracle_layout(k, &L)) return;
    const int F = L.F, E = L.E, M = L.M, eta = L.eta, V2 = 2 * k, NA = KO_MK + V2;
    kosk_rng_reset(seed, rng_mode);

    /* ---- kyber_keygen: kosk.cpp:4-70 ---- */
    uint8_t buf[64];
    u16 *A = malloc(2 * 256 * (size_t)k * k), s[4][256], e[4][256], sh_[4][256], eh[4][256], th[4][256];
    randombytes(buf, 64); buf[32] = (uint8_t)k; { uint8_t h[64]; ok_sha3_512(h, buf, 33); memcpy(buf, h, 64); }
    ko_gen_matrix(k, A, buf);
    for (int i = 0; i < 2 * k; i++) {
        uint8_t nb[192]; prf(nb, (size_t)eta * 64, buf + 32, (uint8_t)i);
        cbd_eta(eta, i < k ? s[i] : e[i - k], nb);
    }
    for (int i = 0; i < k; i++) { memcpy(sh_[i], s[i], 512); ko_ntt(sh_[i]); memcpy(eh[i], e[i], 512); ko_ntt(eh[i]); }
    for (int i = 0; i < k; i++) {
        ko_basemul_acc(k, th[i], A + 256 * i * k, &sh_[0][0]);
        for (int c = 0; c < 256; c++) th[i][c] = gf_add(th[i][c], eh[i][c]);
    }
    for (int i = 0; i < k; i++) poly_tobytes(pk + 384 * i, th[i]);
    memcpy(pk + 384 * k, buf, 32);
    for (int i = 0; i < k; i++) poly_tobytes(sk + 384 * i, sh_[i]);
    memcpy(sk + 384 * k, pk, L.pk_bytes);
    ok_sha3_256(sk + L.sk_bytes - 64, pk, L.pk_bytes);
    memcpy(sk + L.sk_bytes - 32, buf + 32, 32);             /* kosk.cpp:67-69: z = noise seed */

    /* ---- prepare_randomness: mlwe_prover.cpp:4-39 ---- */
    u16 (*f)[256] = malloc(512 * (size_t)F), (*Tf)[256] = malloc(512 * (size_t)F);
    u16 (*f_sh)[N] = malloc(2 * N * (size_t)F), (*Tf_sh)[N] = malloc(2 * N * (size_t)F);
    for (int i = 0; i < F; i++) {
        uint8_t sd[32], pb[512]; randombytes(sd, 32); prf(pb, 512, sd, (uint8_t)i);
        for (int j = 0; j < 256; j++) f[i][j] = (u16)(((pb[2 * j] << 8) | pb[2 * j + 1]) % Q);
    }
    for (int i = 0; i < F; i++) { memcpy(Tf[i], f[i], 512); ko_ntt(Tf[i]); }
    for (int i = 0; i < F; i++) { share_fresh(f_sh[i], f[i]); share_fresh(Tf_sh[i], Tf[i]); }
    memcpy(g_trace.first_share, f_sh[0], 2 * N); memcpy(g_trace.first_secret, f[0], 512); memcpy(g_trace.first_ntt, Tf[0], 512);

    /* ---- prepare_range_proof: mlwe_prover.cpp:41-59 ---- */
    u16 (*seta)[N] = malloc(2 * N * (size_t)k * E), (*eeta)[N] = malloc(2 * N * (size_t)k * E);
    for (int i = 0; i < k; i++) for (int j = 0; j < E; j++) {
        u16 c[256], ev = (u16)((j - eta + Q) % Q); for (int x = 0; x < 256; x++) c[x] = ev;
        share_fresh(seta[i * E + j], c); share_fresh(eeta[i * E + j], c);
    }

    /* ---- prove: mlwe_prover.cpp:81-538 ---- */
    u16 (*s_sh)[N] = malloc(2 * N * 4), (*e_sh)[N] = malloc(2 * N * 4);
    for (int i = 0; i < k; i++) { share_fresh(s_sh[i], s[i]); share_fresh(e_sh[i], e[i]); }           /* :96-101 */
    uint8_t *tcomm = malloc((size_t)N * 32), *views = malloc((size_t)N * 32);
    for (int p = 0; p < N; p++) {                                                                   /* :116-127 */
        uint8_t rec[2 * 2 * (4 + 79)]; int o = 0;
        for (int j = 0; j < k; j++) put16(rec + 2 * (o++), s_sh[j][p]);
        for (int j = 0; j < k; j++) put16(rec + 2 * (o++), e_sh[j][p]);
        for (int j = 0; j < F; j++) put16(rec + 2 * (o++), f_sh[j][p]);
        for (int j = 0; j < F; j++) put16(rec + 2 * (o++), Tf_sh[j][p]);
        ok_sha3_256(tcomm + 32 * p, rec, 2 * (size_t)o);
    }
    u16 alpha[78], *pw = malloc(2 * (size_t)NA * F);
    derive_alpha(NA, alpha, g_trace.fs1_digest, tcomm); memcpy(g_trace.alpha, alpha, sizeof alpha);
    power_table(F, pw, alpha, NA);
    u16 (*beta)[KO_MK] = malloc(2 * KO_MK * (size_t)N), (*gamma)[KO_MK] = malloc(2 * KO_MK * (size_t)N);
    u16 (*r_sh)[N] = malloc(2 * N * 8), (*Tr_sh)[N] = malloc(2 * N * 8);
    for (int p = 0; p < N; p++) {                                                                   /* :159-214 */
        u16 fp[79], tp[79];
        for (int j = 0; j < F; j++) { fp[j] = f_sh[j][p]; tp[j] = Tf_sh[j][p]; }
        for (int j = 0; j < KO_MK; j++) { beta[p][j] = eval_comb(F, pw + j * F, fp, 0); gamma[p][j] = eval_comb(F, pw + j * F, tp, 0); }
        for (int j = 0; j < V2; j++) { r_sh[j][p] = eval_comb(F, pw + (KO_MK + j) * F, fp, KO_MK + 1); Tr_sh[j][p] = eval_comb(F, pw + (KO_MK + j) * F, tp, KO_MK + 1); }
    }
    u16 (*sr_sh)[N] = malloc(2 * N * 4), (*er_sh)[N] = malloc(2 * N * 4), sr_rnd[4][D1], er_rnd[4][D1], sr[4][256];
    for (int i = 0; i < k; i++) {                                                                   /* :222-277 */
        u16 er[256];
        vadd(sr_sh[i], s_sh[i], r_sh[i]);     ko_recon_ddeg(sr[i], sr_sh[i]);
        vadd(er_sh[i], e_sh[i], r_sh[i + k]); ko_recon_ddeg(er, er_sh[i]);
        for (int j = 256; j < D1; j++) { sr_rnd[i][j] = sr_sh[i][j - 256]; er_rnd[i][j] = er_sh[i][j - 256]; }
        ko_ntt(sr[i]); ko_ntt(er);
        memcpy(sr_rnd[i], sr[i], 512); memcpy(er_rnd[i], er, 512);
    }
    u16 (*Ts_sh)[N] = malloc(2 * N * 4), (*Te_sh)[N] = malloc(2 * N * 4), (*As_sh)[N] = malloc(2 * N * 4),
        (*Ar_sh)[N] = malloc(2 * N * 4), (*t_sh)[N] = malloc(2 * N * 4), *tmp = malloc(2 * N);
    for (int i = 0; i < k; i++) {                                                                   /* :296-304 */
        ko_share_ddeg(tmp, sr_rnd[i]); vsub(Ts_sh[i], tmp, Tr_sh[i]);
        ko_share_ddeg(tmp, er_rnd[i]); vsub(Te_sh[i], tmp, Tr_sh[i + k]);
    }
    for (int i = 0; i < k; i++) {                                                                   /* :279-318 */
        u16 As[256], Asr_rnd[D1];
        ko_basemul_acc(k, As, A + 256 * i * k, &sh_[0][0]);
        ko_basemul_acc(k, Asr_rnd, A + 256 * i * k, &sr[0][0]);
        for (int j = 256; j < D1; j++) Asr_rnd[j] = sr_rnd[i][j];
        ko_share_ddeg(tmp, Asr_rnd);
        share_fresh(As_sh[i], As);
        vsub(Ar_sh[i], tmp, As_sh[i]);
    }
    for (int i = 0; i < k; i++) vadd(t_sh[i], As_sh[i], Te_sh[i]);                                    /* :321-323 */

    /* range proof: mlwe_prover.cpp:338-392 */
    u16 (*ssub)[N] = malloc(2 * N * (size_t)k * E), (*esub)[N] = malloc(2 * N * (size_t)k * E);
    u16 (*zs)[N] = malloc(2 * N * (size_t)k * M), (*ze)[N] = malloc(2 * N * (size_t)k * M);
    u16 (*us)[N] = malloc(2 * N * (size_t)k * M), (*ue)[N] = malloc(2 * N * (size_t)k * M);
    for (int i = 0; i < k; i++) for (int j = 0; j < E; j++) { vsub(ssub[i * E + j], s_sh[i], seta[i * E + j]); vsub(esub[i * E + j], e_sh[i], eeta[i * E + j]); }
    for (int i = 0; i < k; i++) for (int j = 0; j < M; j++) {
        u16 z2s[N], z2e[N], sec[256];
        vmul(z2s, j == 0 ? ssub[i * E] : zs[i * M + j - 1], ssub[i * E + j + 1]);
        vmul(z2e, j == 0 ? esub[i * E] : ze[i * M + j - 1], esub[i * E + j + 1]);
        ko_recon_2ddeg(sec, z2s); share_fresh(zs[i * M + j], sec);
        ko_recon_2ddeg(sec, z2e); share_fresh(ze[i * M + j], sec);
        vsub(us[i * M + j], z2s, zs[i * M + j]); vsub(ue[i * M + j], z2e, ze[i * M + j]);
    }
    /* views: mlwe_prover.cpp:395-444 */
    for (int p = 0; p < N; p++) {
        uint8_t vw[32 + 2 * (6 * 4 + 2 * 79 + 8 * 3 * 4)]; int o = 32; memcpy(vw, tcomm + 32 * p, 32);
#define PUT(v) do { put16(vw + o, (v)); o += 2; } while (0)
        for (int j = 0; j < k; j++) PUT(s_sh[j][p]);
        for (int j = 0; j < k; j++) PUT(e_sh[j][p]);
        for (int j = 0; j < F; j++) PUT(f_sh[j][p]);
        for (int j = 0; j < F; j++) PUT(Tf_sh[j][p]);
        for (int j = 0; j < k; j++) PUT(beta[p][j]);
        for (int j = 0; j < k; j++) PUT(gamma[p][j]);
        for (int j = 0; j < k; j++) PUT(sr_sh[j][p]);
        for (int j = 0; j < k; j++) PUT(er_sh[j][p]);
        for (int j = 0; j < k; j++) {
            for (int m = 0; m < M; m++) PUT(zs[j * M + m][p]);
            for (int m = 0; m < M; m++) PUT(ze[j * M + m][p]);
            for (int m = 0; m < M; m++) PUT(us[j * M + m][p]);
            for (int m = 0; m < M; m++) PUT(ue[j * M + m][p]);
        }
#undef PUT
        ok_sha3_256(views + 32 * p, vw, (size_t)o);
    }
    memcpy(g_trace.tcomm0, tcomm, 32); memcpy(g_trace.view0, views, 32);
    u16 I[T]; derive_I(I, g_trace.fs2_digest, views); memcpy(g_trace.I, I, sizeof I);
    /* assemble: mlwe_prover.cpp:480-537 */
    uint8_t inI[N]; memset(inI, 0, N); u16 rest[R];
    for (int i = 0; i < T; i++) inI[I[i]] = 1;
    for (int p = 0, j = 0; p < N; p++) if (!inI[p]) rest[j++] = (u16)p;
    for (int i = 0; i < T; i++) {
        int p = I[i];
        put16(pi + L.o_I + 2 * i, I[i]);
        for (int j = 0; j < F; j++) { put16(pi + L.o_f + 2 * (i * F + j), f_sh[j][p]); put16(pi + L.o_Tf + 2 * (i * F + j), Tf_sh[j][p]); }
        for (int j = 0; j < k; j++) {
            size_t x = 2 * (size_t)(i * k + j);
            put16(pi + L.o_s + x, s_sh[j][p]); put16(pi + L.o_e + x, e_sh[j][p]);
            put16(pi + L.o_NTTs + x, Ts_sh[j][p]); put16(pi + L.o_NTTe + x, Te_sh[j][p]);
            put16(pi + L.o_NTTAr + x, Ar_sh[j][p]); put16(pi + L.o_NTTAs + x, As_sh[j][p]);
            for (int m = 0; m < E; m++) { put16(pi + L.o_ssub + 2 * ((i * k + j) * E + m), ssub[j * E + m][p]); put16(pi + L.o_esub + 2 * ((i * k + j) * E + m), esub[j * E + m][p]); }
            for (int m = 0; m < M; m++) { put16(pi + L.o_zs + 2 * ((i * k + j) * M + m), zs[j * M + m][p]); put16(pi + L.o_ze + 2 * ((i * k + j) * M + m), ze[j * M + m][p]); }
        }
    }
    for (int i = 0; i < R; i++) {
        int p = rest[i];
        for (int j = 0; j < KO_MK; j++) { put16(pi + L.o_beta + 2 * (i * KO_MK + j), beta[p][j]); put16(pi + L.o_gamma + 2 * (i * KO_MK + j), gamma[p][j]); }
        for (int j = 0; j < k; j++) {
            size_t x = 2 * (size_t)(i * k + j);
            put16(pi + L.o_sr + x, sr_sh[j][p]); put16(pi + L.o_er + x, er_sh[j][p]); put16(pi + L.o_t + x, t_sh[j][p]);
            for (int m = 0; m < E; m++) { put16(pi + L.o_seta + 2 * ((i * k + j) * E + m), seta[j * E + m][p]); put16(pi + L.o_eeta + 2 * ((i * k + j) * E + m), eeta[j * E + m][p]); }
            for (int m = 0; m < M; m++) { put16(pi + L.o_us + 2 * ((i * k + j) * M + m), us[j * M + m][p]); put16(pi + L.o_ue + 2 * ((i * k + j) * M + m), ue[j * M + m][p]); }
        }
        memcpy(pi + L.o_Tcomm + 32 * (size_t)i, tcomm + 32 * p, 32);
        memcpy(pi + L.o_comm + 32 * (size_t)i, views + 32 * p, 32);
    }
    free(A); free(f); free(Tf); free(f_sh); free(Tf_sh); free(seta); free(eeta); free(s_sh); free(e_sh); free(tcomm); free(views);
    free(pw); free(beta); free(gamma); free(r_sh); free(Tr_sh); free(sr_sh); free(er_sh); free(Ts_sh); free(Te_sh); free(As_sh);
    free(Ar_sh); free(t_sh); free(tmp); free(ssub); free(esub); free(zs); free(ze); free(us); free(ue);
}

/* ================= verifier: kosk.cpp:88-117 + mlwe_verifier.cpp:4-686 ================= */
static void apply_rows(u16 *out, const u16 *Lm, int rows, int n, const u16 *y)  /* NTL interpolate+eval, restated */
{
    for (int t = 0; t < rows; t++) { uint32_t a = 0; for (int j = 0; j < n; j++) a = (a + (uint32_t)Lm[t * n + j] * (y[j] % Q)) % Q; out[t] = (u16)a; }
}
#define FAIL(tag) do { ok = 0; goto done; } while (0)

int kosk_oracle_verify(int k, const uint8_t *pi, const uint8_t *pk)
{
    ko_layout L; if (kosk_oracle_layout(k, &L)) return 0;
    const int F = L.F, E = L.E, M = L.M, eta = L.eta, V2 = 2 * k, NA = KO_MK + V2;
    int ok = 1;
    /* kosk.cpp:94-112 : t from pk (12-bit raw), A from seed */
    u16 *A = malloc(2 * 256 * (size_t)k * k), tpk[4][256];
    for (int i = 0; i < k; i++) for (int c = 0; c < 128; c++) {                 /* poly.c:151-158 */
        const uint8_t *a = pk + 384 * i + 3 * c;
        tpk[i][2 * c] = (u16)((a[0] | ((u16)a[1] << 8)) & 0xFFF); tpk[i][2 * c + 1] = (u16)(((a[1] >> 4) | ((u16)a[2] << 4)) & 0xFFF);
    }
    ko_gen_matrix(k, A, pk + 384 * k);
#define PI16(off, idx) get16(pi + (off) + 2 * (size_t)(idx))
    /* V1: mlwe_verifier.cpp:9-19. Out-of-range or duplicate I is UB there; restated as reject. */
    u16 I[T], rest[R]; uint8_t inI[N]; int posI[N]; memset(inI, 0, N);
    u16 *Lm1 = malloc(2 * (size_t)D1 * D1), *Lm2 = malloc(2 * (size_t)256 * D2), *pw = malloc(2 * (size_t)NA * F);
    u16 (*beta)[KO_MK] = malloc(2 * KO_MK * (size_t)N), (*gamma)[KO_MK] = malloc(2 * KO_MK * (size_t)N);
    uint8_t *tcomm = malloc((size_t)N * 32), *views = malloc((size_t)N * 32);
    u16 (*sr_sh)[N] = malloc(2 * N * 4), (*er_sh)[N] = malloc(2 * N * 4), *tmp = malloc(2 * N), *tmp2 = malloc(2 * N);
    u16 (*us)[N] = malloc(2 * N * (size_t)k * M), (*ue)[N] = malloc(2 * N * (size_t)k * M);
    for (int i = 0; i < T; i++) { I[i] = PI16(L.o_I, i); if (I[i] >= N || inI[I[i]]) FAIL("I"); inI[I[i]] = 1; posI[I[i]] = i; }
    for (int p = 0, j = 0; p < N; p++) if (!inI[p]) rest[j++] = (u16)p;
    /* V2: :22-38 */
    for (int i = 0; i < T; i++) {
        uint8_t rec[2 * 2 * (4 + 79)]; int o = 0;
        for (int j = 0; j < k; j++) put16(rec + 2 * (o++), PI16(L.o_s, i * k + j));
        for (int j = 0; j < k; j++) put16(rec + 2 * (o++), PI16(L.o_e, i * k + j));
        for (int j = 0; j < F; j++) put16(rec + 2 * (o++), PI16(L.o_f, i * F + j));
        for (int j = 0; j < F; j++) put16(rec + 2 * (o++), PI16(L.o_Tf, i * F + j));
        ok_sha3_256(tcomm + 32 * I[i], rec, 2 * (size_t)o);
    }
    for (int i = 0; i < R; i++) memcpy(tcomm + 32 * rest[i], pi + L.o_Tcomm + 32 * (size_t)i, 32);
    /* V3: :41-65 */
    u16 alpha[78]; uint8_t dg[32]; derive_alpha(NA, alpha, dg, tcomm); power_table(F, pw, alpha, NA);
    /* V4: :67-96 */
    u16 r_op[T][8], Tr_op[T][8];
    for (int i = 0; i < T; i++) {
        u16 fp[79], tp[79];
        for (int j = 0; j < F; j++) { fp[j] = PI16(L.o_f, i * F + j); tp[j] = PI16(L.o_Tf, i * F + j); }
        for (int j = 0; j < KO_MK; j++) { beta[I[i]][j] = eval_comb(F, pw + j * F, fp, 0); gamma[I[i]][j] = eval_comb(F, pw + j * F, tp, 0); }
        for (int j = 0; j < V2; j++) { r_op[i][j] = eval_comb(F, pw + (KO_MK + j) * F, fp, KO_MK + 1); Tr_op[i][j] = eval_comb(F, pw + (KO_MK + j) * F, tp, KO_MK + 1); } /* V8 :148-170 */
    }
    for (int i = 0; i < R; i++) for (int j = 0; j < KO_MK; j++) { beta[rest[i]][j] = PI16(L.o_beta, i * KO_MK + j); gamma[rest[i]][j] = PI16(L.o_gamma, i * KO_MK + j); }
    /* V5+V6: :97-124 (V7 :126-142 is a tautology) */
    for (int j = 0; j < KO_MK; j++) {
        u16 bs[D1], gs[D1], bsec[256], gsec[256];
        for (int p = 0; p < D1; p++) { bs[p] = beta[p][j]; gs[p] = gamma[p][j]; }
        ko_recon_ddeg(bsec, bs); ko_recon_ddeg(gsec, gs); ko_ntt(bsec);
        if (memcmp(bsec, gsec, 512)) FAIL("beta/gamma");
    }
    /* per-proof Lagrange matrices over the rest-party nodes (NTL interpolate+eval restated) */
    { u16 nodes[D2], tg[D1];
      for (int j = 0; j < D2; j++) nodes[j] = (u16)(rest[j] + 256);
      for (int j = 0; j < D1; j++) tg[j] = (u16)j;
      ok_lagrange_matrix(Lm1, nodes, D1, tg, D1); ok_lagrange_matrix(Lm2, nodes, D2, tg, 256); }
    /* V9: :173-247 */
    u16 sr_rnd[4][D1], er_rnd[4][D1], srp[4][256];
    for (int i = 0; i < k; i++) {
        u16 y[D1], yv[D1];
        for (int j = 0; j < D1; j++) y[j] = PI16(L.o_sr, j * k + i);
        apply_rows(yv, Lm1, D1, D1, y); ko_share_ddeg(sr_sh[i], yv); memcpy(sr_rnd[i], yv, sizeof yv); memcpy(srp[i], yv, 512);
        for (int j = 0; j < D1; j++) y[j] = PI16(L.o_er, j * k + i);
        apply_rows(yv, Lm1, D1, D1, y); ko_share_ddeg(er_sh[i], yv); memcpy(er_rnd[i], yv, sizeof yv);
        for (int j = 0; j < R; j++) { if (sr_sh[i][rest[j]] != PI16(L.o_sr, j * k + i)) FAIL("s+r"); if (er_sh[i][rest[j]] != PI16(L.o_er, j * k + i)) FAIL("e+r"); }
    }
    /* V10: :257-284 */
    for (int i = 0; i < k; i++) { ko_ntt(sr_rnd[i]); ko_ntt(er_rnd[i]); memcpy(srp[i], sr_rnd[i], 512); }
    for (int i = 0; i < k; i++) {
        ko_share_ddeg(tmp, sr_rnd[i]); ko_share_ddeg(tmp2, er_rnd[i]);
        for (int j = 0; j < T; j++) {
            if (PI16(L.o_NTTs, j * k + i) != gf_sub(tmp[I[j]], Tr_op[j][i])) FAIL("NTT s");
            if (PI16(L.o_NTTe, j * k + i) != gf_sub(tmp2[I[j]], Tr_op[j][i + k])) FAIL("NTT e");
        }
    }
    /* V11: :287-312 */
    for (int i = 0; i < k; i++) {
        u16 y[D1]; ko_basemul_acc(k, y, A + 256 * i * k, &srp[0][0]);
        for (int j = 256; j < D1; j++) y[j] = sr_rnd[i][j];
        ko_share_ddeg(tmp, y);
        for (int j = 0; j < T; j++) if (tmp[I[j]] != gf_add(PI16(L.o_NTTAs, j * k + i), PI16(L.o_NTTAr, j * k + i))) FAIL("A(s+r)");
    }
    /* V12: :316-376 */
    for (int i = 0; i < k; i++) {
        u16 y[D1], yv[D1];
        for (int j = 0; j < D1; j++) y[j] = PI16(L.o_t, j * k + i);
        apply_rows(yv, Lm1, D1, D1, y);
        for (int c = 0; c < 256; c++) if (yv[c] != tpk[i][c]) FAIL("t");       /* :354-363: canonical vs raw 12-bit */
        ko_share_ddeg(tmp, yv);
        for (int j = 0; j < T; j++) if (tmp[I[j]] != gf_add(PI16(L.o_NTTAs, j * k + i), PI16(L.o_NTTe, j * k + i))) FAIL("t=As+e");
    }
    /* V13: :382-466 */
    for (int i = 0; i < k; i++) for (int m = 0; m < E; m++) for (int w = 0; w < 2; w++) {
        size_t oe = w ? L.o_eeta : L.o_seta, os = w ? L.o_esub : L.o_ssub, ov = w ? L.o_e : L.o_s;
        u16 y[D1], yv[D1], cur = gf_sub((u16)m, (u16)eta);
        for (int j = 0; j < D1; j++) y[j] = PI16(oe, (j * k + i) * E + m);
        apply_rows(yv, Lm1, D1, D1, y);
        for (int c = 0; c < 256; c++) if (yv[c] != cur) FAIL("eta const");
        ko_share_ddeg(tmp, yv);
        for (int j = 0; j < T; j++) if (PI16(os, (j * k + i) * E + m) != gf_sub(PI16(ov, j * k + i), tmp[I[j]])) FAIL("x-eta");
    }
    /* V14+V15: :469-571 */
    for (int i = 0; i < k; i++) for (int m = 0; m < M; m++) for (int w = 0; w < 2; w++) {
        size_t osub = w ? L.o_esub : L.o_ssub, oz = w ? L.o_ze : L.o_zs, ou = w ? L.o_ue : L.o_us;
        u16 *ush = w ? ue[i * M + m] : us[i * M + m], y[D2], sec[256];
        for (int j = 0; j < T; j++) {
            u16 a = (m == 0) ? PI16(osub, (j * k + i) * E) : PI16(oz, (j * k + i) * M + m - 1);
            u16 z2 = gf_mul(a, PI16(osub, (j * k + i) * E + m + 1));
            ush[I[j]] = gf_sub(z2, PI16(oz, (j * k + i) * M + m));
        }
        for (int j = 0; j < D2; j++) y[j] = PI16(ou, (j * k + i) * M + m);
        apply_rows(sec, Lm2, 256, D2, y);
        for (int c = 0; c < 256; c++) if (sec[c] != 0) FAIL("u");
        for (int j = 0; j < R; j++) ush[rest[j]] = PI16(ou, (j * k + i) * M + m);
        ko_recon_2ddeg(sec, ush);
        for (int c = 0; c < 256; c++) if (sec[c] != 0) FAIL("u2d");
    }
    /* V16: :584-683 */
    for (int i = 0; i < T; i++) {
        uint8_t vw[32 + 2 * (6 * 4 + 2 * 79 + 8 * 3 * 4)]; int o = 32, p = I[i]; memcpy(vw, tcomm + 32 * p, 32);
#define PUT(v) do { put16(vw + o, (v)); o += 2; } while (0)
        for (int j = 0; j < k; j++) PUT(PI16(L.o_s, i * k + j));
        for (int j = 0; j < k; j++) PUT(PI16(L.o_e, i * k + j));
        for (int j = 0; j < F; j++) PUT(PI16(L.o_f, i * F + j));
        for (int j = 0; j < F; j++) PUT(PI16(L.o_Tf, i * F + j));
        for (int j = 0; j < k; j++) PUT(beta[p][j]);
        for (int j = 0; j < k; j++) PUT(gamma[p][j]);
        for (int j = 0; j < k; j++) PUT(sr_sh[j][p]);
        for (int j = 0; j < k; j++) PUT(er_sh[j][p]);
        for (int j = 0; j < k; j++) {
            for (int m = 0; m < M; m++) PUT(PI16(L.o_zs, (i * k + j) * M + m));
            for (int m = 0; m < M; m++) PUT(PI16(L.o_ze, (i * k + j) * M + m));
            for (int m = 0; m < M; m++) PUT(us[j * M + m][p]);
            for (int m = 0; m < M; m++) PUT(ue[j * M + m][p]);
        }
#undef PUT
        ok_sha3_256(views + 32 * p, vw, (size_t)o);
    }
    for (int i = 0; i < R; i++) memcpy(views + 32 * rest[i], pi + L.o_comm + 32 * (size_t)i, 32);
    { u16 I2[T]; uint8_t ch[32]; derive_I(I2, ch, views); if (memcmp(I2, I, sizeof I)) FAIL("I"); }
done:
    (void)posI; (void)r_op;
    free(A); free(Lm1); free(Lm2); free(pw); free(beta); free(gamma); free(tcomm); free(views); free(sr_sh); free(er_sh);
    free(tmp); free(tmp2); free(us); free(ue);
    return ok;
}
