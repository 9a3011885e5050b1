/* oracle/ok_tables.c -- TEST INFRASTRUCTURE. See ok_tables.h. */
#include "ok_tables.h"
#include <stdlib.h>
#include <pthread.h>

static uint16_t inv_tab[OK_Q];
static uint16_t *tab_S, *tab_R1, *tab_R2;
static pthread_once_t once = PTHREAD_ONCE_INIT;

static uint32_t mulq(uint32_t a, uint32_t b) { return a * b % OK_Q; }
static uint32_t subq(uint32_t a, uint32_t b) { return (a + OK_Q - b) % OK_Q; }

static void lagrange_impl(uint16_t *out, const uint16_t *nodes, int n, const uint16_t *targets, int nt)
{
    /* barycentric weights w_k = 1/prod_{m!=k}(x_k-x_m) */
    uint16_t *w = (uint16_t *)malloc(sizeof(uint16_t) * n);
    for (int k = 0; k < n; k++) {
        uint32_t d = 1;
        for (int m = 0; m < n; m++) if (m != k) d = mulq(d, subq(nodes[k] % OK_Q, nodes[m] % OK_Q));
        w[k] = inv_tab[d];
    }
    for (int t = 0; t < nt; t++) {
        uint32_t z = targets[t] % OK_Q; int hit = -1; uint32_t full = 1;
        for (int m = 0; m < n; m++) {
            uint32_t d = subq(z, nodes[m] % OK_Q);
            if (d == 0) hit = m; else full = mulq(full, d);
        }
        for (int k = 0; k < n; k++) {
            if (hit >= 0) out[t * n + k] = (k == hit);
            else out[t * n + k] = (uint16_t)mulq(mulq(full, w[k]), inv_tab[subq(z, nodes[k] % OK_Q)]);
        }
    }
    free(w);
}

static void build(void)
{
    inv_tab[0] = 0;
    for (uint32_t a = 1; a < OK_Q; a++) {      /* a^(q-2) */
        uint32_t r = 1, b = a, e = OK_Q - 2;
        while (e) { if (e & 1) r = mulq(r, b); b = mulq(b, b); e >>= 1; }
        inv_tab[a] = (uint16_t)r;
    }
    uint16_t nodes[813], targets[1303];
    tab_S = (uint16_t *)malloc(2 * 1303 * 407); tab_R1 = (uint16_t *)malloc(2 * 256 * 407); tab_R2 = (uint16_t *)malloc(2 * 256 * 813);
    for (int j = 0; j < 407; j++) nodes[j] = j;
    for (int x = 0; x < 1303; x++) targets[x] = x + 407;
    lagrange_impl(tab_S, nodes, 407, targets, 1303);
    for (int j = 0; j < 813; j++) nodes[j] = 256 + j;
    for (int i = 0; i < 256; i++) targets[i] = i;
    lagrange_impl(tab_R1, nodes, 407, targets, 256);
    lagrange_impl(tab_R2, nodes, 813, targets, 256);
}

void ok_lagrange_matrix(uint16_t *out, const uint16_t *nodes, int n, const uint16_t *targets, int nt)
{ pthread_once(&once, build); lagrange_impl(out, nodes, n, targets, nt); }
uint16_t ok_gf_inv(uint16_t a) { pthread_once(&once, build); return inv_tab[a % OK_Q]; }
const uint16_t *ok_table_share_ddeg(void)  { pthread_once(&once, build); return tab_S; }
const uint16_t *ok_table_recon_ddeg(void)  { pthread_once(&once, build); return tab_R1; }
const uint16_t *ok_table_recon_2ddeg(void) { pthread_once(&once, build); return tab_R2; }
