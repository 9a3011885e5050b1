/* oracle/ok_tables.h -- TEST INFRASTRUCTURE. Regenerates the Lagrange tables of the reference's
 * missing utils/precomputed_kyber.c (accessors declared at
 * reference utils/precomputed_kyber.h:10-13, used at ss.cpp:26-27,47,66,91-92).
 *   S [x][j] = l_j^{0..406}(x+407)          x<1303, j<407   (share_coeff_ddeg)
 *   R1[i][j] = l_{256+j}^{256..662}(i)      i<256,  j<407   (recon_coeff_ddeg)
 *   R2[i][j] = l_{256+j}^{256..1068}(i)     i<256,  j<813   (recon_coeff_2ddeg)
 * with l_j^X(z) = prod_{m in X, m!=j} (z-m)/(j-m) mod 3329.
 */
#ifndef OK_TABLES_H
#define OK_TABLES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define OK_Q 3329
const uint16_t *ok_table_share_ddeg(void);   /* [1303][407] */
const uint16_t *ok_table_recon_ddeg(void);   /* [256][407]  */
const uint16_t *ok_table_recon_2ddeg(void);  /* [256][813]  */
uint16_t ok_gf_inv(uint16_t a);
/* generic: out[t*n+k] = l_k^{nodes}(targets[t]) for distinct nodes; a target equal to a node gives a delta row */
void ok_lagrange_matrix(uint16_t *out, const uint16_t *nodes, int n, const uint16_t *targets, int nt);
#ifdef __cplusplus
}
#endif
#endif
