/* oracle/ref_supp/precomputed_kyber_gen.c -- TEST INFRASTRUCTURE.
 * Stand-in for the reference's missing utils/precomputed_kyber.c (.MISSING_LARGE_BLOBS:1):
 * defines the four accessors of utils/precomputed_kyber.h:10-13 on top of the regenerated
 * Lagrange tables (oracle/ok_tables.c). Table content is forced by the Lagrange identities. */
#include "../ok_tables.h"
uint16_t get_precomputed_share_coeff_ddeg(int x, int i)  { return ok_table_share_ddeg()[x * 407 + i]; }
uint16_t get_precomputed_recon_coeff_ddeg(int x, int i)  { return ok_table_recon_ddeg()[x * 407 + i]; }
uint16_t get_precomputed_recon_coeff_2ddeg(int x, int i) { return ok_table_recon_2ddeg()[x * 813 + i]; }
uint16_t get_precomputed_share_coeff_2ddeg(int x, int i) { (void)x; (void)i; return 0; } /* never called */
