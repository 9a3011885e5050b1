/* include/dropin/kyber/indcpa.h -- placeholder for the reference's kyber/indcpa.h (see kem.h next to this file). */
#ifndef KOSK_DROPIN_KYBER_INDCPA_H
#define KOSK_DROPIN_KYBER_INDCPA_H
#include <stdint.h>
#endif
