// include/dropin/ss.hpp -- share_vec / secret_vec (reference ss.hpp:33-42) come from ../kosk_dropin.hpp.  The sharing functions
// themselves (share_secrets_ddeg etc., ss.hpp:44-54) are internal to the device pipeline; kosk_b200_share_eval is their C-ABI form.
#ifndef KOSK_DROPIN_SS_HPP
#define KOSK_DROPIN_SS_HPP
#include "../kosk_dropin.hpp"
#endif
