"""mpcith_kyber_kosk_b200 -- B200-native (sm_100a) MPC-in-the-head prover/verifier core for the Kyber
MLWE knowledge-of-secret-key relation, drop-in behind the reference's kosk.hpp API and proof layout."""
from .api import (KoskContext, KoskError, kyber_keypair, kyber_kosk_verify, kyber_verifiable_keygen, load_library,
                  pk_bytes, proof_bytes, sk_bytes, wire_bytes, wire_pack, wire_unpack, wire_simd, EXPORTS, LIB_PATH)

__all__ = ["KoskContext", "KoskError", "kyber_keypair", "kyber_kosk_verify", "kyber_verifiable_keygen", "load_library",
           "pk_bytes", "proof_bytes", "sk_bytes", "wire_bytes", "wire_pack", "wire_unpack", "wire_simd", "EXPORTS", "LIB_PATH"]
