"""Generates tests/golden/kosk_kem_golden.json from the UNMODIFIED reference compiled into oracle/_ref (build container only):
crypto_kem_enc_derand / crypto_kem_dec (kyber/kem.c:76-169) on keys made by the reference's kyber_verifiable_keygen, including
decapsulation of tampered ciphertexts (implicit rejection) and of a ciphertext under the wrong key."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    out = {"seed_rule": "key i: kyber_verifiable_keygen seed sha256(b'kosk-b200:' + str(i)); coins j: sha256(b'kem-coins:' + str(j))", "cases": []}
    for k in (2, 3, 4):
        assert O.ref(k) is not None, "build oracle/_ref first"
        for i in (0, 1):
            pk, sk, _ = O.ref_prove(k, O.seed_of(i))
            for j in (0, 1):
                coins = O.seed_of(j, b"kem-coins")
                ct, ss = O.ref_kem_enc_derand(k, pk, coins)
                assert bytes(O.ref_kem_dec(k, ct, sk)) == bytes(ss)
                bad = ct.copy(); bad[7] ^= 0x10
                bad2 = ct.copy(); bad2[-1] ^= 0x80
                out["cases"].append({
                    "k": k, "key_index": i, "coins_index": j, "pk_sha256": hashlib.sha256(bytes(pk)).hexdigest(),
                    "ct_sha256": hashlib.sha256(bytes(ct)).hexdigest(), "ss": bytes(ss).hex(),
                    "ss_reject_byte7": bytes(O.ref_kem_dec(k, bad, sk)).hex(), "ss_reject_last": bytes(O.ref_kem_dec(k, bad2, sk)).hex(),
                })
            print(k, i, out["cases"][-1]["ss"][:16], flush=True)
    with open(os.path.join(HERE, "kosk_kem_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
