"""Generates tests/golden/kosk_golden.json from the UNMODIFIED reference compiled into oracle/_ref (run in the
build container, where /root/reference exists: `make -C oracle ref && python tests/golden/make_golden.py`).
Everything is produced by the reference's own kyber_verifiable_keygen / kyber_kosk_verify under the KOSK
counter-mode DRBG (oracle/ok_rng.c)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    out = {"rng": "SHAKE256(seed||LE32(call))", "seed_rule": "sha256(b'kosk-b200:' + str(i))", "cases": [], "tamper": {}}
    for k in (2, 3, 4):
        assert O.ref(k) is not None, "build oracle/_ref first"
        L = O.layout(k)
        for i in range(3):
            seed = O.seed_of(i)
            pk, sk, pi = O.ref_prove(k, seed)
            I = pi[L.o_I:L.o_I + 300].view(np.uint16)
            out["cases"].append({
                "k": k, "seed_index": i, "seed": seed.hex(),
                "pk_sha256": hashlib.sha256(pk).hexdigest(), "sk_sha256": hashlib.sha256(sk).hexdigest(),
                "proof_sha256": hashlib.sha256(pi).hexdigest(), "pk_head": bytes(pk[:48]).hex(),
                "I": [int(x) for x in I], "proof_head": bytes(pi[:32]).hex(),
                "verify": bool(O.ref_verify(k, pi, pk)),
            })
        # accept/reject of the reference under single-bit tampering, seed 0
        pk, sk, pi = O.ref_prove(k, O.seed_of(0))
        offs = [(n, getattr(L, n)) for n in O.FIELDS] + [("end", L.proof_bytes)]
        tam = {}
        for (n, o), (_, e) in zip(offs[:-1], offs[1:]):
            step = 1 if n in ("o_Tcomm", "o_comm") else 2
            for tag, off in (("first", o), ("last", e - step)):
                t = pi.copy(); t[off] ^= 1
                tam[f"{n[2:]}:{tag}"] = bool(O.ref_verify(k, t, pk))
        t = pk.copy(); t[5] ^= 1
        tam["pk:t"] = bool(O.ref_verify(k, pi, t))
        t = pk.copy(); t[-1] ^= 1
        tam["pk:seed"] = bool(O.ref_verify(k, pi, t))
        out["tamper"][str(k)] = tam
        print(k, "accepted:", [n for n, v in tam.items() if v], flush=True)
    with open(os.path.join(HERE, "kosk_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
