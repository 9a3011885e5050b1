"""Generates tests/golden/kosk_struct_golden.json from the UNMODIFIED reference compiled into oracle/_ref (build container
only: `make -C oracle ref && python tests/golden/make_struct_golden.py`): SHA-256 of the struct images the reference's own
prepare_randomness / prepare_range_proof / kyber_keygen / prove leave behind when called in main.cpp's order (main.cpp:16-47)
under the KOSK counter-mode DRBG, with the never-initialised share_vec.len fields zeroed."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    out = {"rng": "SHAKE256(seed||LE32(call)), one counter across the five calls", "order": "prepare_randomness, prepare_range_proof, kyber_keygen, prove, verify",
           "seed_rule": "sha256(b'kosk-b200:' + str(i))", "cases": []}
    for k in (2, 3, 4):
        assert O.ref(k) is not None, "build oracle/_ref first"
        for i in (0, 1):
            seed = O.seed_of(i)
            r = O.ref_struct_sequence(k, seed)
            case = {"k": k, "seed_index": i, "seed": seed.hex(), "verify": r["ok"]}
            for n in ("rand", "eta", "inst", "pk", "sk", "pi"):
                case[n + "_sha256"] = hashlib.sha256(bytes(r[n])).hexdigest()
            out["cases"].append(case)
            print(k, i, case["verify"], case["pi_sha256"][:16], flush=True)
    with open(os.path.join(HERE, "kosk_struct_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
