"""Generates tests/golden/kosk_batch_golden.json from the UNMODIFIED reference (oracle/_ref; run where /root/reference exists:
`make -C oracle ref && python tests/golden/make_batch_golden.py`): SHA-256 of pk || sk || proof of the reference's own
kyber_verifiable_keygen under the KOSK counter-mode DRBG for
  * every proof of BASELINE configs[1] (Kyber512, batch of 1024; seed of proof i = seeds_for_range(BASE, 0, 1024)[i]),
  * sampled proofs of 1024-proof batches of Kyber768 / Kyber1024 (the bench path: chunk 1024),
  * sampled proofs of an 8192-proof Kyber768 slice (BASELINE configs[3]'s shape: chunk 4096, two lanes).
The GPU suite proves the full batches and compares digests."""
import hashlib
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402
from mpcith_kyber_kosk_b200.sharding import seeds_for_range  # noqa: E402

BASE = 0xC0F2
SAMPLES_1024 = [0, 1, 31, 32, 255, 256, 511, 512, 767, 1000, 1022, 1023]
SAMPLES_8192 = [0, 1, 4095, 4096, 4097, 6000, 8190, 8191]


def digest(k, seed):
    pk, sk, pi = O.ref_prove(k, seed)
    return hashlib.sha256(bytes(pk) + bytes(sk) + bytes(pi)).hexdigest()


def main():
    for k in (2, 3, 4):
        assert O.ref(k) is not None, "build oracle/_ref first"
    threads = os.cpu_count() or 1
    pool = ThreadPoolExecutor(threads)          # ctypes releases the GIL
    out = {"rng": "SHAKE256(seed||LE32(call))", "seed_rule": f"mpcith_kyber_kosk_b200.sharding.seeds_for_range({BASE:#x}, 0, n)[i]",
           "digest": "sha256(pk || sk || proof)"}
    s1024 = seeds_for_range(BASE, 0, 1024)
    out["k2_1024_all"] = list(pool.map(lambda i: digest(2, s1024[i]), range(1024)))
    print("k2 done", flush=True)
    for k in (3, 4):
        out[f"k{k}_1024_sampled"] = {str(i): d for i, d in zip(SAMPLES_1024, pool.map(lambda i: digest(k, s1024[i]), SAMPLES_1024))}
    s8192 = seeds_for_range(BASE, 0, 8192)
    out["k3_8192_sampled"] = {str(i): d for i, d in zip(SAMPLES_8192, pool.map(lambda i: digest(3, s8192[i]), SAMPLES_8192))}
    with open(os.path.join(HERE, "kosk_batch_golden.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
