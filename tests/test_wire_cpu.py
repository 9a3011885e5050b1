"""CPU suite for the compact wire format (include/kosk_b200.h): the host codec of libkosk_b200.so against an
independent numpy restatement built from the reference's struct mpcith_proof field table (mlwe_prover.hpp:57-75, SURVEY
Appendix B), on real reference-generated proofs (the CPU oracle) and on boundary patterns, for every SIMD path."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mpcith_kyber_kosk_b200 as pkg
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NP_, NT, MK = 1454, 150, 70
NR = NP_ - NT


def field_table(k):
    """(name, bytes, is_u16) of struct mpcith_proof in declaration order (mlwe_prover.hpp:57-75)."""
    eta = 3 if k == 2 else 2
    F, E, M = MK + 2 * k + 1, 2 * eta + 1, 2 * eta
    t = [("f", 2 * NT * F, 1), ("Tf", 2 * NT * F, 1), ("beta", 2 * NR * MK, 1), ("gamma", 2 * NR * MK, 1), ("Tcomm", 32 * NR, 0), ("I", 2 * NT, 1),
         ("s", 2 * NT * k, 1), ("e", 2 * NT * k, 1), ("t", 2 * NR * k, 1)]
    t += [(n, 2 * NT * k, 1) for n in ("NTTs", "NTTe", "NTTAr", "NTTAs")]
    t += [("sr", 2 * NR * k, 1), ("er", 2 * NR * k, 1), ("seta", 2 * NR * k * E, 1), ("eeta", 2 * NR * k * E, 1), ("ssub", 2 * NT * k * E, 1), ("esub", 2 * NT * k * E, 1),
          ("zs", 2 * NT * k * M, 1), ("ze", 2 * NT * k * M, 1), ("us", 2 * NR * k * M, 1), ("ue", 2 * NR * k * M, 1), ("comm", 32 * NR, 0)]
    return t


def np_pack12(v):
    v = v.astype(np.uint32)
    assert v.size % 2 == 0
    a0, a1 = v[0::2], v[1::2]
    out = np.empty((v.size // 2, 3), np.uint8)
    out[:, 0] = a0 & 0xFF
    out[:, 1] = ((a0 >> 8) | (a1 << 4)) & 0xFF
    out[:, 2] = (a1 >> 4) & 0xFF
    return out.reshape(-1)


def np_wire(k, pi):
    """numpy restatement: consecutive u16 fields are one 12-bit run, byte arrays are copied, segments 16-byte aligned."""
    pi = np.asarray(pi, np.uint8)
    segs, off, run = [], 0, []
    for name, nbytes, is16 in field_table(k):
        if is16:
            run.append(pi[off:off + nbytes])
        else:
            if run:
                segs.append(np_pack12(np.concatenate(run).view("<u2"))); run = []
            segs.append(pi[off:off + nbytes])
        off += nbytes
    assert off == pi.size and not run
    out = []
    for s in segs:
        out.append(s)
        pad = (-sum(x.size for x in out)) % 16
        out.append(np.zeros(pad, np.uint8))
    return np.concatenate(out)


@pytest.mark.parametrize("k", [2, 3, 4])
def test_wire_sizes(built_lib, k):
    assert sum(b for _, b, _ in field_table(k)) == pkg.proof_bytes(k)
    n16 = sum(b for _, b, i in field_table(k) if i) // 2
    assert pkg.wire_bytes(k) == {2: 519136, 3: 531616, 4: 578992}[k]
    assert abs(pkg.wire_bytes(k) - (n16 * 3 // 2 + 2 * 32 * NR)) < 48
    assert pkg.wire_bytes(5) == 0 and pkg.wire_bytes(k) % 16 == 0


@pytest.mark.parametrize("k", [2, 3, 4])
def test_wire_codec_matches_numpy_model_on_reference_proofs(built_lib, k):
    pis = np.stack([np.frombuffer(bytes(O.oracle_prove(k, O.seed_of(900 + i))[2]), np.uint8) for i in range(3)])
    w = pkg.wire_pack(k, pis, threads=1)
    for i in range(3):
        assert (w[i] == np_wire(k, pis[i])).all()
    assert (pkg.wire_unpack(k, w, threads=1) == pis).all()
    # the threaded pool gives the same bytes
    assert (pkg.wire_pack(k, pis, threads=3) == w).all()
    assert (pkg.wire_unpack(k, w, threads=3) == pis).all()


@pytest.mark.parametrize("k", [2, 4])
def test_wire_codec_boundary_patterns(built_lib, k):
    nb = pkg.proof_bytes(k)
    rng = np.random.default_rng(5)
    pats = []
    for fill in (0, 3328, 4095):                          # all-zero, all q-1, all 12-bit ones (non-canonical but representable)
        p = np.zeros(nb // 2, np.uint16) + fill
        pats.append(p)
    pats.append(rng.integers(0, 4096, nb // 2).astype(np.uint16))
    pats.append((np.arange(nb // 2) % 4096).astype(np.uint16))
    pis = np.stack([p.view(np.uint8) for p in pats]).copy()
    # digests are bytes: any value
    off = 0
    for name, nbytes, is16 in field_table(k):
        if not is16:
            pis[:, off:off + nbytes] = rng.integers(0, 256, (len(pats), nbytes))
        off += nbytes
    w = pkg.wire_pack(k, pis)
    for i in range(len(pats)):
        assert (w[i] == np_wire(k, pis[i])).all()
    assert (pkg.wire_unpack(k, w) == pis).all()
    # unaligned source / destination buffers (proof i of a batch is only 4-byte aligned; callers may pass any address)
    big = np.zeros(len(pats) * nb + 7, np.uint8)
    for shift in (1, 2, 4):
        lib = pkg.load_library()
        import ctypes
        rc = lib.kosk_b200_wire_unpack(k, len(pats), ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(big.ctypes.data + shift), 1)
        assert rc == 0 and (big[shift:shift + len(pats) * nb].reshape(len(pats), nb) == pis).all()


def test_wire_pack_rejects_unrepresentable(built_lib):
    k = 2
    pi = np.frombuffer(bytes(O.oracle_prove(k, O.seed_of(3))[2]), np.uint8).copy()
    tab, off = field_table(k), 0
    for name, nbytes, is16 in tab:
        if is16:
            for pos in (off, off + nbytes - 2):          # first and last element of every u16 field
                bad = pi.copy()
                bad[pos + 1] |= 0x10                        # bit 12 of the element
                with pytest.raises(pkg.KoskError, match="4096"):
                    pkg.wire_pack(k, bad[None])
        else:                                             # digest bytes are unconstrained
            ok = pi.copy(); ok[off] ^= 0xFF; ok[off + nbytes - 1] ^= 0xFF
            assert (pkg.wire_unpack(k, pkg.wire_pack(k, ok[None]))[0] == ok).all()
        off += nbytes


@pytest.mark.parametrize("simd", ["0", "1"])
def test_wire_codec_lower_simd_paths(built_lib, simd):
    """KOSK_B200_WIRE_SIMD forces the scalar (0) / AVX2 (1) code path; the bytes must not depend on it."""
    code = ("import numpy as np, mpcith_kyber_kosk_b200 as p\n"
            "rng = np.random.default_rng(11)\n"
            "for k in (2, 3, 4):\n"
            "    pi = rng.integers(0, 4096, (2, p.proof_bytes(k) // 2)).astype(np.uint16).view(np.uint8).reshape(2, -1)\n"
            "    w = p.wire_pack(k, pi, 2)\n"
            "    assert (p.wire_unpack(k, w, 2) == pi).all()\n"
            "    print(p.wire_simd(), int(w.astype(np.uint64).sum()))\n")
    env = dict(os.environ, KOSK_B200_WIRE_SIMD=simd, PYTHONPATH=ROOT)
    forced = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split("\n")
    env.pop("KOSK_B200_WIRE_SIMD")
    default = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split("\n")
    assert [l.split()[1:] for l in forced if l] == [l.split()[1:] for l in default if l]
    assert forced[0].split()[0] in ("scalar", "avx2")
