"""GPU suite (-m gpu): parity at the sizes that are measured (parity gaps 1-5).
  * BASELINE configs[1]: ALL 1024 Kyber512 proofs of one batch against SHA-256 digests produced by the unmodified reference
    (tests/golden/kosk_batch_golden.json, tests/golden/make_batch_golden.py);
  * the bench path (chunk 1024, one lane) for Kyber768 / Kyber1024: sampled proofs against reference digests, all accepted by the
    verifier at that chunk, and the reference's tamper matrix replayed inside a 1024-proof verify batch;
  * BASELINE configs[3]'s shape: a Kyber768 slice of 8192 proofs at chunk 4096 on two lanes, sampled against reference digests;
  * the reconstruction / interpolation kernels element-wise against the oracle (ko_recon_ddeg, ko_recon_2ddeg, ok_lagrange_matrix).
"""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import mpcith_kyber_kosk_b200 as pkg
import oracle_lib as O
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BG = json.load(open(os.path.join(HERE, "golden", "kosk_batch_golden.json")))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "kosk_golden.json")))
BASE = 0xC0F2


def digest(pk, sk, pi):
    return hashlib.sha256(bytes(pk) + bytes(sk) + bytes(pi)).hexdigest()


def test_config2_all_1024_proofs_match_reference_digests():
    ctx = pkg.KoskContext(2, 0, 1024, 1)                    # the bench configuration: one chunk of 1024 on one lane
    seeds = seeds_for_range(BASE, 0, 1024)
    ctx.set_wire(0)
    pk, sk, pi = ctx.prove_batch(seeds)
    got = [digest(pk[i], sk[i], pi[i]) for i in range(1024)]
    bad = [i for i in range(1024) if got[i] != BG["k2_1024_all"][i]]
    assert not bad, f"{len(bad)} of 1024 proofs differ from the reference, first {bad[:5]}"
    assert ctx.verify_batch(pi, pk).all()
    # the same batch through the 12-bit wire path: identical caller-visible bytes
    ctx.set_wire(100)
    pk2, sk2, pi2 = ctx.prove_batch(seeds)
    assert (pi2 == pi).all() and (pk2 == pk).all() and (sk2 == sk).all()
    ctx.close()


@pytest.mark.parametrize("k", [3, 4])
def test_bench_path_k3_k4_chunk_1024(k):
    ctx = pkg.KoskContext(k, 0, 1024, 1)
    seeds = seeds_for_range(BASE, 0, 1024)
    pk, sk, pi = ctx.prove_batch(seeds)
    for i, want in BG[f"k{k}_1024_sampled"].items():
        i = int(i)
        assert digest(pk[i], sk[i], pi[i]) == want, i
    assert ctx.verify_batch(pi, pk).all()
    # the reference's tamper matrix (golden) replayed inside one 1024-proof verify batch: proof 0's (pi, pk) of the
    # golden seed at every position, tamper t applied to position 5 + 7 t
    tam = GOLDEN["tamper"][str(k)]
    L = O.layout(k)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(0))
    pis = np.tile(np.asarray(opi, np.uint8), (1024, 1))
    pks = np.tile(np.asarray(opk, np.uint8), (1024, 1))
    want = np.ones(1024, bool)
    offs = [(n, getattr(L, n)) for n in O.FIELDS] + [("end", L.proof_bytes)]
    pos = 5
    for (n, o), (_, e) in zip(offs[:-1], offs[1:]):
        step = 1 if n in ("o_Tcomm", "o_comm") else 2
        for tag, off in (("first", o), ("last", e - step)):
            pis[pos, off] ^= 1
            want[pos] = tam[f"{n[2:]}:{tag}"]
            pos += 7
    pks[pos, 5] ^= 1; want[pos] = tam["pk:t"]; pos += 7
    pks[pos, -1] ^= 1; want[pos] = tam["pk:seed"]
    for pct in (0, 100):
        ctx.set_wire(pct)
        got = ctx.verify_batch(pis, pks)
        assert (got == want).all(), (pct, np.nonzero(got != want)[0][:8])
    ctx.close()


def test_config4_shape_k3_8192_chunk_4096_two_lanes():
    k, n = 3, 8192
    ctx = pkg.KoskContext(k, 0, 4096, 2)
    seeds = torch.from_numpy(seeds_for_range(BASE, 0, n)).cuda()
    d_pk = torch.empty(n * ctx.pk_bytes, dtype=torch.uint8, device="cuda")
    d_sk = torch.empty(n * ctx.sk_bytes, dtype=torch.uint8, device="cuda")
    d_pi = torch.empty(n * ctx.proof_bytes, dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx.prove_batch_device(n, seeds.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr())
    torch.cuda.synchronize()
    pkv, skv, piv = d_pk.view(n, -1), d_sk.view(n, -1), d_pi.view(n, -1)
    for i, want in BG["k3_8192_sampled"].items():
        i = int(i)
        assert digest(pkv[i].cpu().numpy(), skv[i].cpu().numpy(), piv[i].cpu().numpy()) == want, i
    ctx.verify_batch_device(n, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr())
    torch.cuda.synchronize()
    assert bool(d_ok.all())
    # a checksum of checksums over the whole slice: every proof distinct, every opened set well formed
    L = O.layout(k)
    I = piv[:, L.o_I:L.o_I + 300].cpu().numpy().view(np.uint16)
    assert (I < 1454).all() and all(len(set(r)) == 150 for r in I[::97])
    assert len({bytes(r[:16]) for r in piv[:, :16].cpu().numpy()}) == n
    ctx.close()


# ---------------- reconstruction / interpolation kernels against the oracle ----------------
def _oracle_recon(rows, degree2):
    lib = O.oracle()
    out = np.zeros((rows.shape[0], 256), np.uint16)
    for r in range(rows.shape[0]):
        row = np.ascontiguousarray(rows[r])
        (lib.ko_recon_2ddeg if degree2 else lib.ko_recon_ddeg)(ctypes.c_void_p(out[r].ctypes.data), ctypes.c_void_p(row.ctypes.data))
    return out


@pytest.mark.parametrize("degree2", [False, True])
def test_recon_rows_match_oracle(ctxs, degree2):
    nn = 813 if degree2 else 407
    rng = np.random.default_rng(nn)
    rows = rng.integers(0, 3329, (37, nn)).astype(np.uint16)
    rows[0] = 3328                                           # worst-case magnitude for the lazy sums
    rows[1] = 0
    rows[2] = np.arange(nn) % 3329
    got = ctxs(2).recon_rows(rows, degree2)
    assert (got == _oracle_recon(rows, degree2)).all()
    if not degree2:                                          # recon(share(y)) = y[:256]: the two tables are inverse on the secrets
        y = rng.integers(0, 3329, (9, 407)).astype(np.uint16)
        sh = ctxs(2).share_eval(y)
        assert (ctxs(2).recon_rows(sh[:, :407]) == y[:, :256]).all()


@pytest.mark.parametrize("degree2", [False, True])
@pytest.mark.parametrize("pattern", ["random", "low", "high"])
def test_interp_rows_match_oracle_lagrange_matrix(ctxs, degree2, pattern):
    """The verifier's interpolation (NTL interpolate + eval in the reference, mlwe_verifier.cpp:188-224 / :510-543) against the
    oracle's Lagrange matrix over the same rest-party nodes; opened sets that put all / none of the holes below the last node."""
    nn, nt = (813, 256) if degree2 else (407, 407)
    rng = np.random.default_rng(7 + nn)
    if pattern == "random":
        opened = rng.permutation(1454)[:150]
    elif pattern == "low":
        opened = np.arange(150)                              # parties 0..149 opened: every target 256..405 is a non-node, 406 a node
    else:
        opened = np.arange(1454 - 150, 1454)                 # no hole below the last node: the nodes are consecutive
    opened = opened.astype(np.uint16)
    rest = np.array(sorted(set(range(1454)) - set(int(x) for x in opened)), np.uint16)
    nodes = (rest[:nn] + 256).astype(np.uint16)
    targets = np.arange(nt, dtype=np.uint16)
    Lm = np.zeros((nt, nn), np.uint16)
    O.oracle().ok_lagrange_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    # targets equal to a node are not defined for the barycentric helper: evaluate those by hand (the interpolant passes through the share)
    node_of = {int(x): j for j, x in enumerate(nodes)}
    free_t = np.array([t for t in targets if int(t) not in node_of], np.uint16)
    Lf = np.zeros((free_t.size, nn), np.uint16)
    O.oracle().ok_lagrange_matrix(ctypes.c_void_p(Lf.ctypes.data), ctypes.c_void_p(nodes.ctypes.data), nn, ctypes.c_void_p(free_t.ctypes.data), int(free_t.size))
    for row, t in zip(Lf, free_t):
        Lm[int(t)] = row
    for t, j in node_of.items():
        if t < nt:
            Lm[t, j] = 1
    shares = rng.integers(0, 3329, (11, nn)).astype(np.uint16)
    shares[0] = 3328
    want = (shares.astype(np.int64) @ Lm.astype(np.int64).T) % 3329
    got = ctxs(2).interp_rows(opened, shares, degree2)
    assert (got == want).all()


# ---------------- struct-level prove() on inconsistent instances (the prover's algebraic shortcuts must not change the bytes) ----------------
@pytest.mark.parametrize("k", [2, 3, 4])
def test_struct_prove_on_inconsistent_instances_matches_live_reference(ctxs, k):
    """DESIGN 3 removes the prover's reconstructions algebraically (the opened masked secret is s + r; the range-proof products come
    straight from s / e).  Those identities hold for ANY instance, not only for t = As + e with s, e in [-eta, eta]: prove() on an instance
    with a wrong t, an s coefficient outside the range and an e coefficient outside the range must give the reference's bytes, and both
    verifiers must reject the result."""
    if O.ref(k) is None or not hasattr(O.ref(k), "ref_prove_struct_at"):
        pytest.fail("oracle/_ref (with ref_prove_struct_at) missing: run __graft_entry__.build() where /root/reference exists")
    seed = O.seed_of(600 + k)
    seq = O.ref_struct_sequence(k, seed)
    S = O.struct_sizes(k)
    call = 3 * S["F"] + 2 * k * S["E"] + 1                   # randombytes() calls made before prove() in main.cpp's order
    oT = k * k * 256; oS = oT + k * 256; oE = oS + k * 256
    ctx = ctxs(k)
    cases = {"honest": lambda a: None,
             "t+1": lambda a: a.__setitem__(oT, a[oT] + 1),
             "s=7": lambda a: a.__setitem__(oS + 5, 7),
             "e=-5": lambda a: a.__setitem__(oE + 200, -5),
             "s=-1000,e=1200": lambda a: (a.__setitem__(oS + 255, -1000), a.__setitem__(oE, 1200))}
    for name, edit in cases.items():
        inst = seq["inst"].copy().view(np.int16)
        edit(inst)
        inst = inst.view(np.uint8)
        want = O.ref_prove_struct_at(k, seed, call, inst, seq["rand"], seq["eta"])
        ctx.rng_reset(seed)
        ctx.prepare_randomness(); ctx.prepare_range_proof(); ctx.kyber_keygen()     # consume the calls that precede prove()
        assert ctx.rng_calls() == call
        got = np.frombuffer(ctx.prove(inst, seq["rand"], seq["eta"]), np.uint8)
        assert (got == want).all(), (name, int((got != want).sum()))
        assert ctx.verify(got, inst) == O.ref_verify_struct(k, want, inst) == (name == "honest"), name


@pytest.mark.parametrize("k", [2, 3])
def test_struct_verify_rejects_non_canonical_t_like_the_reference(ctxs, k):
    """encode_to_gf3329 (gf3329.c:308-310) only adds q to negative coefficients, so an instance whose t carries t_i + q (same residue, not
    canonical) does not match the recomputed shares and the reference rejects it (mlwe_verifier.cpp:358-376); so must verify() here."""
    if O.ref(k) is None:
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    seq = O.ref_struct_sequence(k, O.seed_of(650 + k))
    oT = k * k * 256
    ctx = ctxs(k)
    assert ctx.verify(seq["pi"], seq["inst"]) is True
    t = seq["inst"].view(np.int16)[oT:oT + k * 256]
    neg, nonneg = int(np.nonzero(t < 0)[0][0]), int(np.nonzero(t >= 0)[0][0])
    verdicts = {}
    for name, pos, delta in (("nonneg+q", nonneg, 3329), ("neg+q", neg, 3329), ("neg-q", neg, -3329), ("last+q", k * 256 - 1, 3329)):
        inst = seq["inst"].copy().view(np.int16)
        inst[oT + pos] = int(inst[oT + pos]) + delta
        inst = inst.view(np.uint8)
        want = O.ref_verify_struct(k, seq["pi"], inst)
        assert ctx.verify(seq["pi"], inst) is want, name
        verdicts[name] = want
    # t_i >= 0 shifted up by q stays non-canonical and is rejected; t_i < 0 shifted up by q is encoded to the canonical residue and accepted
    assert verdicts["nonneg+q"] is False and verdicts["neg+q"] is True and verdicts["neg-q"] is False
    # A is used arithmetically only: a coefficient shifted by q is the same matrix entry for both verifiers
    inst = seq["inst"].copy().view(np.int16)
    inst[3] = int(inst[3]) + 3329 if int(inst[3]) < 0 else int(inst[3]) - 3329
    inst = inst.view(np.uint8)
    assert ctx.verify(seq["pi"], inst) is O.ref_verify_struct(k, seq["pi"], inst)
