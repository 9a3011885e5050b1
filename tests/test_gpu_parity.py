"""GPU suite (-m gpu): the CUDA path through the C ABI against the oracle and the reference-generated golden
vectors: bit-exact pk / sk / proof bytes and verify result (integer work, so no tolerance)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kosk_golden.json")))


def test_native_library_is_loaded(ctxs):
    ctx = ctxs(2)
    assert any("libkosk_b200.so" in l for l in open("/proc/self/maps"))
    n0 = ctx.kernel_launches()
    ctx.ntt_rows(np.zeros((1, 256), np.uint16))
    assert ctx.kernel_launches() == n0 + 1


# ---------------- components (BASELINE config 5) ----------------
@pytest.mark.parametrize("rows", [1, 8, 214, 300])
def test_share_eval_sweep(ctxs, rows):
    rng = np.random.default_rng(rows)
    y = rng.integers(0, 3329, size=(rows, 407), dtype=np.uint16)
    y[0] = 3328                                    # worst-case magnitude for the lazy int32 accumulation
    if rows > 1:
        y[1] = 0
    got = ctxs(2).share_eval(y)
    for i in sorted(set([0, 1 % rows, rows - 1, rows // 2])):
        assert (got[i] == O.oracle_share(y[i])).all()
    assert (got[:, :151] == y[:, 256:]).all()


def test_share_eval_work_tickets_across_launch_sizes(ctxs):
    """k_share_ntt2 draws its rows from a per-stream device counter that is never reset (share_ntt.cuh, SnTicket): launches of very different
    sizes back to back on one context, below, at and above the number of resident warps (148 SMs x 4 CTAs x 7 warps = 4144), must each
    produce every row."""
    ctx = ctxs(2)
    rng = np.random.default_rng(77)
    for rows in (5000, 3, 4144, 4145, 1, 9000, 4143):
        y = rng.integers(0, 3329, size=(rows, 407), dtype=np.uint16)
        got = ctx.share_eval(y)
        assert (got[:, :151] == y[:, 256:]).all()
        for i in sorted(set([0, rows - 1, rows // 2, rows // 3])):
            assert (got[i] == O.oracle_share(y[i])).all(), (rows, i)
        # every row was written: a row skipped by the scheduler would keep the previous launch's shares
        z = ctx.share_eval(np.zeros((rows, 407), np.uint16))
        assert not z.any()


def test_share_eval_linearity_full_size(ctxs):
    """Size-independent property at a batch the oracle could not finish quickly: S is linear over GF(3329)."""
    rng = np.random.default_rng(5)
    n = 214 * 64
    a = rng.integers(0, 3329, size=(n, 407), dtype=np.uint16)
    b = rng.integers(0, 3329, size=(n, 407), dtype=np.uint16)
    ctx = ctxs(2)
    sa, sb = ctx.share_eval(a), ctx.share_eval(b)
    sab = ctx.share_eval(((a.astype(np.uint32) + b) % 3329).astype(np.uint16))
    assert (sab == (sa.astype(np.uint32) + sb) % 3329).all()
    assert (sa[7] == O.oracle_share(a[7])).all() and (sa[n - 1] == O.oracle_share(a[n - 1])).all()


@pytest.mark.parametrize("length", [0, 2, 134, 136, 138, 272, 308, 320, 332, 452, 472, 524])
def test_sha3_rows(ctxs, length):
    rng = np.random.default_rng(length)
    rows = rng.integers(0, 256, size=(33, max(length, 1)), dtype=np.uint8)[:, :length]
    got = ctxs(2).sha3_256_rows(np.ascontiguousarray(rows))
    for h, r in zip(got, rows):
        assert bytes(h) == hashlib.sha3_256(bytes(r)).digest()


def test_ntt_rows(ctxs):
    rng = np.random.default_rng(9)
    a = rng.integers(0, 3329, size=(17, 256), dtype=np.uint16)
    a[0] = 3328
    got = ctxs(2).ntt_rows(a)
    assert (got == np.stack([O.oracle_ntt(r) for r in a])).all()


# ---------------- prove ----------------
@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"k{c['k']}s{c['seed_index']}")
def test_prove_matches_reference_golden(ctxs, case):
    ctx = ctxs(case["k"])
    pk, sk, pi = ctx.verifiable_keygen(bytes.fromhex(case["seed"]))
    assert hashlib.sha256(pk).hexdigest() == case["pk_sha256"]
    assert hashlib.sha256(sk).hexdigest() == case["sk_sha256"]
    assert hashlib.sha256(pi).hexdigest() == case["proof_sha256"]
    assert ctx.kosk_verify(pi, pk) is True


@pytest.mark.parametrize("k", [2, 3, 4])
def test_prove_batch_matches_oracle(ctxs, k):
    """A ragged batch (not a multiple of the chunk, spans two chunks) against the oracle, byte for byte."""
    ctx = ctxs(k, 8, 1)      # one lane: chunks of 8 run back to back, so the scratch holds the last chunk
    n = 11
    seeds = np.stack([np.frombuffer(O.seed_of(1000 + i), np.uint8) for i in range(n)])
    pk, sk, pi = ctx.prove_batch(seeds)
    for i in range(n):
        opk, osk, opi = O.oracle_prove(k, seeds[i])
        assert (pk[i] == opk).all() and (sk[i] == osk).all() and (pi[i] == opi).all(), i
    # intermediates of the last chunk (proofs 8..10) against the oracle's trace of proof 10
    tr = O.oracle_trace()
    F, NA = 70 + 2 * k + 1, 70 + 2 * k
    pw = ctx.debug_fetch("alpha_pow", 8 * NA * F * 2, np.uint16).reshape(8, NA, F)
    assert (pw[2, :, 1] == np.array(tr.alpha[:NA])).all() and (pw[2, :, 0] == 1).all()
    I = ctx.debug_fetch("I", 8 * 150 * 2, np.uint16).reshape(8, 150)
    assert (I[2] == np.array(tr.I[:])).all() and len(set(I[2])) == 150


def test_prove_empty_batch(ctxs):
    pk, sk, pi = ctxs(2).prove_batch(np.zeros((0, 32), np.uint8))
    assert pk.shape == (0, 800) and pi.shape == (0, 664340)


def test_prove_is_placement_independent(ctxs):
    """Same seeds through different chunkings / batch positions give the same bytes."""
    seeds = seeds_for_range(5, 0, 6)
    a = ctxs(2, 8, 1).prove_batch(seeds)
    b = ctxs(2, 2, 3).prove_batch(seeds[::-1].copy())     # three lanes, sub-batches of two
    for x, y in zip(a, b):
        assert (x == y[::-1]).all()


# ---------------- verify ----------------
@pytest.mark.parametrize("k", [2, 3, 4])
def test_verify_tamper_matrix_matches_reference(ctxs, k):
    """Accept/reject parity with the reference (golden), incl. the 7 lax accepts."""
    tam = GOLDEN["tamper"][str(k)]
    ctx = ctxs(k)
    L = O.layout(k)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(0))
    names, cases = [], []
    offs = [(n, getattr(L, n)) for n in O.FIELDS] + [("end", L.proof_bytes)]
    for (n, o), (_, e) in zip(offs[:-1], offs[1:]):
        step = 1 if n in ("o_Tcomm", "o_comm") else 2
        for tag, off in (("first", o), ("last", e - step)):
            t = opi.copy(); t[off] ^= 1
            names.append(f"{n[2:]}:{tag}"); cases.append(t)
    got = ctx.verify_batch(np.stack(cases), np.repeat(opk[None], len(cases), 0))
    for nme, g in zip(names, got):
        assert bool(g) == tam[nme], nme
    assert int(got.sum()) == 7
    for tag, off in (("pk:t", 5), ("pk:seed", -1)):
        t = opk.copy(); t[off] ^= 1
        assert ctx.kosk_verify(bytes(opi), bytes(t)) == tam[tag]


@pytest.mark.parametrize("k", [2, 3])
def test_verify_random_tampering_matches_oracle(ctxs, k):
    """Random single-byte corruption anywhere in the proof (incl. non-canonical field elements >= q)."""
    rng = np.random.default_rng(40 + k)
    ctx = ctxs(k)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(2))
    cases = []
    for _ in range(24):
        t = opi.copy(); off = int(rng.integers(0, t.size)); t[off] ^= int(rng.integers(1, 256)); cases.append(t)
    got = ctx.verify_batch(np.stack(cases), np.repeat(opk[None], len(cases), 0))
    exp = [O.oracle_verify(k, t, opk) for t in cases]
    assert [bool(g) for g in got] == exp


def test_verify_rejects_malformed_open_set(ctxs):
    k = 2
    L = O.layout(k)
    ctx = ctxs(k)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(1))
    t = opi.copy(); t[L.o_I:L.o_I + 2] = np.frombuffer((1454).to_bytes(2, "little"), np.uint8)
    assert ctx.kosk_verify(bytes(t), bytes(opk)) is False
    t = opi.copy(); t[L.o_I + 2:L.o_I + 4] = t[L.o_I:L.o_I + 2]
    assert ctx.kosk_verify(bytes(t), bytes(opk)) is False
    t = opi.copy(); t[L.o_I:L.o_I + 300] = 0xFF
    assert ctx.kosk_verify(bytes(t), bytes(opk)) is False
    assert ctx.kosk_verify(bytes(opi), bytes(opk)) is True            # context still healthy afterwards


def test_verify_noncanonical_values_match_oracle(ctxs):
    """Shares shifted by q in fields the reference reduces silently (t, eta, u) and in fields it compares raw."""
    k = 2
    L = O.layout(k)
    ctx = ctxs(k)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(3))
    cases = []
    for off in (L.o_t, L.o_seta, L.o_us, L.o_sr, L.o_beta, L.o_f, L.o_s, L.o_zs, L.o_ssub, L.o_NTTAs):
        t = opi.copy(); v = int(t[off:off + 2].view(np.uint16)[0]) + 3329
        t[off:off + 2] = np.frombuffer(v.to_bytes(2, "little"), np.uint8); cases.append(t)
    got = ctx.verify_batch(np.stack(cases), np.repeat(opk[None], len(cases), 0))
    exp = [O.oracle_verify(k, t, opk) for t in cases]
    assert [bool(g) for g in got] == exp
    assert any(exp) and not all(exp)


# ---------------- BASELINE config 2 at full size: properties + sampled bit-exactness ----------------
def test_full_batch_1024_k2(ctxs):
    ctx = ctxs(2, 512, 2)
    n = 1024
    seeds = seeds_for_range(20240, 0, n)
    pk, sk, pi = ctx.prove_batch(seeds)
    ok = ctx.verify_batch(pi, pk)
    assert ok.all()                                              # prove -> verify round trip for every proof
    L = O.layout(2)
    I = pi[:, L.o_I:L.o_I + 300].copy().view(np.uint16)
    assert (I < 1454).all() and all(len(set(r)) == 150 for r in I)
    assert len({hashlib.sha256(p).digest() for p in pk}) == n    # all keys distinct
    for i in (0, 1, 511, 1023):                                  # sampled bit-exactness against the oracle
        opk, osk, opi = O.oracle_prove(2, seeds[i])
        assert (pk[i] == opk).all() and (sk[i] == osk).all() and (pi[i] == opi).all()
    # erase-and-check: swapping two proofs' keys must be rejected
    swapped = pk.copy(); swapped[[0, 1]] = swapped[[1, 0]]
    ok2 = ctx.verify_batch(pi[:4], swapped[:4])
    assert list(ok2) == [False, False, True, True]


# ---------------- drop-in boundary: the reference-shaped C++ API over the C ABI ----------------
def _fnv(b):
    h = 14695981039346656037
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.mark.parametrize("k", [2, 4])
def test_dropin_cpp_program(built_lib, tmp_path, k):
    """examples/main_dropin.cpp uses the reference's names (kyber_keypair, kyber_verifiable_keygen,
    kyber_kosk_verify, MPCITH_PROOF_SIZE) and must reproduce the oracle's bytes for an injected seed."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / f"main_dropin_k{k}")
    libdir = os.path.join(root, "mpcith_kyber_kosk_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", f"-DKYBER_K={k}", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "main_dropin.cpp"), "-L" + libdir, "-lkosk_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    seed = O.seed_of(4242 + k)
    out = subprocess.run([exe], env=dict(os.environ, KOSK_SEED_HEX=seed.hex()), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "kosk verify success" in out.stdout and "[tamper] rejected" in out.stdout
    opk, osk, opi = O.oracle_prove(k, seed)
    assert f"pk={_fnv(opk)} sk={_fnv(osk)} proof={_fnv(opi)}" in out.stdout


def test_strict_verifier_is_a_superset_of_checks(ctxs):
    """the switchable hardened decoder rejects what the reference laxly accepts (non-canonical elements,
    unread t / eta rest shares) and still accepts every honest proof; default (off) stays on the reference's accept set."""
    k = 2
    L = O.layout(k)
    ctx = ctxs(k, 64, 1)
    opk, osk, opi = O.oracle_prove(k, O.seed_of(5))
    last = lambda name, nxt: getattr(L, nxt) - 2
    cases = {}
    for name, nxt in (("o_t", "o_NTTs"), ("o_seta", "o_eeta"), ("o_eeta", "o_ssub"), ("o_beta", "o_gamma")):
        t = opi.copy(); t[last(name, nxt)] ^= 1; cases[name + ":last"] = t
    t = opi.copy(); v = int(t[L.o_t:L.o_t + 2].view(np.uint16)[0]) + 3329
    t[L.o_t:L.o_t + 2] = np.frombuffer(v.to_bytes(2, "little"), np.uint8); cases["t:+q"] = t
    cases["honest"] = opi.copy()
    names = list(cases)
    batch = np.stack([cases[n] for n in names]); pks = np.repeat(opk[None], len(names), 0)
    lax = dict(zip(names, ctx.verify_batch(batch, pks)))
    assert all(lax.values())                                    # reference behaviour: all of these are accepted
    assert all(O.oracle_verify(k, cases[n], opk) for n in names)
    ctx.set_strict(True)
    try:
        strict = dict(zip(names, ctx.verify_batch(batch, pks)))
    finally:
        ctx.set_strict(False)
    assert strict["honest"] and strict["o_beta:last"]           # beta's unread tail is outside the hardened checks (documented)
    assert not strict["o_t:last"] and not strict["o_seta:last"] and not strict["o_eeta:last"] and not strict["t:+q"]


@pytest.mark.parametrize("k", [2, 3])
def test_offline_online_split_matches_one_shot(ctxs, k):
    """preprocessing pool + online-only prove gives the same bytes as kyber_verifiable_keygen."""
    ctx = ctxs(k, 16, 1)
    seeds = seeds_for_range(31337, 0, 5)
    pool = ctx.pool_create(seeds)
    pk, sk, pi = pool.prove()
    pk2, sk2, pi2 = pool.prove()                       # the online phase is a pure function of (pool, seeds)
    pool.close()
    a = ctx.prove_batch(seeds)
    for x, y, z in zip(a, (pk, sk, pi), (pk2, sk2, pi2)):
        assert (x == y).all() and (x == z).all()
    opk, osk, opi = O.oracle_prove(k, seeds[3])
    assert (pi[3] == opi).all() and (pk[3] == opk).all() and (sk[3] == osk).all()
    assert ctx.verify_batch(pi, pk).all()


@pytest.mark.parametrize("k", [2, 4])
def test_pool_export_import_round_trip(ctxs, k):
    """The working (de)serialiser of the preprocessing (reference: mlwe_prover.cpp:61-79, unused and lossy): a pool exported from one
    context and imported into another proves the same bytes as the one-shot call; truncated / foreign images are refused."""
    from mpcith_kyber_kosk_b200 import KoskContext, KoskError
    seeds = seeds_for_range(4242, 0, 6)
    ctx = ctxs(k, 16, 1)
    pool = ctx.pool_create(seeds)
    img = pool.export()
    pool.close()
    other = KoskContext(k, 0, 8, 2)
    p2 = other.pool_import(img)
    pk, sk, pi = p2.prove()
    want = ctx.prove_batch(seeds)
    for x, y in zip(want, (pk, sk, pi)):
        assert (x == y).all()
    opk, osk, opi = O.oracle_prove(k, seeds[5])
    assert (pi[5] == opi).all() and (pk[5] == opk).all()
    with pytest.raises(KoskError):
        other.pool_import(img[:len(img) // 2])
    bad = img.copy(); bad[0] ^= 1
    with pytest.raises(KoskError):
        other.pool_import(bad)
    with pytest.raises(KoskError):
        ctxs(3 if k == 2 else 2, 16, 1).pool_import(img)      # another KYBER_K
    p2.close()
    other.close()


@pytest.mark.parametrize("k", [2, 3, 4])
def test_experimental_tensor_path_is_bit_identical(ctxs, k):
    """Opt-in KOSK_F_TENSOR: share evaluation on int8 tensor cores (limb-split residues); same bytes as the INT32 pipe."""
    ctx = ctxs(k, 16, 1, True)
    rng = np.random.default_rng(70 + k)
    y = rng.integers(0, 3329, size=(300, 407), dtype=np.uint16)
    y[0] = 3328; y[1] = 1664; y[2] = 1665; y[3] = 0
    got = ctx.share_eval(y)
    for i in (0, 1, 2, 3, 150, 299):
        assert (got[i] == O.oracle_share(y[i])).all()
    seeds = seeds_for_range(900 + k, 0, 3)
    pk, sk, pi = ctx.prove_batch(seeds)
    for i in range(3):
        opk, osk, opi = O.oracle_prove(k, seeds[i])
        assert (pk[i] == opk).all() and (sk[i] == osk).all() and (pi[i] == opi).all()
    assert ctx.verify_batch(pi, pk).all()


def test_async_pipelined_host_api(ctxs):
    """kosk_b200_prove_batch_async: three batches in flight over two lanes (kernels back to back, copies overlapped) give
    the same bytes as the synchronous call, also when a batch spans several sub-batches."""
    ctx = ctxs(2, 4, 2)
    lib, h = ctx.lib, ctx._h
    outs = []
    for j, n in enumerate((4, 7, 2)):
        seeds = seeds_for_range(4000 + 100 * j, 0, n)
        bufs = (np.empty((n, ctx.pk_bytes), np.uint8), np.empty((n, ctx.sk_bytes), np.uint8), np.empty((n, ctx.proof_bytes), np.uint8))
        assert lib.kosk_b200_prove_batch_async(h, n, seeds.ctypes.data, bufs[0].ctypes.data, bufs[1].ctypes.data, bufs[2].ctypes.data) == 0
        outs.append((seeds, bufs))
    ctx.sync()
    for seeds, bufs in outs:
        ref = ctx.prove_batch(seeds)
        for x, y in zip(ref, bufs):
            assert (x == y).all()
    opk, osk, opi = O.oracle_prove(2, outs[1][0][6])
    assert (outs[1][1][2][6] == opi).all()


def test_c_abi_from_plain_c(built_lib, tmp_path):
    """examples/c_abi_demo.c: the C ABI called from C99 (no C++/Python in between) proves, verifies and rejects a tampered
    proof; the digest of its proofs equals the oracle's for the same fixed seeds."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(root, "mpcith_kyber_kosk_b200")
    subprocess.run(["gcc", "-std=c99", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_demo.c"),
                    "-L" + libdir, "-lkosk_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe, "3", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "accepted=3 tampered_rejected=1" in out.stdout
    h = 14695981039346656037
    for i in range(3):
        seed = bytes([i + 1, 0xC0]) + bytes(30)
        _, _, opi = O.oracle_prove(3, seed)
        for x in bytes(opi):
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert "fnv=%016x" % h in out.stdout


@pytest.mark.parametrize("k", [3, 4])
def test_larger_batches_k3_k4(ctxs, k):
    """Kyber768 / Kyber1024: a 96-proof batch over two lanes (non-fused kernels at chunk 256 would need >128; here the fused
    latency-mode path) round-trips through the verifier, with sampled bit-exactness and a strict-mode pass."""
    ctx = ctxs(k, 48, 2)
    n = 96
    seeds = seeds_for_range(5000 + k, 0, n)
    pk, sk, pi = ctx.prove_batch(seeds)
    assert ctx.verify_batch(pi, pk).all()
    ctx.set_strict(True)
    try:
        assert ctx.verify_batch(pi, pk).all()
    finally:
        ctx.set_strict(False)
    for i in (0, 47, 48, 95):
        opk, osk, opi = O.oracle_prove(k, seeds[i])
        assert (pk[i] == opk).all() and (sk[i] == osk).all() and (pi[i] == opi).all()
    bad = pi.copy(); bad[:, 0] ^= 1
    assert not ctx.verify_batch(bad, pk).any()


def test_unfused_path_large_chunk_k3(ctxs):
    """Sub-batches above the latency-mode threshold take the separate hash / FS kernels: 160 proofs in one chunk, Kyber768."""
    ctx = ctxs(3, 160, 1)
    seeds = seeds_for_range(6000, 0, 160)
    pk, sk, pi = ctx.prove_batch(seeds)
    assert ctx.verify_batch(pi, pk).all()
    for i in (0, 159):
        opk, osk, opi = O.oracle_prove(3, seeds[i])
        assert (pi[i] == opi).all() and (pk[i] == opk).all()


# ---------------- struct-level API (reference main.cpp:16-59) ----------------
STRUCT_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kosk_struct_golden.json")))


def _struct_sequence(ctx, seed):
    ctx.rng_reset(seed)
    rand = ctx.prepare_randomness()
    eta = ctx.prepare_range_proof()
    pk, sk, inst = ctx.kyber_keygen()
    pi = ctx.prove(inst, rand, eta)
    return {"rand": rand, "eta": eta, "inst": inst, "pk": pk, "sk": sk, "pi": pi}


@pytest.mark.parametrize("case", STRUCT_GOLDEN["cases"], ids=lambda c: f"k{c['k']}s{c['seed_index']}")
def test_struct_sequence_matches_reference_golden(ctxs, case):
    """prepare_randomness, prepare_range_proof, kyber_keygen, prove in main.cpp's order: every struct image equals the reference's."""
    k = case["k"]
    ctx = ctxs(k)
    got = _struct_sequence(ctx, bytes.fromhex(case["seed"]))
    for n in ("inst", "pk", "sk", "rand", "eta", "pi"):
        assert hashlib.sha256(bytes(got[n])).hexdigest() == case[n + "_sha256"], n
    S = O.struct_sizes(k)
    assert ctx.rng_calls() == 3 * S["F"] + 2 * k * S["E"] + 1 + 3 * k + 4 * (3 if k == 2 else 2) * k
    assert ctx.verify(got["pi"], got["inst"]) is True
    bad = bytearray(got["pi"]); bad[0] ^= 1           # an opened party's f share (hashed into its commitment)
    assert ctx.verify(bytes(bad), got["inst"]) is False
    inst = got["inst"].copy(); inst[(k * k) * 512 + 10] ^= 1          # one coefficient of t
    assert ctx.verify(got["pi"], inst) is False


@pytest.mark.parametrize("k", [2, 3, 4])
def test_struct_sequence_against_live_reference(ctxs, k):
    if O.ref(k) is None:
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists (oracle/_ref travels with the repo)")
    seed = O.seed_of(77 + k)
    ctx = ctxs(k)
    got = _struct_sequence(ctx, seed)
    want = O.ref_struct_sequence(k, seed)
    assert want["ok"]
    for n in ("inst", "pk", "sk", "rand", "eta", "pi"):
        assert bytes(got[n]) == bytes(want[n]), n
    assert O.ref_verify_struct(k, got["pi"], got["inst"]) is True       # the reference's verify() accepts the GPU proof
    assert ctx.verify(want["pi"], want["inst"]) is True                 # and the GPU verify() accepts the reference's


@pytest.mark.parametrize("k", [2, 3, 4])
def test_struct_calls_in_keygen_order_equal_verifiable_keygen(ctxs, k):
    """kosk.cpp:72-86 is keygen, prepare_randomness, prepare_range_proof, prove: the same four calls must give the same bytes."""
    ctx = ctxs(k)
    seed = O.seed_of(300 + k)
    pk0, sk0, pi0 = ctx.verifiable_keygen(seed)
    ctx.rng_reset(seed)
    pk, sk, inst = ctx.kyber_keygen()
    rand = ctx.prepare_randomness()
    eta = ctx.prepare_range_proof()
    pi = ctx.prove(inst, rand, eta)
    assert (pk, sk, pi) == (pk0, sk0, pi0)
    assert ctx.kosk_verify(pi, pk) is True and ctx.verify(pi, inst) is True
    # preprocessing is key-independent (main.cpp:18-31): the same rand/eta prove a second, different key
    pk2, sk2, inst2 = ctx.kyber_keygen()
    assert pk2 != pk
    pi2 = ctx.prove(inst2, rand, eta)
    assert ctx.kosk_verify(pi2, pk2) is True and ctx.kosk_verify(pi2, pk) is False


# ---------------- Kyber KEM on the generated keys (reference kyber/kem.c:76-169, main.cpp:98-113) ----------------
KEM_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kosk_kem_golden.json")))


@pytest.mark.parametrize("case", KEM_GOLDEN["cases"], ids=lambda c: f"k{c['k']}key{c['key_index']}c{c['coins_index']}")
def test_kem_matches_reference_golden(ctxs, case):
    k = case["k"]
    ctx = ctxs(k)
    pk, sk, _ = ctx.verifiable_keygen(O.seed_of(case["key_index"]))
    assert hashlib.sha256(pk).hexdigest() == case["pk_sha256"]
    coins = O.seed_of(case["coins_index"], b"kem-coins")
    ct, ss = ctx.kem_enc_derand_batch(np.frombuffer(pk, np.uint8), np.frombuffer(coins, np.uint8))
    assert ct.shape == (1, O.CT_BYTES[k]) == (1, ctx.ct_bytes)
    assert hashlib.sha256(bytes(ct[0])).hexdigest() == case["ct_sha256"]
    assert bytes(ss[0]).hex() == case["ss"]
    assert ctx.crypto_kem_dec(bytes(ct[0]), sk).hex() == case["ss"]
    bad = ct[0].copy(); bad[7] ^= 0x10                                    # implicit rejection: SHAKE256(z || ct)
    assert ctx.crypto_kem_dec(bytes(bad), sk).hex() == case["ss_reject_byte7"]
    bad = ct[0].copy(); bad[-1] ^= 0x80
    assert ctx.crypto_kem_dec(bytes(bad), sk).hex() == case["ss_reject_last"]


@pytest.mark.parametrize("k", [2, 3, 4])
def test_kem_batch_roundtrip_and_live_reference(ctxs, k):
    """encaps -> decaps round trip on a batch of GPU-made keys; sampled rows against the live reference when it is present."""
    ctx = ctxs(k)
    n = 96
    pk, sk, _ = ctx.prove_batch(seeds_for_range(500 + k, 0, n))
    rng = np.random.default_rng(40 + k)
    coins = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ct, ss = ctx.kem_enc_derand_batch(pk, coins)
    assert (ctx.kem_dec_batch(ct, sk) == ss).all()
    assert len({bytes(r) for r in ss}) == n
    # wrong key / tampered ciphertext: decapsulation returns the rejection key, never the encapsulated secret
    rolled = np.roll(sk, 1, axis=0)
    assert not (ctx.kem_dec_batch(ct, rolled) == ss).all(axis=1).any()
    t = ct.copy(); t[:, 3] ^= 1
    rej = ctx.kem_dec_batch(t, sk)
    assert not (rej == ss).all(axis=1).any()
    z = sk[:, -32:]
    for i in (0, n - 1):
        assert bytes(rej[i]) == hashlib.shake_256(bytes(z[i]) + bytes(t[i])).digest(32)
    if O.ref(k) is not None:
        for i in (0, 17, n - 1):
            rct, rss = O.ref_kem_enc_derand(k, pk[i], coins[i])
            assert bytes(rct) == bytes(ct[i]) and bytes(rss) == bytes(ss[i])
            assert bytes(O.ref_kem_dec(k, t[i], sk[i])) == bytes(rej[i])
            assert bytes(O.ref_kem_dec(k, ct[i], rolled[i])) == bytes(ctx.kem_dec_batch(ct[i:i + 1], rolled[i:i + 1])[0])


@pytest.mark.parametrize("k", [2, 3, 4])
def test_main_cpp_sequence_with_kem(ctxs, k):
    """The whole of main.cpp under one DRBG: struct-level prove/verify, kyber_verifiable_keygen + kosk_verify, then
    crypto_kem_enc (coins = the next randombytes call) and crypto_kem_dec."""
    ctx = ctxs(k)
    seed = O.seed_of(900 + k)
    got = _struct_sequence(ctx, seed)
    assert ctx.verify(got["pi"], got["inst"]) is True
    calls = ctx.rng_calls()
    ct, ss = ctx.crypto_kem_enc(got["pk"])
    assert ctx.rng_calls() == calls + 1
    assert ctx.crypto_kem_dec(ct, got["sk"]) == ss
    if O.ref(k) is not None:
        rct, rss = O.ref_kem_enc_at(k, seed, calls, got["pk"])
        assert bytes(rct) == ct and bytes(rss) == ss


@pytest.mark.parametrize("k", [2, 3, 4])
def test_reference_main_cpp_runs_against_the_library(k):
    """The reference's UNMODIFIED main.cpp (struct-level prove/verify, kyber_verifiable_keygen + kyber_kosk_verify, KEM
    encaps/decaps) compiled against include/dropin/ and linked with libkosk_b200.so (`make -C oracle dropin-main`, done by
    __graft_entry__.build() where /root/reference exists; the binary travels to the GPU box)."""
    import subprocess
    exe = os.path.join(O.ORACLE_DIR, "_ref", f"main_dropin_k{k}")
    if not os.path.exists(exe):
        pytest.fail("oracle/_ref/main_dropin_k* missing: run __graft_entry__.build() where /root/reference exists (the binary travels with the repo)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    for line in ("[result] mlwe verify success", "[result] kyber kosk verify success", "[result] decapsulated ss is the same as the encapsulated",
                 f"[proof size] {ctypes_proof_kib(k)} kilobytes"):
        assert line in out.stdout, out.stdout


@pytest.mark.parametrize("k", [2, 3, 4])
def test_reference_main_object_links_against_the_shim_only(k):
    """BINARY drop-in: main.cpp compiled against the reference's OWN headers (oracle/_ref/main_ref_k*.o, `make -C oracle binlink-main`)
    linked with nothing but libkosk_kyber{512,768,1024}.so, which exports the reference's C++-mangled symbols (kosk.hpp:17-23,
    mlwe_prover.hpp:77-99, mlwe_verifier.hpp:14-15) and the pqcrystals_* KEM names (kyber/kem.h:20-33)."""
    import subprocess
    exe = os.path.join(O.ORACLE_DIR, "_ref", f"main_binlink_k{k}")
    if not os.path.exists(exe):
        pytest.fail("oracle/_ref/main_binlink_k* missing: run __graft_entry__.build() where /root/reference exists (the binary travels with the repo)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    for line in ("[result] mlwe verify success", "[result] kyber kosk verify success", "[result] decapsulated ss is the same as the encapsulated",
                 f"[proof size] {ctypes_proof_kib(k)} kilobytes"):
        assert line in out.stdout, out.stdout
    maps = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert f"libkosk_kyber{256 * k}.so" in maps and "libkosk_ref" not in maps


def ctypes_proof_kib(k):
    from mpcith_kyber_kosk_b200 import proof_bytes
    return proof_bytes(k) // 1024


@pytest.mark.parametrize("k", [2, 3, 4])
def test_kem_arbitrary_bytes_match_live_reference(ctxs, k):
    """Inputs a well-formed key pair never produces: 12-bit coefficients >= q in pk / sk (the reference reduces them implicitly in
    its Montgomery arithmetic) and random ciphertext bytes (every bit pattern decompresses): same bytes out as the reference."""
    if O.ref(k) is None:
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists (oracle/_ref travels with the repo)")
    ctx = ctxs(k)
    rng = np.random.default_rng(70 + k)
    n = 6
    pk = rng.integers(0, 256, size=(n, ctx.pk_bytes), dtype=np.uint8)
    sk = rng.integers(0, 256, size=(n, ctx.sk_bytes), dtype=np.uint8)
    coins = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ct, ss = ctx.kem_enc_derand_batch(pk, coins)
    rct = rng.integers(0, 256, size=(n, ctx.ct_bytes), dtype=np.uint8)
    dss = ctx.kem_dec_batch(rct, sk)
    for i in range(n):
        wct, wss = O.ref_kem_enc_derand(k, pk[i], coins[i])
        assert bytes(wct) == bytes(ct[i]) and bytes(wss) == bytes(ss[i])
        assert bytes(O.ref_kem_dec(k, rct[i], sk[i])) == bytes(dss[i])


@pytest.mark.parametrize("k", [2, 3, 4])
def test_dense_table_share_eval_matches_ntt_convolution(k):
    """The share evaluation runs as a blocked NTT convolution by default (share_ntt.cuh: k_share_ntt2, unequal blocks + IDP.2A + radix-2 DFT
    networks; KOSK_B200_SHARE_NTT=1 selects the generic equal-block kernel) and as the dense 1303 x 407 table GEMM (gf_gemm.cuh) with
    KOSK_B200_SHARE_NTT=0: same shares, same proofs, same verify results, and all equal the oracle."""
    from mpcith_kyber_kosk_b200 import KoskContext
    rng = np.random.default_rng(90 + k)
    y = rng.integers(0, 3329, size=(37, 407), dtype=np.uint16)
    y[0] = 3328; y[1] = 0; y[2, :256] = 5
    y[3, 0::2] = 3328; y[3, 1::2] = 1; y[4, 0::2] = 1664; y[4, 1::2] = 1665       # sign-alternating extremes for the lazy products of the DFT networks
    seeds = seeds_for_range(4000 + k, 0, 5)
    out = {}
    for mode in ("2", "1", "0"):
        old = os.environ.get("KOSK_B200_SHARE_NTT")
        os.environ["KOSK_B200_SHARE_NTT"] = mode
        try:
            ctx = KoskContext(k, 0, 8)
        finally:
            if old is None:
                os.environ.pop("KOSK_B200_SHARE_NTT", None)
            else:
                os.environ["KOSK_B200_SHARE_NTT"] = old
        sh = ctx.share_eval(y)
        pk, sk, pi = ctx.prove_batch(seeds)
        ok = ctx.verify_batch(pi, pk)
        bad = pi.copy(); bad[:, 0] ^= 1
        rej = ctx.verify_batch(bad, pk)
        out[mode] = (sh, pk.copy(), sk.copy(), pi.copy(), ok, rej)
        ctx.close()
    for other in ("1", "0"):
        for a, b in zip(out["2"], out[other]):
            assert (a == b).all()
    assert out["2"][4].all() and not out["2"][5].any()
    for i in (0, 1, 2, 3, 4, 36):
        assert (out["2"][0][i] == O.oracle_share(y[i])).all()
    opk, osk, opi = O.oracle_prove(k, bytes(seeds[0]))
    assert (out["2"][3][0] == opi).all() and (out["2"][1][0] == opk).all()


@pytest.mark.parametrize("mode", ["2", "1", "0"])
def test_share_eval_noncanonical_inputs(mode):
    """u16 inputs >= q act as their residue, as in the reference's gf3329_mul ((uint32_t)a * b % 3329, utils/gf3329.c:282-284), on every
    share-evaluation variant (KOSK_B200_SHARE_NTT = 2: k_share_ntt2, 1: generic convolution kernel, 0: dense table GEMM)."""
    from mpcith_kyber_kosk_b200 import KoskContext
    rng = np.random.default_rng(11)
    y = rng.integers(0, 65536, size=(9, 407), dtype=np.uint16)
    y[0] = 65535
    old = os.environ.get("KOSK_B200_SHARE_NTT")
    os.environ["KOSK_B200_SHARE_NTT"] = mode
    try:
        ctx = KoskContext(2, 0, 8)
    finally:
        if old is None:
            os.environ.pop("KOSK_B200_SHARE_NTT", None)
        else:
            os.environ["KOSK_B200_SHARE_NTT"] = old
    got = ctx.share_eval(y)
    ctx.close()
    for i in range(9):
        want = O.oracle_share(y[i])
        assert (got[i, 151:] == want[151:]).all()
        assert (got[i, :151] == y[i, 256:]).all()          # the tail is copied verbatim (ss.cpp:77-80)


@pytest.mark.parametrize("k", [2, 3, 4])
def test_kem_keypair_matches_reference(ctxs, k):
    """crypto_kem_keypair_derand / crypto_kem_keypair (kyber/kem.c:23-58): same pk as kyber_keygen for the same d, z from the coins;
    keys work with encaps / decaps; bytes equal the live reference's when it is present."""
    ctx = ctxs(k)
    rng = np.random.default_rng(120 + k)
    coins = rng.integers(0, 256, size=(70, 64), dtype=np.uint8)          # > chunk (64): exercises the chunk loop
    pk, sk = ctx.kem_keypair_derand_batch(coins)
    assert (sk[:, -32:] == coins[:, 32:]).all()                             # z
    assert (sk[:, 384 * k:384 * k + ctx.pk_bytes] == pk).all()
    for i in (0, 69):
        assert bytes(sk[i, -64:-32]) == hashlib.sha3_256(bytes(pk[i])).digest()
    ct, ss = ctx.kem_enc_derand_batch(pk, coins[:, :32])
    assert (ctx.kem_dec_batch(ct, sk) == ss).all()
    seed = O.seed_of(640 + k)
    ctx.rng_reset(seed)
    ctx.prepare_range_proof()                                               # move the DRBG off call 0
    calls = ctx.rng_calls()
    pk1, sk1 = ctx.crypto_kem_keypair()
    assert ctx.rng_calls() == calls + 1
    if O.ref(k) is not None:
        for i in (0, 33, 69):
            rpk, rsk = O.ref_kem_keypair_derand(k, coins[i])
            assert bytes(rpk) == bytes(pk[i]) and bytes(rsk) == bytes(sk[i])
        rpk, rsk = O.ref_kem_keypair_at(k, seed, calls)
        assert bytes(rpk) == pk1 and bytes(rsk) == sk1
