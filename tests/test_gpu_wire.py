"""GPU suite (-m gpu) for the compact wire format: the device pack / unpack kernels and the host-buffer batch
calls in wire mode against (i) the host codec, which the CPU suite pins to an independent numpy restatement, (ii) the oracle's
proof bytes, (iii) the raw-link mode (KOSK_B200_WIRE=0 behaviour of round 1).  Bit-exact: byte work."""
import numpy as np
import pytest
import torch

import mpcith_kyber_kosk_b200 as pkg
import oracle_lib as O
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [2, 3, 4])
def test_device_pack_unpack_match_host_codec(ctxs, k):
    ctx = ctxs(k, 16, 2)
    n = 5
    rng = np.random.default_rng(k)
    pis = rng.integers(0, 4096, (n, ctx.proof_bytes // 2)).astype(np.uint16).view(np.uint8).reshape(n, -1).copy()
    pis[0] = np.frombuffer(bytes(O.oracle_prove(k, O.seed_of(41))[2]), np.uint8)
    d_pi = torch.from_numpy(pis).cuda()
    d_w = torch.full((n, ctx.wire_bytes), 0xAA, dtype=torch.uint8, device="cuda")     # padding must be written too
    ctx.wire_pack_device(n, d_pi.data_ptr(), d_w.data_ptr())
    torch.cuda.synchronize()
    assert (d_w.cpu().numpy() == pkg.wire_pack(k, pis)).all()
    d_back = torch.zeros_like(d_pi)
    ctx.wire_unpack_device(n, d_w.data_ptr(), d_back.data_ptr())
    torch.cuda.synchronize()
    assert (d_back.cpu().numpy() == pis).all()


@pytest.mark.parametrize("k", [2, 3, 4])
def test_wire_mode_prove_equals_oracle_and_raw_mode(ctxs, k):
    """prove_batch through the packed link (device pack -> sliced D2H -> worker-pool expansion) gives the oracle's bytes; ragged
    batch spanning several sub-batches and lanes, caller buffers not pinned."""
    n = 21
    seeds = seeds_for_range(77, 0, n)
    ctx = pkg.KoskContext(k, 0, 8, 2)
    ctx.set_wire(100, 3)
    assert ctx.wire_info()["mode"] == 100
    l0 = ctx.kernel_launches()
    pk, sk, pi = ctx.prove_batch(seeds)
    assert ctx.kernel_launches() > l0
    ctx.set_wire(0)
    pk0, sk0, pi0 = ctx.prove_batch(seeds)
    assert (pk == pk0).all() and (sk == sk0).all() and (pi == pi0).all()
    for i in (0, 7, 8, 20):
        opk, osk, opi = O.oracle_prove(k, seeds[i])
        assert (pi[i] == opi).all() and (pk[i] == opk).all() and (sk[i] == osk).all(), i
    # packed API: the caller keeps the compact bytes
    pk2, sk2, w = ctx.prove_batch_packed(seeds)
    assert (pk2 == pk).all() and (sk2 == sk).all()
    assert (w == pkg.wire_pack(k, pi, 2)).all()
    assert ctx.verify_batch_packed(w, pk).all()
    # verify_batch in wire mode (host pack -> H2D -> device unpack) agrees with raw mode, on good and tampered proofs
    bad = pi.copy()
    bad[3, 0] ^= 1                     # f_shares[0][0]: a field element
    bad[5, O.layout(k).o_Tcomm + 9] ^= 0x80   # a digest byte
    bad[9, 2] = 0x00; bad[9, 3] = 0x20  # an element >= 4096: not representable, the sub-batch goes raw
    want = np.ones(n, bool); want[[3, 5, 9]] = False
    ctx.set_wire(0)
    got_raw = ctx.verify_batch(bad, pk)
    assert (got_raw == want).all()
    for pct in (100, 60):              # all packed / a split of packed and struct-byte proofs inside every sub-batch
        ctx.set_wire(pct)
        assert (ctx.verify_batch(bad, pk) == want).all(), pct
    ctx.set_wire(60)
    pk6, sk6, pi6 = ctx.prove_batch(seeds)
    assert (pi6 == pi).all() and (pk6 == pk).all() and (sk6 == sk).all()
    wb = pkg.wire_pack(k, np.delete(bad, 9, axis=0), 2)
    assert (ctx.verify_batch_packed(wb, np.delete(pk, 9, axis=0)) == np.delete(want, 9)).all()
    ctx.close()


def test_wire_mode_async_pipeline_many_steps(ctxs):
    """Consecutive async calls reuse the lanes' staging buffers: every step's bytes must still be its own (proof i of step s is
    checked through the device verifier and, sampled, against the oracle)."""
    k, B, steps = 2, 48, 5
    ctx = pkg.KoskContext(k, 0, B, 2)
    ctx.set_wire(70, 4)               # slices alternate between the packed and the struct-byte route
    outs = []
    seeds = [seeds_for_range(1 << 20, s * B, (s + 1) * B) for s in range(steps)]
    pin = [torch.from_numpy(s).pin_memory() for s in seeds]
    for s in range(steps):
        o = tuple(torch.zeros((B, nb), dtype=torch.uint8).pin_memory().numpy() for nb in (ctx.pk_bytes, ctx.sk_bytes, ctx.proof_bytes))
        rc = ctx.lib.kosk_b200_prove_batch_async(ctx._h, B, pin[s].data_ptr(), o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data)
        assert rc == 0
        outs.append(o)
    ctx.sync()
    for s in range(steps):
        pk, sk, pi = outs[s]
        assert ctx.verify_batch(pi, pk).all(), s
        for i in (0, B - 1):
            opk, osk, opi = O.oracle_prove(k, seeds[s][i])
            assert (pi[i] == opi).all() and (pk[i] == opk).all()
    ctx.close()


def test_context_serialises_concurrent_callers():
    """Every entry point takes the context's lock: two host threads proving on ONE context get correct, complete results (the
    reference's functions are re-entrant; a context is merely thread-safe)."""
    import threading
    k = 2
    ctx = pkg.KoskContext(k, 0, 8, 2)
    seeds = [seeds_for_range(900 + t, 0, 12) for t in range(2)]
    outs = [None, None]

    def work(t):
        outs[t] = ctx.prove_batch(seeds[t])
    ths = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for t in range(2):
        pk, sk, pi = outs[t]
        for i in (0, 11):
            opk, osk, opi = O.oracle_prove(k, seeds[t][i])
            assert (pi[i] == opi).all() and (pk[i] == opk).all() and (sk[i] == osk).all(), (t, i)
        assert ctx.verify_batch(pi, pk).all()
    ctx.close()


def test_destroy_refuses_while_a_pool_is_alive():
    ctx = pkg.KoskContext(2, 0, 8, 1)
    pool = ctx.pool_create(seeds_for_range(1, 0, 2))
    h = ctx._h
    ctx.lib.kosk_b200_destroy(h)                       # refused: the pool still references the context's streams and tables
    assert b"pools" in ctx.lib.kosk_b200_last_error()
    pk, sk, pi = pool.prove()                          # ... and the context is still usable
    assert ctx.verify_batch(pi, pk).all()
    pool.close()
    ctx.close()
