"""CPU suite: batch-mode sharding of independent proofs across ranks (no data-path collective), exercised with a
world_size-2 gloo group; the oracle stands in for the device so that the host-side plumbing is what is tested."""
import os
import socket

import numpy as np
import pytest

from mpcith_kyber_kosk_b200.sharding import seeds_for_range, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_seeds_are_placement_independent():
    a = seeds_for_range(1000, 0, 16)
    b = np.concatenate([seeds_for_range(1000, *shard_range(16, r, 4)) for r in range(4)])
    assert (a == b).all() and a.shape == (16, 32)
    assert int.from_bytes(bytes(a[5, :8]), "little") == 1005 and not a[:, 8:].any()


def _worker(rank, world, port, n, q):
    import hashlib
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    seeds = seeds_for_range(77, lo, hi)
    digs = []
    for s in seeds:                     # oracle in place of the device: this test is about the host-side sharding
        pk, sk, pi = O.oracle_prove(2, bytes(s))
        digs.append(np.frombuffer(hashlib.sha256(bytes(pk) + bytes(pi)).digest(), np.uint8))
    mine = torch.zeros(n, 32, dtype=torch.uint8)
    if digs:
        mine[lo:hi] = torch.from_numpy(np.stack(digs))
    dist.barrier()
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)        # test-side gather only; the data path itself has no collective
    t = torch.tensor([float(hi - lo)]); dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tm = torch.tensor([0.5 + rank]); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((mine.numpy().copy(), float(t[0]), float(tm[0])))
    dist.destroy_process_group()


def test_two_rank_gloo_batch():
    import hashlib
    import torch.multiprocessing as mp
    import oracle_lib as O
    n, world = 4, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, total, tmax = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == n and tmax == 1.5
    for i, s in enumerate(seeds_for_range(77, 0, n)):
        pk, sk, pi = O.oracle_prove(2, bytes(s))
        assert bytes(got[i]) == hashlib.sha256(bytes(pk) + bytes(pi)).digest()
