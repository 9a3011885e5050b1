"""End-to-end sweep of the host-buffer batch calls over the link mode (raw struct bytes / 12-bit wire images), D2H slice size,
host worker threads and lanes.  Run under gpurun:  python tools/wire_sweep.py '[{"mode":1,"slice":16,"threads":8}, ...]'
Each config prints one JSON line: prove e2e (kosk_b200_prove_batch_async + sync) and verify e2e (kosk_b200_verify_batch_async)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext  # noqa: E402
from mpcith_kyber_kosk_b200.sharding import seeds_for_range  # noqa: E402


def run(k=2, B=1024, lanes=2, chunk=0, mode=100, slice=16, threads=8, steps=8, pinned=True, verify=True, packed=False):
    os.environ["KOSK_B200_WIRE_SLICE"] = str(slice)
    ctx = KoskContext(k, 0, chunk or B, lanes)
    ctx.set_wire(mode, threads)
    mk = (lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()) if pinned else (lambda n: torch.empty(n, dtype=torch.uint8))
    hs = [torch.from_numpy(seeds_for_range(7, s * B, (s + 1) * B)).pin_memory() for s in range(steps + 1)]
    outs = [(mk(B * ctx.pk_bytes), mk(B * ctx.sk_bytes), mk(B * ctx.proof_bytes)) for _ in range(2)]
    fn = ctx.lib.kosk_b200_prove_batch_packed_async if packed else ctx.lib.kosk_b200_prove_batch_async

    def step(s):
        o = outs[s % 2]
        assert fn(ctx._h, B, hs[s].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) == 0
    step(steps); step(steps); ctx.sync()
    ctx.wire_stats()
    t0 = time.perf_counter()
    tq = []
    for s in range(steps):
        tq0 = time.perf_counter(); step(s); tq.append(round((time.perf_counter() - tq0) * 1e3, 2))
    ctx.sync()
    dt = (time.perf_counter() - t0) / steps
    ws = ctx.wire_stats()
    res = {"k": k, "B": B, "lanes": lanes, "chunk": chunk or B, "mode": mode, "slice": slice, "threads": threads, "pinned": pinned, "packed": packed,
           "prove_ms": round(dt * 1e3, 2), "prove_e2e": round(B / dt), "enqueue_ms": tq, "wire_stats": ws}
    if verify and not packed:
        step(0); ctx.sync()
        pk, _, pi = outs[0]
        oks = [mk(B) for _ in range(2)]
        assert ctx.lib.kosk_b200_verify_batch_async(ctx._h, B, pi.data_ptr(), pk.data_ptr(), oks[0].data_ptr()) == 0
        ctx.sync()
        assert bool(oks[0].all())
        t0 = time.perf_counter()
        for s in range(steps):
            assert ctx.lib.kosk_b200_verify_batch_async(ctx._h, B, pi.data_ptr(), pk.data_ptr(), oks[s % 2].data_ptr()) == 0
        ctx.sync()
        dv = (time.perf_counter() - t0) / steps
        res.update({"verify_ms": round(dv * 1e3, 2), "verify_e2e": round(B / dv)})
    print(json.dumps(res), flush=True)
    ctx.close()


if __name__ == "__main__":
    for c in json.loads(sys.argv[1]):
        run(**c)
