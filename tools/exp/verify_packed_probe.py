"""Probe of the packed end-to-end verify path (kosk_b200_verify_batch_packed_async, two lanes): calls per second, host time per call, the plain H2D rate of
the same pinned wire images, and the H2D rate while verify kernels run on another stream.  B200, 1024 Kyber512 proofs: 84.4 k verifies/s = 12.1 ms per call
against 9.6 - 10.2 ms for the H2D copy alone (52 - 56 GB/s, unaffected by running kernels); host calls return in 0.1 ms.  Run: python tools/exp/verify_packed_probe.py"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range
B = 1024
ctx = KoskContext(2, 0, B, 2)
seeds = torch.from_numpy(seeds_for_range(5, 0, B)).pin_memory()
h_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8).pin_memory(); h_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8).pin_memory()
h_w = torch.empty(B * ctx.wire_bytes, dtype=torch.uint8).pin_memory()
assert ctx.lib.kosk_b200_prove_batch_packed(ctx._h, B, seeds.data_ptr(), h_pk.data_ptr(), h_sk.data_ptr(), h_w.data_ptr()) == 0
h_ok = [torch.zeros(B, dtype=torch.uint8).pin_memory() for _ in range(2)]
fn = ctx.lib.kosk_b200_verify_batch_packed_async
for lanes_note in ("",):
    for w in range(2): assert fn(ctx._h, B, h_w.data_ptr(), h_pk.data_ptr(), h_ok[w].data_ptr()) == 0
    ctx.sync()
    t0 = time.perf_counter()
    n = 12
    per = []
    for s in range(n):
        t1 = time.perf_counter()
        assert fn(ctx._h, B, h_w.data_ptr(), h_pk.data_ptr(), h_ok[s % 2].data_ptr()) == 0
        per.append(round((time.perf_counter() - t1) * 1e3, 2))
    t2 = time.perf_counter()
    ctx.sync(); dt = time.perf_counter() - t0
    print("host ms per call", per, "final sync ms", round((time.perf_counter() - t2) * 1e3, 2))
    print("packed verify e2e", B * n / dt, "per call ms", dt / n * 1e3, bool(h_ok[0].all()))
# plain H2D of the same bytes
d = torch.empty(B * ctx.wire_bytes, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(12): d.copy_(h_w, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D GB/s", 12 * B * ctx.wire_bytes / dt / 1e9, "ms per copy", dt / 12 * 1e3)
# device-only packed verify: unpack + verify
ctx.close()
# is the H2D copy slower while verify kernels run?  (device-resident verify loop on one stream, H2D on another)
ctx = KoskContext(2, 0, B, 1)
d_w = torch.empty(B * ctx.wire_bytes, dtype=torch.uint8, device="cuda"); d_w.copy_(h_w)
d_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8, device="cuda"); d_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8, device="cuda"); d_pk.copy_(h_pk)
d_ok = torch.empty(B, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
assert ctx.lib.kosk_b200_wire_unpack_device(ctx._h, B, d_w.data_ptr(), d_pi.data_ptr(), s1.cuda_stream) == 0
torch.cuda.synchronize()
for busy in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if busy:
        for _ in range(40): ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), s1.cuda_stream)
    with torch.cuda.stream(s2):
        e0.record(s2)
        for _ in range(8): d.copy_(h_w, non_blocking=True)
        e1.record(s2)
    torch.cuda.synchronize()
    print("H2D GB/s", "with verify kernels running" if busy else "alone", 8 * B * ctx.wire_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, bool(d_ok.all()) if busy else "")
ctx.close()
