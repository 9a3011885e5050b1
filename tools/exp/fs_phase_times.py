"""fs1 / fs2 phase times of a 1024-proof prove step and verify pass times (schedule experiments on the FS sponge kernels)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range
dev = torch.device("cuda", 0)
for B in (1024, 1):
    ctx = KoskContext(2, 0, 1024, 1)
    st = torch.cuda.current_stream().cuda_stream
    d_seed = torch.from_numpy(seeds_for_range(3, 0, B).copy()).to(dev)
    d_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8, device=dev); d_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8, device=dev)
    d_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8, device=dev); d_ok = torch.empty(B, dtype=torch.uint8, device=dev)
    run = lambda: ctx.prove_batch_device(B, d_seed.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ctx.set_profiling(True); ctx.phase_times(reset=True)
    for _ in range(5): run()
    torch.cuda.synchronize()
    ph = ctx.phase_times()
    ver = lambda: ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), st)
    for _ in range(5): ver()
    torch.cuda.synchronize()
    vt = ctx.phase_times()
    assert bool(d_ok.all())
    print(B, {n: round(ms / 5, 3) for n, (ms, c) in ph.items() if c and n in ("commit", "fs1", "view", "fs2")}, "verify_ms", round(vt["verify"][0] / 5, 3))
    ctx.close()
