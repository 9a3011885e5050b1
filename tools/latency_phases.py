"""Per-phase device time of a single proof (B = 1) and a single verification: where the 2.7 ms go (run under gpurun)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

dev = torch.device("cuda", 0)
out = {}
for k in (2, 3, 4):
    ctx = KoskContext(k, 0, 8, 1)
    st = torch.cuda.current_stream().cuda_stream
    d_seed = torch.from_numpy(seeds_for_range(3, 0, 1).copy()).to(dev)
    d_pk = torch.empty(ctx.pk_bytes, dtype=torch.uint8, device=dev); d_sk = torch.empty(ctx.sk_bytes, dtype=torch.uint8, device=dev)
    d_pi = torch.empty(ctx.proof_bytes, dtype=torch.uint8, device=dev)
    run = lambda: ctx.prove_batch_device(1, d_seed.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ctx.set_profiling(True); ctx.phase_times(reset=True)
    n = 20
    for _ in range(n): run(); torch.cuda.synchronize()
    ph = ctx.phase_times()
    out[f"kyber{256*k}"] = {name: round(ms / n * 1e3, 1) for name, (ms, c) in ph.items() if c}
    out[f"kyber{256*k}"]["sum_us"] = round(sum(ms for ms, c in ph.values()) / n * 1e3, 1)
    ctx.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/latency_phases.json", "w"), indent=1)
