"""Print the per-lane phase timeline of one device-resident batch (debug). Run under gpurun."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range
k, B, lanes, chunk = (int(x) for x in sys.argv[1:5])
ctx = KoskContext(k, 0, chunk, lanes)
dev = torch.device("cuda", 0)
seeds = torch.from_numpy(seeds_for_range(3, 0, B)).to(dev)
d_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8, device=dev); d_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8, device=dev)
d_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): ctx.prove_batch_device(B, seeds.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
torch.cuda.synchronize()
ctx.set_profiling(True)
ctx.prove_batch_device(B, seeds.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
tr = ctx.debug_trace()
for lane in range(lanes):
    print("lane", lane, " ".join(f"{n}@{t:.2f}" for l, n, t in tr if l == lane))
