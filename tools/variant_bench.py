"""Time one build variant of the library (mpcith_kyber_kosk_b200/libkosk_b200_<name>.so from build.build_variant, or the default
build): share evaluation alone on 214 x 1024 sharings and the phases of a 1024-proof prove step; prints one JSON line.  The share
evaluation's output is hashed so that variants can be compared for equality.  Usage: python tools/variant_bench.py [name] [k]"""
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpcith_kyber_kosk_b200 import api  # noqa: E402
from mpcith_kyber_kosk_b200.sharding import seeds_for_range  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else ""
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if name:
    api._lib = api.load_library(os.path.join(ROOT, "mpcith_kyber_kosk_b200", f"libkosk_b200_{name}.so"))
B = 1024
ctx = api.KoskContext(k, 0, B, 1)
dev = torch.device("cuda", 0)
rows = 214 * 1024
g = torch.Generator(device=dev); g.manual_seed(1)
y = torch.randint(0, 3329, (rows, 416), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
y[:, 407:] = 0
pl = torch.zeros((rows, 1456), dtype=torch.int16, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.share_eval_device(rows, y.data_ptr(), pl.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.share_eval_device(rows, y.data_ptr(), pl.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
share_ms = e0.elapsed_time(e1) / 5
digest = hashlib.sha256(pl[:4096].cpu().numpy().tobytes()).hexdigest()[:16]
del y, pl
seeds = [torch.from_numpy(seeds_for_range(9, s * B, (s + 1) * B)).to(dev) for s in range(13)]
o = [torch.empty(B * n, dtype=torch.uint8, device=dev) for n in (ctx.pk_bytes, ctx.sk_bytes, ctx.proof_bytes)]
for s in range(3):
    ctx.prove_batch_device(B, seeds[s].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), st)
torch.cuda.synchronize()
ctx.set_profiling(True); ctx.phase_times(reset=True)
e0.record()
for s in range(3, 13):
    ctx.prove_batch_device(B, seeds[s].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ph = ctx.phase_times()
pd = hashlib.sha256(o[2][:ctx.proof_bytes * 4].cpu().numpy().tobytes()).hexdigest()[:16]
print(json.dumps({"variant": name or "default", "k": k, "share_eval_ms": round(share_ms, 4), "share_digest": digest, "step_ms": round(e0.elapsed_time(e1) / 10, 4),
                  "proofs_per_s": round(B * 10 / (e0.elapsed_time(e1) * 1e-3)), "proof_digest": pd,
                  "phases": {n: round(v[0] / 10, 4) for n, v in ph.items() if v[1]}}), flush=True)
ctx.close()
