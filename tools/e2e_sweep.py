"""e2e (host-buffer API, pinned memory) timing over lanes/chunk/priority mode. Run under gpurun."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

def run(k, B, lanes, chunk, steps=4, **_ignored):
    ctx = KoskContext(k, 0, chunk, lanes)
    hs = [torch.from_numpy(seeds_for_range(7, s * B, (s + 1) * B)).pin_memory() for s in range(steps + 1)]
    h_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8).pin_memory(); h_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8).pin_memory()
    h_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8).pin_memory()
    def step(s):
        assert ctx.lib.kosk_b200_prove_batch(ctx._h, B, hs[s].data_ptr(), h_pk.data_ptr(), h_sk.data_ptr(), h_pi.data_ptr()) == 0
    step(steps); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps): step(s)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"k": k, "B": B, "lanes": lanes, "chunk": chunk, "ms": round(dt * 1e3, 2), "e2e_proofs_s": round(B / dt)}), flush=True)
    ctx.close()

if __name__ == "__main__":
    for c in json.loads(sys.argv[1]): run(**c)
