"""Timing sweep of the device-resident prove (and verify) path over batch / chunk / lanes / tensor flag. Run under gpurun."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range

def run(k, B, lanes, chunk, steps=4, verify=False, tensor=0, **_ignored):
    ctx = KoskContext(k, 0, chunk, lanes, bool(tensor))
    dev = torch.device("cuda", 0)
    seeds = [torch.from_numpy(seeds_for_range(99, s * B, (s + 1) * B)).to(dev) for s in range(steps + 2)]
    d_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8, device=dev); d_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8, device=dev)
    d_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8, device=dev); d_ok = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for s in range(2): ctx.prove_batch_device(B, seeds[s].data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps): ctx.prove_batch_device(B, seeds[2 + s].data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"k": k, "B": B, "lanes": lanes, "chunk": chunk, "tensor": tensor, "ms": round(ms, 3), "proofs_s": round(B / ms * 1e3)}
    if verify:
        ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), st); torch.cuda.synchronize()
        e0.record()
        for s in range(steps): ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        out["verify_ms"] = round(e0.elapsed_time(e1) / steps, 3); out["verifies_s"] = round(B / out["verify_ms"] * 1e3); out["all_ok"] = bool(d_ok.all())
    ctx.close()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    cfgs = json.loads(sys.argv[1])
    for c in cfgs:
        run(**c)
