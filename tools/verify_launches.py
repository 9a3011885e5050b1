"""One prove + two verify passes of a 1024-proof batch (for an ncu launch list of the verifier kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcith_kyber_kosk_b200 import KoskContext
from mpcith_kyber_kosk_b200.sharding import seeds_for_range
k, B = int(sys.argv[1]), int(sys.argv[2])
ctx = KoskContext(k, 0, B, 1)
dev = torch.device("cuda", 0)
seeds = torch.from_numpy(seeds_for_range(3, 0, B)).to(dev)
d_pk = torch.empty(B * ctx.pk_bytes, dtype=torch.uint8, device=dev); d_sk = torch.empty(B * ctx.sk_bytes, dtype=torch.uint8, device=dev)
d_pi = torch.empty(B * ctx.proof_bytes, dtype=torch.uint8, device=dev); d_ok = torch.empty(B, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
ctx.prove_batch_device(B, seeds.data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), st)
for _ in range(2): ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), st)
torch.cuda.synchronize()
assert bool(d_ok.all())
print("ok")
