#!/usr/bin/env python
"""Platform ceiling of the end-to-end path: what N concurrent ranks can move between device and pinned host memory when NO
kernel runs, and what the host can expand.  The e2e number of bench.py is bounded by these, not by the GPU:

  d2h / h2d      each rank loops cudaMemcpyAsync of one 1024-proof output (Kyber512: 683 MB raw, 532 MB as wire images) between its
                 GPU and a pinned buffer; all ranks start together (barrier), the figure is total bytes / max-over-ranks time
  variants       cudaHostAlloc (torch pin_memory) vs cudaHostRegister over an anonymous mmap with MADV_HUGEPAGE / MAP_HUGETLB
  unpack         the host codec (wire -> struct mpcith_proof bytes) on T threads per rank, all ranks at once, no GPU traffic
  d2h+unpack     both at once (the pipeline of kosk_b200_prove_batch_async in wire mode)

Run:  python tools/d2h_ceiling.py                       (1 GPU)
      python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/d2h_ceiling.py
      python tools/d2h_ceiling.py --sweep 1,2,4,8       (launches itself under torchrun for each N)
One JSON line per N on stdout (rank 0); commit the output under profiles/.
"""
import argparse
import ctypes
import json
import mmap
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", default="")
    ap.add_argument("--kyber-k", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--threads", default="", help="comma list of unpack thread counts per rank (default: 2,4,8,16 capped by cpus/world)")
    args = ap.parse_args()
    if args.sweep:
        for i, n in enumerate(int(x) for x in args.sweep.split(",")):
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29611 + i), os.path.abspath(__file__), "--kyber-k", str(args.kyber_k), "--batch", str(args.batch), "--reps", str(args.reps)]
            if args.threads:
                cmd += ["--threads", args.threads]
            r = subprocess.run(cmd, capture_output=True, text=True)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(lines[-1] if lines else json.dumps({"n_gpus": n, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import mpcith_kyber_kosk_b200 as pkg
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    k, B = args.kyber_k, args.batch
    raw_bytes, wire_bytes = B * pkg.proof_bytes(k), B * pkg.wire_bytes(k)
    cudart = torch.cuda.cudart()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def copy_loop(dst, src, reps):
        """seconds for `reps` async copies on one stream, all ranks started together; max over ranks"""
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            dst.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
        st.synchronize()
        return allmax(time.perf_counter() - t0)

    res = {"n_gpus": world, "kyber_k": k, "batch": B, "cpus": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(),
           "raw_mb": raw_bytes / 1e6, "wire_mb": wire_bytes / 1e6, "simd": pkg.wire_simd()}
    d_raw = torch.empty(raw_bytes, dtype=torch.uint8, device=dev)
    h_raw = torch.empty(raw_bytes, dtype=torch.uint8).pin_memory()
    for name, nbytes in (("raw", raw_bytes), ("wire", wire_bytes)):
        s = copy_loop(h_raw[:nbytes], d_raw[:nbytes], args.reps)
        res[f"d2h_{name}_gbs"] = world * nbytes * args.reps / s / 1e9
        s = copy_loop(d_raw[:nbytes], h_raw[:nbytes], args.reps)
        res[f"h2d_{name}_gbs"] = world * nbytes * args.reps / s / 1e9

    # cudaHostRegister over an mmap with transparent / explicit huge pages
    def registered(flags_name):
        length = (raw_bytes + (1 << 21) - 1) & ~((1 << 21) - 1)
        try:
            if flags_name == "hugetlb":
                mm = mmap.mmap(-1, length, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | getattr(mmap, "MAP_HUGETLB", 0x40000))
            else:
                mm = mmap.mmap(-1, length, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
                mm.madvise(getattr(mmap, "MADV_HUGEPAGE", 14))
        except Exception as e:
            return None, f"mmap: {e}"
        arr = np.frombuffer(mm, dtype=np.uint8)
        arr[::4096] = 1                                           # touch
        rc = cudart.cudaHostRegister(arr.ctypes.data, length, 0)
        if int(rc) != 0:
            return None, f"cudaHostRegister rc={int(rc)}"
        return (mm, arr, torch.from_numpy(arr)), None
    for flavour in ("thp", "hugetlb"):
        got, err = registered(flavour)
        ok = torch.tensor([1 if got else 0])
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            res[f"d2h_raw_registered_{flavour}_gbs"] = None
            res[f"registered_{flavour}_error"] = err or "failed on another rank"
        else:
            s = copy_loop(got[2][:raw_bytes], d_raw, args.reps)
            res[f"d2h_raw_registered_{flavour}_gbs"] = world * raw_bytes * args.reps / s / 1e9
        if got:
            cudart.cudaHostUnregister(got[1].ctypes.data)
            del got

    # host codec alone, then together with the D2H of the wire bytes
    lib = pkg.load_library()
    rng = np.random.default_rng(rank)
    n_un = min(B, 256)
    pi = rng.integers(0, 3329, (n_un, pkg.proof_bytes(k) // 2)).astype(np.uint16).view(np.uint8).reshape(n_un, -1)
    w = pkg.wire_pack(k, pi, 4)
    out = np.empty_like(pi)
    cap = max(1, res["cpus"] // world)
    tlist = [int(x) for x in args.threads.split(",")] if args.threads else sorted({t for t in (1, 2, 4, 8, 16, cap) if t <= max(cap, 2)})
    res["unpack"] = {}
    for t in tlist:
        lib.kosk_b200_wire_unpack(k, n_un, ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data), t)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            lib.kosk_b200_wire_unpack(k, n_un, ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data), t)
        s = allmax(time.perf_counter() - t0)
        entry = {"proofs_per_s": world * 3 * n_un / s, "out_gbs": world * 3 * n_un * pkg.proof_bytes(k) / s / 1e9}
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            lib.kosk_b200_wire_pack(k, n_un, ctypes.c_void_p(pi.ctypes.data), ctypes.c_void_p(w.ctypes.data), t)
        entry["pack_proofs_per_s"] = world * 3 * n_un / allmax(time.perf_counter() - t0)
        # ... with the link busy: a thread keeps copying wire-sized buffers D2H while the codec runs
        stop = threading.Event()
        moved = [0]

        def pump():
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                while not stop.is_set():
                    h_raw[:wire_bytes].copy_(d_raw[:wire_bytes], non_blocking=True)
                    st.synchronize()
                    moved[0] += wire_bytes
        th = threading.Thread(target=pump)
        barrier()
        th.start()
        t0 = time.perf_counter()
        for _ in range(6):
            lib.kosk_b200_wire_unpack(k, n_un, ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data), t)
        s_local = time.perf_counter() - t0
        stop.set(); th.join()
        d2h_local = moved[0] / s_local
        s = allmax(s_local)
        tt = torch.tensor([d2h_local], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        entry.update({"with_d2h_proofs_per_s": world * 6 * n_un / s, "with_d2h_link_gbs": float(tt.item()) / 1e9})
        res["unpack"][str(t)] = entry
    if rank == 0:
        res["ceiling_proofs_per_s"] = {"raw_link": res["d2h_raw_gbs"] * 1e9 / pkg.proof_bytes(k), "wire_link": res["d2h_wire_gbs"] * 1e9 / pkg.wire_bytes(k)}
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
