#!/usr/bin/env python
"""bench.py -- KOSK proofs/s on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation (oracle/_ref, else the C port)

A "step" is one pass of the hot path over one batch: kyber_verifiable_keygen for `--batch` independent proofs
(default: BASELINE configs[1], Kyber512 x 1024) on every rank; ranks are independent (batch mode shards whole proofs,
no data-path collective), so scaling is weak and `value` = all ranks' proofs / max-over-ranks device time.
`value` is timed with inputs (seeds) and outputs resident in HBM; `e2e` goes through the host-buffer C-ABI call
(kosk_b200_prove_batch) with pinned host buffers, so H2D of the seeds and D2H of pk/sk/proof are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# Algorithmic work per prove: field MACs and Keccak-f permutations
MACS_PER_PROVE = {2: 135.14e6, 3: 142.69e6, 4: 160.41e6}
KECCAK_PER_PROVE = {2: 11196, 3: 11220, 4: 11254}
SHARE_MACS_PER_ROW = 1303 * 407          # ss.cpp:23-32: one sharing = 1303 x 407 MACs
# share_ntt.cuh: FMA-pipe instructions per sharing (one warp): 2 forward passes x 328 + 6 inverse passes x 472 (DFT mat-vecs,
# twiddles, pointwise products, two IMADs per Montgomery reduction), x 32 lanes
NTT_IMAD_PER_SHARING = (2 * 328 + 6 * 472) * 32
INT_OPS_PER_KECCAK = 7440                # 24 rounds x 155 64-bit logic ops x 2 (32-bit lanes)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kyber-k", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1024, help="proofs per rank per step")
    ap.add_argument("--chunk", type=int, default=0, help="proofs per kernel wave (0 = batch / lanes)")
    ap.add_argument("--lanes", type=int, default=1, help="pipeline lanes (CUDA streams with their own scratch) of the device-resident run")
    ap.add_argument("--e2e-lanes", type=int, default=2, help="lanes of the host-buffer (e2e) run: the D2H copies of one step overlap the kernels of the next")
    ap.add_argument("--wire-threads", type=int, default=0, help="host worker threads per rank expanding wire images (0 = cpus / local ranks, 2..16)")
    ap.add_argument("--wire-percent", default="", help="comma list of packed shares (0..100) the e2e calibration tries (default 0,50,75,100)")
    ap.add_argument("--cpu-sample", type=int, default=8, help="proofs of the bounded single-core CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tensor-probe", action="store_true", help="skip the short measurement of the opt-in tensor-core path")
    ap.add_argument("--verify", action="store_true", help="also time kyber_kosk_verify on the produced proofs")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  The sampler runs from
    process start (nvidia-smi needs a moment to come up); only samples stamped inside [mark_start, mark_stop] are used."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.02]
        use = inside if inside else [r for _, r in self.rows[-3:]]
        sm = sorted(int(float(r[1])) for r in use if len(r) > 1 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[2])) for r in use if len(r) > 2 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in use if len(r) > 3 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(inside), "power_w_max": max(pw) if pw else None}


def cpu_oracle():
    """(kind, prove_fn(k, seed) -> (pk, sk, pi), verify_fn): the unmodified reference if oracle/_ref travelled, else the C port."""
    import oracle_lib as O
    if all(O.ref(k) is not None for k in (2, 3, 4)):
        return "reference", O.ref_prove, O.ref_verify
    return "port", O.oracle_prove, O.oracle_verify


def cpu_baseline(k, nsample):
    kind, prove, verify = cpu_oracle()
    import oracle_lib as O
    prove(k, O.seed_of(0))                                   # warm-up (table generation excluded)
    t0 = time.perf_counter()
    outs = [prove(k, O.seed_of(10 + i)) for i in range(nsample)]
    t1 = time.perf_counter()
    nv = min(nsample, 4)
    for pk, sk, pi in outs[:nv]:
        assert verify(k, pi, pk)
    t2 = time.perf_counter()
    return {"value": nsample / (t1 - t0), "unit": "proofs/s", "cores": 1, "kind": kind,
            "sample": f"{nsample} sequential Kyber{256 * k} kyber_verifiable_keygen calls on one host core (of {os.cpu_count()})",
            "verify_per_s": nv / (t2 - t1)}


def run_reference(args):
    """The reference's own CPU implementation with every host thread (ctypes releases the GIL)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    import oracle_lib as O
    kind, prove, verify = cpu_oracle()
    k, threads = args.kyber_k, os.cpu_count() or 1
    per_step = threads                                      # bounded sample: one proof per host thread per step
    prove(k, O.seed_of(0))
    pool = ThreadPoolExecutor(threads)

    def step(s):
        list(pool.map(lambda i: prove(k, O.seed_of(1000 * s + i)), range(per_step)))
    for w in range(min(args.warmup, 1)):
        step(-1 - w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(s)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = f"{per_step} proofs per step, one per host thread ({threads} threads), of the {args.batch}-proof batch"
    print(json.dumps({
        "impl": "reference", "metric": "KOSK proofs/sec (prove)", "value": val, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": f"Kyber{256 * k} kyber_verifiable_keygen, batch {args.batch} (bounded CPU sample)", "kyber_k": k, "batch": args.batch},
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and so its pinned staging buffers, first-touch) to the CPUs next to its GPU: at 8 GPUs the
    e2e path moves 8 x 683 MB per step through host memory and a rank left on the far socket pays the inter-socket link twice."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # 00000000:1b:00.0 -> 0000:1b:00.0
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return {"bus": bus, "cpus": cpus}
    except Exception as e:                                  # best effort: the bench runs unbound if the topology is not exposed
        return {"error": str(e)[:80]}
    return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mpcith_kyber_kosk_b200 import build as _build
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        _build.build()                                     # no-op when libkosk_b200.so is up to date (it normally travels with the repo)
    else:                                                  # other local ranks wait for rank 0's build, if one was needed
        t_wait = time.time()
        while _build.needs_build() and time.time() - t_wait < 600:
            time.sleep(2)
        if time.time() - t_wait > 1:
            time.sleep(3)
    from mpcith_kyber_kosk_b200 import KoskContext
    from mpcith_kyber_kosk_b200.sharding import seeds_for_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the KOSK core has no CPU path; use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        # keep stdout to the one JSON line: NCCL's banner / debug output (it prints to stdout by default) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if "KOSK_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["KOSK_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    k, B = args.kyber_k, args.batch
    chunk = args.chunk or -(-B // args.lanes)
    ctx = KoskContext(k, local, chunk, args.lanes)
    npk, nsk, npi = ctx.pk_bytes, ctx.sk_bytes, ctx.proof_bytes

    # device-resident inputs/outputs; a different seed range every step and rank (placement-independent seeds)
    total_steps = args.warmup + args.steps
    d_seeds = [torch.from_numpy(seeds_for_range(1 << 32, (s * world + rank) * B, (s * world + rank + 1) * B)).to(dev) for s in range(total_steps)]
    d_pk = torch.empty(B * npk, dtype=torch.uint8, device=dev)
    d_sk = torch.empty(B * nsk, dtype=torch.uint8, device=dev)
    d_pi = torch.empty(B * npi, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(s):
        ctx.prove_batch_device(B, d_seeds[s].data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), stream)

    peaks = ctx.int_peak() if rank == 0 else None
    for s in range(args.warmup):
        step(s)
    barrier()
    ctx.set_profiling(True)
    ctx.phase_times(reset=True)
    sampler.mark_start()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sampler.mark_stop()
    launches = ctx.kernel_launches() - l0
    phases = ctx.phase_times(reset=True)
    ctx.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end to end through the host-buffer C-ABI call (kosk_b200_prove_batch_async + kosk_b200_sync), pinned host memory; every
    # step's H2D (seeds) and D2H (pk, sk, proofs) copies are inside the timed region; consecutive steps alternate over the lanes
    # and two host buffer sets so that the copies of step i overlap the kernels of step i+1.  Measured twice: proofs crossing the
    # link as struct bytes ("raw", round 1) and as 12-bit wire images expanded into the same reference-layout buffers by host worker
    # threads ("wire"); the caller-visible bytes are identical.  The headline `e2e` is the faster of the two (named in `e2e.link`).
    h_seeds = [torch.from_numpy(seeds_for_range(1 << 33, (s * world + rank) * B, (s * world + rank + 1) * B)).pin_memory() for s in range(args.steps + 1)]
    h_out = [(torch.empty(B * npk, dtype=torch.uint8).pin_memory(), torch.empty(B * nsk, dtype=torch.uint8).pin_memory(),
              torch.empty(B * npi, dtype=torch.uint8).pin_memory()) for _ in range(2)]
    ctx_e = KoskContext(k, local, B, args.e2e_lanes)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    wire_threads = args.wire_threads or max(2, min(16, len(os.sched_getaffinity(0)) // max(1, local_world)))

    def e2e_run(fn_name, steps):
        fn = getattr(ctx_e.lib, fn_name)

        def e2e_step(s):
            o = h_out[s % 2]
            rc = fn(ctx_e._h, B, h_seeds[s % len(h_seeds)].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
            assert rc == 0, ctx_e.lib.kosk_b200_last_error()
        for w in range(max(2, args.e2e_lanes)):             # every lane (and its staging buffers) has run once
            e2e_step(args.steps - w)
        ctx_e.sync()
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            e2e_step(s)
        ctx_e.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())                              # seconds per step, max over ranks
    # calibration: which share of the proofs should travel packed on THIS host (link bytes vs host memory bandwidth); short runs
    cal_steps = max(3, args.steps // 4)
    e2e_cal = {}
    for pct in ([int(x) for x in args.wire_percent.split(",")] if args.wire_percent else [0, 50, 75, 100]):
        ctx_e.set_wire(pct, wire_threads)
        e2e_cal[pct] = e2e_run("kosk_b200_prove_batch_async", cal_steps)
    wire_pct = min(e2e_cal, key=e2e_cal.get)
    ctx_e.set_wire(wire_pct, wire_threads)
    e2e_step_s = e2e_run("kosk_b200_prove_batch_async", args.steps)
    e2e_max = e2e_step_s * args.steps
    h_pk, h_sk, h_pi = h_out[(args.steps - 1) % 2]
    h_pi_copy = h_pi.clone()
    # the caller keeps the compact bytes (kosk_b200_prove_batch_packed_async): same buffers, only wire_bytes per proof are written
    e2e_packed = e2e_run("kosk_b200_prove_batch_packed_async", args.steps) * args.steps
    h_wire_last = h_pi[:B * ctx_e.wire_bytes].clone().pin_memory()
    h_pi.copy_(h_pi_copy)

    # ---- sanity on the measured outputs: every proof of the last e2e step verifies on the device; rank 0 checks one against the oracle
    pi_np = h_pi.numpy().reshape(B, npi)
    pk_np = h_pk.numpy().reshape(B, npk)
    nver = min(B, 64)
    tv0 = time.perf_counter()
    ok = ctx.verify_batch(pi_np[:nver], pk_np[:nver])
    tv = time.perf_counter() - tv0
    assert ok.all(), "a measured proof failed verification"
    verify_stats = None
    if True:                                               # kyber_kosk_verify throughput on the proofs just produced (device-resident)
        d_ok = torch.empty(B, dtype=torch.uint8, device=dev)
        d_pi.copy_(h_pi.to(dev)); d_pk.copy_(h_pk.to(dev))
        ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), stream)
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(max(1, args.steps // 2)):
            ctx.verify_batch_device(B, d_pi.data_ptr(), d_pk.data_ptr(), d_ok.data_ptr(), stream)
        v1.record(); torch.cuda.synchronize()
        assert bool(d_ok.all())
        tvm = torch.tensor([v0.elapsed_time(v1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tvm, op=dist.ReduceOp.MAX)
        verify_stats = {"verifies_per_s": world * B * max(1, args.steps // 2) / (float(tvm.item()) * 1e-3), "batch_per_gpu": B,
                        "note": "kyber_kosk_verify, device-resident proofs, CUDA events, max over ranks"}
        # end to end: proofs and public keys in pinned HOST buffers (reference layout), accept bits back on the host, through
        # kosk_b200_verify_batch (synchronous; its sub-batches alternate over two lanes so the H2D of one overlaps the kernels of the
        # previous one); link = raw struct bytes vs 12-bit wire images packed by the host workers; and the packed API
        ctx_v = KoskContext(k, local, B, 2)
        h_oks = [torch.zeros(B, dtype=torch.uint8).pin_memory() for _ in range(2)]
        vsteps = max(4, args.steps // 2)

        def verify_e2e(fn_name, src, steps):
            fn = getattr(ctx_v.lib, fn_name)
            for w in range(2):
                assert fn(ctx_v._h, B, src.data_ptr(), h_pk.data_ptr(), h_oks[w].data_ptr()) == 0
            ctx_v.sync()
            barrier()
            t0v = time.perf_counter()
            for s_ in range(steps):
                assert fn(ctx_v._h, B, src.data_ptr(), h_pk.data_ptr(), h_oks[s_ % 2].data_ptr()) == 0
            ctx_v.sync()
            dtv = time.perf_counter() - t0v
            assert bool(h_oks[0].all()) and bool(h_oks[1].all())
            tt = torch.tensor([dtv], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return world * B * steps / float(tt.item())
        ve = {}
        for pct in e2e_cal:
            ctx_v.set_wire(pct, wire_threads)
            ve[pct] = verify_e2e("kosk_b200_verify_batch_async", h_pi, max(3, vsteps // 2))
        vpct = max(ve, key=ve.get)
        ctx_v.set_wire(vpct, wire_threads)
        ve_best = verify_e2e("kosk_b200_verify_batch_async", h_pi, vsteps)
        ve_packed = verify_e2e("kosk_b200_verify_batch_packed_async", h_wire_last, vsteps)
        verify_stats["e2e"] = {"value": ve_best, "unit": "verifies/s", "wire_percent": vpct, "calibration_verifies_per_s": {str(p_): v for p_, v in ve.items()},
                               "h2d_bytes_per_step": B * npk + (B * vpct // 100) * ctx_v.wire_bytes + (B - B * vpct // 100) * npi, "d2h_bytes_per_step": B,
                               "packed_api": {"value": ve_packed, "unit": "verifies/s", "h2d_bytes_per_step": B * (npk + ctx_v.wire_bytes)},
                               "api": "kosk_b200_verify_batch_async + kosk_b200_sync (host buffers in the reference layout, pinned; 2 lanes: the H2D of one call overlaps the kernels of the previous one)"}
        ctx_v.close()

    # ---- single-proof latency (BASELINE configs[2]): host API, one seed in -> pk, sk, proof out / proof in -> accept bit out
    ctx_l = KoskContext(k, local, 8, 1)
    ls = seeds_for_range(1 << 34, rank * 64, rank * 64 + 24)
    lp = ctx_l.prove_batch(ls[:1])
    tp, tvv = [], []
    for i in range(1, 17):
        t0l = time.perf_counter(); lp = ctx_l.prove_batch(ls[i:i + 1]); tp.append(time.perf_counter() - t0l)
    for i in range(16):
        t0l = time.perf_counter(); okl = ctx_l.verify_batch(lp[2], lp[0]); tvv.append(time.perf_counter() - t0l)
    assert okl.all()
    ctx_l.close()
    lat = torch.tensor([float(np.median(tp)) * 1e3, float(np.median(tvv)) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lat, op=dist.ReduceOp.MAX)
    latency_stats = {"prove_ms": float(lat[0].item()), "verify_ms": float(lat[1].item()),
                     "note": "median of 16 single calls through the host-buffer C ABI (copies included), max over ranks; one proof uses one GPU"}

    tensor_stats = None
    if rank == 0 and not args.no_tensor_probe:
        # opt-in experimental path (NOT the headline): share evaluation on int8 tensor cores, same bytes
        ctx_t = KoskContext(k, local, chunk, 1, True)
        for s in range(2):
            ctx_t.prove_batch_device(B, d_seeds[s].data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), stream)
        torch.cuda.synchronize()
        ctx_t.set_profiling(True); ctx_t.phase_times(reset=True)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for s in range(3):
            ctx_t.prove_batch_device(B, d_seeds[s].data_ptr(), d_pk.data_ptr(), d_sk.data_ptr(), d_pi.data_ptr(), stream)
        t1e.record(); torch.cuda.synchronize()
        pht = ctx_t.phase_times()
        pi_t = d_pi[:npi].cpu().numpy().copy()
        tensor_stats = {"proofs_per_s_one_gpu": B * 3 / (t0e.elapsed_time(t1e) * 1e-3), "ms_per_step": t0e.elapsed_time(t1e) / 3,
                        "share1_ms_per_step": pht["share1"][0] / 3, "peak_int8_mac_per_s_mma_sync": peaks["imma_int8_mac"],
                        "note": "KOSK_F_TENSOR: limb-split int8 mma.sync share evaluation; experimental, not the plan of record"}
        ctx_t.close()

    if rank == 0:
        import oracle_lib as O
        kind, prove, verify = cpu_oracle()
        if tensor_stats is not None:
            opk_t, osk_t, opi_t = prove(k, bytes(d_seeds[2][0].cpu().numpy()))
            assert (opi_t == pi_t).all(), "tensor-path proof differs from the CPU oracle"
            tensor_stats["bit_exact_vs_oracle"] = True
        opk, osk, opi = prove(k, bytes(h_seeds[args.steps - 1][0].numpy()))
        assert (opi == pi_np[0]).all() and (opk == pk_np[0]).all(), "measured proof differs from the CPU oracle"
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = (json.load(open(peaks_file))["hbm_gbs"], "measured") if os.path.exists(peaks_file) else (6650.0, "fallback")
        sh_ms, sh_calls = phases["share1"]
        rows = min(chunk, B) * ctx_rows(k)                    # sharings per launch
        ms_per_launch = sh_ms / max(sh_calls, 1)
        macs_per_launch = rows * SHARE_MACS_PER_ROW
        achieved_tmac = macs_per_launch / (ms_per_launch * 1e-3) / 1e12 if sh_calls else None
        peak_tmac = peaks["imad"] / 1e12
        # HBM view of the same kernel: algorithmic bytes = Y rows in (407 x 2 B) + planes out (1454 x 2 B) per sharing
        bytes_per_launch = macs_per_launch / SHARE_MACS_PER_ROW * (407 + 1454) * 2
        step_ms = ms_max / args.steps
        # Dominant kernel = the share evaluation.  `achieved` is the ALGORITHMIC rate (530 321 field MACs per sharing, the
        # reference's table mat-vec) over the measured time.  With the default NTT-convolution kernel (share_ntt.cuh) the device executes
        # only NTT_IMAD_PER_SHARING multiply-adds per sharing, so the algorithmic rate exceeds the IMAD issue peak (frac > 1); `executed`
        # is what the integer pipe actually does.  KOSK_B200_SHARE_NTT=0 selects the dense-table GEMM, for which both coincide.
        use_ntt = os.environ.get("KOSK_B200_SHARE_NTT", "1") != "0"
        executed_tops = rows * NTT_IMAD_PER_SHARING / (ms_per_launch * 1e-3) / 1e12 if (sh_calls and use_ntt) else achieved_tmac
        roofline = {"bound": "int32-pipe",
                    "kernel": ("k_share_ntt (share evaluation as a blocked NTT convolution over GF(3329), ss.cpp:23-32; first share-eval phase, all sharings of the step)"
                               if use_ntt else "k_gf_gemm<8> (share evaluation, ss.cpp:23-32; first share-eval phase = 3 launches: f/NTT_f | eta constants | s,e,z)"),
                    "achieved": achieved_tmac, "peak": peak_tmac, "unit": "TMAC/s", "frac": (achieved_tmac / peak_tmac) if achieved_tmac else None,
                    "executed": {"imad_per_sharing": NTT_IMAD_PER_SHARING if use_ntt else SHARE_MACS_PER_ROW, "achieved": executed_tops, "unit": "T IMAD/s",
                                 "frac": (executed_tops / peak_tmac) if executed_tops else None,
                                 "note": "multiply-adds the kernel executes (16-point DFT mat-vecs, pointwise products, Montgomery reductions) against the IMAD issue peak"},
                    "traffic": ncu_traffic(k, B),
                    "peak_source": "IMAD issue-rate microbenchmark run in this process (MEASURED_PEAKS.json has no integer entry)",
                    "ms_per_launch": ms_per_launch, "share_of_step": sh_ms / ms if ms else None}
        out = {
            "metric": "KOSK proofs/sec (prove)", "value": world * B * args.steps / (ms_max * 1e-3), "unit": "proofs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": f"Kyber{256 * k} kyber_verifiable_keygen, batch of {B} independent proofs per GPU (BASELINE configs[1])",
                       "kyber_k": k, "batch_per_gpu": B, "chunk": chunk, "lanes": args.lanes, "parallelism": f"proof-sharded x{world}, no collective",
                       "l2": f"per-step working set {B * (npi + 1_500_000) / 1e6:.0f} MB >> 126 MB L2, fresh seeds every step"},
            "e2e": {"value": world * B * args.steps / e2e_max, "unit": "proofs/s", "h2d_bytes_per_step": B * 32,
                    "d2h_bytes_per_step": B * (npk + nsk) + (B * wire_pct // 100) * ctx_e.wire_bytes + (B - B * wire_pct // 100) * npi,
                    "lanes": args.e2e_lanes, "link": f"{wire_pct}% of the proofs as 12-bit wire images, the rest as struct bytes",
                    "wire_percent": wire_pct, "wire_threads": wire_threads, "wire_simd": ctx_e.wire_info()["simd"],
                    "calibration_proofs_per_s": {str(p_): world * B / v for p_, v in e2e_cal.items()}, "calibration_steps": cal_steps,
                    "packed_api": {"value": world * B * args.steps / e2e_packed, "unit": "proofs/s", "d2h_bytes_per_step": B * (npk + nsk + ctx_e.wire_bytes),
                                   "api": "kosk_b200_prove_batch_packed_async: the caller keeps the 12-bit wire images"},
                    "api": "kosk_b200_prove_batch_async + kosk_b200_sync (host buffers in the reference layout, pinned; step i+1 computes while step i copies out)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_hbm": {"bound": "hbm", "achieved": bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 if sh_calls else None, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 / hbm_peak) if sh_calls else None, "peak_source": hbm_src},
            "int_pipe": {"imad_tops": peaks["imad"] / 1e12, "lop3_tops": peaks["lop3"] / 1e12, "shf_tops": peaks["shf"] / 1e12,
                         "algorithmic_int_ops_per_proof": 2 * MACS_PER_PROVE[k] + KECCAK_PER_PROVE[k] * INT_OPS_PER_KECCAK,
                         "whole_job_frac_of_fma_plus_alu_pipe_peak": (world * B * args.steps / (ms_max * 1e-3)) * (2 * MACS_PER_PROVE[k] + KECCAK_PER_PROVE[k] * INT_OPS_PER_KECCAK)
                         / (world * (peaks["imad"] + peaks["shf"])),
                         "note": "lop3 = 3-register-input LOP3 chain (register-port bound), shf = 2-register ALU op: the ALU pipe issue peak"},
            "phases_ms_per_step": {n: v[0] / args.steps for n, v in phases.items() if v[1]},
            "verify_check": {"proofs": nver, "all_accept": True, "wall_s": tv},
        }
        if numa:
            out["e2e"]["host_binding"] = numa
        if verify_stats:
            out["verify"] = verify_stats
        out["single_proof_latency"] = latency_stats
        if tensor_stats:
            out["experimental_tensor_path"] = tensor_stats
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(k, args.cpu_sample)
        print(json.dumps(out))
    if ctx_e is not ctx:
        ctx_e.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(k, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the share-evaluation launches of one step, from the committed
    `ncu --set full` capture (profiles/ncu_share_eval.json); only valid for the configuration that was profiled."""
    path = os.path.join(ROOT, "profiles", "ncu_share_eval.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    if d.get("kyber_k") != k or d.get("batch") != B:
        return None
    return {"dram_bytes_per_step": d["dram_bytes_read"] + d["dram_bytes_write"], "unit": "B", "source": d.get("source")}


def ctx_rows(k):
    """sharings evaluated by the first (dominant) share-eval launch per proof: 2F + 2K(2eta+1) + 2K + 4 eta K."""
    eta = 3 if k == 2 else 2
    F = 70 + 2 * k + 1
    return 2 * F + 2 * k * (2 * eta + 1) + 2 * k + 4 * eta * k


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
