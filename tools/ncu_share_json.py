"""profiles/ncu_share_eval.json (read by bench.py: roofline.traffic) from an `ncu --page raw --csv` export of tools/ncu_r02.sh.
Usage: python tools/ncu_share_json.py gpurun_out/ncu_r02_raw.csv profiles/ncu_share_eval.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
sh = [r for r in data if "k_share_ntt2" in r[hdr.index("Kernel Name")]][:2]        # the two launches of one prove step
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def g(r, k):
    return float(r[hdr.index(k)].replace(",", ""))


def nbytes(r, k):
    return g(r, k) * SCALE[units[hdr.index(k)]]


m = sh[0]
rows_main = 206 * 1024
out = {
    "kyber_k": 2, "batch": 1024,
    "dram_bytes_read": sum(nbytes(r, "dram__bytes_read.sum") for r in sh), "dram_bytes_write": sum(nbytes(r, "dram__bytes_write.sum") for r in sh),
    "launches_per_step": len(sh), "launch_grids": [r[hdr.index("launch__grid_size")] for r in sh],
    "time_ms_under_ncu": sum(g(r, "gpu__time_duration.sum") for r in sh),
    "fmaheavy_pct_main_launch": g(m, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "alu_pct_main_launch": g(m, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "issue_active_pct_main_launch": g(m, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "l1_data_pipe_pct_main_launch": g(m, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts_main_launch": g(m, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "warp_instructions_main_launch": g(m, "smsp__inst_executed.sum"),
    "warp_instructions_per_sharing": g(m, "smsp__inst_executed.sum") / rows_main,
    "shared_wavefronts_per_sharing": g(m, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / rows_main,
    "algorithmic_bytes": 214 * 1024 * (407 + 1454) * 2,
    "source": "ncu --set full --clock-control none on B200 (profiles/ncu_r02_metrics.csv, tools/ncu_r02.sh), command: python bench.py --steps 4 --warmup 3 "
              "--no-cpu-baseline --no-tensor-probe --no-extras --sustained-s 0",
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
