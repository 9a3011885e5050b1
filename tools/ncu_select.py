"""Selected columns of an `ncu --page raw --csv` export (tools/ncu_r02.sh) -> a small CSV for profiles/.
Usage: python tools/ncu_select.py gpurun_out/ncu_r02_raw.csv profiles/ncu_r02_metrics.csv"""
import csv
import sys

WANT = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(w) for w in WANT if w in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])
print(f"{len(data)} kernels, {len(idx)} columns -> {sys.argv[2]}")
