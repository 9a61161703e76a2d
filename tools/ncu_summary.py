"""Markdown table of the headline metrics of an .ncu-rep:  python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum.per_second", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]
print("| metric | " + " | ".join(f"launch {i+1}" for i in range(len(data))) + " | unit |")
print("|---|" + "---|" * (len(data) + 1))
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i][:70].replace("|", "/") for r in data]
        print(f"| {k} | " + " | ".join(vals) + f" | {units[i]} |")
