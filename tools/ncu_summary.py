"""Condense an .ncu-rep (ncu --set full) into the per-launch summary CSV kept under profiles/:
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_summary.csv"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, data = rows[0], rows[1], rows[2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
w.writerow(["Kernel Name", ""] + [d[h.index("Kernel Name")] for d in data])
for k in KEYS:
    if k in h:
        i = h.index(k)
        w.writerow([k, u[i]] + [d[i] for d in data])
