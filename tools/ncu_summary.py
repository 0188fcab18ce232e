"""Condense an ncu report into a small metric,unit,value CSV: python tools/ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv, subprocess, sys
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "launch__shared_mem_config_size", "launch__shared_mem_per_block",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "smsp__issue_active.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__average_warp", "smsp__warp_issue_stalled")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for name, unit, val in zip(h, u, v):
        if name in ("Kernel Name",) or any(name.startswith(k) for k in KEEP):
            if ".peak_sustained" in name or ".per_second" in name and "dram" not in name:
                continue
            w.writerow([name, unit, val])
