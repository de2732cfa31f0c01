"""Summarise an .ncu-rep (run here, no GPU needed): python scripts/ncu_summary.py gpurun_out/x.ncu-rep [pattern ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or [
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__pipe_fp64_cycles_active_realtime.avg.pct", "sm__throughput.avg.pct",
    "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit", "launch__waves_per_multiprocessor", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
    "lts__t_bytes.sum ", "lts__t_sector_hit_rate", "smsp__average_warps_issue_stalled", "smsp__issue_active.avg.pct",
    "smsp__inst_executed.sum ", "sm__inst_executed_pipe_tensor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg ", "smsp__warps_eligible.avg.per_cycle", "sm__cycles_active.avg",
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print("== kernel:", name[:80])
    for h, u, v in zip(hdr, units, r):
        hh = h + " "
        if any(p in hh for p in pats):
            print("  %-88s %-10s %s" % (h, u, v))
