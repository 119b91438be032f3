"""Print the headline metrics and the top stall reasons / source lines of an .ncu-rep (run here, no GPU)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.sum",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:60])
    for k in keys:
        if k in hdr:
            print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    stalls = [(h, r[i]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_ratio") is False and "pct" in h]
    stalls = sorted(stalls, key=lambda kv: -float(kv[1] or 0))[:8]
    for h, v in stalls:
        print(f"  stall {h:66s} {v}")
