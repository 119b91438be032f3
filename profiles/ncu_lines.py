"""Per-CUDA-source-line samples / instruction counts of an .ncu-rep captured with --import-source on."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[2]
i_line, i_src, i_addr = 0, 1, 2
i_samp = hdr.index("# Samples")
i_inst = hdr.index("Instructions Executed")
i_lsb = hdr.index("stall_long_sb")
i_bar = hdr.index("stall_barrier")
i_ssb = hdr.index("stall_short_sb")
i_wait = hdr.index("stall_wait")
lines = []
for r in rows[3:]:
    if len(r) <= i_inst or r[i_addr] != "-":
        continue  # keep the per-source-line aggregate rows only
    try:
        lines.append((int(r[i_samp]), int(r[i_inst]), int(r[i_lsb]), int(r[i_bar]), int(r[i_ssb]), int(r[i_wait]),
                      r[i_line], r[i_src].strip()))
    except ValueError:
        pass
ts, ti = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"total samples {ts}, warp instructions {ti}")
print(" samples   %   |  warp-inst    %   | long_sb barrier short_sb wait | line: source")
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{l[0]:8d} {100 * l[0] / ts:5.1f} | {l[1]:10d} {100 * l[1] / ti:5.1f} | {l[2]:7d} {l[3]:7d} {l[4]:7d} {l[5]:5d} | {l[6]}: {l[7][:100]}")
