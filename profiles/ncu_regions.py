"""Aggregate samples / warp instructions of an .ncu-rep (--import-source on) over source-line ranges of one file.
usage: python profiles/ncu_regions.py rep.ncu-rep name:lo-hi [name:lo-hi ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    name, rng = a.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[2]
i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = {name: [0, 0] for name, _, _ in regions}
tot["(other)"] = [0, 0]
for r in rows[3:]:
    if len(r) <= i_inst or r[2] != "-":
        continue
    try:
        line, samp, inst = int(r[0]), int(r[i_samp]), int(r[i_inst])
    except ValueError:
        continue
    for name, lo, hi in regions:
        if lo <= line <= hi:
            tot[name][0] += samp
            tot[name][1] += inst
            break
    else:
        tot["(other)"][0] += samp
        tot["(other)"][1] += inst
ts = sum(v[0] for v in tot.values())
ti = sum(v[1] for v in tot.values())
for name, (s, i) in tot.items():
    print(f"{name:24s} samples {s:8d} {100 * s / max(ts, 1):5.1f} %   warp-inst {i:12d} {100 * i / max(ti, 1):5.1f} %")
