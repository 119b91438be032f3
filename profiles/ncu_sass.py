"""Top SASS instructions by stall samples of an .ncu-rep captured with --import-source on (with their CUDA line)."""
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
i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
cur, src, allr = None, "", []
for r in rows[3:]:
    if len(r) <= i_inst:
        continue
    if r[2] == "-":
        cur, src = r[0], r[1].strip()
        continue
    try:
        allr.append((int(r[i_samp]), cur, r[3].strip()[:60], int(r[i_inst]), src[:70]))
    except ValueError:
        pass
total = sum(x[0] for x in allr)
print(f"total samples {total}")
for s, line, sass, n, src in sorted(allr, key=lambda x: -x[0])[:top]:
    print(f"{s:6d} {100 * s / total:5.1f}% | x{n:<8d} | {sass:60s} | {line}: {src}")
