"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        rows.setdefault(x["Kernel Name"].split("(")[0], []).append(float(x["Metric Value"].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in rows.values())
print("# ncu launch list (gpu__time_duration.sum, --clock-control none): python bench.py --steps 3 --warmup 3 "
      "--no-extras --no-cpu-baseline (C3); cold-cache, serialised")
print("# includes the one-off cold evaluation (spatial_fill + psf_conv launches for the 50 sources, twice: fake() + first stat)")
for k, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:60]:60s} n={len(v):4d} total={sum(v):10.1f}us mean={sum(v) / len(v):9.2f}us share={100 * sum(v) / tot:5.1f}%")
steady = {k: sum(v) / len(v) for k, v in rows.items()
          if "fold_cash_tma_kernel<0" in k or "fold_cash_t64_kernel<0" in k or "reduce_stat" in k or "spectral_flux" in k}
s = sum(steady.values())
print("# steady-state step (one launch of each; prep_items_kernel only runs when a cutout or the source list changes): " +
      ", ".join(f"{k.split('::')[-1][:26]} {v:.1f}us ({100 * v / s:.1f}%)" for k, v in steady.items()))
