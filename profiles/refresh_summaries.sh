#!/bin/sh
# Rebuild the tracked summaries from the scratch captures in gpurun_out/ (run in the repo root, no GPU needed):
#   sh profiles/refresh_summaries.sh <tag>      e.g. r1j -> gpurun_out/fold_<tag>.ncu-rep, gpurun_out/launches_<tag>.csv
set -e
tag=$1
(
  echo "# ncu --set full --clock-control none --import-source on, fold_cash_tma_kernel as shipped at the end of round 1 (C3, B=1)"
  echo "# command: python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline, 5th launch of the kernel (capture $tag)"
  python profiles/ncu_summary.py gpurun_out/fold_$tag.ncu-rep
  python profiles/ncu_lines.py gpurun_out/fold_$tag.ncu-rep 30
  echo "# top SASS instructions by stall samples"
  python profiles/ncu_sass.py gpurun_out/fold_$tag.ncu-rep 16
) > profiles/r01_fold_final_ncu_summary.txt
cp gpurun_out/launches_$tag.csv profiles/r01_launches_final.csv
python profiles/launch_summary.py profiles/r01_launches_final.csv > profiles/r01_launches_final_summary.txt
