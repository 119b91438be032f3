#!/bin/sh
# Rebuild the tracked round-2 summaries from the scratch captures of profiles/capture_r02.sh (run in the repo root, no GPU):
#   gpurun_out/fold_r02.ncu-rep, gpurun_out/psf_r02.ncu-rep, gpurun_out/launches_r02.csv
set -e
(
  echo "# ncu --set full --clock-control none --import-source on, fold_cash_t64_kernel<0,1,0,1> as shipped at the end of round 2 (C3, B = 1)"
  echo "# command: python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline, 5th launch of the kernel (profiles/capture_r02.sh)"
  python profiles/ncu_summary.py gpurun_out/fold_r02.ncu-rep
  python profiles/ncu_lines.py gpurun_out/fold_r02.ncu-rep 30
  echo "# top SASS instructions by stall samples"
  python profiles/ncu_sass.py gpurun_out/fold_r02.ncu-rep 16
) > profiles/r02_fold_ncu_summary.txt
(
  echo "# ncu --set full --clock-control none --import-source on, psf_conv_batch_kernel (K2) as shipped at the end of round 2: the PSF stencil of all 50 sources"
  echo "# of C3 in one launch (cutouts ~ 150 x 150 px, 60 planes each, kernel half widths 33 .. 11 px), command: python bench.py --steps 3 --warmup 3"
  echo "# --no-extras --no-cpu-baseline, first launch (profiles/capture_r02.sh)"
  python profiles/ncu_summary.py gpurun_out/psf_r02.ncu-rep
  python profiles/ncu_lines.py gpurun_out/psf_r02.ncu-rep 12
) > profiles/r02_psf_conv_ncu_summary.txt
# (the hand-written analysis lines of r02_psf_conv_ncu_summary.txt -- launch shape, fp32 FMA rate -- are re-inserted by hand)
cp gpurun_out/launches_r02.csv profiles/r02_launches.csv
python profiles/launch_summary.py profiles/r02_launches.csv > profiles/r02_launches_summary.txt
