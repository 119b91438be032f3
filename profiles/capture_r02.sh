# Round-2 captures of the shipped kernels (run under gpurun from the repo root):
#   plain run first (must exit 0), then the launch list, then --set full of the fold kernel and of the PSF stencil.
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/plain_r02.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_l_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fold_cash_t -s 4 -c 1 -f -o gpurun_out/fold_r02 $CMD > gpurun_out/ncu_f_r02.log 2>&1
# (the stencil of all 50 sources is ONE launch of psf_conv_batch_kernel since the grouped spatial launches)
ncu --set full --clock-control none --import-source on -k regex:psf_conv_batch -s 0 -c 1 -f -o gpurun_out/psf_r02 $CMD > gpurun_out/ncu_p_r02.log 2>&1
