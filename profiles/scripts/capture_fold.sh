# ncu --set full of the fold kernel (5th launch) on the default bench workload; TAG names the report
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fold_cash_t -s 4 -c 1 -f -o gpurun_out/fold_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
