# Per-phase / per-item-class clock64 buckets of the fold kernel (GPY_FOLD_TIMING build; not a bench number)
mkdir -p gpurun_out
GPY_NVCC_EXTRA=-DGPY_FOLD_TIMING python -m gammapy_b200._build --force > /dev/null 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/timing_run.log 2>&1
grep "fold classes\|fold timing" gpurun_out/timing_run.log | tail -6 > gpurun_out/fold_timing.txt
cat gpurun_out/fold_timing.txt
