mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline_parity.py -x -q -s -m gpu > gpurun_out/mma_parity.log 2>&1; echo "parity rc=$?" 
tail -15 gpurun_out/mma_parity.log
for w in 11 15; do
  GPY_MMA_WARPS=$w timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/mma_bench_w$w.json 2> gpurun_out/mma_bench_w$w.err; echo "bench w=$w rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/mma_bench_w$w.json'));print(d['ms_per_step'], d['roofline']['duration_us'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
done
GPY_FOLD_KERNEL=1 timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/mma_bench_old.json 2> gpurun_out/mma_bench_old.err; echo "bench old rc=$?"
python -c "import json;d=json.load(open('gpurun_out/mma_bench_old.json'));print(d['ms_per_step'], d['roofline']['duration_us'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
