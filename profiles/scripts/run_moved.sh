# GPU suite, then the many-sources-moved timings with batched spatial launches on and off, then the bench with extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/moved_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/moved.log
tail -3 gpurun_out/moved_tests.log >> gpurun_out/moved.log
for s in 1 0 1 0; do
  GPY_SPATIAL_BATCH=$s timeout 300 python profiles/scripts/moved_many.py >> gpurun_out/moved.log 2>&1
done
timeout 300 python profiles/scripts/moved_source.py >> gpurun_out/moved.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/moved_bench.json 2> gpurun_out/moved_bench.err
python - >> gpurun_out/moved.log 2>&1 <<'PY'
import json
d = json.load(open('gpurun_out/moved_bench.json'))
print('step_us', round(d['ms_per_step'] * 1e3, 1), 'e2e', d['e2e']['value'], 'launches', d.get('gpu_launches'))
ex = d.get('extras', {})
for k in ('ms_per_eval_one_source_moved', 'ms_per_eval_one_source_moved_median', 'batched_scan', 'fd_gradient'):
    print(k, ex.get(k))
PY
cat gpurun_out/moved.log
