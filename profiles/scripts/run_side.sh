# GPU suite with the side streams on (default), then the many-sources-moved timings with them on and off, then the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/side_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/side.log
tail -3 gpurun_out/side_tests.log >> gpurun_out/side.log
for s in 1 0 1 0; do
  GPY_SIDE_STREAMS=$s timeout 300 python profiles/scripts/side_streams.py >> gpurun_out/side.log 2>&1
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/side_bench.json 2> gpurun_out/side_bench.err
python - >> gpurun_out/side.log 2>&1 <<'PY'
import json
d = json.load(open('gpurun_out/side_bench.json'))
print('step_us', round(d['ms_per_step'] * 1e3, 1), 'e2e', d['e2e']['value'], 'parity', d.get('parity'))
ex = d.get('extras', {})
for k in ('ms_per_eval_one_source_moved', 'ms_per_eval_one_source_moved_median', 'batched_scan', 'fd_gradient'):
    print(k, ex.get(k) if k in ex else {kk: vv.get(k) for kk, vv in ex.items() if isinstance(vv, dict) and k in vv})
PY
cat gpurun_out/side.log
