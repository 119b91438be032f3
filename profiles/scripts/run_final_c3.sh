# GPU suite, then the default bench run (the line the driver records) and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err; echo "c3 rc=$? $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/final_c3.json'))
print('value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'fold_us', d['roofline']['duration_us'], 'parity', d.get('parity'), 'launches', d.get('gpu_launches'))
print('cpu_baseline', d.get('cpu_baseline'))
ex = d.get('extras', {})
for k, v in ex.items():
    print(k, v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if not isinstance(vv, (dict, list))})
PY
