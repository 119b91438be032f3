mkdir -p gpurun_out; rm -f gpurun_out/t64_wl.log
for wl in c3-flat c5-one c4-large c2 c1; do
  for t in 0 1; do
    GPY_FOLD_T64=$t timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline > gpurun_out/t64_wl.json 2> gpurun_out/t64_wl.err
    python - "$wl T64=$t" >> gpurun_out/t64_wl.log 2>&1 <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/t64_wl.json'))
    print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'], 4), 'kernel_us', round(d['roofline']['duration_us'], 1), 'frac', round(d['roofline']['frac'], 3))
except Exception as e:
    print(sys.argv[1], 'FAILED', repr(e)[:100])
PY
  done
done
cat gpurun_out/t64_wl.log
