# usage: CONFIGS="A=1,B=2 C=3" bash profiles/micro/run_env.sh   (comma-separated env assignments per run)
mkdir -p gpurun_out
rm -f gpurun_out/env_all.log
for cfg in $CONFIGS; do
  envs=$(echo "$cfg" | tr ',' ' ')
  env $envs timeout 60 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/env.json 2> gpurun_out/env.err
  python - "$cfg" >> gpurun_out/env_all.log 2>&1 <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/env.json'))
    print(sys.argv[1], round(d['ms_per_step'] * 1e3, 1), round(d['roofline']['duration_us'], 1))
except Exception as e:
    print(sys.argv[1], 'FAILED', repr(e)[:80])
PY
done
cat gpurun_out/env_all.log
