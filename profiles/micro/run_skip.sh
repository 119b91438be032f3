mkdir -p gpurun_out
rm -f gpurun_out/skip_all.log
run() {
  env "$@" timeout 60 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/skip.json 2> gpurun_out/skip.err
  python - "$*" >> gpurun_out/skip_all.log 2>&1 <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/skip.json'))
    print(sys.argv[1], d['ms_per_step'], d['roofline']['duration_us'])
except Exception as e:
    print(sys.argv[1], 'FAILED', repr(e)[:80])
PY
}
for s in ${SKIPS:-0}; do run GPY_MMA_WARPS=11 GPY_MMA_SKIP=$s; done
cat gpurun_out/skip_all.log
