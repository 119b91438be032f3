# A/B of fold-kernel builds: default build = parity tests + bench; then each GPY_NVCC_EXTRA variant in VARIANTS (";"-separated) = bench only
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
run_bench() {
  timeout 200 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - "$1" >> gpurun_out/ab.log 2>&1 <<'PY'
import json, sys
try:
    d = json.load(open(f'gpurun_out/ab_{sys.argv[1]}.json'))
    print(sys.argv[1], 'step_us', round(d['ms_per_step'] * 1e3, 1), 'fold_us', round(d['roofline']['duration_us'], 1), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity'))
except Exception as e:
    print(sys.argv[1], 'FAILED', repr(e)[:200])
PY
}
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q ${TESTSEL} > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab.log
  tail -3 gpurun_out/ab_tests.log >> gpurun_out/ab.log
fi
run_bench default
i=0
IFS=';' read -ra VARS <<< "$VARIANTS"
for v in "${VARS[@]}"; do
  i=$((i+1))
  GPY_NVCC_EXTRA="$v" python -m gammapy_b200._build --force > /dev/null 2>&1
  echo "variant $i: $v" >> gpurun_out/ab.log
  run_bench v$i
done
cat gpurun_out/ab.log
