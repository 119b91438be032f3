# timing-bucket build first, then A/B of the variants
bash profiles/scripts/run_timing.sh > /dev/null 2>&1
python -m gammapy_b200._build --force > /dev/null 2>&1
SKIP_TESTS=${SKIP_TESTS-1} bash profiles/scripts/run_ab.sh
cat gpurun_out/fold_timing.txt
