# parity + A/B of the 64-pixel sub-tile kernel (GPY_FOLD_T64=1); every step under its own timeout
mkdir -p gpurun_out
GPY_FOLD_T64=${T64:-1} timeout 300 python -m pytest tests/test_gpu_headline_parity.py -x -q -m gpu > gpurun_out/t64_tests_a.log 2>&1; echo "headline rc=$?"
tail -5 gpurun_out/t64_tests_a.log
GPY_FOLD_T64=${T64:-1} timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py tests/test_gpu_fullsize.py tests/test_gpu_fit.py -x -q -m gpu > gpurun_out/t64_tests_b.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/t64_tests_b.log
CONFIGS="GPY_FOLD_T64=0 GPY_FOLD_T64=1 GPY_FOLD_T64=2" bash profiles/micro/run_env.sh
