# Last GPU call of the round: the GPU suite, smoke(), then the ncu captures of the shipped code (profiles/capture_r02.sh)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/finale_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/finale_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 bash profiles/capture_r02.sh; echo "capture rc=$?"
ls -la gpurun_out/*.ncu-rep
