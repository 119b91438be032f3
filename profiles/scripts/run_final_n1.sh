# Single-GPU numbers quoted in DESIGN.md / README.md (logs kept under profiles/logs/ by the caller)
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err; echo "c3 rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/final_c3_ref.json 2> gpurun_out/final_c3_ref.err; echo "ref rc=$? $(( $(date +%s) - t0 )) s"
timeout 300 python bench.py --workload c1 --no-cpu-baseline --no-extras > gpurun_out/final_c1.json 2> gpurun_out/final_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-extras > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c4 --steps 100 > gpurun_out/final_c4_n1.json 2> gpurun_out/final_c4_n1.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c4-large --steps 20 --no-extras > gpurun_out/final_c4l_n1.json 2> gpurun_out/final_c4l_n1.err; echo "c4-large rc=$?"
timeout 600 python bench.py --workload c5-one --no-cpu-baseline --no-extras > gpurun_out/final_c5one.json 2> gpurun_out/final_c5one.err; echo "c5-one rc=$?"
timeout 900 python bench.py --workload c5 --datasets 24 --steps 10 > gpurun_out/final_c5_d24.json 2> gpurun_out/final_c5_d24.err; echo "c5 rc=$?"
