# N = 2 sanity after host-layer changes: the multi-GPU check script and the bench as the driver launches it
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu_check.py > gpurun_out/mg2s_check.log 2>&1; echo "check rc=$?"
tail -4 gpurun_out/mg2s_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-extras > gpurun_out/mg2s_bench.json 2> gpurun_out/mg2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
line = [l for l in open('gpurun_out/mg2s_bench.json') if l.startswith('{')][-1]
d = json.loads(line)
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'collective', d.get('collective'), 'parity', (d.get('parity') or {}).get('ok'))
PY
