# bench.py as the driver launches it at N GPUs (NG=4 by default), final code
NG=${NG:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 50 --warmup 5 --no-extras > gpurun_out/mgN_bench.json 2> gpurun_out/mgN_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
line = [l for l in open('gpurun_out/mgN_bench.json') if l.startswith('{')][-1]
d = json.loads(line)
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'collective', d.get('collective'))
PY
