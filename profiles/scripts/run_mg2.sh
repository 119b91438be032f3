set -x
mkdir -p gpurun_out
export GPY_PEER_REDUCE_DEBUG=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu_check.py > gpurun_out/mg2_check.log 2>&1; echo "check rc=$?" >> gpurun_out/mg2_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/mg2_bench.log 2>&1; echo "rc=$?" >> gpurun_out/mg2_bench.log
GPY_PEER_REDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/mg2_bench_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/mg2_bench_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --workload c4 > gpurun_out/mg2_c4.log 2>&1; echo "rc=$?" >> gpurun_out/mg2_c4.log
tail -3 gpurun_out/mg2_check.log; tail -c 600 gpurun_out/mg2_bench.log
