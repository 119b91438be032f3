# N-GPU check of the joint statistic + the driver's scaling bench command (N from $NG)
set -x
mkdir -p gpurun_out
NG=${NG:-8}
export GPY_PEER_REDUCE_DEBUG=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu_check.py > gpurun_out/mg${NG}_check.log 2>&1; echo "check rc=$?" >> gpurun_out/mg${NG}_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/mg${NG}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/mg${NG}_bench.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 --impl reference > gpurun_out/mg${NG}_ref.log 2>&1; echo "rc=$?" >> gpurun_out/mg${NG}_ref.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --steps 50 --warmup 5 --workload c4 --no-extras > gpurun_out/mg${NG}_c4.log 2>&1; echo "rc=$?" >> gpurun_out/mg${NG}_c4.log
tail -3 gpurun_out/mg${NG}_check.log; tail -c 300 gpurun_out/mg${NG}_bench.log
