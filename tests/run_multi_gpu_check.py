"""Multi-rank check of the joint statistic on real GPUs (not collected by pytest: needs one GPU per rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_multi_gpu_check.py

Every rank evaluates its own observation; the joint statistic from ``ShardedDatasets.joint_stat_device`` (NVLink
one-shot reduce inside the library, or NCCL when symmetric memory is unavailable) must equal the rank-ordered sum of
the per-rank statistics gathered with NCCL, bit for bit, on every rank, for single vectors and for a batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    from gammapy_b200 import Datasets, configs
    from gammapy_b200.distributed import ShardedDatasets

    rng = np.random.RandomState(1000 + rank)
    ds = configs.make_c3(name=f"obs-{rank}", width=(8.0, 4.0), n_sources=6, exposure_scale=float(rng.uniform(0.5, 1.5)))
    ds.fake(random_state=2 + rank)
    datasets = Datasets([ds])
    sharded = ShardedDatasets(datasets)
    eng = sharded._local_engine()
    base = eng.values()
    cols = [i for i, p in enumerate(eng.parameters) if p.name in ("index", "alpha")]
    ok = True
    for trial in range(6):
        B = 1 if trial < 4 else 5
        grid = np.tile(base, (B, 1))
        grid[:, cols] += 0.01 * (trial + 1) * np.linspace(1.0, 2.0, B)[:, None]
        local_stat = np.atleast_1d(eng.stat_sum(grid if B > 1 else grid[0]))
        gathered = [torch.zeros(B, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(local_stat, dtype=torch.float64, device="cuda"))
        want = np.zeros(B)
        for t in gathered:  # rank order
            want = want + t.cpu().numpy()
        got = sharded.joint_stat_device(grid)
        eng.ctx.stream.synchronize()  # the result is ordered on the library's stream, not on torch's current one
        got = got.cpu().numpy()
        same = np.array_equal(got, want)
        ok &= same
        if rank == 0:
            print(f"trial {trial} B={B}: joint {got[0]!r} want {want[0]!r} equal={same}", flush=True)
    peer = sharded._peer is not None
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world}: collective = {'nvlink peer one-shot' if peer else 'nccl'}; all ranks equal: {bool(flag.item())}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
