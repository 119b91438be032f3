"""Host-side logic of the sharded joint fit on CPU: size-balanced partition + one all-reduce per evaluation
(gloo, world_size 2), mirroring test_map_fit_ray's "same numbers as the serial fit" (datasets/tests/test_map.py:891-954)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
from numpy.testing import assert_allclose

from gammapy_b200.distributed import ShardedDatasets, partition


def test_partition_balanced_and_complete():
    sizes = [510, 510, 510, 510, 7, 7, 7, 7, 100]
    shards = partition(sizes, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 100
    assert partition([1] * 64, 8) == [list(range(r, 64, 8)) for r in range(8)]
    assert partition([5, 3], 4)[2:] == [[], []]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dataset_stat(i, values):
    """Deterministic stand-in for one dataset's Cash sum as a function of the shared parameter vector."""
    w = np.sin(np.arange(values.shape[1]) + i)
    return (values * w).sum(axis=1) ** 2 + i


def _worker(rank, world, port, n_datasets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = partition([1 + (i % 3) for i in range(n_datasets)], world)[rank]

        def local_stat(values):
            total = np.zeros(values.shape[0])
            for i in mine:
                total += _dataset_stat(i, values)
            return total

        sharded = ShardedDatasets(local_datasets=None, local_stat=local_stat)
        values = np.linspace(0.5, 1.5, 5 * 3).reshape(5, 3)
        got = sharded.stat_sum_batch(values)
        single = sharded.stat_sum_batch(values[2])
        out.put((rank, got.tolist(), single))
    finally:
        dist.destroy_process_group()


def test_sharded_stat_sum_gloo_world2():
    world, n_datasets = 2, 7
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_datasets, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    values = np.linspace(0.5, 1.5, 5 * 3).reshape(5, 3)
    want = sum(_dataset_stat(i, values) for i in range(n_datasets))
    for rank, got, single in results:
        assert_allclose(got, want, rtol=1e-13)  # every rank holds the joint statistic
        assert_allclose(single, want[2], rtol=1e-13)
