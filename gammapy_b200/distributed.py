"""Joint fits over many observations sharded across GPUs: one process per GPU, datasets never move.

The reference's precedent is ``DatasetsActor`` (datasets/actors.py:16-241): free parameter values go to every
Ray actor, each returns ``stat_sum_likelihood()`` and the driver sums.  Here every rank evaluates its shard with
one fused launch, leaves the per-vector Cash sums on the device, and a single NCCL all-reduce (fp64, count =
number of parameter vectors) over NVLink produces the joint statistic on every rank -- one collective per
evaluation, never per dataset.
"""
import numpy as np

__all__ = ["partition", "allreduce_stat", "ShardedDatasets"]


def partition(sizes, world_size):
    """Size-balanced assignment of datasets to ranks (longest-processing-time greedy); ties keep input order.

    ``sizes`` are the per-dataset byte counts ``(4 nT + 9 nR) P``.  Returns a list of index lists, one per rank."""
    sizes = list(sizes)
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += sizes[i]
    return [sorted(s) for s in shards]


def dataset_bytes(ds):
    nR, ny, nx = ds.data_shape
    nT = ds.exposure.data.shape[0]
    return (4 * nT + 9 * nR) * ny * nx


def allreduce_stat(tensor, stream=None):
    """Sum the per-rank statistics in place; the collective is ordered after the kernels on ``stream``."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensor
    if tensor.is_cuda and stream is not None:
        with torch.cuda.stream(stream):
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    else:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


class ShardedDatasets:
    """The local shard of a joint ``Datasets`` plus the collective that completes ``stat_sum``.

    Every rank must hold the *same models in the same order* for the shared sources (the parameter vector is
    replicated, as the reference replicates ``models.parameters.free_parameters.value`` to its actors);
    per-dataset background models live with their dataset and are appended after the shared block.

    ``local_stat`` is pluggable so the host-side logic can be exercised on CPU with gloo: it maps a
    ``[B, n_par]`` array to the ``[B]`` local statistic.
    """

    def __init__(self, local_datasets, local_stat=None):
        import torch.distributed as dist

        self.datasets = local_datasets
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1
        self._local_stat = local_stat
        self._engine = None
        self._buf = None

    def _get_engine(self):
        if self._engine is None:
            self._engine = self.datasets._get_engine()
        return self._engine

    def stat_sum_batch(self, values):
        """Joint statistic of ``values`` ([n_par] or [B, n_par]) over the datasets of ALL ranks."""
        import torch

        values = np.ascontiguousarray(values, dtype=float)
        v2 = values.reshape(1, -1) if values.ndim == 1 else values
        B = v2.shape[0]
        if self._local_stat is not None:  # CPU / gloo path of the host logic
            t = torch.from_numpy(np.asarray(self._local_stat(v2), dtype=float).reshape(B).copy())
            allreduce_stat(t)
            out = t.numpy()
        else:
            eng = self._get_engine()
            D = len(eng.datasets)
            if self._buf is None or self._buf.shape != (B, D):
                self._buf = torch.zeros(B, D, dtype=torch.float64, device=eng.ctx.torch_device)
                self._host = torch.zeros(B, dtype=torch.float64).pin_memory()
            eng.stat_sum_device(v2, self._buf)
            with torch.cuda.stream(eng.ctx.stream):
                total = self._buf[:, 0] if D == 1 else self._buf.sum(dim=1)
                allreduce_stat(total, None)  # NCCL enqueues behind the kernels of the current (library) stream
                self._host.copy_(total, non_blocking=True)
            eng.ctx.stream.synchronize()
            out = self._host.numpy().copy()
        return float(out[0]) if values.ndim == 1 else out

    def stat_sum(self):
        """``Datasets.stat_sum`` of the joint fit with the current local Parameter values."""
        if self._local_stat is not None:
            raise ValueError("stat_sum() needs GPU datasets; use stat_sum_batch(values) with local_stat")
        return self.stat_sum_batch(self._get_engine().values())
