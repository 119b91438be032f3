"""Joint fits over many observations sharded across GPUs: one process per GPU, datasets never move.

The reference's precedent is ``DatasetsActor`` (datasets/actors.py:16-241): free parameter values go to every
Ray actor, each returns ``stat_sum_likelihood()`` and the driver sums.  Here every rank evaluates its shard with
one fused launch, leaves the per-vector Cash sums on the device, and a single NCCL all-reduce (fp64, count =
number of parameter vectors) over NVLink produces the joint statistic on every rank -- one collective per
evaluation, never per dataset.
"""
import numpy as np

import os

from .modeling.parameter import bind_mirror, release_mirror

__all__ = ["partition", "allreduce_stat", "ShardedDatasets", "PeerReduce"]

PEER_REDUCE_CAPACITY = 1024  # parameter vectors per evaluation the NVLink one-shot reduce is sized for


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


class PeerReduce:
    """One-shot all-reduce of the joint statistic over NVLink peer memory (``gpy_peer_reduce_init``): every rank's
    buffer is symmetric memory mapped by all ranks (``torch.distributed._symmetric_memory``), the library's last
    kernel of an evaluation stores this rank's sums into every peer and adds the peers' in rank order.

    ``PeerReduce.for_context(ctx)`` returns the per-context instance or ``None`` when the process group has one rank,
    symmetric memory is unavailable (then NCCL does the all-reduce) or ``GPY_PEER_REDUCE=0``.  Collective: all ranks
    must call it at the same point."""

    _by_ctx = {}

    def __init__(self, ctx, handle, tensor):
        self.ctx, self.handle, self.tensor = ctx, handle, tensor

    @classmethod
    def for_context(cls, ctx):
        import ctypes as C

        import torch
        import torch.distributed as dist

        key = id(ctx)
        if key in cls._by_ctx:
            return cls._by_ctx[key]
        inst = None
        ok = (dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
              and os.environ.get("GPY_PEER_REDUCE", "1") != "0")
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                world, rank = dist.get_world_size(), dist.get_rank()
                nbytes = int(ctx.lib.gpy_peer_reduce_bytes(world, PEER_REDUCE_CAPACITY))
                t = symm_mem.empty(nbytes // 8, dtype=torch.float64, device=ctx.torch_device)
                t.zero_()
                hdl = symm_mem.rendezvous(t, dist.group.WORLD)
                torch.cuda.synchronize()
                dist.barrier()  # every rank's buffer is zero before anybody's first evaluation
                ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
                from . import _lib

                _lib.check(ctx.lib.gpy_peer_reduce_init(ctx.handle, ptrs, rank, world, PEER_REDUCE_CAPACITY))
                inst = cls(ctx, hdl, t)
            except Exception as exc:  # no symmetric memory on this system: NCCL path
                if os.environ.get("GPY_PEER_REDUCE_DEBUG"):
                    print(f"[gammapy_b200] peer reduce unavailable: {exc!r}")
                inst = None
            # all ranks must agree (a rank on the NCCL path would wait for a collective the others never issue)
            flag = torch.tensor([1 if inst is not None else 0], device=ctx.torch_device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                inst = None
        cls._by_ctx[key] = inst
        return inst


class _JointEngine:
    """What ``Fit`` needs from an engine (``parameters``, ``values``, ``stat_sum``), over the datasets of ALL ranks.

    The column order is the *global* one: the parameters of ``global_models`` in order, identical on every rank.
    A rank's own engine knows a subset of them (the shared sources + its own datasets' background models);
    ``stat_sum`` maps the global vectors to the local columns, evaluates the shard and all-reduces."""

    def __init__(self, owner, global_models):
        self.owner = owner
        self.parameters, seen = [], set()
        for model in global_models:
            for par in model.parameters:
                if id(par) not in seen:
                    seen.add(id(par))
                    self.parameters.append(par)
        self.n_par = len(self.parameters)
        self._local_cols = None
        self._vec = bind_mirror(self.parameters)

    def values(self):
        return self._vec.copy()

    def __del__(self):
        try:
            release_mirror(self.parameters, self._vec)
        except Exception:
            pass

    def local_columns(self):
        """Global column of every local parameter-vector column."""
        if self._local_cols is None:
            where = {id(p): i for i, p in enumerate(self.parameters)}
            local = self.owner._local_parameters()
            missing = [p.name for p in local if id(p) not in where]
            if missing:
                raise ValueError(f"local parameters missing from the global model list: {missing}")
            self._local_cols = np.array([where[id(p)] for p in local], dtype=int)
        return self._local_cols

    def stat_sum(self, values=None):
        if values is None:
            values = self.values()
        values = np.ascontiguousarray(values, dtype=float)
        v2 = values.reshape(1, -1) if values.ndim == 1 else values
        if v2.shape[1] != self.n_par:
            raise ValueError(f"expected {self.n_par} parameter values per vector, got {v2.shape[1]}")
        out = self.owner._joint_stat(np.ascontiguousarray(v2[:, self.local_columns()]))
        return float(out[0]) if values.ndim == 1 else out


class ShardedDatasets:
    """The local shard of a joint ``Datasets`` plus the collective that completes ``stat_sum``.

    Every rank must hold the *same models in the same order* for the shared sources (the parameter vector is
    replicated, as the reference replicates ``models.parameters.free_parameters.value`` to its actors);
    per-dataset background models live with their dataset.

    ``global_models``: the models of ALL datasets of the joint fit in one order that is the same on every rank
    (the shared sources, then every dataset's background model -- for datasets on other ranks a model object
    without data).  With it the object is a drop-in for ``Datasets`` in ``Fit.run / optimize / covariance /
    stat_profile``: every rank runs the same optimiser on the same all-reduced numbers and ends with the same
    parameters.  Without it only ``stat_sum_batch`` over the local column order is available.

    ``local_stat`` is pluggable so the host-side logic can be exercised on CPU with gloo: it maps a
    ``[B, n_local]`` array to the ``[B]`` local statistic; ``local_parameters`` then names its columns.
    """

    def __init__(self, local_datasets, local_stat=None, global_models=None, local_parameters=None):
        import torch.distributed as dist

        self.datasets = local_datasets
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1
        self._local_stat = local_stat
        self._local_pars = local_parameters
        self._engine = None
        self._buf = None
        self._peer = False       # PeerReduce, None (NCCL path) or False (not decided yet: first GPU evaluation)
        self._joint_out = None
        self._joint = _JointEngine(self, global_models) if global_models is not None else None

    # ---- Datasets protocol used by Fit -------------------------------------------------------------------
    @property
    def parameters(self):
        from .modeling.parameter import Parameters

        if self._joint is None:
            raise ValueError("ShardedDatasets needs global_models to expose the joint parameter list")
        return Parameters(self._joint.parameters)

    def _get_engine(self):
        if self._joint is not None:
            return self._joint
        return self._local_engine()

    def _local_engine(self):
        if self._engine is None:
            self._engine = self.datasets._get_engine()
        return self._engine

    def _local_parameters(self):
        if self._local_stat is not None:
            if self._local_pars is None:
                raise ValueError("local_stat needs local_parameters to name its columns")
            return list(self._local_pars)
        return list(self._local_engine().parameters)

    # ---- evaluation ------------------------------------------------------------------------------------------
    def _joint_stat(self, v2):
        """[B, n_local] local-order vectors -> [B] joint statistic (all ranks call this together)."""
        import torch

        B = v2.shape[0]
        if self._local_stat is not None:  # CPU / gloo path of the host logic
            t = torch.from_numpy(np.asarray(self._local_stat(v2), dtype=float).reshape(B).copy())
            allreduce_stat(t)
            return t.numpy()
        eng = self._local_engine()
        D = len(eng.datasets)
        if self._peer is False:
            self._peer = PeerReduce.for_context(eng.ctx)
        if self._peer is not None and B <= PEER_REDUCE_CAPACITY:
            return eng.stat_sum_joint(v2)  # evaluation + all-reduce + copy + synchronisation inside the library
        # (NCCL path) Device and pinned buffers are allocated once at a capacity (a fit alternates between batch sizes 1, 5, 11,
        # 2 n + 1 ...; pinning memory costs milliseconds) and sliced; they are created on the library's stream, the
        # one every later write to them is ordered on.
        if self._buf is None or self._buf.shape[0] < B or self._buf.shape[1] != D:
            cap = max(B, 64, 2 * (self._buf.shape[0] if self._buf is not None else 0))
            with torch.cuda.stream(eng.ctx.stream):
                self._buf = torch.empty(cap, D, dtype=torch.float64, device=eng.ctx.torch_device)
            self._host = torch.empty(cap, dtype=torch.float64).pin_memory()
        buf, host = self._buf[:B], self._host[:B]
        with torch.cuda.stream(eng.ctx.stream):
            total = self.joint_stat_device(v2, buf)
            host.copy_(total, non_blocking=True)
        eng.ctx.stream.synchronize()
        return host.numpy().copy()

    def joint_stat_device(self, values, scratch=None):
        """Asynchronous joint statistic of ``values`` ([n_local] or [B, n_local]): returns a CUDA tensor [B] that holds,
        in the library's stream order, the sum over the datasets of all ranks.  With NVLink peer memory the library's
        own last kernel does the all-reduce (``gpy_stat_sum_joint_device``); otherwise one NCCL all-reduce follows
        the evaluation on the same stream.  ``scratch``: optional CUDA float64 buffer [>= B, n_local_datasets]."""
        import torch

        eng = self._local_engine()
        v2 = np.ascontiguousarray(values, dtype=float).reshape(-1, eng.n_par)
        B, D = v2.shape[0], len(eng.datasets)
        if self._peer is False:
            self._peer = PeerReduce.for_context(eng.ctx)
        if self._peer is not None and B <= PEER_REDUCE_CAPACITY:
            if self._joint_out is None or self._joint_out.shape[0] < B:
                with torch.cuda.stream(eng.ctx.stream):
                    self._joint_out = torch.empty(max(B, 64), dtype=torch.float64, device=eng.ctx.torch_device)
            out = self._joint_out[:B]
            eng.stat_sum_joint_device(v2, out)
            return out
        if scratch is None or scratch.shape[0] < B or scratch.shape[1] != D:
            with torch.cuda.stream(eng.ctx.stream):
                scratch = torch.empty(B, D, dtype=torch.float64, device=eng.ctx.torch_device)
        buf = scratch[:B]
        eng.stat_sum_device(v2, buf)
        with torch.cuda.stream(eng.ctx.stream):
            total = buf[:, 0].contiguous() if D == 1 else buf.sum(dim=1)
            allreduce_stat(total, None)  # NCCL enqueues behind the kernels of the current (library) stream
        return total

    def stat_sum_batch(self, values):
        """Joint statistic of ``values`` ([n] or [B, n]) over the datasets of ALL ranks; columns in the global
        order when ``global_models`` was given, else in the local engine's order."""
        values = np.ascontiguousarray(values, dtype=float)
        if self._joint is not None:  # likelihood of all ranks + the priors, as stat_sum()
            out = self._joint.stat_sum(values)
            if any(p.prior is not None for p in self._joint.parameters):
                vals = np.atleast_2d(values)
                prior = np.zeros(vals.shape[0])
                for col, par in enumerate(self._joint.parameters):
                    if par.prior is not None:
                        prior += par.prior.evaluate_many(vals[:, col])
                out = out + (float(prior[0]) if np.ndim(out) == 0 else prior)
            return out
        v2 = values.reshape(1, -1) if values.ndim == 1 else values
        out = self._joint_stat(v2)
        return float(out[0]) if values.ndim == 1 else out

    def stat_sum(self):
        """``Datasets.stat_sum`` of the joint fit with the current Parameter values."""
        if self._joint is not None:  # likelihood of all ranks + the priors (evaluated identically on every rank)
            return self._joint.stat_sum() + sum(p.prior(p) for p in self._joint.parameters if p.prior is not None)
        if self._local_stat is not None:
            raise ValueError("stat_sum() needs GPU datasets; use stat_sum_batch(values) with local_stat")
        return self.stat_sum_batch(self._local_engine().values())
