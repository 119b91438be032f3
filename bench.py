#!/usr/bin/env python
"""bench.py -- npred + Cash evaluations/s on the Galactic-plane survey MapDataset (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3-flat|c3-small|c1|c2|c4|c4-large|c5-one|c5 [--datasets N]]

A *step* is one full likelihood evaluation (one parameter vector, every source's spectral parameters moved,
spatial / PSF caches warm exactly as in the reference's steady-state Minuit step) of one MapDataset per GPU.
At N > 1 every rank holds its own observation of the field (a joint fit over N observations, weak scaling) and
the per-rank Cash sums are all-reduced with NCCL inside every step.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "npred+Cash evals/s on 3D MapDataset"
UNIT = "evals/s"

WORKLOADS = {
    # BASELINE.json configs[2]: the north-star target shape (inputs 510 MB + cached PSF cubes >> 126 MB L2)
    "c3": dict(width=(40.0, 10.0), n_sources=50, n_reco=30, n_true=60,
               desc="Galactic-plane survey MapDataset 2000x500 px, 30 reco / 60 true bins, 50 sources"),
    "c3-small": dict(width=(10.0, 5.0), n_sources=12, n_reco=30, n_true=60,
                     desc="reduced survey MapDataset 500x250 px, 30 reco / 60 true bins, 12 sources (debug)"),
    # the same shape with exposure / background flat along the plane (falloff in latitude only): every low-energy bin of
    # the map holds counts, so the Cash pass takes a log for most bins
    "c3-flat": dict(width=(40.0, 10.0), n_sources=50, n_reco=30, n_true=60, falloff="latitude",
                    desc="Galactic-plane survey MapDataset 2000x500 px, 30 reco / 60 true bins, 50 sources, exposure and "
                         "background flat in longitude (dense counts)"),
    # one dataset of BASELINE.json configs[4] (CTAO-like resolution, 4000x1000 px, 2.04 GB of cubes): secondary measurement
    "c5-one": dict(width=(80.0, 20.0), n_sources=100, n_reco=30, n_true=60,
                   desc="one CTAO-resolution MapDataset 4000x1000 px, 30 reco / 60 true bins, 100 sources (one of the "
                        "stacked-fit datasets)"),
    # BASELINE.json configs[4]: stacked fit of many datasets per GPU in the compact storage format (uint16 counts, bit-packed
    # mask, one exposure map shared by the datasets of the pointing grid); --datasets N of them on this GPU
    "c5": dict(width=(80.0, 20.0), n_sources=100, n_reco=30, n_true=60,
               desc="stacked fit: N CTAO-resolution MapDatasets 4000x1000 px, 30 reco / 60 true bins, 100 shared sources, "
                    "per-dataset background norm, compact storage (uint16 counts, bit mask, shared exposure)"),
    # BASELINE.json configs[0] / configs[1]: L2-resident (6.8 MB), launch-latency bound; secondary measurements
    "c1": dict(desc="analysis_3d shape: 200x200 px, 10 reco / 20 true bins, Gaussian x PowerLaw, 1 vector per launch"),
    "c2": dict(desc="200x200 px, 10 reco / 20 true bins, 5 sources, profile scan of 256 parameter vectors per launch",
               batch=256),
    # BASELINE.json configs[3]: joint fit of 64 observations sharded over the ranks (strong scaling: 64 / n_gpus each)
    "c4": dict(desc="joint fit of 64 observations, 200x200 px, 10 reco / 20 true bins, shared Gaussian x PowerLaw source, "
                    "per-observation background norm, datasets sharded over the GPUs", n_obs=64, geometry={}),
    "c4-large": dict(desc="joint fit of 64 observations, 1000x250 px, 30 reco / 60 true bins (C3/4 size, 127 MB each), "
                          "shared Gaussian x PowerLaw source, per-observation background norm, datasets sharded over "
                          "the GPUs", n_obs=64,
                     geometry=dict(width=(20.0, 5.0), n_reco=30, n_true=60, e_reco=(0.1, 100.0), e_true=(0.03, 300.0),
                                   mask_radius=None)),
}


def config_for(workload):
    """The `config` object of the JSON line: identical in the repo arm and the reference arm (the driver compares them)."""
    w = WORKLOADS[workload]
    cfg = {"workload": w["desc"], "key": workload,
           "per_gpu": "one observation (MapDataset) per GPU; joint fit over n_gpus observations",
           "step": "one parameter vector, spectral shape parameters of all sources moved, spatial/PSF caches warm",
           "l2": "inputs per evaluation exceed the 126 MB L2: no flush needed" if workload in ("c3", "c5-one", "c3-flat")
                 else "L2-resident workload (secondary measurement, no flush)"}
    if w.get("batch", 1) > 1:
        cfg["vectors_per_launch"] = w["batch"]
    return cfg


def committed_traffic(workload):
    """DRAM bytes of one launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of this kernel on this workload (profiles/r02_fold_ncu_summary.txt; a run under ncu is
    never a bench value, so the capture is a separate, committed run of the same command).  None when there is none."""
    if workload != "c3":
        return None, None
    path = os.path.join(ROOT, "profiles", "r02_fold_ncu_summary.txt")
    try:
        total = 0.0
        with open(path) as fh:
            for line in fh:
                parts = line.split()
                if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[parts[2]]
                    total += float(parts[1]) * scale
        return (total or None), "profiles/r02_fold_ncu_summary.txt"
    except OSError:
        return None, None


def _dbg(msg):
    if os.environ.get("GPY_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def extra_steps(timed_wall, steps, distributed, device):
    """Untimed steps that keep the timed loop's load up until ~0.4 s for the 50 ms clock sampler; identical on all ranks."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([timed_wall], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    if wall >= 0.4:
        return 0
    return int(np.ceil((0.4 - wall) / max(wall / steps, 1e-6)))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--datasets", type=int, default=8, help="datasets on this GPU (workload c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def build_dataset(workload, rank=0):
    from gammapy_b200 import configs

    w = WORKLOADS[workload]
    if workload == "c1":
        return configs.make_c1(name=f"obs-{rank:03d}")
    if workload == "c2":
        return configs.make_c2(name=f"obs-{rank:03d}")
    if workload.startswith("c4"):  # reference arm: one observation of the joint fit (the CPU path loops over them)
        return configs.make_c4(n_obs=w["n_obs"], indices=[rank], **w["geometry"])[0]
    rng = np.random.RandomState(1000 + rank)
    kwargs = {}
    if rank > 0:  # other observations of the same field: exposure scale U(0.5, 1.5), pointing offset N(0, 0.5 deg)
        kwargs = dict(exposure_scale=float(rng.uniform(0.5, 1.5)), pointing_offset=tuple(rng.normal(0, 0.5, size=2)))
    if "falloff" in w:
        kwargs["falloff"] = w["falloff"]
    return configs.make_c3(name=f"obs-{rank:03d}", width=w["width"], n_sources=w["n_sources"], n_reco=w["n_reco"],
                           n_true=w["n_true"], **kwargs)


def algorithmic_bytes(ds):
    """4 nT P (exposure) + 4 nR P (background) + 4 nR P (counts) + nR P (mask) per evaluation."""
    nR, ny, nx = ds.data_shape
    nT = ds.exposure.data.shape[0]
    P = ny * nx
    return 4 * nT * P + 4 * nR * P + 4 * nR * P + 1 * nR * P


def spectral_columns(engine):
    """Columns of the spectral shape parameters that move every step (index / alpha of every source)."""
    cols = []
    for i, p in enumerate(engine.parameters):
        if p.name in ("index", "alpha") and p.type == "spectral":
            cols.append(i)
    return np.array(cols, dtype=int)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement + the reference's compiled Cython Cash on host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_evals(ds, steps, warmup, max_sources=None):
    """Steady-state evaluations of the oracle on the host.  Returns (seconds per eval, sample description, kind)."""
    from oracle import adapter, cash

    models = list(ds.models)
    de = adapter.evaluator_for(ds, conv="fft32")
    n_src = len(de.evaluators)
    used_all = True
    if max_sources is not None and n_src > max_sources:
        keep = list(de.evaluators)[:max_sources]
        de.evaluators = {k: de.evaluators[k] for k in keep}
        used_all = False
    if ds.counts is None:
        de.fake(0)
    kind = "port"  # numpy restatement of the evaluator; the Cash loop is the reference's own compiled Cython
    have_ref = cash.reference_module() is not None
    for _ in range(max(warmup, 1)):  # cold evaluation fills the spatial / PSF caches like the reference
        cpu_evals.stat_at_start = de.stat_sum() if used_all else None
    t0 = time.perf_counter()
    for k in range(steps):
        for ev in de.evaluators.values():
            spec = ev.model["spectral"]
            key = "index" if "index" in spec else "alpha"
            spec[key] += 1e-3 if k % 2 == 0 else -1e-3
        de.stat_sum()
    dt = (time.perf_counter() - t0) / steps
    used = len(de.evaluators)
    # the same with ONE source's position moved each step (spatial model + float32-FFT PSF convolution redone for it)
    evs = list(de.evaluators.values())
    t0 = time.perf_counter()
    for k in range(2):
        evs[k % len(evs)].model["spatial"]["lon_0"] += 1e-3
        de.stat_sum()
    cpu_evals.one_source_moved_s = (time.perf_counter() - t0) / 2
    sample = (f"{steps} steady-state evaluations (spatial/PSF caches warm, spectral parameters of all sources moved), "
              f"{used} of {n_src} sources, float32-FFT convolution as the reference, "
              f"Cash = {'reference Cython (oracle/_ref)' if have_ref else 'numpy restatement'}")
    del models
    return dt, sample, kind


def parity_check(ds, gpu_stat):
    """The headline configuration itself against the oracle, on the same counts and parameter values: the
    exact-arithmetic oracle npred (float64 convolution of the same float32 operands) with the Cash sum in extended
    precision is the number the 1e-3 budget is checked against; the reference-arithmetic oracle (float32 FFT, serial
    double Cash) carries its own ~3e-3 of rounding at this size (tests/test_gpu_headline_parity.py)."""
    from oracle import adapter

    de = adapter.evaluator_for(ds, conv="exact64")
    npred = de.npred()
    counts = np.asarray(ds.counts.data, dtype=float)
    mask = de.ds.mask
    mu = np.where(npred > 1e-25, npred, 1e-25)
    log_mu = np.where(npred > 1e-25, np.log(mu), np.log(np.float64(np.float32(1e-25))))
    terms = mu - np.where(counts > 0, counts * log_mu, 0.0)
    if mask is not None:
        terms = terms[np.asarray(mask, bool)]
    exact = 2.0 * float(np.sum(terms.astype(np.longdouble)))
    gpu_npred = ds.npred().data
    rel = float(np.max(np.abs(gpu_npred - npred) / np.maximum(np.abs(npred), 1e-300)))
    return {"stat_oracle_exact": exact, "abs_diff_exact": abs(gpu_stat - exact), "npred_max_rel_err": rel,
            "tolerance": {"npred_rel": 1e-5, "stat_abs": 1e-3},
            "ok": bool(rel < 1e-5 and abs(gpu_stat - exact) < 1e-3),
            "what": "stat_gpu = Datasets.stat_sum(); stat_oracle = oracle with the reference's own arithmetic (float32-FFT "
                    "convolution, Cython Cash); stat_oracle_exact = oracle with float64 convolution + extended-precision "
                    "Cash sum; npred_max_rel_err = per-bin, GPU cube vs exact oracle cube; full workload, same counts"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps or 3
    warmup = 1 if args.warmup is None else args.warmup
    ds = build_dataset(args.workload)
    dt, sample, kind = cpu_evals(ds, steps, warmup)
    value = 1.0 / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_for(args.workload),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle restatement of gammapy's arithmetic without astropy Quantity/Map overhead: a lower bound on "
                "real gammapy time; single Python thread, BLAS threads as numpy chooses",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# TS map: per-pixel norm fits of a C3-sized image, GPU launch vs the oracle on the host cores
# ------------------------------------------------------------------------------------------------------------
def ts_map_inputs(ny=500, nx=2000, k=41, seed=5):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:ny, :nx]
    g = np.exp(-0.5 * ((np.arange(k) - k // 2) / 4.0) ** 2)   # PSF-like kernel, sigma 0.08 deg at 0.02 deg pixels
    kernel = (np.outer(g, g) / np.outer(g, g).sum())[None]
    exposure = (1e3 * np.exp(-0.5 * ((yy - ny / 2) / (0.6 * ny)) ** 2))[None]
    background = (1.5 + np.cos(xx / 150.0) ** 2)[None]
    counts = rng.poisson(background).astype(float)
    for _ in range(200):  # point sources
        j, i = rng.randint(k, ny - k), rng.randint(k, nx - k)
        counts[0, j - k // 2:j + k // 2 + 1, i - k // 2:i + k // 2 + 1] += rng.poisson(
            kernel[0] * rng.uniform(20, 400))
    return counts, background, exposure, kernel


def ts_map_bench(cpu_pixels=600, with_cpu=True):
    import torch

    from gammapy_b200.estimators import compute_ts_map

    counts, background, exposure, kernel = ts_map_inputs()
    ny, nx = counts.shape[1:]
    compute_ts_map(counts[:, :64, :64], background[:, :64, :64], exposure[:, :64, :64], kernel)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = compute_ts_map(counts, background, exposure, kernel, rtol=0.01, max_niter=100)
    gpu_s = time.perf_counter() - t0
    out = {"image": f"{nx}x{ny} px, 1 plane, kernel {kernel.shape[1]}x{kernel.shape[2]}, rtol 0.01",
           "gpu_s_incl_copies": gpu_s, "gpu_pixels_per_s": ny * nx / gpu_s,
           "mean_niter": float(np.nanmean(res["niter"])), "max_ts": float(np.nanmax(res["ts"])),
           "success_fraction": float(res["success"].mean())}
    if with_cpu:
        from oracle import cash, tsmap

        rng = np.random.RandomState(0)
        pos = [(int(rng.randint(ny)), int(rng.randint(nx))) for _ in range(cpu_pixels)]
        t0 = time.perf_counter()
        worst = 0.0
        for p in pos:
            r = tsmap.ts_value(p, counts, background, exposure, kernel, rtol=0.01, max_niter=100)
            worst = max(worst, abs(r["ts"] - res["ts"][p]))
        cpu_s = time.perf_counter() - t0
        out.update({"cpu_pixels_per_s": cpu_pixels / cpu_s, "cpu_sample": f"{cpu_pixels} random pixels, 1 thread",
                    "cpu_kind": "reference" if cash.reference_module() is not None else "port",
                    "max_abs_ts_difference_on_sample": worst,
                    "speedup": (ny * nx / gpu_s) / (cpu_pixels / cpu_s)})
    return out


# ------------------------------------------------------------------------------------------------------------
# C4: joint fit of 64 observations sharded over the ranks (strong scaling), evaluation rate + fit wall time
# ------------------------------------------------------------------------------------------------------------
def run_c4(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = args.steps or 300
    warmup = 5 if args.warmup is None else max(args.warmup, 3)

    from gammapy_b200 import configs
    from gammapy_b200.distributed import ShardedDatasets, allreduce_stat, partition
    from gammapy_b200.modeling import Fit

    w = WORKLOADS[args.workload]
    n_obs = w["n_obs"]
    t_build = time.perf_counter()
    mine = partition([1] * n_obs, world)[rank]
    datasets, models = configs.make_c4(n_obs=n_obs, indices=mine, return_models=True, **w["geometry"])
    for i, ds in zip(mine, datasets):
        ds.fake(random_state=100 + i)  # independent seeds 100 + i
    sharded = ShardedDatasets(datasets, global_models=models)
    joint = sharded._get_engine()
    engine = sharded._local_engine()
    ctx, stream = engine.ctx, engine.ctx.stream
    t_build = time.perf_counter() - t_build
    base = engine.values()
    col = [i for i, p in enumerate(engine.parameters) if p.name == "index"][0]
    D = len(engine.datasets)
    buf = torch.zeros(1, D, dtype=torch.float64, device=ctx.torch_device)

    def step_device(k):
        v = base.copy()
        v[col] += 1e-3 if k % 2 == 0 else -1e-3
        if distributed:
            sharded.joint_stat_device(v, buf)
        else:
            engine.stat_sum_device(v, buf)

    for k in range(warmup):
        step_device(k)
    stream.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        ev0.record(stream)
        for k in range(steps):
            step_device(k)
        ev1.record(stream)
        stream.synchronize()
        timed_wall = time.perf_counter() - t_wall
        # every rank must run the SAME number of extra steps (each step holds a collective): the count comes from
        # the slowest rank's timed region
        extra = extra_steps(timed_wall, steps, distributed, ctx.torch_device)
        for k in range(steps, steps + extra):
            step_device(k)
            if k % 16 == 0:
                stream.synchronize()
        stream.synchronize()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_per_step = max_over_ranks(ms) / steps

    # end to end: Parameter objects -> joint stat_sum() (H2D descriptors, kernels, all-reduce, D2H), every step
    par = engine.parameters[col]
    v0 = par.value
    for k in range(3):
        sharded.stat_sum()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        par.value = v0 + (1e-3 if k % 2 == 0 else -1e-3)
        sharded.stat_sum()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / steps)
    par.value = v0

    # joint fit from a displaced start: 5 source parameters + 64 background norms free
    fit = None
    if not args.no_extras:
        start = {p: p.value for p in joint.parameters}
        for p in joint.parameters:
            if p.name == "index":
                p.value += 0.2
            elif p.name == "amplitude":
                p.value *= 1.5
            elif p.name == "norm":
                p.value *= 1.1
        if distributed:
            dist.barrier()
        t0 = time.perf_counter()
        result = Fit().run(sharded)
        fit_s = max_over_ranks(time.perf_counter() - t0)
        free = [p for p in joint.parameters if not p.frozen]
        fit = {"fit_s": fit_s, "nfev": int(result.nfev), "success": bool(result.success),
               "total_stat": float(result.total_stat), "free_parameters": len(free),
               "index": [float(free[[p.name for p in free].index("index")].value),
                         float(free[[p.name for p in free].index("index")].error)],
               "what": "Fit.run (batched BFGS + Hesse) on ShardedDatasets: every rank runs the optimiser on the "
                       "all-reduced joint statistic"}
        for p, v in start.items():
            p.value = v

    alg_bytes = sum(algorithmic_bytes(ds) for ds in datasets)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": n_obs * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "key": args.workload, "observations": n_obs,
                   "per_gpu": f"{len(mine)} observations on rank 0; one NCCL all-reduce of the joint statistic per step",
                   "l2": "per-rank inputs exceed the 126 MB L2" if alg_bytes > 2 * 126e6 else "per-rank inputs are L2-resident",
                   "step": "one joint evaluation (all 64 observations, one parameter vector), shared spectral index moved",
                   "value_is": "dataset evaluations/s = 64 x joint evaluations/s"},
        "joint_evals_per_s": 1e3 / ms_per_step,
        "gpu_launches": int(launches),
        "clocks": dict(clocks.summary(), timed_region_s=timed_wall),
        "e2e": {"value": n_obs / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(engine.n_par * 8 + D * 400), "d2h_bytes_per_step": 8,
                "what": "ShardedDatasets.stat_sum() with Parameter objects on the host"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": "fold_cash_kernel", "duration_us": ms_per_step * 1e3,
                     "note": "rank 0's shard; duration = whole step incl. the all-reduce"},
        "setup_s": t_build,
    }
    if fit is not None:
        line["extras"] = {"joint_fit": fit}
    if rank == 0:
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def run_c5(args):
    """Stacked fit of `--datasets` C5-shaped datasets on ONE GPU in the compact storage format.  Builder-run secondary
    line (the driver's line stays C3): per-dataset HBM footprint, joint evaluation time, how many fit 180 GB."""
    import torch

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("workload c5 is a single-GPU measurement (datasets shard over ranks as in c4)")
    torch.cuda.set_device(0)
    from gammapy_b200 import Datasets, Map, configs
    from gammapy_b200.modeling.models import FoVBackgroundModel

    w = WORKLOADS["c5"]
    n_ds = int(args.datasets)
    steps = args.steps or 20
    warmup = 3 if args.warmup is None else max(args.warmup, 3)
    t_build = time.perf_counter()
    first = configs.make_c3(name="stack-000", width=w["width"], n_sources=w["n_sources"], n_reco=w["n_reco"],
                            n_true=w["n_true"])
    first.fake(random_state=500)
    sources = [m for m in first.models if not isinstance(m, FoVBackgroundModel)]
    datasets = [first]
    rng = np.random.RandomState(501)
    for i in range(1, n_ds):
        name = f"stack-{i:03d}"
        ds = configs.MapDataset(name=name, exposure=first.exposure, psf=first.psf, edisp=first.edisp,
                                mask_safe=first.mask_safe, mask_fit=first.mask_fit)
        scale = np.float32(rng.uniform(0.7, 1.3))
        ds.background = Map.from_geom(first._geom, data=first.background.data * scale)
        # distinct integer counts without another 120 M Poisson draws on the host: the first cube shifted in longitude
        ds.counts = Map.from_geom(first._geom, data=np.roll(first.counts.data, 7 * i, axis=2), dtype=float)
        ds.models = sources + [FoVBackgroundModel(dataset_name=name)]
        datasets.append(ds)
    for ds in datasets:
        ds.storage = "compact"
    joint = Datasets(datasets)
    engine = joint._get_engine()
    ctx, stream = engine.ctx, engine.ctx.stream
    ref = joint.stat_sum()
    t_build = time.perf_counter() - t_build
    base = engine.values()
    cols = spectral_columns(engine)
    buf = torch.zeros(1, n_ds, dtype=torch.float64, device=ctx.torch_device)

    def step(k):
        v = base.copy()
        v[cols] += 1e-3 if k % 2 == 0 else -1e-3
        engine.stat_sum_device(v, buf)

    for k in range(warmup):
        step(k)
    stream.synchronize()
    launches0 = ctx.launch_count
    ms_blocks = []
    with ClockSampler(0) as clocks:
        for _ in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for k in range(steps):
                step(k)
            ev1.record(stream)
            stream.synchronize()
            ms_blocks.append(ev0.elapsed_time(ev1) / steps)
    launches = (ctx.launch_count - launches0) // 3
    ms_per_step = float(np.median(ms_blocks))
    ctx.kernel_timing(True)
    for k in range(5):
        step(k)
    kms, kn = ctx.kernel_time()
    ctx.kernel_timing(False)
    # end to end: Parameter objects on the host -> joint stat_sum()
    pars = [engine.parameters[c] for c in cols]
    v0 = [p.value for p in pars]
    for k in range(2):
        joint.stat_sum()
    t0 = time.perf_counter()
    for k in range(steps):
        for p, v in zip(pars, v0):
            p.value = v + (1e-3 if k % 2 == 0 else -1e-3)
        joint.stat_sum()
    e2e_s = (time.perf_counter() - t0) / steps
    for p, v in zip(pars, v0):
        p.value = v
    nR, ny, nx = first.data_shape
    nT, P = first.exposure.data.shape[0], ny * nx
    per_ds = 4 * nR * P + 2 * nR * P + nR * ((P + 127) // 128 * 16)   # background f32 + counts u16 + mask bits
    shared = 4 * nT * P
    alg = n_ds * (per_ds + shared)  # every dataset's evaluation reads the (shared) exposure once
    free, total = torch.cuda.mem_get_info()
    used = total - free
    peak, peak_src = measured_peak()
    kernel_us = kms / max(kn, 1) * 1e3
    line = {
        "metric": METRIC, "value": n_ds * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"], "key": "c5", "datasets": n_ds,
                   "l2": "inputs per evaluation exceed the 126 MB L2: no flush needed",
                   "step": "one joint evaluation (all datasets, one parameter vector), spectral shape of all sources moved",
                   "value_is": "dataset evaluations/s = datasets x joint evaluations/s"},
        "joint_evals_per_s": 1e3 / ms_per_step, "blocks_ms_per_step": ms_blocks, "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "e2e": {"value": n_ds / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(engine.n_par * 8 + n_ds * 400), "d2h_bytes_per_step": 8 * n_ds,
                "what": "Datasets.stat_sum() with Parameter objects on the host"},
        "roofline": {"bound": "hbm", "achieved": alg / (kernel_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (kernel_us * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "kernel": "fold_cash_t64_kernel", "duration_us": kernel_us,
                     "note": "compact format: 4 nT P (shared exposure, read once per dataset) + 4 nR P + 2 nR P + nR P / 8"},
        "memory": {"bytes_per_dataset": per_ds, "shared_exposure_bytes": shared, "device_bytes_in_use": int(used),
                   "datasets_that_fit_180GB": int((180e9 - shared) // (per_ds + (used - shared - n_ds * per_ds) / n_ds)),
                   "note": "in use = cubes + per-source PSF-convolved slot cubes + work descriptors of this process"},
        "joint_stat": ref, "setup_s": t_build,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def timed_blocks(step, steps, blocks, stream, distributed, device):
    """`blocks` back-to-back blocks of exactly `steps` steps, each bracketed by CUDA events on the library's stream
    (+ barrier and device synchronize on both sides of the whole series); returns the per-block milliseconds (max
    over ranks per block) and the wall-clock seconds of the series.  The reported step time is the MEDIAN block:
    with the driver's `--steps 20` one block is a 6 ms region, which a single stray host stall can double."""
    import torch
    import torch.distributed as dist

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(blocks)]
    k = 0
    t_wall = time.perf_counter()
    for e0, e1 in evs:
        e0.record(stream)
        for _ in range(steps):
            step(k)
            k += 1
        e1.record(stream)
    stream.synchronize()
    wall = time.perf_counter() - t_wall
    ms = torch.tensor([e0.elapsed_time(e1) for e0, e1 in evs], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return [float(x) for x in ms.tolist()], wall


def c4_large_extras(world, rank, steps=40):
    """BASELINE.json configs[3] inside the default run: the 64-observation joint fit (c4-large geometry) sharded over
    the ranks, strong scaling -- joint evaluations/s at this N, so that the driver's N = 1, 2, 4, 8 records hold the
    strong-scaling series (the headline `value` stays one C3 observation per GPU, weak scaling)."""
    import torch
    import torch.distributed as dist

    from gammapy_b200 import configs
    from gammapy_b200.distributed import ShardedDatasets, allreduce_stat, partition

    w = WORKLOADS["c4-large"]
    n_obs = w["n_obs"]
    distributed = world > 1
    mine = partition([1] * n_obs, world)[rank]
    datasets, models = configs.make_c4(n_obs=n_obs, indices=mine, return_models=True, **w["geometry"])
    for i, ds in zip(mine, datasets):
        ds.fake(random_state=100 + i)
    sharded = ShardedDatasets(datasets, global_models=models)
    engine = sharded._local_engine()
    ctx, stream = engine.ctx, engine.ctx.stream
    base = engine.values()
    col = [i for i, p in enumerate(engine.parameters) if p.name == "index"][0]
    D = len(engine.datasets)
    buf = torch.zeros(1, D, dtype=torch.float64, device=ctx.torch_device)

    def step(k):
        v = base.copy()
        v[col] += 1e-3 if k % 2 == 0 else -1e-3
        if distributed:
            sharded.joint_stat_device(v, buf)
        else:
            engine.stat_sum_device(v, buf)

    for k in range(5):
        step(k)
    stream.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    ms, _ = timed_blocks(step, steps, 3, stream, distributed, ctx.torch_device)
    ms_per_eval = float(np.median(ms)) / steps
    # end to end: Parameter objects -> joint stat_sum() (H2D descriptors, kernels, all-reduce, D2H)
    par = engine.parameters[col]
    v0 = par.value
    for k in range(3):
        sharded.stat_sum()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        par.value = v0 + (1e-3 if k % 2 == 0 else -1e-3)
        sharded.stat_sum()
    e2e_s = (time.perf_counter() - t0) / steps
    par.value = v0
    if distributed:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    alg = sum(algorithmic_bytes(ds) for ds in datasets)
    out = {"workload": w["desc"], "scaling": "strong", "observations": n_obs, "observations_on_rank0": len(mine),
           "ms_per_joint_eval": ms_per_eval, "joint_evals_per_s": 1e3 / ms_per_eval,
           "dataset_evals_per_s": n_obs * 1e3 / ms_per_eval, "e2e_ms_per_joint_eval": e2e_s * 1e3,
           "rank0_GBps_algorithmic": alg / (ms_per_eval * 1e-3) / 1e9, "steps": steps, "blocks_ms": ms}
    del sharded, datasets
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    if args.workload == "c5":
        return run_c5(args)
    if args.workload.startswith("c4"):
        return run_c4(args)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = args.steps or 1000
    warmup = 5 if args.warmup is None else max(args.warmup, 3)
    blocks = 5

    from gammapy_b200 import Datasets, Map
    from gammapy_b200.distributed import ShardedDatasets, allreduce_stat

    t_build = time.perf_counter()
    ds = build_dataset(args.workload, rank)
    datasets = Datasets([ds])
    # counts: Poisson realisation of the GPU npred (MapDataset.fake, legacy RandomState seed 2 + rank)
    ds.fake(random_state=2 + rank)
    engine = datasets._get_engine()
    ctx = engine.ctx
    sharded = ShardedDatasets(datasets) if distributed else None
    t_build = time.perf_counter() - t_build
    base = engine.values()
    cols = spectral_columns(engine)

    batch = WORKLOADS[args.workload].get("batch", 1)

    def vector(k):
        v = base.copy()
        v[cols] += 1e-3 if k % 2 == 0 else -1e-3
        if batch > 1:  # profile scan of the first source's spectral shape parameter, all other parameters as above
            grid = np.tile(v, (batch, 1))
            grid[:, cols[0]] += np.linspace(-0.3, 0.3, batch)
            return grid
        return v

    stat_dev = torch.zeros(batch, 1, dtype=torch.float64, device=ctx.torch_device)
    stream = ctx.stream

    def step_device(k):
        if distributed:  # this rank's observation + the all-reduce of the joint statistic (NVLink one-shot, else NCCL)
            sharded.joint_stat_device(vector(k), stat_dev)
        else:
            engine.stat_sum_device(vector(k), stat_dev)

    _dbg("setup done")
    for k in range(warmup):
        step_device(k)
    stream.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    _dbg("warm-up done")
    launches0 = ctx.launch_count
    with ClockSampler(local_rank) as clocks:
        block_ms, timed_wall = timed_blocks(step_device, steps, blocks, stream, distributed, ctx.torch_device)
        launches = (ctx.launch_count - launches0) // blocks
        # nvidia-smi samples every 50 ms: when the timed series is shorter than ~0.4 s keep the very same loop
        # running (untimed) under the sampler so that the clocks line describes this workload under load; every
        # rank must run the SAME number of extra steps (each step holds a collective)
        extra = extra_steps(timed_wall, steps * blocks, distributed, ctx.torch_device)
        for k in range(extra):
            step_device(k)
            if k % 16 == 0:
                stream.synchronize()
        stream.synchronize()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    _dbg("timed loop done")
    ms = float(np.median(block_ms))
    ms_per_step = ms / steps
    value = world * batch * 1e3 / ms_per_step  # dataset evaluations (parameter vectors) per second, all ranks

    # dominant kernel alone (fold + Cash): the library brackets that launch with CUDA events on its own stream
    # (gpy_kernel_timing) while the very same steps run again; max over ranks of the mean launch duration.
    alg_bytes = algorithmic_bytes(ds)
    peak, peak_src = measured_peak()
    ctx.kernel_timing(True)
    for k in range(max(steps, 50)):
        step_device(k)
    kernel_ms, kernel_launches = ctx.kernel_time()
    ctx.kernel_timing(False)
    _dbg("kernel timing done")
    # (shapes that take the any-shape kernel are not bracketed: charge the whole step)
    kernel_us = 1e3 * kernel_ms / kernel_launches if kernel_launches else ms_per_step * 1e3
    if distributed:
        t = torch.tensor([kernel_us], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_us = float(t.item())
    achieved = alg_bytes / (kernel_us * 1e-6) / 1e9
    traffic, traffic_src = committed_traffic(args.workload)

    # end to end through the public API: Parameter objects -> Datasets.stat_sum() -> python float.  The very call
    # that is timed is warmed up first (at N > 1 that is ShardedDatasets.stat_sum: its device / pinned buffers).
    pars = [engine.parameters[c] for c in cols]
    vals = [p.value for p in pars]
    api = sharded if distributed else datasets
    for k in range(5):
        api.stat_sum()
    if distributed:
        dist.barrier()
    e2e_blocks = []
    for _ in range(blocks):
        t0 = time.perf_counter()
        for k in range(steps):
            d = 1e-3 if k % 2 == 0 else -1e-3
            for p, v in zip(pars, vals):
                p.value = v + d
            s = api.stat_sum()
        e2e_blocks.append((time.perf_counter() - t0) / steps)
    e2e_s = float(np.median(e2e_blocks))
    _dbg("e2e loop done")
    for p, v in zip(pars, vals):
        p.value = v
    if distributed:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d_bytes = int(engine.n_par * 8 + 64 * (len(ds._source_names) + 4) + 80 * len(ds._source_names))
    e2e = {"value": world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
           "ms_per_step": e2e_s * 1e3, "blocks_ms_per_step": [x * 1e3 for x in e2e_blocks],
           "what": "Datasets.stat_sum() with Parameter objects on the host: parameter vector + descriptor block copied "
                   "H2D from pinned memory, Cash sum copied D2H, every step; the cubes are dataset state resident in "
                   "HBM exactly as the reference keeps them in host RAM between Minuit steps; median of "
                   f"{blocks} blocks of {steps} steps"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_for(args.workload),
        "timing": {"blocks": blocks, "block_ms": block_ms, "reported": "median block / steps",
                   "note": f"{blocks} back-to-back blocks of exactly {steps} steps, each bracketed by CUDA events on the "
                           "library's stream, max over ranks per block"},
        "gpu_launches": int(launches),
        "collective": (None if not distributed else
                       "one-shot all-reduce over NVLink peer memory inside the library's last kernel (gpy_stat_sum_joint_device)"
                       if getattr(sharded, "_peer", None) else "ncclAllReduce of the joint statistic (torch.distributed), 8 bytes per step"),
        "clocks": dict(clocks.summary(), timed_region_s=timed_wall,
                       note="sampled over the timed blocks and, if shorter than 0.4 s, their untimed continuation"),
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum "
                                                         f"of the committed capture {traffic_src})",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": "fold_cash_t64_kernel", "duration_us": kernel_us, "launches_timed": int(kernel_launches),
                     "step_us": ms_per_step * 1e3,
                     "note": "duration = mean of the kernel's own launches, CUDA events recorded by the library on its "
                             "stream around that launch during a repeat of the timed steps; step_us = whole step "
                             "(K0 + fold + reduce; the work descriptors are reused while no cutout changes), the "
                             "value above; the kernel's share of the step agrees with the ncu launch list "
                             "(profiles/r02_launches_summary.txt)"},
        "setup_s": t_build,
    }
    if rank == 0 and not args.no_extras and not distributed and args.workload.startswith("c3"):
        line["extras"] = extras(engine, datasets, ds, base, cols)
        line["extras"]["fraction_of_bins_with_counts"] = float((ds.counts.data > 0).mean())
        try:
            line["extras"]["c3_flat"] = c3_flat_extras()
        except Exception as exc:
            line["extras"]["c3_flat"] = {"error": repr(exc)}
        try:
            line["extras"]["fit"] = fit_wall_time(with_cpu=not args.no_cpu_baseline)
        except Exception as exc:
            line["extras"]["fit"] = {"error": repr(exc)}
        try:
            line["extras"]["ts_map"] = ts_map_bench(with_cpu=not args.no_cpu_baseline)
        except Exception as exc:
            line["extras"]["ts_map"] = {"error": repr(exc)}
    if rank == 0 and not args.no_cpu_baseline and not distributed:
        try:
            # parity of the headline configuration itself: the oracle (reference arithmetic: float32-FFT convolution,
            # reference Cython Cash) on the same counts and parameter values as the GPU
            gpu_stat = datasets.stat_sum()
            dt, sample, kind = cpu_evals(ds, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                                    "sample": sample,
                                    "ms_per_eval_one_source_moved": 1e3 * getattr(cpu_evals, "one_source_moved_s", float("nan")),
                                    "note": "single Python thread (the reference's N_JOBS_DEFAULT = 1); numpy's BLAS "
                                            "may use more cores inside the edisp matmul"}
            oracle_stat = getattr(cpu_evals, "stat_at_start", None)
            if oracle_stat is not None:
                line["parity"] = dict(parity_check(ds, gpu_stat), stat_gpu=gpu_stat, stat_oracle=oracle_stat,
                                      abs_diff=abs(gpu_stat - oracle_stat))
        except Exception as exc:  # never lose the GPU numbers to the checker
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc}"}
    # BASELINE.json configs[3] at this N (every rank takes part)
    if not args.no_extras and args.workload == "c3":
        del datasets, engine, sharded, api
        ds = None
        torch.cuda.empty_cache()
        try:
            c4 = c4_large_extras(world, rank)
        except Exception as exc:
            c4 = {"error": repr(exc)}
        line.setdefault("extras", {})["c4_large"] = c4
    if rank == 0:
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def extras(engine, datasets, ds, base, cols):
    """Secondary measurements: cold / spatial-change evaluation, batched scans (C2-style 256 vectors per launch)."""
    import torch

    ctx = engine.ctx
    stream = ctx.stream
    out = {}

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        e0.record(stream)
        for k in range(n):
            fn(k)
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / n

    # one source's position moves every step: K1 + K2 (PSF re-convolution) + fold
    lon_cols = [i for i, p in enumerate(engine.parameters) if p.name == "lon_0"]
    stat_dev = torch.zeros(1, 1, dtype=torch.float64, device=ctx.torch_device)

    def move(k, sign=1.0):
        v = base.copy()
        v[lon_cols[k % len(lon_cols)]] += sign * 1e-3 * (1 + k // len(lon_cols))
        engine.stat_sum_device(v, stat_dev)

    for k in range(len(lon_cols)):  # untimed, other positions than the timed ones: every spatial kernel has run
        move(k, sign=-1.0)
    # one evaluation at a time, as an optimiser issues them (wall clock around call + host wait)
    per_call = []
    for k in range(20):
        t0 = time.perf_counter()
        move(k)
        stream.synchronize()
        per_call.append((time.perf_counter() - t0) * 1e3)
    _dbg(f"one-source-moved per call [ms]: {[round(x, 3) for x in per_call]}")
    out["ms_per_eval_one_source_moved"] = float(np.mean(per_call))
    out["ms_per_eval_one_source_moved_median"] = float(np.median(per_call))
    # batched profile scan: 32 vectors per launch
    B = 32
    grid = np.tile(base, (B, 1))
    grid[:, cols[0]] += np.linspace(-0.3, 0.3, B)
    sd = torch.zeros(B, 1, dtype=torch.float64, device=ctx.torch_device)
    engine.stat_sum_device(grid, sd)  # untimed first call (buffers of the batched path)
    ms = timed(lambda k: engine.stat_sum_device(grid, sd), 3)
    out["batched_scan"] = {"vectors_per_launch": B, "ms_per_launch": ms, "vectors_per_s": B * 1e3 / ms}
    # central-difference gradient over every spectral + position parameter (what Minuit computes each iteration):
    # vector 0 = the base point, then +-h per parameter; items equal to the base point's are evaluated once
    gcols = list(cols) + [i for i, p in enumerate(engine.parameters) if p.name in ("lon_0", "lat_0")]
    grid = np.tile(base, (1 + 2 * len(gcols), 1))
    for j, cidx in enumerate(gcols):
        h = 1e-3 * max(abs(base[cidx]), 1.0) if engine.parameters[cidx].name not in ("lon_0", "lat_0") else 1e-3
        grid[1 + 2 * j, cidx] += h
        grid[2 + 2 * j, cidx] -= h
    sd = torch.zeros(grid.shape[0], 1, dtype=torch.float64, device=ctx.torch_device)
    engine.stat_sum_device(grid, sd)  # builds the PSF-convolved cubes of the displaced positions once
    ms = timed(lambda k: engine.stat_sum_device(grid, sd), 3)
    out["fd_gradient"] = {"parameters": len(gcols), "vectors_per_launch": int(grid.shape[0]), "ms_per_gradient": ms}
    return out


def c3_flat_extras(steps=100):
    """The C3 shape with dense counts (`c3-flat`): whole step and fold kernel alone, same measurement as the headline."""
    import torch

    from gammapy_b200 import Datasets

    ds = build_dataset("c3-flat")
    datasets = Datasets([ds])
    ds.fake(random_state=2)
    engine = datasets._get_engine()
    ctx, stream = engine.ctx, engine.ctx.stream
    base = engine.values()
    cols = spectral_columns(engine)
    stat_dev = torch.zeros(1, 1, dtype=torch.float64, device=ctx.torch_device)

    def step(k):
        v = base.copy()
        v[cols] += 1e-3 if k % 2 == 0 else -1e-3
        engine.stat_sum_device(v, stat_dev)

    for k in range(5):
        step(k)
    stream.synchronize()
    ms, _ = timed_blocks(step, steps, 3, stream, False, ctx.torch_device)
    ctx.kernel_timing(True)
    for k in range(steps):
        step(k)
    kernel_ms, n = ctx.kernel_time()
    ctx.kernel_timing(False)
    kernel_us = 1e3 * kernel_ms / max(n, 1)
    peak, _ = measured_peak()
    counts = ds.counts.data
    out = {"workload": WORKLOADS["c3-flat"]["desc"], "ms_per_step": float(np.median(ms)) / steps, "fold_kernel_us": kernel_us,
           "roofline_frac": algorithmic_bytes(ds) / (kernel_us * 1e-6) / 1e9 / peak,
           "fraction_of_bins_with_counts": float((counts > 0).mean())}
    del datasets, engine, ds
    torch.cuda.empty_cache()
    return out


def fit_wall_time(with_cpu=True):
    """BASELINE.json metric (iii): wall time of a full Cash fit of the C1 dataset (analysis_3d shape, 6 free
    parameters: lon_0, lat_0, sigma, index, amplitude, background norm), same optimiser (batched BFGS + Hesse in
    factor space) on the GPU likelihood and on the CPU oracle."""
    from gammapy_b200 import Datasets, Fit, Map, configs
    from oracle import adapter

    def perturb(ds):
        m = ds.models["model-simu"]
        m.spatial_model.lon_0.value, m.spatial_model.lat_0.value, m.spatial_model.sigma.value = 0.23, 0.08, 0.27
        m.spectral_model.index.value, m.spectral_model.amplitude.value = 2.8, 1.3e-11
        ds.background_model.spectral_model.norm.value = 1.05

    ds = configs.make_c1(name="c1-fit")
    ds.fake(random_state=0)
    datasets = Datasets([ds])
    perturb(ds)
    datasets.stat_sum()  # device mirror + first (cold) evaluation outside the timed fit, as for the CPU arm
    t0 = time.perf_counter()
    res = Fit().run(datasets)
    gpu_s = time.perf_counter() - t0
    out = {"workload": "C1: 200x200 px, 10/20 energy bins, Gaussian x PowerLaw + background, 6 free parameters",
           "gpu_fit_s": gpu_s, "gpu_nfev": int(res.nfev), "gpu_success": bool(res.success),
           "gpu_total_stat": float(res.total_stat)}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from tests.test_gpu_fit import OracleDatasets

        de = adapter.evaluator_for(ds, conv="fft32")
        perturb(ds)
        od = OracleDatasets(ds, de)
        od._get_engine().stat_sum()
        t0 = time.perf_counter()
        res_c = Fit().run(od)
        out.update({"cpu_fit_s": time.perf_counter() - t0, "cpu_nfev": int(res_c.nfev), "cpu_success": bool(res_c.success),
                    "cpu_total_stat": float(res_c.total_stat), "cpu_cores": os.cpu_count()})
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
