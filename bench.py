#!/usr/bin/env python
"""Headline benchmark: neighbourhoods/s of the fused k=50 solve + posterior (mean AND
variance) on BASELINE.json config C2 -- 2-D spatial, 1 M train / 100 k test, Matern 3/2,
k = 50, tau^2 = 1e-3 -- at N GPUs of one node (weak scaling: every rank owns its own
100 k-row test batch, the training set is replicated).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA)
    python bench.py --impl reference [...]                        # CPU reference arm

One JSON line on stdout (rank 0).  A step is one pass of the hot path over the batch:
`value` times the fused kernel with everything resident in HBM (L2 flushed between steps);
`e2e` times the public call `regress_any` (the reference's features-in call,
S/examples/regress.py) with pinned HOST buffers, host<->device copies inside the timed region:
test FEATURES in, exact KNN on the device, fused kernel, mean / variance out -- the neighbour
indices never cross the host link, so it scales with the GPUs; `e2e_from_indices` times
`regress_from_indices` with the precomputed int64 neighbour lists in pinned host memory (what
the reference arm is handed; bound by the 42 MB index upload per step, and by the shared host
link at N > 2), `e2e_int32_indices` the same with int32 lists.  `roofline` is the fused kernel against the FP64 issue rate
MEASURED on this GPU (MEASURED_PEAKS.json has no FP64 entry, see tools/fp64_probe.py);
`cpu_baseline` is the unmodified reference (MuyGPyS numpy backend) timed on the host cores on
the same inputs; `loo` is the second half of the BASELINE metric (LOO objective evaluations/s)
with its own CPU baseline; `configs` holds one record per other named shape (C1, C3, C4, C5)
with its own roofline, KNN time and CPU baseline.  SKIP_CONFIGS=1 / SKIP_CPU=1 shorten a
development run.
"""

from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[1] ------------------
N_TRAIN = 1_000_000
N_TEST = 100_000
D, K, R = 2, 50, 1
LENGTH_SCALE, NOISE = 0.1, 1e-3
KERNEL_RBF, KERNEL_M05, KERNEL_M15, KERNEL_M25 = 0, 1, 2, 3
METRIC_L2, METRIC_F2 = 0, 1
FLOP_PER_NBHD = 62_692  # F(k=50, d=2, r=1, Matern 3/2)
BYTES_PER_NBHD = 1_632  # B(k=50, d=2, r=1)
LOO_BATCH = 10_000      # LOO-mse objective evaluations are timed on this batch per rank
WORKLOAD = ("C2: 2-D spatial, 1M train / 100k test per GPU, Matern nu=3/2 Isotropy(l2, 0.1), "
            "k=50, tau^2=1e-3, posterior mean+variance, neighbours precomputed")


def flops_per_nbhd(k, d, r, c_kappa, extra_rhs=0):
    """P (3d + c_k) + k^3/3 + k^2/2 + 2k^2 + 2kr + 2k (+ 2k^2 + 2k per extra
    right-hand side, e.g. the y^T K^-1 y of a lool / analytic-scale evaluation)."""
    P = k * (k + 1) // 2 + k
    return (P * (3 * d + c_kappa) + k ** 3 / 3 + k ** 2 / 2 + 2 * k * k + 2 * k * r + 2 * k
            + extra_rhs * (2 * k * k + 2 * k))


def bytes_per_nbhd(k, d, r):
    return 8 * k + 8 * k * d + 8 * k * r + 8 * d + 8 * (r + 1)


def config_dict(world):
    """Identical for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "l2": "256 MB flush between timed steps",
            "parallelism": f"dp{world}: test rows sharded, training set replicated"}


def make_data(seed, n=N_TRAIN, t=N_TEST):
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, D))
    q = rng.uniform(size=(t, D))
    y = (np.sin(4 * x[:, 0]) + np.cos(3 * x[:, 1]) + 0.3 * np.sin(11 * x[:, 0] * x[:, 1])
         + 0.05 * rng.normal(size=n))
    return x, y, q


def test_queries(rank, t=N_TEST):
    return np.random.default_rng(1000 + rank).uniform(size=(t, D))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def get_chunk_sizes(count, size):
    """The reference's row partition (S/_src/mpi_utils.py:36-41)."""
    base = count // size
    extra = count - base * size
    return [base + 1 if i >= size - extra else base for i in range(size)]


# ---- CPU reference arm ------------------------------------------------------------------
# The UNMODIFIED reference (MuyGPyS, numpy backend) through its own public calls when it is
# importable (the build container installs it under baseline/_ref, which travels to the GPU
# box); otherwise the numpy restatement of the same pipeline from oracle/.  Workers are forked
# AFTER the inputs exist, so every worker sees the full training set (copy-on-write) and is
# handed only its row range -- the fork-per-core substitute for the MPI backend that
# BASELINE.md section 3 item 2 prescribes when mpi4py is missing.
_REF = {}
_SHARED = {}


def _reference_modules():
    if "mods" in _REF:
        return _REF["mods"]
    for extra in (os.path.join(ROOT, "oracle", "ref_shims"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.append(extra)
    try:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from MuyGPyS.examples.from_indices import regress_from_indices
            from MuyGPyS.gp import MuyGPS
            from MuyGPyS.gp.deformation import F2, Anisotropy, Isotropy, l2
            from MuyGPyS.gp.hyperparameter import (AnalyticScale, FixedScale, Parameter,
                                                   VectorParameter)
            from MuyGPyS.gp.kernels import RBF, Matern
            from MuyGPyS.gp.noise import HomoscedasticNoise
            from MuyGPyS.optimize import L_BFGS_B_optimize
            from MuyGPyS.optimize.loss import lool_fn, mse_fn
        _REF["mods"] = dict(locals())
    except Exception:  # noqa: BLE001  (not installed, or a missing optional dependency)
        _REF["mods"] = None
    return _REF["mods"]


def reference_model(kernel, length_scale, noise, analytic=False, bounds=False):
    """A genuine MuyGPyS.gp.MuyGPS (kernel: 'rbf' | 0.5 | 1.5 | 2.5)."""
    m = _reference_modules()
    P = m["Parameter"]

    def par(v):
        return P(v, (v * 0.1, v * 10.0)) if bounds else P(v)

    if isinstance(length_scale, (list, tuple)):
        deformation = m["Anisotropy"](m["l2"], m["VectorParameter"](*[par(v) for v in length_scale]))
    elif kernel == "rbf":
        deformation = m["Isotropy"](m["F2"], par(length_scale))
    else:
        deformation = m["Isotropy"](m["l2"], par(length_scale))
    kern = (m["RBF"](deformation=deformation) if kernel == "rbf"
            else m["Matern"](smoothness=P(kernel), deformation=deformation))
    return m["MuyGPS"](kernel=kern, noise=m["HomoscedasticNoise"](noise),
                       scale=m["AnalyticScale"]() if analytic else m["FixedScale"]())


def _predict_rows(args):
    """Worker: posterior mean + variance of rows [lo, hi) in chunks of <= `chunk` rows."""
    lo, hi, chunk = args
    s = _SHARED
    acc = 0.0
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        if s["model"] is not None:
            mean, var = _REF["mods"]["regress_from_indices"](
                s["model"], np.arange(a, b), s["nn"][a:b], s["q"], s["x"], s["y"])
        else:
            from oracle import numpy_oracle as O

            mean, var = O.predict(s["kid"], s["mid"], s["ls"], s["noise"], 1.0, s["x"],
                                  s["y"][:, 0] if s["y"].ndim == 2 and s["y"].shape[1] == 1
                                  else s["y"], s["q"], np.arange(a, b), s["nn"][a:b])
        acc += float(np.ravel(mean)[0] + np.ravel(var)[0])
    return acc


def _loo_rows(args):
    """Worker: one LOO objective evaluation (reference obj_fn) over its precomputed tensors."""
    theta = args
    return float(_SHARED["obj"](**theta))


class CpuPredict:
    """Reference posterior mean + variance over `rows` query rows, fork-per-core."""

    def __init__(self, x, y, q, nn, model_args, oracle_ids, cores, chunk=50_000):
        import multiprocessing as mp

        mods = _reference_modules()
        self.kind = "reference" if mods is not None else "port"
        _SHARED.clear()
        _SHARED.update(x=x, y=y if y.ndim == 2 else y[:, None], q=q, nn=nn,
                       model=reference_model(*model_args) if mods is not None else None,
                       kid=oracle_ids[0], mid=oracle_ids[1], ls=oracle_ids[2], noise=oracle_ids[3])
        self.rows = q.shape[0] if nn.shape[0] == q.shape[0] else nn.shape[0]
        self.cores = cores
        sizes = get_chunk_sizes(self.rows, cores)
        starts = np.concatenate(([0], np.cumsum(sizes)))
        self.jobs = [(int(starts[i]), int(starts[i + 1]), chunk) for i in range(cores)
                     if sizes[i] > 0]
        self.pool = mp.get_context("fork").Pool(cores)
        self.pool.map(_predict_rows, [(lo, min(hi, lo + 4), chunk) for lo, hi, _ in self.jobs])

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_predict_rows, self.jobs)
        return time.perf_counter() - t0

    def single_process(self, rows):
        """The numpy backend as a user runs it: one process, BLAS threads = all cores."""
        t0 = time.perf_counter()
        _predict_rows((0, rows, 50_000))
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_predict_rate(x, y, q, nn, model_args, oracle_ids, cores, repeats=2, single_rows=0):
    ref = CpuPredict(x, y, q, nn, model_args, oracle_ids, cores)
    best = min(ref.step() for _ in range(repeats))
    single = ref.single_process(single_rows) if single_rows else None
    ref.close()
    out = {"value": ref.rows / best, "unit": "neighbourhoods/s", "cores": cores, "kind": ref.kind,
           "seconds": best, "rows": ref.rows}
    if single:
        out["single_process_numpy_backend"] = {
            "value": single_rows / single, "rows": single_rows, "seconds": single,
            "threads": f"OMP/OPENBLAS default = {cores} (batched numpy.linalg.solve is not "
                       "multi-threaded)"}
    return out


def cpu_loo_rate(x, y, bi, bnn, model_args, loss_name, theta_list):
    """Reference `make_loo_crossval_fn` objective (S/optimize/objective.py:20-105) on the host:
    tensors built once (not timed), then one obj_fn call per theta, single process (the
    objective is one python closure over batch-wide tensors; its MPI form is the only parallel
    one the reference has)."""
    mods = _reference_modules()
    if mods is None:
        return None
    model = reference_model(*model_args, analytic=True, bounds=True)
    cw, pw, b_t, b_nn_t = model.make_train_tensors(bi, bnn, x, y)
    obj = mods["L_BFGS_B_optimize"].make_obj_fn(model, b_t, b_nn_t, cw, pw,
                                                loss_fn=mods[f"{loss_name}_fn"])
    obj(**theta_list[0])
    t0 = time.perf_counter()
    for th in theta_list:
        obj(**th)
    secs = (time.perf_counter() - t0) / len(theta_list)
    return {"evals_per_s": 1.0 / secs, "seconds_per_eval": secs, "batch_rows": int(len(bi)),
            "cores": 1, "kind": "reference",
            "sample": f"MuyGPyS numpy backend make_loo_crossval_fn obj_fn ({loss_name}), "
                      f"{len(bi)} batch rows, tensors prebuilt, single process"}


def mpi_status():
    try:
        import mpi4py  # noqa: F401
        have_mpi4py = True
    except Exception:  # noqa: BLE001
        have_mpi4py = False
    launcher = shutil.which("mpirun") or shutil.which("mpiexec")
    if have_mpi4py and launcher:
        return "available but not run: the fork-per-core numbers below use the same row partition"
    return ("unavailable: " + ("mpi4py not installed" if not have_mpi4py else "mpi4py present")
            + (", no mpirun/mpiexec on PATH" if not launcher else "")
            + " (no network to install them); cpu_baseline is the fork-per-core substitute "
              "BASELINE.md section 3 item 2 prescribes (reference chunk rule, one process per core)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from scipy.spatial import cKDTree

    cores = host_cores()
    x, y, _ = make_data(2)           # the SAME seed-2 training set as the GPU arm
    q = test_queries(0)              # and rank 0's 100 k test rows
    _, nn = cKDTree(x).query(q, k=K, workers=-1)   # neighbours precomputed, not timed
    nn = nn.astype(np.int64)
    ref = CpuPredict(x, y, q, nn, (1.5, LENGTH_SCALE, NOISE),
                     (KERNEL_M15, METRIC_L2, LENGTH_SCALE, NOISE), cores)
    times = []
    for i in range(args.warmup + args.steps):
        secs = ref.step()
        if i >= args.warmup:
            times.append(secs)
    ref.close()
    ms = 1e3 * float(np.mean(times))
    value = N_TEST / (ms / 1e3)
    impl_name = ("MuyGPyS 0.9.0 numpy backend, MuyGPyS.examples.from_indices.regress_from_indices"
                 if ref.kind == "reference" else "numpy restatement (oracle/) of the reference")
    sample = (f"{impl_name}: the full C2 step -- the same seed-2 1M-point training set and the "
              f"same 100k test rows as the GPU arm, neighbours precomputed (cKDTree, not timed) "
              f"-- fork-per-core over {cores} cores with the reference's chunk rule "
              f"({N_TEST // cores} rows per process, (b,k,k,d) temporary "
              f"{N_TEST // cores * K * K * D * 8 / 1e9:.2f} GB each); best-of is not taken: "
              f"mean of {args.steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": "neighbourhoods/s (k=50 fused solve+posterior)",
        "value": value, "unit": "neighbourhoods/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": "neighbourhoods/s", "cores": cores,
                         "kind": ref.kind, "sample": sample},
        "mpi_baseline": mpi_status(),
        "e2e": {"value": value, "unit": "neighbourhoods/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


# ---- GPU arm ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []  # (arrival time, fields)
        self.proc = None
        self.index = index
        self.t_begin = None

    def start(self):
        """Launch the nvidia-smi loop (takes up to seconds to produce its first line on an
        8-GPU box, so this is called before the warm-up)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def begin(self):
        """The timed region starts now: only later samples count."""
        self.t_begin = time.monotonic()

    def stop(self, keep_loaded=None):
        """Samples since begin().  If the timed region was too short for even one sample,
        `keep_loaded()` (the same device step, untimed) is run for up to 3 s until two arrive,
        so that the clocks are still read under this load; the JSON says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = self.t_begin if self.t_begin is not None else 0.0
        note = "timed region"
        if keep_loaded is not None and not [1 for t, _ in self.rows if t >= t0]:
            note = "same load continued after a timed region shorter than the sampling period"
            deadline = time.monotonic() + 3.0
            while time.monotonic() < deadline and len([1 for t, _ in self.rows if t >= t0]) < 2:
                keep_loaded()
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.note = note
        for r in [f for t, f in self.rows if t >= t0]:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "sampled_during": self.note, "reasons": sorted(reasons)}


def ncu_traffic():
    """dram bytes per launch of the timed kernel from the committed ncu --set full capture
    (profiles/r2_traffic.json is written next to the summary it is taken from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            rec = json.load(f)
        return rec["dram_bytes_per_launch"], rec["source"]
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture committed for this kernel"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from muygpys_b200 import ops
    from muygpys_b200.examples.from_indices import regress_from_indices
    from muygpys_b200.examples.regress import regress_any
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import F2, Anisotropy, Isotropy, l2
    from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter, VectorParameter
    from muygpys_b200.gp.kernels import RBF, Matern
    from muygpys_b200.gp.noise import HomoscedasticNoise
    from muygpys_b200.neighbors import NN_Wrapper
    from muygpys_b200.optimize.loss import lool_fn, mse_fn
    from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: muygpys_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for N>1"
    skip_configs = os.environ.get("SKIP_CONFIGS") == "1"
    # CPU baselines: rank 0 at N = 1 only (the other ranks would idle at the next barrier)
    skip_cpu = os.environ.get("SKIP_CPU") == "1" or world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def timed_ms(fn, reps, warm=2, do_flush=True):
        """Mean CUDA-event milliseconds of `fn` over `reps` launches (max over ranks)."""
        for _ in range(warm):
            fn()
        barrier()
        evs = []
        for _ in range(reps):
            if do_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return allmax(sum(a.elapsed_time(b) for a, b in evs) / reps)

    # ---- inputs: replicated training set, per-rank test batch, neighbours precomputed ----
    x_h, y_h, _ = make_data(2)                       # same training set on every rank
    q_h = test_queries(rank)
    x, y, q = (torch.as_tensor(a).to(dev) for a in (x_h, y_h, q_h))
    t0 = time.perf_counter()
    nbrs = NN_Wrapper(x, K)  # uniform-grid exact KNN index (d = 2), training set resident
    torch.cuda.synchronize()
    knn_build_s = time.perf_counter() - t0
    nn, _ = nbrs.get_nns(q)
    knn_ms = timed_ms(lambda: nbrs.get_nns(q), 5, warm=1, do_flush=False)
    model = MuyGPS(kernel=Matern(smoothness=Parameter(1.5),
                                 deformation=Isotropy(l2, Parameter(LENGTH_SCALE, (0.01, 1.0)))),
                   noise=HomoscedasticNoise(NOISE), scale=AnalyticScale())
    fused_kw = dict(kernel_id=KERNEL_M15, metric_id=METRIC_L2, length_scale=LENGTH_SCALE,
                    noise=NOISE, scale=1.0)

    def step_device():
        return ops.fused_posterior(x, q, None, nn, y, **fused_kw)

    # ---- device-resident timing: K steps, L2 flushed between steps ------------------
    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs a moment before its first line: start before warm-up
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    sampler.begin()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device()
        b.record()
        evs.append((a, b))
    barrier()
    ms_per_step = allmax(sum(a.elapsed_time(b) for a, b in evs) / args.steps)

    # ---- end to end through the public API with pinned host buffers ------------------
    q_pin = torch.as_tensor(q_h).pin_memory()
    nn_pin = nn.cpu().pin_memory()
    idx_pin = torch.arange(N_TEST).pin_memory()
    mean_pin = torch.empty(N_TEST, dtype=torch.float64).pin_memory()
    var_pin = torch.empty(N_TEST, dtype=torch.float64).pin_memory()

    def step_e2e():
        m, v = regress_from_indices(model, idx_pin, nn_pin, q_pin, x, y)
        mean_pin.copy_(m, non_blocking=True)
        var_pin.copy_(v, non_blocking=True)

    nn_pin32 = nn.cpu().to(torch.int32).pin_memory()

    def step_e2e_i32():  # the same call for a caller that keeps its neighbour lists as int32
        m, v = regress_from_indices(model, idx_pin, nn_pin32, q_pin, x, y)
        mean_pin.copy_(m, non_blocking=True)
        var_pin.copy_(v, non_blocking=True)

    def step_e2e_knn():  # test FEATURES in: KNN on the device, indices never cross the link
        m, v, _ = regress_any(model, q_pin, x, nbrs, y, sync_timing=False)
        mean_pin.copy_(m, non_blocking=True)
        var_pin.copy_(v, non_blocking=True)

    # the end-to-end step is paced by the HOST as much as by the GPU (≈70 driver calls and the
    # Python path per step): it keeps getting faster for ~10 steps after a cold start (host
    # clocks, page-locked buffers first touched by the copy engine), so it gets W + 8 warm-ups
    for _ in range(args.warmup + 8):
        step_e2e_knn()
    barrier()
    e_evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_e2e_knn()
        b.record()
        e_evs.append((a, b))
    barrier()

    def keep_loaded():
        for _ in range(20):
            step_device()
        torch.cuda.synchronize()

    clocks = sampler.stop(keep_loaded)
    e2e_steps_ms = [a.elapsed_time(b) for a, b in e_evs]
    e2e_ms_per_step = allmax(sum(e2e_steps_ms) / args.steps)
    e2e_idx_ms = timed_ms(step_e2e, args.steps, warm=args.warmup + 8)
    e2e_i32_ms = timed_ms(step_e2e_i32, args.steps, warm=5)
    # what the host link of this box gives for the same pinned index buffer (explains e2e)
    nn_stage = torch.empty_like(nn)
    nn_stage.copy_(nn_pin, non_blocking=True)
    torch.cuda.synchronize()
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ca.record()
    for _ in range(5):
        nn_stage.copy_(nn_pin, non_blocking=True)
    cb.record()
    torch.cuda.synchronize()
    h2d_gbs = 5 * nn_pin.numel() * 8 / (ca.elapsed_time(cb) * 1e-3) / 1e9
    del nn_stage

    # ---- LOO objective evaluations (second half of the BASELINE metric) ---------------
    bi_h = np.random.default_rng(50 + rank).choice(N_TRAIN, LOO_BATCH, replace=False)
    bi = torch.as_tensor(bi_h).to(dev)
    bnn, _ = nbrs.get_batch_nns(bi)
    loo = {}
    n_eval = 40
    loo_collective = "none (1 GPU)"
    for lname, lfn in (("mse", mse_fn), ("lool", lool_fn)):
        obj = make_fused_loo_crossval_fn(model, lfn, bi, bnn, x, y, distributed=world > 1)
        if world > 1:
            # what the evaluations actually use: the one-shot NVLink peer-memory sum in the
            # objective kernel's epilogue when symmetric memory is available, else NCCL
            from muygpys_b200.distributed import PartialsReducer
            fused_peers = PartialsReducer(x.device).peers is not None
            loo_collective = ("SUM of the 8-double record over NVLink peer memory inside the "
                              "objective kernel (mgp_fused_loo_peers)" if fused_peers else
                              "1 NCCL SUM all-reduce of 8 doubles per eval")
        for _ in range(5):
            obj(length_scale=0.1)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_eval):
            obj(length_scale=0.05 + 0.005 * i)
        barrier()
        secs = allmax(time.perf_counter() - t0)
        loo[lname] = {"evals_per_s": n_eval / secs, "us_per_eval": 1e6 * secs / n_eval,
                      "neighbourhoods_per_s": n_eval * LOO_BATCH * world / secs}

    value = world * N_TEST / (ms_per_step * 1e-3)

    # ---- the other named shapes ---------------------------------------------------------
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from fp64_probe import measure

    peak = measure(iters=4000)
    fp64_peak = peak["fp64_peak_tflops"]
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = json.load(f)["hbm_gbs"], "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6650.0, "fallback stated in B200_PROFILING.md"
    cores = host_cores()

    def roof_fp64(rows_per_s_per_gpu, flop):
        ach = rows_per_s_per_gpu * flop / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach / fp64_peak, "flop_per_neighbourhood": flop}

    def roof_hbm(rows_per_s_per_gpu, nbytes):
        ach = rows_per_s_per_gpu * nbytes / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "bytes_per_row": nbytes, "peak_source": hbm_src}

    configs = {}
    if not skip_configs:
        del nn_pin, q_pin, nn_pin32
        torch.cuda.empty_cache()
        gen = torch.Generator(device=dev)

        def c2_like_targets(xx):
            return (torch.sin(4 * xx[:, 0]) + torch.cos(3 * xx[:, 1])
                    + 0.3 * torch.sin(11 * xx[:, 0] * xx[:, 1])
                    + 0.05 * torch.randn(xx.shape[0], generator=gen, device=dev,
                                         dtype=torch.float64))

        # -- C1: univariate tutorial, 10k train / 1k test, RBF(F2, 0.05), k = 30 -------------
        rng = np.random.default_rng(1)
        x1 = rng.uniform(size=(10_000, 1))
        y1 = np.sin(2 * np.pi * 4 * x1[:, 0]) + 1e-2 * rng.normal(size=10_000)
        q1 = rng.uniform(size=(1_000, 1))
        x1d, y1d, q1d = (torch.as_tensor(a).to(dev) for a in (x1, y1, q1))
        nb1 = NN_Wrapper(x1d, 30)
        nn1, _ = nb1.get_nns(q1d)
        kw1 = dict(kernel_id=KERNEL_RBF, metric_id=METRIC_F2, length_scale=0.05, noise=1e-3)
        ms1 = timed_ms(lambda: ops.fused_posterior(x1d, q1d, None, nn1, y1d, **kw1), 20)
        knn1 = timed_ms(lambda: nb1.get_nns(q1d), 5, warm=1, do_flush=False)
        f1 = flops_per_nbhd(30, 1, 1, 2)
        rec = {"workload": "C1: 1-D sine, 10k train / 1k test per GPU, RBF Isotropy(F2, 0.05), "
                           "k=30, mean+variance (one 1k-row launch: launch-latency-bound)",
               "rows_per_gpu": 1000, "ms": ms1, "value": world * 1000 / (ms1 * 1e-3),
               "unit": "neighbourhoods/s", "knn_ms": knn1,
               "roofline": roof_fp64(1000 / (ms1 * 1e-3), f1)}
        if not skip_cpu:
            c = cpu_predict_rate(x1, y1, q1, nn1.cpu().numpy(), ("rbf", 0.05, 1e-3),
                                 (KERNEL_RBF, METRIC_F2, 0.05, 1e-3), cores)
            c["sample"] = "MuyGPyS numpy backend, the full 1k test rows, fork-per-core"
            rec["cpu_baseline"] = c
        configs["C1"] = rec
        del x1d, y1d, q1d, nn1, nb1

        # -- C3: MNIST-shaped, 60k x 784, 10k test, r = 10, RBF(F2, 28), k = 30 ---------------
        rng = np.random.default_rng(3)
        n3, t3, d3, r3, k3 = 60_000, 10_000, 784, 10, 30
        cent = rng.normal(0, 0.5, size=(r3, d3))
        lab = rng.integers(0, r3, size=n3)
        x3 = cent[lab] + rng.normal(size=(n3, d3))
        q3 = cent[rng.integers(0, r3, size=t3)] + rng.normal(size=(t3, d3))
        y3 = -0.1 * np.ones((n3, r3))
        y3[np.arange(n3), lab] = 0.9
        x3d, y3d, q3d = (torch.as_tensor(a).to(dev) for a in (x3, y3, q3))
        nn3, _ = ops.knn(x3d, q3d, k3)
        knn3 = timed_ms(lambda: ops.knn(x3d, q3d, k3), 3, warm=1, do_flush=False)
        kw3 = dict(kernel_id=KERNEL_RBF, metric_id=METRIC_F2, length_scale=28.0, noise=1e-3)
        ms3 = timed_ms(lambda: ops.fused_posterior(x3d, q3d, None, nn3, y3d, **kw3), 10)
        f3, b3 = flops_per_nbhd(k3, d3, r3, 2), bytes_per_nbhd(k3, d3, r3)
        rec = {"workload": "C3: MNIST-shaped 784-d, 60k train / 10k test per GPU, r=10, RBF "
                           "Isotropy(F2, 28), k=30, mean+variance (DMMA Gram assembly)",
               "rows_per_gpu": t3, "ms": ms3, "value": world * t3 / (ms3 * 1e-3),
               "unit": "neighbourhoods/s", "knn_ms": knn3,
               "knn": {"queries_per_s": world * t3 / (knn3 * 1e-3),
                       "gram_tflops": 2.0 * n3 * t3 * d3 / (knn3 * 1e-3) / 1e12,
                       "fp64_peak_tflops": fp64_peak,
                       "frac": 2.0 * n3 * t3 * d3 / (knn3 * 1e-3) / 1e12 / fp64_peak},
               "roofline": roof_hbm(t3 / (ms3 * 1e-3), b3),
               "roofline_fp64": roof_fp64(t3 / (ms3 * 1e-3), f3)}
        if not skip_cpu:
            rows3 = 40 * cores
            c = cpu_predict_rate(x3, y3, q3[:rows3], nn3[:rows3].cpu().numpy(), ("rbf", 28.0, 1e-3),
                                 (KERNEL_RBF, METRIC_F2, 28.0, 1e-3), cores, repeats=1)
            c["sample"] = (f"MuyGPyS numpy backend, {rows3} of the 10k test rows (40 per core: "
                           f"the (b,k,k,784) temporary is 5.6 MB per row), fork-per-core")
            rec["cpu_baseline"] = c
        configs["C3"] = rec
        del x3d, y3d, q3d, nn3, x3, q3, y3
        torch.cuda.empty_cache()

        # -- C4: anisotropic Matern 5/2, 10M train, k = 100, lool objective, 10k rows per GPU --
        n4, b4, k4 = 10_000_000, 10_000, 100
        gen.manual_seed(4)
        x4 = torch.rand((n4, 2), generator=gen, device=dev, dtype=torch.float64)
        y4 = (torch.sin(2 * np.pi * x4[:, 0] / 0.1) * torch.cos(2 * np.pi * x4[:, 1] / 0.5)
              + 0.05 * torch.randn(n4, generator=gen, device=dev, dtype=torch.float64))
        t0 = time.perf_counter()
        nb4 = NN_Wrapper(x4, k4)
        torch.cuda.synchronize()
        knn4_build = time.perf_counter() - t0
        gen.manual_seed(40 + rank)
        bi4 = torch.randperm(n4, generator=gen, device=dev)[:b4].sort().values
        bnn4, _ = nb4.get_batch_nns(bi4)
        knn4 = timed_ms(lambda: nb4.get_batch_nns(bi4), 3, warm=1, do_flush=False)
        model4 = MuyGPS(kernel=Matern(smoothness=Parameter(2.5), deformation=Anisotropy(
            l2, VectorParameter(Parameter(0.1, (0.01, 1.0)), Parameter(0.5, (0.05, 5.0))))),
            noise=HomoscedasticNoise(NOISE), scale=AnalyticScale())
        obj4 = make_fused_loo_crossval_fn(model4, lool_fn, bi4, bnn4, x4, y4,
                                          distributed=world > 1)
        for _ in range(3):
            obj4(length_scale0=0.1, length_scale1=0.5)
        barrier()
        t0 = time.perf_counter()
        n4e = 20
        for i in range(n4e):
            obj4(length_scale0=0.08 + 0.002 * i, length_scale1=0.5)
        barrier()
        secs4 = allmax(time.perf_counter() - t0)
        # analytic gradient (both length scales) in one launch, against the 1 + p = 3 plain
        # evaluations the reference's finite differences need
        from muygpys_b200.optimize.objective import make_fused_loo_value_and_grad_fn
        vg4 = make_fused_loo_value_and_grad_fn(model4, lool_fn, bi4, bnn4, x4, y4,
                                               distributed=world > 1)
        for _ in range(3):
            vg4(length_scale0=0.1, length_scale1=0.5)
        barrier()
        t0 = time.perf_counter()
        for i in range(n4e):
            vg4(length_scale0=0.08 + 0.002 * i, length_scale1=0.5)
        barrier()
        secs4g = allmax(time.perf_counter() - t0)
        kw4 = dict(kernel_id=KERNEL_M25, metric_id=METRIC_L2, length_scale=[0.1, 0.5], noise=NOISE,
                   want_yky=True)
        ms4 = timed_ms(lambda: ops.fused_posterior(x4, x4, bi4, bnn4, y4, **kw4), 10)
        f4 = flops_per_nbhd(k4, 2, 1, 8, extra_rhs=1)
        rec = {"workload": "C4: anisotropic 2-D Matern nu=5/2 (l = 0.1, 0.5), 10M train, k=100, "
                           "lool objective with analytic scale, 10k batch rows per GPU",
               "rows_per_gpu": b4, "ms": ms4, "value": world * b4 / (ms4 * 1e-3),
               "unit": "neighbourhoods/s (fused kernel: mean, variance, y^T K^-1 y)",
               "loo_evals_per_s": n4e / secs4, "loo_us_per_eval": 1e6 * secs4 / n4e,
               "loo_value_and_grad_us_per_eval": 1e6 * secs4g / n4e,
               "loo_finite_difference_us_per_gradient": 3e6 * secs4 / n4e,
               "knn_ms": knn4, "knn_index_build_s": knn4_build,
               "roofline": roof_fp64(b4 / (ms4 * 1e-3), f4)}
        if not skip_cpu:
            rows4 = 125 * cores if cores <= 16 else 2000
            sel = bi4[:rows4].cpu().numpy()
            snn = bnn4[:rows4].cpu().numpy()
            uniq, inv = np.unique(np.concatenate((sel, snn.ravel())), return_inverse=True)
            xs, ys = x4[torch.as_tensor(uniq).to(dev)].cpu().numpy(), y4[torch.as_tensor(uniq).to(dev)].cpu().numpy()
            c = cpu_loo_rate(xs, ys, inv[:rows4], inv[rows4:].reshape(snn.shape),
                             (2.5, [0.1, 0.5], NOISE), "lool",
                             [{"length_scale0": 0.08 + 0.01 * i, "length_scale1": 0.5}
                              for i in range(2)])
            if c:
                c["scaled_to_10k_rows_evals_per_s"] = c["evals_per_s"] * rows4 / b4
                rec["cpu_baseline"] = c
        configs["C4"] = rec
        del x4, y4, nb4, bnn4, obj4, vg4
        torch.cuda.empty_cache()

        # -- C5: scale-out posterior, 100M train / 10M test, Matern 1/2, k = 50 ----------------
        n5, t5, k5 = 100_000_000, 10_000_000, 50
        gen.manual_seed(5)
        x5 = torch.rand((n5, 2), generator=gen, device=dev, dtype=torch.float64)
        y5 = c2_like_targets(x5)
        gen.manual_seed(500 + rank)
        q5 = torch.rand((t5, 2), generator=gen, device=dev, dtype=torch.float64)
        t0 = time.perf_counter()
        nb5 = NN_Wrapper(x5, k5)
        torch.cuda.synchronize()
        knn5_build = time.perf_counter() - t0
        nn5, _ = nb5._query(q5, k5)
        knn5 = timed_ms(lambda: nb5._query(q5, k5), 2, warm=0, do_flush=False)
        kw5 = dict(kernel_id=KERNEL_M05, metric_id=METRIC_L2, length_scale=0.1, noise=NOISE)
        mean5 = torch.empty((t5, 1), dtype=torch.float64, device=dev)
        var5 = torch.empty((t5,), dtype=torch.float64, device=dev)
        ms5 = timed_ms(lambda: ops.fused_posterior(x5, q5, None, nn5, y5, out_mean=mean5,
                                                   out_var=var5, **kw5), 3, warm=1)
        # fast posterior mean (S/examples/fast_posterior_mean.py): coefficients of the training
        # points that are some test point's nearest neighbour, then a k-dot per test point
        closest = torch.unique(nn5[:, 0])
        t0 = time.perf_counter()
        cnn, _ = nb5._query(x5[closest], k5)   # the point itself first == fast_nn_update
        coeffs = ops.fused_posterior(x5, x5, closest, cnn, y5, want_mean=False, want_var=False,
                                     want_coeffs=True, **kw5)["coeffs"]
        torch.cuda.synchronize()
        pre_s = allmax(time.perf_counter() - t0)
        slot = torch.searchsorted(closest, nn5[:, 0].contiguous())
        nn_fast = cnn[slot]
        fkw = dict(kernel_id=KERNEL_M05, metric_id=METRIC_L2, length_scale=0.1)
        msf = timed_ms(lambda: ops.fast_mean(x5, q5, None, nn_fast, slot, coeffs, **fkw), 3,
                       warm=1)
        f5 = flops_per_nbhd(k5, 2, 1, 2)
        bfast = 8 + 8 * k5 + 8 * k5 * 2 + 8 * k5 + 8 * 2 + 8
        rec = {"workload": "C5: 2-D, 100M train / 10M test per GPU, Matern nu=1/2 Isotropy(l2, "
                           "0.1), k=50: (i) mean+variance, (ii) fast posterior mean",
               "rows_per_gpu": t5, "ms": ms5, "value": world * t5 / (ms5 * 1e-3),
               "unit": "neighbourhoods/s", "knn_ms": knn5, "knn_index_build_s": knn5_build,
               "knn_queries_per_s": world * t5 / (knn5 * 1e-3),
               "roofline": roof_fp64(t5 / (ms5 * 1e-3), f5),
               "fast_mean": {"apply_ms": msf, "value": world * t5 / (msf * 1e-3),
                             "unit": "test points/s",
                             "precompute_s": pre_s, "precompute_rows": int(closest.numel()),
                             "precompute_note": "KNN + coefficients of the training points that "
                                                "are some test point's nearest neighbour",
                             "roofline": roof_hbm(t5 / (msf * 1e-3), bfast)},
               "data": "generated on the device (torch.Generator seeds 5 / 500+rank), same "
                       "distribution and targets as C2"}
        if not skip_cpu:
            rows5 = 400 * cores
            sel_nn = nn5[:rows5].cpu().numpy()
            uniq, inv = np.unique(sel_nn, return_inverse=True)
            ud = torch.as_tensor(uniq).to(dev)
            c = cpu_predict_rate(x5[ud].cpu().numpy(), y5[ud].cpu().numpy(),
                                 q5[:rows5].cpu().numpy(), inv.reshape(sel_nn.shape),
                                 (0.5, 0.1, NOISE), (KERNEL_M05, METRIC_L2, 0.1, NOISE), cores)
            c["sample"] = (f"MuyGPyS numpy backend, {rows5} of the 10M test rows (their "
                           f"neighbours gathered from the 100M-point set), fork-per-core; "
                           f"linear extrapolation to 10M rows: {t5 / c['value']:.0f} s")
            rec["cpu_baseline"] = c
        configs["C5"] = rec
        del x5, y5, q5, nn5, nb5, coeffs, cnn, nn_fast, mean5, var5
        torch.cuda.empty_cache()

    out = None
    if rank == 0:
        per_gpu_nbhd = N_TEST / (ms_per_step * 1e-3)
        achieved = per_gpu_nbhd * FLOP_PER_NBHD / 1e12
        traffic, traffic_src = ncu_traffic()
        cpu = None
        cpu_loo = None
        if not skip_cpu:
            from scipy.spatial import cKDTree

            # the SAME inputs as the timed GPU step: seed-2 training set, rank 0's test rows
            cpu = cpu_predict_rate(x_h, y_h, q_h, nn.cpu().numpy(), (1.5, LENGTH_SCALE, NOISE),
                                   (KERNEL_M15, METRIC_L2, LENGTH_SCALE, NOISE), cores,
                                   repeats=3, single_rows=5000)
            cpu["sample"] = (
                ("MuyGPyS 0.9.0 numpy backend (regress_from_indices), " if cpu["kind"] == "reference"
                 else "numpy restatement (oracle/), ")
                + f"the full C2 step on identical inputs (seed-2 1M-point training set, the same "
                  f"100k test rows and neighbour lists as the GPU step), fork-per-core with the "
                  f"reference's chunk rule, best of 3: {cpu['seconds']:.2f} s")
            rows_loo = 2000
            cpu_loo = {}
            for lname in ("mse", "lool"):
                c = cpu_loo_rate(x_h, y_h, bi_h[:rows_loo], bnn[:rows_loo].cpu().numpy(),
                                 (1.5, LENGTH_SCALE, NOISE), lname,
                                 [{"length_scale": 0.05 + 0.02 * i} for i in range(3)])
                if c:
                    c["scaled_to_10k_rows_evals_per_s"] = c["evals_per_s"] * rows_loo / LOO_BATCH
                    cpu_loo[lname] = c
        setup = {"knn_ms_100k_queries": knn_ms, "knn_index_build_seconds": knn_build_s,
                 "neighbours": "exact KNN precomputed on device for `value` / `e2e_from_indices`; `e2e` runs it inside the timed step"}
        out = {
            "metric": "neighbourhoods/s (k=50 fused solve+posterior)",
            "value": value, "unit": "neighbourhoods/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world), "setup": setup,
            "e2e": {"value": world * N_TEST / (e2e_ms_per_step * 1e-3),
                    "unit": "neighbourhoods/s", "ms_per_step": e2e_ms_per_step,
                    # (rank 0's individual steps: the host enqueues the copies / launches of a
                    # step one by one, so host jitter shows up here and not in `value`)
                    "ms_steps_rank0_first20": [round(t, 4) for t in e2e_steps_ms[:20]],
                    "warmup_steps": args.warmup + 8,
                    "h2d_bytes_per_step": int(N_TEST * D * 8),
                    "d2h_bytes_per_step": int(2 * N_TEST * 8),
                    "api": "muygpys_b200.examples.regress.regress_any (the reference's "
                           "features-in call, S/examples/regress.py): pinned host test FEATURES "
                           "in, exact KNN on the device (mgp_knn_grid), fused kernel, mean/var "
                           "out to pinned host memory -- the neighbour indices never cross the "
                           "host link.  Does MORE than the reference arm's step, whose "
                           "neighbours are precomputed and untimed"},
            "e2e_from_indices": {
                "value": world * N_TEST / (e2e_idx_ms * 1e-3), "unit": "neighbourhoods/s",
                "ms_per_step": e2e_idx_ms,
                "h2d_bytes_per_step": int(N_TEST * D * 8 + N_TEST * K * 8 + N_TEST * 8),
                "d2h_bytes_per_step": int(2 * N_TEST * 8),
                "h2d_link_gbs": h2d_gbs,
                "h2d_ms_at_link_rate": (N_TEST * D + N_TEST * K + N_TEST) * 8 / h2d_gbs / 1e6,
                "api": "muygpys_b200.examples.from_indices.regress_from_indices -> "
                       "mgp_fused_posterior_host (chunked copy/compute pipeline in the C-ABI "
                       "library); pinned host test features + precomputed int64 neighbour "
                       "indices in (exactly what the reference arm is handed), mean/var out; "
                       "bound by the index upload, which shares the host link at N > 1"},
            "e2e_int32_indices": {
                "value": world * N_TEST / (e2e_i32_ms * 1e-3), "unit": "neighbourhoods/s",
                "ms_per_step": e2e_i32_ms,
                "h2d_bytes_per_step": int(N_TEST * D * 8 + N_TEST * K * 4 + N_TEST * 8),
                "d2h_bytes_per_step": int(2 * N_TEST * 8),
                "api": "the `e2e_from_indices` call with the neighbour indices held as int32 on "
                       "the host (numpy index arrays of any integer dtype are valid reference "
                       "inputs): mgp_fused_posterior_host32 uploads them as they are and widens "
                       "each chunk on the device"},
            "gpu_launches": args.steps,  # one fused_tp_kernel launch per timed `value` step per rank
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": "measured here: max(DFMA, mma.sync DMMA) issue rate, "
                               "tools/fp64_probe.py (MEASURED_PEAKS.json has no FP64 entry)",
                "flop_per_neighbourhood": FLOP_PER_NBHD,
                "traffic": traffic, "traffic_source": traffic_src,
                "hbm": {"algorithmic_bytes_per_launch": BYTES_PER_NBHD * N_TEST,
                        "achieved_gbs": per_gpu_nbhd * BYTES_PER_NBHD / 1e9,
                        "peak_gbs": hbm_peak,
                        "frac": per_gpu_nbhd * BYTES_PER_NBHD / 1e9 / hbm_peak},
                "probe": peak},
            "cpu_baseline": cpu if cpu is not None else {"skipped": "SKIP_CPU=1"},
            "mpi_baseline": mpi_status(),
            "loo": {"metric": "LOO objective evaluations/s (fused obj_fn, one launch per "
                              "evaluation, k=50, 10k batch rows per GPU)",
                    "evals_per_s": loo["mse"]["evals_per_s"],
                    "batch_rows_total": LOO_BATCH * world,
                    "neighbourhoods_per_s": loo["mse"]["neighbourhoods_per_s"],
                    "mse": loo["mse"], "lool": loo["lool"],
                    "cpu_baseline": cpu_loo,
                    "collective": loo_collective},
            "configs": configs,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
