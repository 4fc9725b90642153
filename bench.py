#!/usr/bin/env python
"""Headline benchmark: neighbourhoods/s of the fused k=50 solve + posterior (mean AND
variance) on BASELINE.json config C2 -- 2-D spatial, 1 M train / 100 k test, Matern 3/2,
k = 50, tau^2 = 1e-3 -- at N GPUs of one node (weak scaling: every rank owns its own
100 k-row test batch, the training set is replicated).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA)
    python bench.py --impl reference [...]                        # CPU reference arm

One JSON line on stdout (rank 0).  A step is one pass of the hot path over the batch:
`value` times the fused kernel with everything resident in HBM (L2 flushed between
steps); `e2e` times the public call `regress_from_indices` with pinned HOST buffers,
host<->device copies inside the timed region.  `roofline` is the fused kernel against
the FP64 issue rate MEASURED on this GPU (MEASURED_PEAKS.json has no FP64 entry, see
tools/fp64_probe.py); `cpu_baseline` is the numpy restatement of the reference timed
on the host cores on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
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
KERNEL_M15, METRIC_L2 = 2, 0
FLOP_PER_NBHD = 62_692  # F(k=50, d=2, r=1, Matern 3/2)
BYTES_PER_NBHD = 1_632  # B(k=50, d=2, r=1)
LOO_BATCH = 10_000      # LOO-mse objective evaluations are timed on this batch per rank
# dram__bytes_read.sum + dram__bytes_write.sum of fused_pipe_kernel<7,52,1> for ONE launch over
# the 100 k batch, from the ncu --set full capture summarised in profiles/ (r1)
NCU_TRAFFIC_BYTES = 68_049_152


def make_data(seed, n=N_TRAIN, t=N_TEST):
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, D))
    q = rng.uniform(size=(t, D))
    y = (np.sin(4 * x[:, 0]) + np.cos(3 * x[:, 1]) + 0.3 * np.sin(11 * x[:, 0] * x[:, 1])
         + 0.05 * rng.normal(size=n))
    return x, y, q


# ---- CPU reference arm ------------------------------------------------------------------
# The UNMODIFIED reference (MuyGPyS, numpy backend) through its own public call
# `MuyGPyS.examples.from_indices.regress_from_indices` when it is importable (the build
# container installs it under baseline/_ref, which travels to the GPU box); otherwise the numpy
# restatement of the same pipeline from oracle/.  Either way: fork-per-core over disjoint row
# chunks, neighbours precomputed and not timed -- the same work the GPU arm times.
_REF = {}


def _reference_model():
    """(MuyGPS object of the real reference, its regress_from_indices) or None."""
    if "model" in _REF:
        return _REF["model"]
    root = os.path.dirname(os.path.abspath(__file__))
    for extra in (os.path.join(root, "oracle", "ref_shims"), os.path.join(root, "baseline", "_ref")):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.append(extra)
    try:
        from MuyGPyS.examples.from_indices import regress_from_indices as ref_regress
        from MuyGPyS.gp import MuyGPS
        from MuyGPyS.gp.deformation import Isotropy, l2
        from MuyGPyS.gp.hyperparameter import Parameter
        from MuyGPyS.gp.kernels import Matern
        from MuyGPyS.gp.noise import HomoscedasticNoise

        model = MuyGPS(kernel=Matern(smoothness=Parameter(1.5),
                                     deformation=Isotropy(l2, length_scale=Parameter(LENGTH_SCALE))),
                       noise=HomoscedasticNoise(NOISE))
        _REF["model"] = (model, ref_regress)
    except Exception:  # noqa: BLE001  (not installed, or a missing optional dependency)
        _REF["model"] = None
    return _REF["model"]


def _ref_chunk(args):
    x, y, q, nn = args
    model, ref_regress = _REF["model"]
    mean, var = ref_regress(model, np.arange(q.shape[0]), nn, q, x, y[:, None])
    return float(np.ravel(mean)[0] + np.ravel(var)[0])


def _cpu_chunk(args):
    from oracle import numpy_oracle as O

    x, y, q, nn = args
    mean, var = O.predict(O.KERNEL_MATERN_15, O.METRIC_L2, LENGTH_SCALE, NOISE, 1.0, x, y, q,
                          np.arange(q.shape[0]), nn)
    return float(mean[0] + var[0])


class CpuReference:
    """The reference pipeline (tensors -> kernel -> LU mean -> LU var) on `cores`
    forked processes over disjoint row chunks (the reference's own chunk rule).  The
    neighbour search is NOT timed, as in the GPU arm.  `step()` returns seconds."""

    def __init__(self, rows_per_core, cores, seed=2, n_train=200_000):
        import multiprocessing as mp

        self.kind = "reference" if _reference_model() is not None else "port"
        chunk_fn = _ref_chunk if self.kind == "reference" else _cpu_chunk
        self.chunk_fn = chunk_fn

        from scipy.spatial import cKDTree

        x, y, _ = make_data(seed, n=n_train, t=1)
        rng = np.random.default_rng(seed + 1)
        self.rows = rows_per_core * cores
        q = rng.uniform(size=(self.rows, D))
        _, nn = cKDTree(x).query(q, k=K, workers=-1)
        nn = nn.astype(np.int64)
        self.jobs = []
        for c in range(cores):  # each worker only gets the training rows its chunk touches
            sl = slice(c * rows_per_core, (c + 1) * rows_per_core)
            uniq, inv = np.unique(nn[sl], return_inverse=True)
            self.jobs.append((x[uniq], y[uniq], q[sl], inv.reshape(nn[sl].shape)))
        self.pool = mp.get_context("fork").Pool(cores)
        self.pool.map(chunk_fn, [(j[0], j[1], j[2][:4], j[3][:4]) for j in self.jobs])

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(self.chunk_fn, self.jobs)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference(rows_per_core, cores, repeats=2):
    ref = CpuReference(rows_per_core, cores)
    best = min(ref.step() for _ in range(repeats))
    ref.close()
    return ref.rows / best, ref.rows, best, ref.kind


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    rows_per_core = 400
    ref = CpuReference(rows_per_core, cores)
    times = []
    for i in range(args.warmup + args.steps):
        secs = ref.step()
        if i >= args.warmup:
            times.append(secs)
    ref.close()
    ms = 1e3 * float(np.mean(times))
    rows = rows_per_core * cores
    value = rows / (ms / 1e3)
    impl_name = ("MuyGPyS 0.9.0 numpy backend, MuyGPyS.examples.from_indices.regress_from_indices"
                 if ref.kind == "reference" else "numpy restatement (oracle/) of the reference")
    sample = (f"{impl_name}: {rows} test rows per step ({rows_per_core} per core) of the "
              f"C2-shaped problem, 200k-point training subsample, neighbours precomputed "
              f"(cKDTree), fork-per-core over disjoint row chunks")
    print(json.dumps({
        "impl": "reference", "metric": "neighbourhoods/s (k=50 fused solve+posterior)",
        "value": value, "unit": "neighbourhoods/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 2-D, 1M train / 100k test, Matern 3/2, k=50, mean+variance",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "neighbourhoods/s", "cores": cores,
                         "kind": ref.kind, "sample": sample},
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


def run_ours(args):
    import torch
    import torch.distributed as dist

    from muygpys_b200 import ops
    from muygpys_b200.examples.from_indices import regress_from_indices
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import Isotropy, l2
    from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter
    from muygpys_b200.gp.kernels import Matern
    from muygpys_b200.gp.noise import HomoscedasticNoise
    from muygpys_b200.optimize.loss import mse_fn
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: replicated training set, per-rank test batch, neighbours precomputed ----
    x_h, y_h, _ = make_data(2)                       # same training set on every rank
    q_h = np.random.default_rng(1000 + rank).uniform(size=(N_TEST, D))
    x, y, q = (torch.as_tensor(a).to(dev) for a in (x_h, y_h, q_h))
    from muygpys_b200.neighbors import NN_Wrapper

    t0 = time.perf_counter()
    nbrs = NN_Wrapper(x, K)  # uniform-grid exact KNN index (d = 2), training set resident
    torch.cuda.synchronize()
    knn_build_s = time.perf_counter() - t0
    nn, _ = nbrs.get_nns(q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        nn, _ = nbrs.get_nns(q)
    torch.cuda.synchronize()
    knn_s = (time.perf_counter() - t0) / 5
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
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
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)

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

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e_evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_e2e()
        b.record()
        e_evs.append((a, b))
    barrier()
    def keep_loaded():
        for _ in range(20):
            step_device()
        torch.cuda.synchronize()

    clocks = sampler.stop(keep_loaded)
    e2e_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in e_evs)], device=dev)
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

    # ---- LOO-mse objective evaluations (second half of the BASELINE metric) -----------
    bi = torch.as_tensor(np.random.default_rng(50 + rank).choice(N_TRAIN, LOO_BATCH,
                                                                 replace=False)).to(dev)
    bnn, _ = nbrs.get_batch_nns(bi)
    obj = make_fused_loo_crossval_fn(model, mse_fn, bi, bnn, x, y, distributed=world > 1)
    for _ in range(3):
        obj(length_scale=0.1)
    barrier()
    t0 = time.perf_counter()
    n_eval = 20
    for i in range(n_eval):
        obj(length_scale=0.05 + 0.01 * i)
    barrier()
    loo_s = torch.tensor([time.perf_counter() - t0], device=dev)

    if world > 1:
        for t in (total_ms, e2e_ms, loo_s):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    e2e_ms_per_step = float(e2e_ms) / args.steps
    value = world * N_TEST / (ms_per_step * 1e-3)

    out = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from fp64_probe import measure

        peak = measure(iters=4000)
        fp64_peak = peak["fp64_peak_tflops"]
        per_gpu_nbhd = N_TEST / (ms_per_step * 1e-3)
        achieved = per_gpu_nbhd * FLOP_PER_NBHD / 1e12
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f)["hbm_gbs"]
        except (OSError, KeyError, ValueError):
            hbm_peak = 6650.0  # fallback stated in B200_PROFILING.md
        cores = host_cores()
        cpu_rows = 20000  # ~5 s per pass on one core's share
        cpu_val, cpu_n, cpu_secs, cpu_kind = cpu_reference(cpu_rows, cores)
        out = {
            "metric": "neighbourhoods/s (k=50 fused solve+posterior)",
            "value": value, "unit": "neighbourhoods/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "C2: 2-D spatial, 1M train / 100k test per GPU, Matern nu=3/2 "
                            "Isotropy(l2, 0.1), k=50, tau^2=1e-3, posterior mean+variance",
                "neighbours": "exact KNN precomputed on device, not timed",
                "knn_seconds_100k_queries": knn_s, "knn_index_build_seconds": knn_build_s,
                "l2": "256 MB flush between timed steps",
                "parallelism": f"dp{world}: test rows sharded, training set replicated"},
            "e2e": {"value": world * N_TEST / (e2e_ms_per_step * 1e-3),
                    "unit": "neighbourhoods/s", "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": int(q_pin.numel() * 8 + nn_pin.numel() * 8
                                              + idx_pin.numel() * 8),
                    "d2h_bytes_per_step": int(2 * N_TEST * 8),
                    "h2d_link_gbs": h2d_gbs,
                    "h2d_ms_at_link_rate": (q_pin.numel() + nn_pin.numel() + idx_pin.numel())
                    * 8 / h2d_gbs / 1e6,
                    "api": "muygpys_b200.examples.from_indices.regress_from_indices -> "
                           "mgp_fused_posterior_host (chunked copy/compute pipeline in the "
                           "C-ABI library); pinned host test features + int64 neighbour "
                           "indices in, mean/var out"},
            "gpu_launches": args.steps,  # one fused_tile_kernel launch per timed step per rank
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": "measured here: max(DFMA, mma.sync DMMA) issue rate, "
                               "tools/fp64_probe.py (MEASURED_PEAKS.json has no FP64 entry)",
                "flop_per_neighbourhood": FLOP_PER_NBHD,
                "traffic": NCU_TRAFFIC_BYTES,
                "hbm": {"algorithmic_bytes_per_launch": BYTES_PER_NBHD * N_TEST,
                        "achieved_gbs": per_gpu_nbhd * BYTES_PER_NBHD / 1e9,
                        "peak_gbs": hbm_peak,
                        "frac": per_gpu_nbhd * BYTES_PER_NBHD / 1e9 / hbm_peak},
                "probe": peak},
            "cpu_baseline": {
                "value": cpu_val, "unit": "neighbourhoods/s", "cores": cores, "kind": cpu_kind,
                "sample": ("MuyGPyS 0.9.0 numpy backend (regress_from_indices), "
                           if cpu_kind == "reference" else "numpy restatement (oracle/), ")
                          + f"{cpu_n} test rows of the same C2-shaped problem (200k-point "
                          f"training subsample), fork-per-core, {cpu_secs:.1f} s"},
            "loo": {"metric": "LOO-mse objective evaluations/s (fused obj_fn, k=50)",
                    "evals_per_s": n_eval / float(loo_s),
                    "batch_rows_total": LOO_BATCH * world,
                    "neighbourhoods_per_s": n_eval * LOO_BATCH * world / float(loo_s),
                    "collective": "1 NCCL SUM all-reduce of 8 doubles per eval" if world > 1
                                  else "none (1 GPU)"},
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
