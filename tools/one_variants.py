#!/usr/bin/env python
"""Dev tool: time several builds of the single-instantiation harness (csrc/_one.cu) on the C2 shape."""
import ctypes as C, glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200 import ops
from muygpys_b200.neighbors import NN_Wrapper
rng = np.random.default_rng(7)
n, b, k = 1_000_000, int(os.environ.get('B', 100_000)), int(os.environ.get('K', 50))
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda(); y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
nn, _ = NN_Wrapper(x, k).get_nns(q)
ops.set_fused_variant(2)
ref = ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = {}
if os.environ.get("TIME_REF"):
    ts = []
    for _ in range(6):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    print("library tile kernel (variant 2):", round(min(ts[1:]), 4), "ms", flush=True)
    ops.set_fused_variant(0)
    ts = []
    for _ in range(8):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    print("library, automatic choice (variant 0):", round(min(ts[1:]), 4), "min", round(float(np.mean(ts[1:])), 4), "mean ms", flush=True)
    ts = []
    for _ in range(8):  # three flushes: the GPU is still busy when the launch is enqueued
        flush.zero_(); flush.zero_(); flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    print("library, launch enqueued under a longer flush:", round(min(ts[1:]), 4), "min", round(float(np.mean(ts[1:])), 4), "mean ms", flush=True)
    ops.set_fused_variant(2)
only = os.environ.get("ONLY")  # comma-separated variant names
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "muygpys_b200", "csrc", "build", "libone_*.so"))):
    if only and os.path.basename(path)[7:-3] not in only.split(","):
        continue
    lib = C.CDLL(path)
    lib.one_run.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_longlong, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.one_err.restype = C.c_char_p
    mean = torch.empty(b, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
    def f():
        rc = lib.one_run(x.data_ptr(), q.data_ptr(), nn.data_ptr(), y.data_ptr(), n, b, k, 0.1, 1e-3, mean.data_ptr(), var.data_ptr(), None)
        assert rc == 0, lib.one_err()
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.default_stream()); f(); e.record(torch.cuda.default_stream()); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    err = float((mean - ref["mean"][:, 0]).abs().max() / ref["mean"].abs().max())
    res[os.path.basename(path)[7:-3]] = {"ms": round(float(np.mean(ts)), 4), "min": round(min(ts), 4), "err": err}
    print(os.path.basename(path), res[os.path.basename(path)[7:-3]], flush=True)
print(json.dumps(res))
