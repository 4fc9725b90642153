#!/usr/bin/env python
"""Dev tool: back-substitution (coefficient) mode of the thread-per-tile kernel against the
lane-parallel column kernel, harness csrc/_one_tp.cu built with -DONE_GRAD."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200.neighbors import NN_Wrapper
rng = np.random.default_rng(7)
n, b, k = 1_000_000, int(os.environ.get("B", 100_000)), int(os.environ.get("K", 50))
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda(); y = torch.as_tensor(rng.normal(size=n)).cuda()
bi = torch.as_tensor(np.sort(rng.choice(n, b, replace=False))).cuda()
nn, _ = NN_Wrapper(x, k).get_batch_nns(bi)
nn = nn.contiguous()
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "muygpys_b200", "csrc", "build", sys.argv[1]))
lib.one_run_coeffs.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_longlong, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p]
lib.one_err.restype = C.c_char_p
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = {}
for use_tp in (0, 1):
    mean = torch.empty(b, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
    co = torch.zeros((b, k), dtype=torch.float64, device="cuda")
    def f():
        rc = lib.one_run_coeffs(x.data_ptr(), bi.data_ptr(), nn.data_ptr(), y.data_ptr(), n, b, k, 0.1, 1e-3,
                                mean.data_ptr(), var.data_ptr(), co.data_ptr(), use_tp, None)
        assert rc == 0, lib.one_err()
    for _ in range(2): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    res[use_tp] = (mean.clone(), var.clone(), co.clone(), min(ts))
    print("tp" if use_tp else "col", "min ms", round(min(ts), 4), flush=True)
for name, i in (("mean", 0), ("var", 1), ("coeffs", 2)):
    a_, b_ = res[0][i], res[1][i]
    print(name, "max rel diff", float((a_ - b_).abs().max() / a_.abs().max()), "nan", int(torch.isnan(b_).sum()))
