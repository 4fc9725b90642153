#!/usr/bin/env python
"""Device timings of the non-headline BASELINE configs (C3, C4, C5 shapes) -- dev tool."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200 import ops

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

rng = np.random.default_rng(0)
out = {}
# C4: k=100 anisotropic Matern 5/2, batch 10k, d=2, n=10M
n = 10_000_000
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda(); y = torch.as_tensor(rng.normal(size=n)).cuda()
grid = ops.KnnGrid(x)
bi = torch.as_tensor(rng.choice(n, 10_000, replace=False)).cuda()
out["c4_knn_ms_10k_k101"] = timeit(lambda: grid.query(x[bi], 101))
nn = grid.query(x[bi], 101)[0][:, 1:].contiguous()
out["c4_fused_ms_10k_k100"] = timeit(lambda: ops.fused_posterior(x, x, bi, nn, y, kernel_id=3, metric_id=0, length_scale=[0.1, 0.5], noise=1e-3, want_yky=True))
# C5: Matern 1/2 k=50 on 10M train, 1M test
q = torch.as_tensor(rng.uniform(size=(1_000_000, 2))).cuda()
out["c5_knn_ms_1M_k50"] = timeit(lambda: grid.query(q, 50))
nn5 = grid.query(q, 50)[0]
out["c5_fused_ms_1M"] = timeit(lambda: ops.fused_posterior(x, q, None, nn5, y, kernel_id=1, metric_id=0, length_scale=0.1, noise=1e-3))
del x, y, q, nn5, grid
# C3: 60k x 784, 10k queries, k=30, r=10
n, t, d, r = 60_000, 10_000, 784, 10
x = torch.as_tensor(rng.normal(size=(n, d))).cuda(); q = torch.as_tensor(rng.normal(size=(t, d))).cuda()
y = torch.as_tensor(rng.normal(size=(n, r))).cuda()
t0 = time.perf_counter(); nn3, _ = ops.knn(x, q[:1000], 30); torch.cuda.synchronize()
out["c3_knn_s_1000q"] = time.perf_counter() - t0
nn3 = torch.randint(0, n, (t, 30), device="cuda")
out["c3_fused_ms_10k"] = timeit(lambda: ops.fused_posterior(x, q, None, nn3, y, kernel_id=0, metric_id=1, length_scale=28.0, noise=1e-3))
print(json.dumps(out))
