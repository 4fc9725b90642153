#!/usr/bin/env python
"""Quick device-only timing of the fused kernel on a C2-shaped batch (dev tool)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200 import ops

def main():
    n = int(os.environ.get("N", 1_000_000)); b = int(os.environ.get("B", 100_000))
    k = int(os.environ.get("K", 50)); kid = int(os.environ.get("KID", 2))
    rng = np.random.default_rng(2)
    x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
    y = torch.as_tensor(rng.normal(size=n)).cuda()
    q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
    # cheap locality-preserving fake neighbours: sort train by a grid key, take windows
    t0 = time.time()
    if os.environ.get("REAL_KNN", "0") == "1":
        nn, _ = ops.knn(x, q, k)
    else:
        nn = torch.randint(0, n, (b, k), device="cuda")
    torch.cuda.synchronize()
    print("knn/indices s", time.time() - t0)
    best=1e9
    ops.set_fused_variant(int(os.environ.get('VARIANT',0)))
    for rep in range(int(os.environ.get('REPS',12))):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.fused_posterior(x, q, None, nn, y, kernel_id=kid, metric_id=0, length_scale=0.1,
                                  noise=1e-3)
        e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e)
        best=min(best,ms)
    print(json.dumps({"best_ms": best, "nbhd_per_s": b / best * 1e3}))
main()
