#!/usr/bin/env python
"""Dev tool: grid KNN (warp-per-query kernel; MGP_KNN_GRID=thread|warp forces one kernel) on
the C2 and C5-like shapes: time, and agreement of the two kernels when run in two processes."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

rng = np.random.default_rng(5)
out = {}
for name, n, q, k in (("c2_1M_100k_k50", 1_000_000, 100_000, 50), ("c5like_10M_1M_k50", 10_000_000, 1_000_000, 50),
                      ("c4like_10M_10k_k101", 10_000_000, 10_000, 101)):
    x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
    qq = torch.as_tensor(rng.uniform(size=(q, 2))).cuda()
    grid = ops.KnnGrid(x)
    for _ in range(2):
        nn, d2 = grid.query(qq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        nn, d2 = grid.query(qq, k)
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = round(e0.elapsed_time(e1) / 5, 4)
    out[name + "_checksum"] = [int(nn.sum()), float(d2.sum())]
print(json.dumps(out))
