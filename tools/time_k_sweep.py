#!/usr/bin/env python
"""Fused posterior throughput against nn_count (d = 2, r = 1, Matern 3/2) -- dev tool."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
n, b = 500_000, 50_000
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
out = {}
for k in (20, 40, 50, 52, 56, 60, 68, 76, 84, 100, 120):
    nn = torch.randint(0, n, (b, k), device="cuda")
    f = lambda: ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1,  # noqa: E731
                                    noise=1e-3)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    out[f"k{k}"] = round(b / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e6, 2)
print(json.dumps(out), "(M neighbourhoods/s)")
