#!/usr/bin/env python
"""Device timing of the high-dimensional exact KNN (C3 shape) -- dev tool."""
import json
import sys

import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
out = {}
for name, (n, q, d, k) in {"c3_10k": (60000, 10000, 784, 30), "c3_1k": (60000, 1000, 784, 30),
                           "d40_50k": (200000, 50000, 40, 50)}.items():
    x = torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g)
    qs = torch.randn((q, d), device="cuda", dtype=torch.float64, generator=g)
    ops.knn(x, qs, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.knn(x, qs, k)
    e1.record()
    torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / 3
print(json.dumps(out))
