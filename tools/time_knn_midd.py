#!/usr/bin/env python
"""Exact sweep vs Gram pre-filter for moderate feature counts -- dev tool (set MGP_GRAM_KNN_MIN_D)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
out = {}
for d in (4, 5, 6, 8, 9, 16, 32):
    n, q, k = 200000, 20000, 50
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
    out[f"d{d}"] = round(e0.elapsed_time(e1) / 3, 2)
print(json.dumps(out))
