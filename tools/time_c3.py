#!/usr/bin/env python
"""Device timing of the d > 8 fused path (C3 shape and a few others) -- dev tool."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
g = torch.Generator(device="cuda").manual_seed(0)
for name, (n, b, k, d, r) in {"c3_784": (60000, 10000, 30, 784, 10), "d40_k50": (200000, 50000, 50, 40, 1),
                              "d16_k100": (100000, 10000, 100, 16, 1), "d9_k30": (100000, 100000, 30, 9, 1)}.items():
    x = torch.rand((n, d), device="cuda", dtype=torch.float64, generator=g)
    y = torch.randn((n, r), device="cuda", dtype=torch.float64, generator=g)
    q = torch.rand((b, d), device="cuda", dtype=torch.float64, generator=g)
    nn = torch.randint(0, n, (b, k), device="cuda", generator=g)
    out[name] = timeit(lambda: ops.fused_posterior(x, q, None, nn, y, kernel_id=0, metric_id=1,
                                                   length_scale=float(np.sqrt(d)), noise=1e-3))
print(json.dumps(out))
# same C3 shape with the gathered rows confined to an L2-resident subset / to sorted runs
n, b, k, d, r = 60000, 10000, 30, 784, 10
x = torch.rand((n, d), device="cuda", dtype=torch.float64, generator=g)
y = torch.randn((n, r), device="cuda", dtype=torch.float64, generator=g)
q = torch.rand((b, d), device="cuda", dtype=torch.float64, generator=g)
extra = {}
for name, hi in (("l2_resident_4k_rows", 4000), ("hbm_60k_rows", n)):
    nn = torch.randint(0, hi, (b, k), device="cuda", generator=g)
    extra[name] = timeit(lambda: ops.fused_posterior(x, q, None, nn, y, kernel_id=0, metric_id=1,
                                                     length_scale=28.0, noise=1e-3))
print(json.dumps(extra))
