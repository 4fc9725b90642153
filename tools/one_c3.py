#!/usr/bin/env python
"""Two launches of the d = 784 fused posterior (C3 shape) for profiling -- dev tool."""
import sys

import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
n, b, k, d, r = 60000, 10000, 30, 784, 10
x = torch.rand((n, d), device="cuda", dtype=torch.float64, generator=g)
y = torch.randn((n, r), device="cuda", dtype=torch.float64, generator=g)
q = torch.rand((b, d), device="cuda", dtype=torch.float64, generator=g)
nn = torch.randint(0, n, (b, k), device="cuda", generator=g)
for _ in range(2):
    ops.fused_posterior(x, q, None, nn, y, kernel_id=0, metric_id=1, length_scale=28.0, noise=1e-3)
torch.cuda.synchronize()
