#!/usr/bin/env python
"""Two fused launches at k = 100 (C4 shape) for profiling -- dev tool."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
n, b, k = 500_000, 20_000, 100
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
nn = torch.randint(0, n, (b, k), device="cuda")
for _ in range(2):
    ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3)
torch.cuda.synchronize()
