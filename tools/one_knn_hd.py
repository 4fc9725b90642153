#!/usr/bin/env python
"""Two high-dimensional KNN calls (60k x 784 train, 4096 queries, k = 30) for profiling -- dev tool."""
import sys

import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((60000, 784), device="cuda", dtype=torch.float64, generator=g)
qs = torch.randn((4096, 784), device="cuda", dtype=torch.float64, generator=g)
for _ in range(2):
    ops.knn(x, qs, 30)
torch.cuda.synchronize()
