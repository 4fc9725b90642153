#!/usr/bin/env python
"""One exact grid-KNN query batch at the C5 shape (10M train, 1M queries, k = 50) -- dev tool."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

rng = np.random.default_rng(5)
x = torch.as_tensor(rng.uniform(size=(10_000_000, 2))).cuda()
q = torch.as_tensor(rng.uniform(size=(1_000_000, 2))).cuda()
grid = ops.KnnGrid(x)
torch.cuda.synchronize()
for _ in range(2):
    grid.query(q, 50)
torch.cuda.synchronize()
