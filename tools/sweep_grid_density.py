#!/usr/bin/env python
"""Grid-KNN query time against points per cell (k = 50, 2-D uniform) -- dev tool."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402

rng = np.random.default_rng(5)
x = torch.as_tensor(rng.uniform(size=(10_000_000, 2))).cuda()
q = torch.as_tensor(rng.uniform(size=(1_000_000, 2))).cuda()
out = {}
for k in (10, 50, 100):
    for ppc in (4.0, 8.0, 12.0, 20.0, 35.0):
        ops.KnnGrid.POINTS_PER_CELL = ppc
        grid = ops.KnnGrid(x)
        grid.query(q, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            grid.query(q, k)
        b.record()
        torch.cuda.synchronize()
        out[f"k{k}_ppc{int(ppc)}"] = round(a.elapsed_time(b) / 3, 2)
        del grid
print(json.dumps(out))
