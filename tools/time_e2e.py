#!/usr/bin/env python
"""Where the end-to-end time goes: H2D alone, kernel alone, host-buffer pipeline -- dev tool."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(2)
n, b, k = 1_000_000, 100_000, 50
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
grid = ops.KnnGrid(x)
nn = grid.query(q, k)[0]
nn_pin = nn.cpu().pin_memory()
nn_dev = torch.empty_like(nn)
mean_pin = torch.empty((b, 1), dtype=torch.float64).pin_memory()
var_pin = torch.empty((b,), dtype=torch.float64).pin_memory()
kw = dict(kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3)
out = {}
out["h2d_40MB_ms"] = timeit(lambda: nn_dev.copy_(nn_pin, non_blocking=True))
out["kernel_ms"] = timeit(lambda: ops.fused_posterior(x, q, None, nn, y, **kw))
out["host_pipeline_ms"] = timeit(lambda: ops.fused_posterior_host(x, q, None, nn_pin, y, **kw))
out["host_pipeline_with_d2h_ms"] = timeit(lambda: ops.fused_posterior_host(
    x, q, None, nn_pin, y, mean_host=mean_pin, var_host=var_pin, **kw))
nn_pin32 = nn.cpu().to(torch.int32).pin_memory()
out["host_pipeline_int32_ms"] = timeit(lambda: ops.fused_posterior_host(x, q, None, nn_pin32, y, **kw))
out["host_pipeline_int32_with_d2h_ms"] = timeit(lambda: ops.fused_posterior_host(
    x, q, None, nn_pin32, y, mean_host=mean_pin, var_host=var_pin, **kw))
print(json.dumps({k_: round(v, 4) for k_, v in out.items()}))
