#!/usr/bin/env python
"""Dev tool: fast-mean apply and coefficient precompute rates on a C5-like shape (10 M train)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from muygpys_b200 import ops
from muygpys_b200.neighbors import NN_Wrapper
g = torch.Generator(device="cuda").manual_seed(5)
n, t, k = 10_000_000, 2_000_000, 50
x = torch.rand((n, 2), generator=g, device="cuda", dtype=torch.float64)
y = torch.sin(4 * x[:, 0]) + torch.cos(3 * x[:, 1])
q = torch.rand((t, 2), generator=g, device="cuda", dtype=torch.float64)
nb = NN_Wrapper(x, k)
nn, _ = nb._query(q, k)
closest = torch.unique(nn[:, 0])
cnn, _ = nb._query(x[closest], k)
kw = dict(kernel_id=1, metric_id=0, length_scale=0.1, noise=1e-3)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
out = {}
for variant in (0, 2):
    ops.set_fused_variant(variant)
    ms, co = timed(lambda: ops.fused_posterior(x, x, closest, cnn, y, want_mean=False, want_var=False, want_coeffs=True, **kw)["coeffs"])
    out[f"coeffs_variant{variant}_Mrows_s"] = closest.numel() / ms / 1e3
    out[f"coeffs_variant{variant}_ms"] = ms
ops.set_fused_variant(0)
slot = torch.searchsorted(closest, nn[:, 0].contiguous())
nn_fast = cnn[slot]
ms, fm = timed(lambda: ops.fast_mean(x, q, None, nn_fast, slot, co, kernel_id=1, metric_id=0, length_scale=0.1))
out["fast_mean_ms"] = ms; out["fast_mean_Mpts_s"] = t / ms / 1e3; out["fast_mean_GBs"] = t * 1632 / ms / 1e6
print(json.dumps(out))
