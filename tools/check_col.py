#!/usr/bin/env python
"""Dev tool: column-direct kernel (variant 4) against the tile kernel (variant 2) on a sweep of
shapes, then timing on the C2 shape with real neighbours."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200 import ops

rng = np.random.default_rng(7)
worst = 0.0
fails = []
for d in (2, 1, 3):
    n, b = 4000, 777
    x = torch.as_tensor(rng.uniform(size=(n, d))).cuda()
    y = torch.as_tensor(rng.normal(size=n)).cuda()
    q = torch.as_tensor(rng.uniform(size=(b, d))).cuda()
    for k in list(range(7, 63)) if d == 2 else (7, 14, 15, 16, 30, 31, 38, 39, 50, 54, 55, 62):
        nn, _ = ops.knn(x, q, k)
        for kid, mid, ls in ((2, 0, 0.3), (1, 0, 0.3), (3, 0, [0.3, 0.5, 0.2][:d] if d > 1 else 0.3), (4, 0, 0.2), (0, 1, 0.2)):
            if k not in (30, 50, 55) and kid != 2:
                continue
            kw = dict(kernel_id=kid, metric_id=mid, length_scale=ls, noise=1e-3, scale=1.7,
                      want_yky=True, want_status=True)
            ops.set_fused_variant(2)
            ref = ops.fused_posterior(x, q, None, nn, y, **kw)
            ops.set_fused_variant(3)
            got = ops.fused_posterior(x, q, None, nn, y, **kw)
            for name in ("mean", "var", "yky"):
                r_, g_ = ref[name].flatten(), got[name].flatten()
                err = float((r_ - g_).abs().max() / r_.abs().max())
                worst = max(worst, err)
                if not err < 1e-11:
                    fails.append((d, k, kid, name, err))
            if int(got["status"].sum()) != 0:
                fails.append((d, k, kid, "status", int(got["status"].sum())))
print("sweep worst rel err", worst, "fails", fails[:10], len(fails))

# LOO-style call: query_x aliases train_x, batch indices
n, b, k = 200_000, 10_000, 50
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(np.sin(4 * x.cpu().numpy()[:, 0]) + rng.normal(size=n) * 0.05).cuda()
bi = torch.as_tensor(np.sort(rng.choice(n, b, replace=False))).cuda()
nn, _ = ops.knn(x, x[bi], k + 1)
nn = nn[:, 1:].contiguous()
kw = dict(kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3, want_yky=True)
ops.set_fused_variant(2); ref = ops.fused_posterior(x, x, bi, nn, y, **kw)
ops.set_fused_variant(3); got = ops.fused_posterior(x, x, bi, nn, y, **kw)
for name in ("mean", "var", "yky"):
    print("loo", name, float((ref[name] - got[name]).abs().max() / ref[name].abs().max()))

# timing, C2 shape
n, b, k = 1_000_000, 100_000, 50
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
from muygpys_b200.neighbors import NN_Wrapper
nn, _ = NN_Wrapper(x, k).get_nns(q)
nn = torch.as_tensor(nn).cuda() if not torch.is_tensor(nn) else nn
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = {}
for variant in (2, 3):
    ops.set_fused_variant(variant)
    f = lambda: ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3)
    for _ in range(3): f()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    res[f"v{variant}"] = {"ms": float(np.mean(ts)), "min_ms": min(ts), "Mnbhd_s": b / np.mean(ts) / 1e3}
for kk in (30, 40, 46, 54, 62):
    nn2 = torch.randint(0, n, (b, kk), device="cuda")
    for variant in (2, 3):
        ops.set_fused_variant(variant)
        f = lambda: ops.fused_posterior(x, q, None, nn2, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3)
        f(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): f()
        e.record(); torch.cuda.synchronize()
        res[f"k{kk}_v{variant}"] = round(b / (a.elapsed_time(e) / 5) / 1e3, 1)
print(json.dumps(res))
