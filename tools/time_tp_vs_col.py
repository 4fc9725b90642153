#!/usr/bin/env python
"""Dev tool: thread-per-tile kernel (variant 3) against the lane-parallel column kernel (variant 4)
over nn_count, kernels and feature counts: throughput and agreement with the generic kernel."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200 import ops  # noqa: E402
from muygpys_b200.neighbors import NN_Wrapper  # noqa: E402

rng = np.random.default_rng(0)
n, b = 500_000, 100_000
out = {}
for d, kid in ((2, 2), (1, 1), (3, 3), (2, 4)):
    x = torch.as_tensor(rng.uniform(size=(n, d))).cuda()
    y = torch.as_tensor(rng.normal(size=n)).cuda()
    q = torch.as_tensor(rng.uniform(size=(b, d))).cuda()
    for k in (7, 10, 14, 20, 22, 30, 38, 46, 50, 54, 62):
        if d != 2 and k not in (7, 14, 30, 50, 62):
            continue
        nn, _ = NN_Wrapper(x, k).get_nns(q)
        res = {}
        for variant in (1, 4, 3):
            ops.set_fused_variant(variant)
            bb = b if variant != 1 else 5_000
            f = lambda: ops.fused_posterior(x, q[:bb], None, nn[:bb], y, kernel_id=kid, metric_id=0,  # noqa: E731
                                            length_scale=0.05 if d < 3 else 0.2, noise=1e-3)
            r = f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                f()
            e1.record()
            torch.cuda.synchronize()
            res[variant] = (r, round(bb / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e6, 1))
        ref = res[1][0]
        errs = []
        for variant in (4, 3):
            r = res[variant][0]
            em = float((r["mean"][:5000] - ref["mean"]).abs().max() / ref["mean"].abs().max())
            ev = float((r["var"][:5000] - ref["var"]).abs().max() / ref["var"].abs().max())
            errs.append(max(em, ev))
        out[f"d{d}_kern{kid}_k{k}"] = {"col_M/s": res[4][1], "tp_M/s": res[3][1], "err_col": errs[0], "err_tp": errs[1]}
        print(f"d{d}_kern{kid}_k{k}", out[f"d{d}_kern{kid}_k{k}"], flush=True)
ops.set_fused_variant(0)
print(json.dumps(out))
