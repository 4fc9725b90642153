import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from muygpys_b200 import ops
from muygpys_b200.neighbors import NN_Wrapper
rng = np.random.default_rng(7)
n, b = 1_000_000, 20_000
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda(); y = torch.as_tensor(rng.normal(size=n)).cuda()
bi = torch.as_tensor(np.sort(rng.choice(n, b, replace=False))).cuda()
for k in (70, 86, 100):
    nn, _ = NN_Wrapper(x, k).get_batch_nns(bi)
    for variant in (0, 2):
        ops.set_fused_variant(variant)
        f = lambda: ops.fused_posterior(x, x, bi, nn, y, kernel_id=2, metric_id=0, length_scale=0.1, noise=1e-3, want_coeffs=True)
        for _ in range(3): o = f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); o = f(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
        print(k, "variant", variant, "ms", round(min(ts), 3), float(o["coeffs"].abs().sum()), flush=True)
ops.set_fused_variant(0)
