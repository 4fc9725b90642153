#!/usr/bin/env python
"""Small launches of every kernel family for compute-sanitizer (memcheck / racecheck /
synccheck): thread-per-tile (plain, LOO / looph epilogue, back substitution and gradient;
T <= 8 and T = 13), int32 host-index pipeline, column-direct with
lane-parallel steps (coefficients, gradient, variant 4), tile (register and shared-memory
factor, Gram d > 8), generic, host pipeline, KNN (grid, small-d, tiled, Gram pre-filter), fast
mean, losses, staged ops, label mask."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200 import ops, _lib as L

rng = np.random.default_rng(0)
def dev(a): return torch.as_tensor(np.ascontiguousarray(a)).cuda()

n, b = 3000, 96
for d, k, r in ((2, 50, 1), (1, 30, 1), (3, 23, 1), (2, 100, 1), (2, 50, 3), (20, 30, 4), (2, 130, 1)):
    x, q = rng.uniform(size=(n, d)), rng.uniform(size=(b, d))
    y = rng.normal(size=(n, r))
    xd, qd, yd = dev(x), dev(q), dev(y)
    nn, _ = ops.knn(xd, qd, k)
    for variant in (0, 4, 2, 1) if (r == 1 and k <= 62 and d <= 3) else (0, 2, 1):
        ops.set_fused_variant(variant)
        ops.fused_posterior(xd, qd, None, nn, yd, kernel_id=2, metric_id=0, length_scale=0.3,
                            noise=1e-3, want_yky=True, want_status=True,
                            want_coeffs=(variant not in (0, 4)))
    ops.set_fused_variant(0)
    if r == 1 and k <= 102 and d <= 3:
        bi = dev(np.sort(rng.choice(n, b, replace=False)))
        bnn, _ = ops.knn(xd, xd[bi], k + 1)
        bnn = bnn[:, 1:].contiguous()
        for want_grad in (False, True):  # (GRAD instantiations: every tile count, T = 13 at k = 100)
            for loss_id in (L.LOSS_LOOL, L.LOSS_LOOPH):
                loo = ops.FusedLoo(xd, yd[:, 0].contiguous(), bi, bnn, kernel_id=2, metric_id=0,
                                   loss_id=loss_id, boundary_scale=3.0, want_grad=want_grad)
                for _ in range(2):
                    loo.record(loo.launch(0.3, 1e-3, scale=0.8))
        # fast-mean coefficients from the back substitution (k = 100: GRAD, T = 13)
        ops.fused_posterior(xd, xd, bi, bnn, yd, kernel_id=2, metric_id=0, length_scale=0.3,
                            noise=1e-3, want_coeffs=True)
# host pipeline
x, q, y = rng.uniform(size=(20000, 2)), rng.uniform(size=(20000, 2)), rng.normal(size=20000)
xd, qd, yd = dev(x), dev(q), dev(y)
from muygpys_b200.neighbors import NN_Wrapper
nb = NN_Wrapper(xd, 50)
nn, _ = nb.get_nns(qd)
out = ops.fused_posterior_host(xd, qd, torch.arange(20000), nn.cpu(), yd, kernel_id=2, metric_id=0,
                               length_scale=0.1, noise=1e-3)
torch.cuda.synchronize()
# int32 host indices: widened on the device chunk by chunk (odd k: unaligned chunk offsets)
for kk in (50, 13):
    nn_k = nn[:, :kk].contiguous()
    out = ops.fused_posterior_host(xd, qd, None, nn_k.cpu().to(torch.int32), yd, kernel_id=2,
                                   metric_id=0, length_scale=0.1, noise=1e-3)
    torch.cuda.synchronize()
nb.get_batch_nns(torch.arange(0, 20000, 7).cuda())
# high-d KNN paths
x, q = rng.normal(size=(5000, 40)), rng.normal(size=(300, 40))
ops.knn(dev(x), dev(q), 30); ops.knn(dev(x), dev(q), 100); ops.knn(dev(x[:1000]), dev(q), 10)
# fast mean, staged ops, losses, label mask
k = 20
x, q, y = rng.uniform(size=(2000, 2)), rng.uniform(size=(100, 2)), rng.normal(size=2000)
xd, qd, yd = dev(x), dev(q), dev(y)
nn, _ = ops.knn(xd, qd, k)
co = ops.fused_posterior(xd, xd, None, ops.knn(xd, xd, k)[0], yd, kernel_id=1, metric_id=0,
                         length_scale=0.2, noise=1e-3, want_mean=False, want_var=False, want_coeffs=True)["coeffs"]
ops.fast_mean(xd, qd, None, nn, nn[:, 0].contiguous(), co, kernel_id=1, metric_id=0, length_scale=0.2)
pw = ops.pairwise_dists(0, xd, nn); cw = ops.crosswise_dists(0, qd, xd, torch.arange(100).cuda(), nn)
Kin = ops.perturb(ops.kernel_apply(2, pw, 5.0), 1e-3); Kc = ops.kernel_apply(2, cw, 5.0)
s = ops.solve(Kin, Kc, yd[nn], want_mean=True, want_var=True, want_yky=True, want_coeffs=True)
ops.loss_partials(L.LOSS_LOOL, s["mean"][:, 0].contiguous(), yd[:100].contiguous(), var=s["var"], yky=s["yky"])
ops.pairwise_diffs(xd, nn); ops.crosswise_diffs(qd, xd, torch.arange(100).cuda(), nn)
ops.nn_label_mask(dev((rng.integers(0, 3, 2000)).astype(np.float64)), nn)
torch.cuda.synchronize()
print("sanitize_run ok")
