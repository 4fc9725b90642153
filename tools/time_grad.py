#!/usr/bin/env python
"""Dev tool: cost of an analytic-gradient evaluation vs a plain objective evaluation, and
L-BFGS-B wall time with finite differences vs the analytic gradient (2 length scales)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200.examples.from_indices import optimize_from_indices
from muygpys_b200.gp import MuyGPS
from muygpys_b200.gp.deformation import Anisotropy, Isotropy, l2
from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter, VectorParameter
from muygpys_b200.gp.kernels import Matern
from muygpys_b200.gp.noise import HomoscedasticNoise
from muygpys_b200.neighbors import NN_Wrapper
from muygpys_b200.optimize.loss import lool_fn, mse_fn
from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn, make_fused_loo_value_and_grad_fn

rng = np.random.default_rng(4)
n, b, k = int(os.environ.get("N", 1_000_000)), 10_000, int(os.environ.get("K", 50))
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
xn = x.cpu().numpy()
y = torch.as_tensor(np.sin(4 * xn[:, 0]) + np.cos(3 * xn[:, 1]) + 0.3 * np.sin(11 * xn[:, 0] * xn[:, 1]) + 0.05 * rng.normal(size=n)).cuda()
nbrs = NN_Wrapper(x, k)
bi = torch.as_tensor(np.sort(rng.choice(n, b, replace=False))).cuda()
bnn, _ = nbrs.get_batch_nns(bi)
out = {}
for name, model, theta in (
    ("iso_m15", MuyGPS(kernel=Matern(smoothness=Parameter(1.5), deformation=Isotropy(l2, Parameter(0.1, (0.01, 1.0)))),
                       noise=HomoscedasticNoise(1e-3), scale=AnalyticScale()), {"length_scale": 0.12}),
    ("aniso_m25", MuyGPS(kernel=Matern(smoothness=Parameter(2.5), deformation=Anisotropy(l2, VectorParameter(
        Parameter(0.1, (0.01, 1.0)), Parameter(0.5, (0.05, 5.0))))), noise=HomoscedasticNoise(1e-3), scale=AnalyticScale()),
     {"length_scale0": 0.12, "length_scale1": 0.4})):
    obj = make_fused_loo_crossval_fn(model, lool_fn, bi, bnn, x, y)
    vg = make_fused_loo_value_and_grad_fn(model, lool_fn, bi, bnn, x, y)
    for fn, tag in ((obj, "plain"), (vg, "grad")):
        for _ in range(5): fn(**theta)
        t0 = time.perf_counter()
        for _ in range(50): fn(**theta)
        out[f"{name}_{tag}_us"] = (time.perf_counter() - t0) / 50 * 1e6
    for tag, kw in (("fd", {}), ("grad", {"use_gradient": True})):
        t0 = time.perf_counter()
        calls = [0]
        import muygpys_b200.ops as _ops
        _orig = _ops.FusedLoo.launch
        def _count(self, *a, **k2):
            calls[0] += 1
            return _orig(self, *a, **k2)
        _ops.FusedLoo.launch = _count
        opt = optimize_from_indices(model, bi, bnn, x, y, loss_fn=lool_fn, **kw)
        _ops.FusedLoo.launch = _orig
        out[f"{name}_lbfgsb_{tag}_s"] = time.perf_counter() - t0
        out[f"{name}_lbfgsb_{tag}_launches"] = calls[0]
        out[f"{name}_lbfgsb_{tag}_opt"] = [float(v) for v in opt.get_opt_params()[1]]
print(json.dumps(out))
