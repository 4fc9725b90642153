#!/usr/bin/env python
"""cProfile of the fused LOO-mse objective at the bench's shape (b = 10 k, k = 50) -- dev tool."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from muygpys_b200.gp import MuyGPS  # noqa: E402
from muygpys_b200.gp.deformation import Isotropy, l2  # noqa: E402
from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter  # noqa: E402
from muygpys_b200.gp.kernels import Matern  # noqa: E402
from muygpys_b200.gp.noise import HomoscedasticNoise  # noqa: E402
from muygpys_b200.neighbors import NN_Wrapper  # noqa: E402
from muygpys_b200.optimize.loss import mse_fn  # noqa: E402
from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn  # noqa: E402

rng = np.random.default_rng(0)
n, b, k = 1_000_000, 10_000, 50
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda()
y = torch.as_tensor(rng.normal(size=n)).cuda()
model = MuyGPS(kernel=Matern(smoothness=Parameter(1.5),
                             deformation=Isotropy(l2, length_scale=Parameter(0.1, (0.01, 1.0)))),
               noise=HomoscedasticNoise(1e-3), scale=AnalyticScale())
nbrs = NN_Wrapper(x, k, nn_method="exact")
bi = torch.as_tensor(rng.choice(n, b, replace=False)).cuda()
bnn, _ = nbrs.get_batch_nns(bi)
obj = make_fused_loo_crossval_fn(model, mse_fn, bi, bnn, x, y)
for v in (0.1, 0.11, 0.12):
    obj(length_scale=v)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 300
for i in range(N):
    obj(length_scale=0.1 + 1e-4 * i)
dt = time.perf_counter() - t0
print(f"{1e6 * dt / N:.1f} us per evaluation")
pr = cProfile.Profile()
pr.enable()
for i in range(N):
    obj(length_scale=0.1 + 1e-4 * i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
