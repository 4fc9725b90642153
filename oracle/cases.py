"""Seeded synthetic cases shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE.

Each case is a scaled-down instance of one BASELINE.json config (C1..C5, see
SURVEY.md section 8d for the generating formulas) plus a few edge cases.  Inputs
are regenerated from the seed (numpy's `default_rng` streams are stable), so the
golden files only need to hold the reference's *outputs*.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple, Union

import numpy as np

from . import numpy_oracle as O


@dataclass
class Case:
    name: str
    seed: int
    n: int  # train count
    t: int  # test count
    d: int  # feature count
    k: int  # nn_count
    r: int  # response count
    kernel_id: int
    metric_id: int
    length_scale: Union[float, Tuple[float, ...]]
    noise: float = 1e-3
    batch: int = 0  # LOO batch size (0 = no training objective)
    losses: Tuple[str, ...] = ()
    flat_features: bool = False  # 1-D feature array (reference reshapes to (n,1))
    hetero: bool = False
    fast: bool = False
    notes: str = ""
    loss_kwargs: dict = field(default_factory=dict)

    @property
    def anisotropic(self) -> bool:
        return not np.isscalar(self.length_scale)


CASES = [
    Case("c1_rbf_1d", 1, 2000, 100, 1, 30, 1, O.KERNEL_RBF, O.METRIC_F2, 0.05,
         batch=120, losses=("mse",), flat_features=False,
         notes="C1: univariate sine, RBF/Isotropy(F2), k=30"),
    Case("c1_rbf_flat", 11, 600, 40, 1, 12, 1, O.KERNEL_RBF, O.METRIC_F2, 0.05,
         flat_features=True, notes="1-D feature arrays (tensors/numpy.py:53-54,65-66)"),
    Case("c2_m15_2d", 2, 4000, 200, 2, 50, 1, O.KERNEL_MATERN_15, O.METRIC_L2, 0.1,
         batch=300, losses=("mse", "lool", "looph", "pseudo_huber"),
         notes="C2: Heaton-shaped 2-D, Matern 3/2, k=50, LOO-mse + analytic scale"),
    Case("c3_rbf_784", 3, 1500, 50, 784, 30, 10, O.KERNEL_RBF, O.METRIC_F2, 28.0,
         batch=100, losses=("cross_entropy", "mse"),
         notes="C3: MNIST-shaped, r=10, cross-entropy"),
    Case("c4_m25_aniso", 4, 5000, 64, 2, 100, 1, O.KERNEL_MATERN_25, O.METRIC_L2,
         (0.1, 0.5), batch=100, losses=("lool", "mse"),
         notes="C4: anisotropic Matern 5/2, k=100, lool"),
    Case("c5_m05_2d", 5, 3000, 200, 2, 50, 1, O.KERNEL_MATERN_05, O.METRIC_L2, 0.1,
         fast=True, notes="C5: Matern 1/2, mean+var and fast posterior mean"),
    Case("minf_3d", 6, 1200, 64, 3, 20, 2, O.KERNEL_MATERN_INF, O.METRIC_L2, 0.4,
         batch=80, losses=("mse",), fast=True, notes="Matern nu=inf, r=2, d=3"),
    Case("m15_hetero", 7, 1000, 60, 2, 16, 1, O.KERNEL_MATERN_15, O.METRIC_L2, 0.2,
         hetero=True, notes="heteroscedastic nugget (noise/numpy.py:56-67)"),
    Case("rbf_aniso_f2", 8, 1500, 48, 4, 24, 3, O.KERNEL_RBF, O.METRIC_F2,
         (0.5, 1.0, 2.0, 0.25), batch=64, losses=("mse",),
         notes="anisotropic RBF on F2 (tests/kernels.py:587-690 shape)"),
    Case("tiny_k3", 9, 40, 7, 2, 3, 1, O.KERNEL_MATERN_25, O.METRIC_L2, 0.7,
         batch=5, losses=("mse", "lool"), notes="ragged tiny sizes"),
]


def by_name(name: str) -> Case:
    for c in CASES:
        if c.name == name:
            return c
    raise KeyError(name)


def make_data(case: Case):
    """Return dict(train_x, train_y, test_x[, hetero_noise]) for a case."""
    rng = np.random.default_rng(case.seed)
    n, t, d, r = case.n, case.t, case.d, case.r
    if case.name.startswith("c3"):
        centroids = rng.normal(0.0, 0.5, size=(r, d))
        lab_tr = rng.integers(0, r, size=n)
        lab_te = rng.integers(0, r, size=t)
        train_x = centroids[lab_tr] + rng.normal(size=(n, d))
        test_x = centroids[lab_te] + rng.normal(size=(t, d))
        train_y = -0.1 * np.ones((n, r))
        train_y[np.arange(n), lab_tr] = 0.9  # one-hot - 0.1 (S/_test/utils.py:143-144)
    else:
        train_x = rng.uniform(size=(n, d))
        test_x = rng.uniform(size=(t, d))
        if d == 1:
            f = np.sin(2 * np.pi * 4 * train_x[:, 0])
        elif case.anisotropic and d == 2:
            f = np.sin(2 * np.pi * train_x[:, 0] / 0.1 * 0.1) * np.cos(
                2 * np.pi * train_x[:, 1] / 0.5 * 0.1
            )
        else:
            x0, x1 = train_x[:, 0], train_x[:, 1]
            f = np.sin(4 * x0) + np.cos(3 * x1) + 0.3 * np.sin(11 * x0 * x1)
        cols = [f + 0.05 * rng.normal(size=n)]
        for j in range(1, r):
            cols.append(np.cos((j + 1) * f) + 0.05 * rng.normal(size=n))
        train_y = np.stack(cols, axis=1)
    out = dict(train_x=np.ascontiguousarray(train_x),
               train_y=np.ascontiguousarray(train_y),
               test_x=np.ascontiguousarray(test_x))
    if case.flat_features:
        out["train_x"] = out["train_x"][:, 0].copy()
        out["test_x"] = out["test_x"][:, 0].copy()
    if case.hetero:
        out["hetero_train_noise"] = rng.uniform(1e-4, 1e-2, size=n)
    if case.batch:
        out["batch_idx"] = np.sort(rng.choice(n, case.batch, replace=False)).astype(np.int64)
    return out
