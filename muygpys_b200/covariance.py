"""Covariance functors `RBF` and `Matern` (S/gp/kernels/{rbf,matern,kernel_fn}.py).

`kernel(diffs, **theta)` maps a distance tensor (Isotropy) or a difference tensor
(Anisotropy) to covariances with one elementwise CUDA kernel; the length-scale
division is folded into that kernel instead of materialising a scaled copy.
Only the closed-form smoothness values the north star names are built
(nu in {1/2, 3/2, 5/2, inf}); general-nu Bessel kernels are out of scope.
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

from . import _lib as L
from . import ops
from ._arrays import fdev, like_input
from .deformation import F2, DeformationFn, Isotropy, l2
from .hyperparameter import Parameter, _Named


class KernelFn:
    kernel_id: int = -1

    def __init__(self, deformation: DeformationFn):
        self.deformation = deformation
        self._hyperparameters: Dict = {}
        self._make()

    def _make(self) -> None:
        self._hyperparameters = {}
        self.deformation.length_scale.populate(self._hyperparameters)

    def set_params(self, **kwargs) -> None:
        for name, val in kwargs.items():
            self._hyperparameters[name]._set_val(val)

    def get_opt_params(self) -> Tuple[List[str], List[float], List[Tuple[float, float]]]:
        names: List[str] = []
        params: List[float] = []
        bounds: List[Tuple[float, float]] = []
        self.deformation.length_scale.append_lists(names, params, bounds)
        return names, params, bounds

    def __call__(self, diffs, **kwargs):
        x = fdev(diffs)
        ls = self.deformation.length_scales(**kwargs)
        if self.deformation.anisotropic:
            if x.shape[-1] != len(ls):
                raise ValueError(
                    f"Difference tensor of shape {tuple(x.shape)} must have final dimension "
                    f"size of {len(ls)}"
                )
            x = ops.metric_reduce(self.deformation.metric.metric_id, x, length_scale=ls)
            pre = 1.0
        else:
            pre = self.deformation.metric.length_scale_factor(ls[0])
        return like_input(ops.kernel_apply(self.kernel_id, x, pre), diffs)

    def get_opt_fn(self):
        return self.__call__

    def Kout(self, **kwargs) -> float:
        """Prior variance at distance zero: 1 for RBF and Matern (rbf.py:113-114)."""
        return 1.0

    def __str__(self) -> str:
        return "\n".join(f"{k} : {p()} - {p.get_bounds()}" for k, p in self._hyperparameters.items())


class RBF(KernelFn):
    """exp(-d_F2 / (2 l^2)); expects the F2 metric like the reference's default."""

    kernel_id = L.KERNEL_RBF

    def __init__(self, deformation: DeformationFn = None):
        super().__init__(deformation or Isotropy(F2, Parameter(1.0)))


_SMOOTHNESS_IDS = {0.5: L.KERNEL_MATERN_05, 1.5: L.KERNEL_MATERN_15, 2.5: L.KERNEL_MATERN_25,
                   math.inf: L.KERNEL_MATERN_INF}


class Matern(KernelFn):
    """Matern covariance at a FIXED smoothness in {0.5, 1.5, 2.5, inf}."""

    def __init__(self, smoothness: Parameter = None, deformation: DeformationFn = None):
        self.smoothness = _Named("smoothness", smoothness or Parameter(0.5))
        super().__init__(deformation or Isotropy(l2, Parameter(1.0)))

    def _make(self) -> None:
        super()._make()
        self.smoothness.populate(self._hyperparameters)
        p = self.smoothness.param
        if not p.fixed() or p() not in _SMOOTHNESS_IDS:
            raise NotImplementedError(
                "muygpys_b200 builds the closed-form Matern kernels only: smoothness must be "
                f"fixed at one of {sorted(_SMOOTHNESS_IDS)} (got {p})"
            )
        self.kernel_id = _SMOOTHNESS_IDS[p()]

    def get_opt_params(self):
        names, params, bounds = super().get_opt_params()
        self.smoothness.append_lists(names, params, bounds)
        return names, params, bounds
