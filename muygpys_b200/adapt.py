"""Read the hot-path description of a model off a MuyGPS object -- ours OR the reference's.

The fused entry points (`*_from_indices`, `make_fused_loo_crossval_fn`, `regress_any`) need five
facts about a model: which covariance function, which metric, the length scale(s), the nugget
and the scale.  `ModelSpec.of(muygps)` extracts them by duck typing from

* a genuine `MuyGPyS.gp.MuyGPS` (S/gp/muygps.py:28-110: `.kernel` an `RBF` / `Matern` with
  `.deformation` an `Isotropy` / `Anisotropy` over the `l2` / `F2` metric, `.noise` a
  `HomoscedasticNoise` / `HeteroscedasticNoise` / `NullNoise`, `.scale` a `FixedScale` /
  `AnalyticScale`), so that a reference user only changes an import to get the one-launch path;
* the mirror objects of this package (`muygpys_b200.gp.MuyGPS`).

Nothing here imports MuyGPyS: the reference classes are recognised by name and by their
public accessors, hyperparameter overrides follow the reference's keyword protocol
(`length_scale` | `length_scale0..`, `noise`; S/gp/hyperparameter/scalar.py:300-330,
vector.py:76-125).
"""

from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L

_SMOOTHNESS_IDS = {0.5: L.KERNEL_MATERN_05, 1.5: L.KERNEL_MATERN_15, 2.5: L.KERNEL_MATERN_25,
                   math.inf: L.KERNEL_MATERN_INF}


def _metric_id(metric) -> int:
    mid = getattr(metric, "metric_id", None)
    if mid is not None:
        return int(mid)
    # reference MetricFn (S/gp/deformation/metric.py:21-66): tell l2 from F2 by evaluating it
    probe = float(np.asarray(metric(np.array([[3.0, 4.0]]))).reshape(-1)[0])
    if abs(probe - 5.0) < 1e-12:
        return L.METRIC_L2
    if abs(probe - 25.0) < 1e-12:
        return L.METRIC_F2
    raise NotImplementedError(f"unknown metric {metric!r}: the fused path knows l2 and F2")


class ModelSpec:
    """Kernel / metric / length scale / nugget / scale of one MuyGPS object."""

    def __init__(self, muygps):
        self.muygps = muygps
        kernel = muygps.kernel
        deformation = kernel.deformation
        kname = type(kernel).__name__
        kid = getattr(kernel, "kernel_id", None)
        if isinstance(kid, int) and kid >= 0:  # mirror objects carry the id
            self.kernel_id = kid
        elif kname == "RBF":
            self.kernel_id = L.KERNEL_RBF
        elif kname == "Matern":
            sm = kernel.smoothness
            nu = float(sm()) if callable(sm) else float(sm)
            fixed = sm.fixed() if hasattr(sm, "fixed") else True
            if not fixed or nu not in _SMOOTHNESS_IDS:
                raise NotImplementedError(
                    "the fused path builds the closed-form Matern kernels only: smoothness must "
                    f"be fixed at one of {sorted(_SMOOTHNESS_IDS)} (got {nu})")
            self.kernel_id = _SMOOTHNESS_IDS[nu]
        else:
            raise NotImplementedError(f"kernel {kname} is not on the fused path (RBF, Matern)")
        dname = type(deformation).__name__
        if dname not in ("Isotropy", "Anisotropy"):
            raise NotImplementedError(f"deformation {dname} is not on the fused path")
        self.anisotropic = dname == "Anisotropy"
        self.deformation = deformation
        self.metric_id = _metric_id(deformation.metric)
        noise = muygps.noise
        nname = type(noise).__name__
        self.heteroscedastic = nname == "HeteroscedasticNoise"
        self._noise = noise
        self._null_noise = nname == "NullNoise"
        scale = muygps.scale
        self._scale = scale
        self.analytic = type(scale).__name__ == "AnalyticScale"
        if type(scale).__name__ == "DownSampleScale":
            raise NotImplementedError("DownSampleScale (stochastic) is not on the fused path")
        self.iteration_count = int(getattr(scale, "iteration_count", 1))
        self._own_metric = self.metric_id
        self._tensor_metric = None

    @staticmethod
    def of(muygps) -> "ModelSpec":
        return muygps if isinstance(muygps, ModelSpec) else ModelSpec(muygps)

    # ---- hyperparameters, with the optimiser's keyword overrides -------------------------
    def length_scales(self, **theta) -> List[float]:
        ls = self.deformation.length_scale
        if hasattr(ls, "resolve"):  # mirror objects
            return [float(v) for v in ls.resolve(theta)]
        if self.anisotropic:
            vals = ls(**{k: v for k, v in theta.items() if k.startswith("length_scale")})
            return [float(v) for v in np.asarray(vals).reshape(-1)]
        if "length_scale" in theta:
            return [float(theta["length_scale"])]
        return [float(ls())]

    def length_scale_arg(self, **theta):
        ls = self.length_scales(**theta)
        if self._tensor_metric is not None and self._tensor_metric != self._own_metric:
            # MultivariateMuyGPS feeds every model the tensor of models[0]'s deformation; an
            # isotropic model then applies ITS metric's length-scale rule (x / l for l2, x / l^2
            # for F2; S/gp/deformation/metric.py:241,264) to the OTHER metric's tensor.  The
            # same factor through the tensor metric's rule needs this effective length scale.
            ell = ls[0]
            factor = 1.0 / ell if self._own_metric == L.METRIC_L2 else 1.0 / (ell * ell)
            return 1.0 / factor if self._tensor_metric == L.METRIC_L2 else factor ** -0.5
        return ls if self.anisotropic else ls[0]

    def seen_through(self, tensor_metric_id: int) -> "ModelSpec":
        """This model evaluated on distance tensors of another metric (see length_scale_arg)."""
        import copy

        if tensor_metric_id == self.metric_id or self.anisotropic:
            return self
        view = copy.copy(self)
        view._own_metric = self.metric_id
        view._tensor_metric = tensor_metric_id
        view.metric_id = tensor_metric_id
        return view

    def noise(self, override: Optional[float] = None):
        """Python float (homoscedastic / null) or a (b,k) tensor / array (heteroscedastic)."""
        if self.heteroscedastic:
            if hasattr(self._noise, "value"):  # mirror objects: already a device tensor
                return self._noise.value(None)
            return self._noise._val if hasattr(self._noise, "_val") else self._noise()
        if self._null_noise:
            return 0.0
        if override is not None:
            return float(override)
        if hasattr(self._noise, "value"):  # mirror objects
            return float(self._noise.value(None))
        return float(self._noise())

    def scale(self) -> float:
        val = self._scale()
        return float(np.asarray(val.cpu() if isinstance(val, torch.Tensor) else val).reshape(-1)[0])

    def sigma_from_mean_quadratic_form(self, sigma0: float) -> float:
        """AnalyticScale's iteration scale <- (scale + f(scale K)) / 2 with f(cK) = f(K) / c
        (S/gp/hyperparameter/scale.py:205-217)."""
        s = float(sigma0)
        for _ in range(1, self.iteration_count if self.analytic else 1):
            s = 0.5 * (s + sigma0 / s)
        return s

    def set_scale(self, value: float) -> None:
        self._scale._set(value)
        if hasattr(self.muygps, "_make"):
            self.muygps._make()
