"""`MuyGPS`: the reference's model object (S/gp/muygps.py:27-551) over CUDA kernels.

Two families of methods:

* the reference's staged API on materialised tensors -- `make_predict_tensors`,
  `make_train_tensors`, `kernel(...)`, `posterior_mean`, `posterior_variance`,
  `fast_coefficients`, `fast_posterior_mean`, `optimize_scale`,
  `get_opt_mean_fn` / `get_opt_var_fn` -- each backed by one K5 kernel;
* fused methods that take indices instead of `(b,k,k[,d])` tensors and run the
  whole neighbourhood pipeline in one K1 launch (`fused_regress`,
  `fused_fast_coefficients`, ...); `muygpys_b200.from_indices` exposes them under
  the reference's `*_from_indices` names.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import fused, ops
from ._arrays import fdev, idev, like_input
from .covariance import KernelFn
from .hyperparameter import AnalyticScale, FixedScale, ScaleFn
from .noise import HomoscedasticNoise, NoiseFn


def _squeeze_response(mean: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """(b,r) -> (b,) when the caller's targets carry no response axis."""
    return mean[:, 0] if targets.dim() == 1 else mean


class MuyGPS:
    def __init__(self, kernel: KernelFn, noise: NoiseFn = None, scale: ScaleFn = None):
        self.kernel = kernel
        self.noise = noise if noise is not None else HomoscedasticNoise(0.0, "fixed")
        self.scale = scale if scale is not None else FixedScale()
        self._make()

    def _make(self) -> None:
        self.kernel._make()

    # ---- bookkeeping -------------------------------------------------------
    def fixed(self) -> bool:
        for p in self.kernel._hyperparameters.values():
            if not p.fixed():
                return False
        return self.noise.fixed()

    def get_opt_params(self) -> Tuple[List[str], np.ndarray, np.ndarray]:
        names, params, bounds = self.kernel.get_opt_params()
        self.noise.append_lists(names, params, bounds)
        return names, np.array(params, dtype=np.float64), np.array(bounds, dtype=np.float64)

    def __str__(self) -> str:
        return f"MuyGPS(kernel={self.kernel}, noise={self.noise}, scale={self.scale})"

    # ---- staged API (materialised tensors) ----------------------------------
    def _solve(self, Kin, Kcross, Y, noise=None, **want):
        pK = ops.perturb(fdev(Kin), self.noise.value(noise))
        return ops.solve(pK, None if Kcross is None else fdev(Kcross),
                         None if Y is None else fdev(Y), self.kernel.Kout(), **want)

    def posterior_mean(self, Kin, Kcross, batch_nn_targets, **kwargs):
        """Kcross (Kin + eps)^-1 Y.  The nugget is added here, as in the reference
        (S/gp/mean.py:25), so an already perturbed Kin gets it twice."""
        y = fdev(batch_nn_targets)
        out = self._solve(Kin, Kcross, y, noise=kwargs.get("noise"), want_mean=True)["mean"]
        out = out[:, 0] if y.dim() == 2 else out
        return like_input(out, Kin, Kcross, batch_nn_targets)

    def _unscaled_variance(self, Kin, Kcross, **kwargs):
        return self._solve(Kin, Kcross, None, noise=kwargs.get("noise"), want_var=True)["var"]

    def posterior_variance(self, Kin, Kcross, **kwargs):
        """scale * (Kout - Kcross (Kin+eps)^-1 Kcross^T)   (S/gp/variance.py:22-52)."""
        scale = kwargs.pop("scale", self.scale())
        return like_input(self._unscaled_variance(Kin, Kcross, **kwargs) * scale, Kin, Kcross)

    def fast_coefficients(self, Kin, train_nn_targets_fast, **kwargs):
        y = fdev(train_nn_targets_fast)
        out = self._solve(Kin, None, y, noise=kwargs.get("noise"), want_coeffs=True)["coeffs"]
        out = out[:, :, 0] if y.dim() == 2 else out
        return like_input(out, Kin, train_nn_targets_fast)

    def fast_posterior_mean(self, Kcross, coeffs_tensor):
        c = fdev(coeffs_tensor)
        out = ops.rowdot(fdev(Kcross), c)
        out = out[:, 0] if c.dim() == 2 else out
        return like_input(out, Kcross, coeffs_tensor)

    def get_opt_mean_fn(self) -> Callable:
        return self.posterior_mean

    def get_opt_var_fn(self) -> Callable:
        def unscaled(Kin, Kcross, **kwargs):
            return like_input(self._unscaled_variance(Kin, Kcross, **kwargs), Kin, Kcross)

        return unscaled

    def get_opt_scale_fn(self) -> Callable:
        """scale_fn(Kin, nn_targets, **theta) used inside lool/looph objectives.

        Reproduces the reference quirk that the analytic scale perturbs with the
        MODEL's stored nugget, ignoring the optimiser's `noise=` keyword
        (S/gp/hyperparameter/scale.py:205-217)."""
        if not self.scale.analytic:
            return lambda Kin, nn_targets, *a, **kw: self.scale()

        def analytic(Kin, nn_targets, *args, **kwargs):
            y = fdev(nn_targets)
            b, k = y.shape[0], y.shape[1]
            yky = self._solve(Kin, None, y, want_yky=True)["yky"]
            sigma0 = float(yky.sum()) / (b * k)
            return self.scale.from_mean_quadratic_form(sigma0)

        return analytic

    def optimize_scale(self, pairwise_diffs, nn_targets):
        """S/gp/muygps.py:373-403."""
        Kin = self.kernel(fdev(pairwise_diffs))
        self.scale._set(self.get_opt_scale_fn()(Kin, nn_targets))
        self._make()
        return self

    def make_predict_tensors(self, batch_indices, batch_nn_indices, test_features,
                             train_features, train_targets, **kwargs):
        """(crosswise, pairwise, batch_nn_targets)    S/gp/muygps.py:405-475."""
        if test_features is None:
            test_features = train_features
        bi, bnn = idev(batch_indices), idev(batch_nn_indices)
        deformation = self.kernel.deformation
        crosswise = deformation.crosswise_tensor(fdev(test_features), fdev(train_features), bi, bnn)
        pairwise = deformation.pairwise_tensor(fdev(train_features), bnn)
        nn_targets = fdev(train_targets)[bnn]
        host = (batch_indices, batch_nn_indices, test_features, train_features, train_targets)
        return (like_input(crosswise, *host), like_input(pairwise, *host),
                like_input(nn_targets, *host))

    def make_train_tensors(self, batch_indices, batch_nn_indices, train_features, train_targets,
                           **kwargs):
        """(crosswise, pairwise, batch_targets, batch_nn_targets)   S/gp/muygps.py:477-551."""
        crosswise, pairwise, nn_targets = self.make_predict_tensors(
            batch_indices, batch_nn_indices, train_features, train_features, train_targets)
        batch_targets = fdev(train_targets)[idev(batch_indices)]
        host = (batch_indices, batch_nn_indices, train_features, train_targets)
        return crosswise, pairwise, like_input(batch_targets, *host), nn_targets

    # ---- fused API (indices in, posterior out; K1): see fused.py ----------------
    def _fused(self, indices, nn_indices, test_features, train_features, train_targets, *,
               theta: Optional[dict] = None, scale: Optional[float] = None, **want):
        return fused.fused_call(self, indices, nn_indices, test_features, train_features,
                                train_targets, theta=theta, scale=scale, **want)

    def fused_regress(self, indices, nn_indices, test_features, train_features, train_targets,
                      want_mean=True, want_var=True):
        return fused.fused_regress(self, indices, nn_indices, test_features, train_features,
                                   train_targets, want_mean=want_mean, want_var=want_var)

    def fused_fast_coefficients(self, nn_indices_fast, train_features, train_targets):
        return fused.fused_fast_coefficients(self, nn_indices_fast, train_features,
                                             train_targets)

    def fused_optimize_scale(self, batch_indices, batch_nn_indices, train_features,
                             train_targets):
        return fused.fused_optimize_scale(self, batch_indices, batch_nn_indices, train_features,
                                          train_targets)


class MultivariateMuyGPS:
    """One MuyGPS per response over shared neighbourhoods (S/gp/multivariate_muygps.py:28-340).

    Built like the reference's: `MultivariateMuyGPS(*model_args)` with one keyword dictionary
    per response.  The `*_from_indices` entry points and `fused_*` methods run one fused launch
    per response on the shared index arrays; the staged methods keep the reference's
    (pairwise_diffs, crosswise_diffs, ...) signatures."""

    def __init__(self, *model_args):
        self.models = [MuyGPS(**args) for args in model_args]

    def fixed(self) -> bool:
        return all(model.fixed() for model in self.models)

    def make_predict_tensors(self, *args, **kwargs):
        return self.models[0].make_predict_tensors(*args, **kwargs)

    def make_train_tensors(self, *args, **kwargs):
        return self.models[0].make_train_tensors(*args, **kwargs)

    def posterior_mean(self, pairwise_diffs, crosswise_diffs, batch_nn_targets):
        y = fdev(batch_nn_targets)
        cols = [fdev(m.posterior_mean(m.kernel(pairwise_diffs), m.kernel(crosswise_diffs),
                                      y[:, :, i].contiguous()))
                for i, m in enumerate(self.models)]
        return like_input(torch.stack(cols, dim=1), pairwise_diffs, crosswise_diffs,
                          batch_nn_targets)

    def posterior_variance(self, pairwise_diffs, crosswise_diffs):
        # scale enters twice, as in the reference (multivariate_muygps.py:183-192)
        cols = [fdev(m.posterior_variance(m.kernel(pairwise_diffs), m.kernel(crosswise_diffs)))
                * m.scale() for m in self.models]
        return like_input(torch.stack(cols, dim=1), pairwise_diffs, crosswise_diffs)

    def fast_coefficients(self, pairwise_diffs_fast, train_nn_targets_fast):
        y = fdev(train_nn_targets_fast)
        # the nugget enters twice, as in the reference (multivariate_muygps.py:224-231)
        cols = [fdev(m.fast_coefficients(m.noise.perturb(m.kernel(pairwise_diffs_fast)),
                                         y[:, :, i].contiguous()))
                for i, m in enumerate(self.models)]
        return like_input(torch.stack(cols, dim=2), pairwise_diffs_fast, train_nn_targets_fast)

    def fast_posterior_mean(self, crosswise_diffs, coeffs_tensor):
        c = fdev(coeffs_tensor)
        cols = [ops.rowdot(fdev(m.kernel(crosswise_diffs)), c[:, :, i].contiguous())[:, 0]
                for i, m in enumerate(self.models)]
        return like_input(torch.stack(cols, dim=1), crosswise_diffs, coeffs_tensor)

    # fused (indices in)
    def fused_regress(self, indices, nn_indices, test_features, train_features, train_targets,
                      want_mean=True, want_var=True):
        return fused.mm_fused_regress(self, indices, nn_indices, test_features, train_features,
                                      train_targets, want_mean=want_mean, want_var=want_var)

    def fused_fast_coefficients(self, nn_indices_fast, train_features, train_targets):
        return fused.mm_fused_fast_coefficients(self, nn_indices_fast, train_features,
                                                train_targets)
