"""Drop-in callables for the reference's own constructor-injection points.

MuyGPyS binds its backend math through private keyword arguments of its L3 objects
(SURVEY.md section 1): `MuyGPS(_backend_mean_fn=..., _backend_var_fn=...,
_backend_fast_mean_fn=..., _backend_fast_precompute_fn=...)` (S/gp/muygps.py:93-110),
`RBF(_backend_fn=...)` (S/gp/kernels/rbf.py:72-87), `Matern(_backend_05_fn=... )`
(S/gp/kernels/matern.py:61-82), `HomoscedasticNoise(_backend_fn=...)`,
`AnalyticScale(_backend_fn=...)`, `MetricFn(...)`, `LossFn(...)`.  Every function below has
the name and signature of the numpy-backend function it replaces
(S/_src/**/numpy.py) and runs one CUDA kernel on torch CUDA tensors; numpy inputs are
uploaded and results returned as numpy, so reference code that expects ndarrays keeps
working.  `reference_objects()` assembles unmodified reference classes with these
callables injected -- the same pattern as tests/backend/torch_correctness.py:244-331.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._arrays import fdev, idev, like_input


# ---- S/_src/gp/tensors/numpy.py --------------------------------------------------------
def _crosswise_tensor(data, nn_data, data_indices, nn_indices):
    out = ops.crosswise_diffs(fdev(data), fdev(nn_data), idev(data_indices), idev(nn_indices))
    return like_input(out, data, nn_data, data_indices, nn_indices)


def _pairwise_tensor(data, nn_indices):
    return like_input(ops.pairwise_diffs(fdev(data), idev(nn_indices)), data, nn_indices)


def _F2(diffs):
    return like_input(ops.metric_reduce(L.METRIC_F2, fdev(diffs)), diffs)


def _l2(diffs):
    return like_input(ops.metric_reduce(L.METRIC_L2, fdev(diffs)), diffs)


def _fast_nn_update(train_nn_indices):
    from .gp.tensors import fast_nn_update

    return fast_nn_update(train_nn_indices)


# ---- S/_src/gp/kernels/numpy.py --------------------------------------------------------
def _kernel(kernel_id):
    def fn(dists, **kwargs):  # Matern fns also receive smoothness= (scalar.py:314-319)
        return like_input(ops.kernel_apply(kernel_id, fdev(dists), 1.0), dists)

    return fn


_rbf_fn = _kernel(L.KERNEL_RBF)
_matern_05_fn = _kernel(L.KERNEL_MATERN_05)
_matern_15_fn = _kernel(L.KERNEL_MATERN_15)
_matern_25_fn = _kernel(L.KERNEL_MATERN_25)
_matern_inf_fn = _kernel(L.KERNEL_MATERN_INF)


def _matern_gen_fn(dists, smoothness, **kwargs):
    raise NotImplementedError(
        "general-smoothness Matern (Bessel) is outside the B200 hot path; fix smoothness at "
        "0.5, 1.5, 2.5 or inf")


# ---- S/_src/gp/noise/numpy.py ----------------------------------------------------------
def _homoscedastic_perturb(Kin, noise_variance):
    return like_input(ops.perturb(fdev(Kin), float(noise_variance)), Kin)


def _heteroscedastic_perturb(Kin, noise_variances):
    return like_input(ops.perturb(fdev(Kin), fdev(noise_variances)), Kin)


# ---- S/_src/gp/muygps/numpy.py ---------------------------------------------------------
def _muygps_posterior_mean(Kin, Kcross, nn_targets, **kwargs):
    y = fdev(nn_targets)
    out = ops.solve(fdev(Kin), fdev(Kcross), y, 1.0, want_mean=True)["mean"]
    out = out[:, 0] if y.dim() == 2 else out
    return like_input(out, Kin, Kcross, nn_targets)


def _muygps_diagonal_variance(Kin, Kcross, Kout, batch_size=1, **kwargs):
    out = ops.solve(fdev(Kin), fdev(Kcross), None, float(Kout), want_var=True)["var"]
    return like_input(out, Kin, Kcross)


def _muygps_fast_posterior_mean(Kcross, coeffs_tensor, **kwargs):
    c = fdev(coeffs_tensor)
    out = ops.rowdot(fdev(Kcross), c)
    out = out[:, 0] if c.dim() == 2 else out
    return like_input(out, Kcross, coeffs_tensor)


def _muygps_fast_posterior_mean_precompute(Kin, train_nn_targets_fast, **kwargs):
    y = fdev(train_nn_targets_fast)
    out = ops.solve(fdev(Kin), None, y, 1.0, want_coeffs=True)["coeffs"]
    out = out[:, :, 0] if y.dim() == 2 else out
    return like_input(out, Kin, train_nn_targets_fast)


# ---- S/_src/optimize/scale/numpy.py ----------------------------------------------------
def _analytic_scale_optim_unnormalized(Kin, nn_targets, **kwargs):
    yky = ops.solve(fdev(Kin), None, fdev(nn_targets), 1.0, want_yky=True)["yky"]
    return float(yky.sum())


def _analytic_scale_optim(Kin, nn_targets, batch_dim_count=1, **kwargs):
    K = fdev(Kin)
    return _analytic_scale_optim_unnormalized(K, nn_targets) / (K.shape[0] * K.shape[1])


# ---- S/_src/optimize/loss/numpy.py -----------------------------------------------------
def _mse_fn(predictions, targets, **kwargs):
    from .losses import mse_fn

    return mse_fn(predictions, targets)


def _cross_entropy_fn(predictions, targets, **kwargs):
    from .losses import cross_entropy_fn

    return cross_entropy_fn(predictions, targets)


def _lool_fn(predictions, targets, variances, scale, **kwargs):
    from .losses import lool_fn

    return lool_fn(predictions, targets, variances, float(scale))


def _pseudo_huber_fn(predictions, targets, boundary_scale=1.5, **kwargs):
    from .losses import pseudo_huber_fn

    return pseudo_huber_fn(predictions, targets, boundary_scale=boundary_scale)


def _looph_fn(predictions, targets, variances, scale, boundary_scale=3.0, **kwargs):
    from .losses import looph_fn

    return looph_fn(predictions, targets, variances, float(scale), boundary_scale=boundary_scale)


# ---- S/_src/math helpers the kernels take (Kout must be a plain scalar) ----------------
def _ones(shape, **kwargs):
    return np.ones(shape)


def _zeros(shape, **kwargs):
    return np.zeros(shape)


def _squeeze(x, **kwargs):
    return float(np.squeeze(x))


def reference_objects():
    """Return a namespace of factories that build UNMODIFIED reference objects with the CUDA
    callables injected.  Needs `MuyGPyS` importable (the reference is not shipped here)."""
    from types import SimpleNamespace

    from MuyGPyS.gp import MuyGPS
    from MuyGPyS.gp.deformation import Anisotropy, Isotropy
    from MuyGPyS.gp.deformation.metric import MetricFn
    from MuyGPyS.gp.hyperparameter import AnalyticScale
    from MuyGPyS.gp.kernels import RBF, Matern
    from MuyGPyS.gp.noise import HeteroscedasticNoise, HomoscedasticNoise
    from MuyGPyS.optimize.loss import (LossFn, make_raw_predict_and_loss_fn,
                                       make_var_predict_and_loss_fn)

    l2 = MetricFn(differences_metric_fn=_l2, crosswise_differences_fn=_crosswise_tensor,
                  pairwise_diffferences_fn=_pairwise_tensor,
                  apply_length_scale_fn=lambda x, y: x / y)
    F2 = MetricFn(differences_metric_fn=_F2, crosswise_differences_fn=_crosswise_tensor,
                  pairwise_diffferences_fn=_pairwise_tensor,
                  apply_length_scale_fn=lambda x, y: x / y**2)
    kernel_kw = dict(_backend_ones=_ones, _backend_zeros=_zeros, _backend_squeeze=_squeeze)

    def matern(**kw):
        return Matern(_backend_05_fn=_matern_05_fn, _backend_15_fn=_matern_15_fn,
                      _backend_25_fn=_matern_25_fn, _backend_inf_fn=_matern_inf_fn,
                      _backend_gen_fn=_matern_gen_fn, **kernel_kw, **kw)

    def rbf(**kw):
        return RBF(_backend_fn=_rbf_fn, **kernel_kw, **kw)

    def muygps(**kw):
        return MuyGPS(_backend_mean_fn=_muygps_posterior_mean,
                      _backend_var_fn=_muygps_diagonal_variance,
                      _backend_fast_mean_fn=_muygps_fast_posterior_mean,
                      _backend_fast_precompute_fn=_muygps_fast_posterior_mean_precompute, **kw)

    return SimpleNamespace(
        l2=l2, F2=F2, Isotropy=Isotropy, Anisotropy=Anisotropy, Matern=matern, RBF=rbf,
        MuyGPS=muygps,
        HomoscedasticNoise=lambda *a, **kw: HomoscedasticNoise(
            *a, _backend_fn=_homoscedastic_perturb, **kw),
        HeteroscedasticNoise=lambda *a, **kw: HeteroscedasticNoise(
            *a, _backend_fn=_heteroscedastic_perturb, **kw),
        AnalyticScale=lambda **kw: AnalyticScale(_backend_fn=_analytic_scale_optim, **kw),
        mse_fn=LossFn(_mse_fn, make_raw_predict_and_loss_fn),
        cross_entropy_fn=LossFn(_cross_entropy_fn, make_raw_predict_and_loss_fn),
        pseudo_huber_fn=LossFn(_pseudo_huber_fn, make_raw_predict_and_loss_fn),
        lool_fn=LossFn(_lool_fn, make_var_predict_and_loss_fn),
        looph_fn=LossFn(_looph_fn, make_var_predict_and_loss_fn),
    )
