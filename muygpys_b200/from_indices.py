"""Fused `*_from_indices` entry points (signatures of S/examples/from_indices.py:22-223).

In the reference these build (b,k,k[,d]) tensors on the way; here each one is a single K1 / K4
launch on the index arrays, which is what makes the C5-sized problems (10^7 rows) possible at
all.  `muygps` may be a genuine `MuyGPyS.gp.MuyGPS` or this package's mirror object
(`adapt.ModelSpec`).  `tensors_from_indices` still returns materialised `Kin`/`Kcross` for
callers that want them.
"""

from __future__ import annotations

from copy import deepcopy

import numpy as np

from . import fused, ops
from ._arrays import fdev, idev, like_input
from .adapt import ModelSpec
from .objective import make_fused_loo_crossval_fn


def tensors_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    crosswise, pairwise, nn_targets = muygps.make_predict_tensors(
        indices, nn_indices, test, train, targets)
    return muygps.kernel(pairwise), muygps.kernel(crosswise), nn_targets


def posterior_mean_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    if fused.is_multivariate(muygps):
        return fused.mm_fused_regress(muygps, indices, nn_indices, test, train, targets,
                                      want_var=False)
    return fused.fused_regress(muygps, indices, nn_indices, test, train, targets, want_var=False)


def posterior_variance_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    if fused.is_multivariate(muygps):
        return fused.mm_fused_regress(muygps, indices, nn_indices, test, train, targets,
                                      want_mean=False)
    return fused.fused_regress(muygps, indices, nn_indices, test, train, targets,
                               want_mean=False)


def regress_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    if fused.is_multivariate(muygps):
        return fused.mm_fused_regress(muygps, indices, nn_indices, test, train, targets)
    return fused.fused_regress(muygps, indices, nn_indices, test, train, targets)


def fast_posterior_mean_from_indices(muygps, indices, nn_indices, test_features,
                                     train_features, closest_index, coeffs_tensor):
    if fused.is_multivariate(muygps):
        return fused.mm_fast_posterior_mean(muygps, indices, nn_indices, test_features,
                                            train_features, closest_index, coeffs_tensor)
    spec = ModelSpec.of(muygps)
    out = ops.fast_mean(
        fdev(train_features), fdev(test_features), idev(indices), idev(nn_indices),
        idev(closest_index), fdev(coeffs_tensor), kernel_id=spec.kernel_id,
        metric_id=spec.metric_id, length_scale=spec.length_scale_arg())
    if fdev(coeffs_tensor).dim() == 2:
        out = out[:, 0]
    return like_input(out, indices, nn_indices, test_features, train_features, coeffs_tensor)


def _lbfgsb(muygps, obj_fn, verbose=False, **kwargs):
    """The outer loop when MuyGPyS itself is not importable: scipy L-BFGS-B over the free
    hyperparameters (what S/_src/optimize/chassis/numpy.py:52-81 drives)."""
    from scipy import optimize as sciopt

    names, x0, bounds = muygps.get_opt_params()
    res = sciopt.minimize(lambda x: -obj_fn(**dict(zip(names, x))), x0, method="L-BFGS-B",
                          bounds=bounds, **kwargs)
    if verbose:
        print(res)
    new = deepcopy(muygps)
    for name, val, (lo, hi) in zip(names, res.x, bounds):
        target = new.noise if name == "noise" else new.kernel._hyperparameters[name]
        target._set_val(min(max(float(val), lo), hi))
    new._make()
    return new


def _lbfgsb_with_gradient(muygps, value_and_grad, verbose=False, **kwargs):
    """scipy L-BFGS-B with jac=True: one kernel launch per iteration instead of 1 + p."""
    from scipy import optimize as sciopt

    names, x0, bounds = muygps.get_opt_params()
    names = [str(n) for n in names]
    for n in names:
        if n != "noise" and not n.startswith("length_scale"):
            raise NotImplementedError(
                f"use_gradient=True: no analytic derivative with respect to {n!r} (length "
                "scales and the nugget only); optimise it by finite differences")

    def fun(xv):
        val, grads = value_and_grad(**dict(zip(names, xv)))
        return -val, -np.array([grads[n] for n in names])

    res = sciopt.minimize(fun, x0, method="L-BFGS-B", jac=True, bounds=bounds, **kwargs)
    if verbose:
        print(res)
    new = deepcopy(muygps)
    for name, val, (lo, hi) in zip(names, res.x, bounds):
        target = new.noise if name == "noise" else new.kernel._hyperparameters[name]
        target._set_val(min(max(float(val), lo), hi))
    new._make()
    return new


def optimize_from_indices(muygps, batch_indices, batch_nn_indices, train_features,
                          train_targets, loss_fn=None, opt_fn=None, verbose: bool = False,
                          loss_kwargs=None, target_mask=None, group=None,
                          distributed: bool = False, use_gradient: bool = False, **kwargs):
    """`optimize_from_indices` (S/examples/from_indices.py:126-223) with the fused objective.

    The outer loop is the reference's own: `opt_fn` is a `MuyGPyS.optimize.OptimizeFn`
    (`L_BFGS_B_optimize`, `Bayes_optimize`; default L-BFGS-B) whose optimiser
    `opt_fn._fn(muygps, obj_fn, verbose=..., **kwargs)` (S/optimize/chassis.py:86-120) is handed
    the one-launch `obj_fn`; only the objective changes."""
    if loss_fn is None:
        from .losses import lool_fn as loss_fn
    if use_gradient:
        # analytic gradient from the same launch: L-BFGS-B with jac=True
        from .objective import make_fused_loo_value_and_grad_fn

        vg = make_fused_loo_value_and_grad_fn(muygps, loss_fn, batch_indices, batch_nn_indices,
                                              train_features, train_targets, group=group,
                                              distributed=distributed, loss_kwargs=loss_kwargs)
        return _lbfgsb_with_gradient(muygps, vg, verbose=verbose, **kwargs)
    obj_fn = make_fused_loo_crossval_fn(
        muygps, loss_fn, batch_indices, batch_nn_indices, train_features, train_targets,
        target_mask=target_mask, loss_kwargs=loss_kwargs, group=group, distributed=distributed)
    if opt_fn is None:
        try:
            from MuyGPyS.optimize import L_BFGS_B_optimize as opt_fn
        except ImportError:
            return _lbfgsb(muygps, obj_fn, verbose=verbose, **kwargs)
    driver = getattr(opt_fn, "_fn", opt_fn)
    return driver(muygps, obj_fn, verbose=verbose, **kwargs)
