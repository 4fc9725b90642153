"""Fused `*_from_indices` entry points (signatures of S/examples/from_indices.py:22-223).

In the reference these build (b,k,k[,d]) tensors on the way; here each one is a
single K1 / K4 launch on the index arrays, which is what makes the C5-sized
problems (10^7 rows) possible at all.  `tensors_from_indices` still returns
materialised `Kin`/`Kcross` for callers that want them.
"""

from __future__ import annotations

from ._arrays import fdev, idev, like_input
from . import ops
from .losses import LossFn, lool_fn
from .optimizers import L_BFGS_B_optimize, OptimizeFn


def tensors_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    crosswise, pairwise, nn_targets = muygps.make_predict_tensors(
        indices, nn_indices, test, train, targets)
    return muygps.kernel(pairwise), muygps.kernel(crosswise), nn_targets


def posterior_mean_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    return muygps.fused_regress(indices, nn_indices, test, train, targets, want_var=False)


def posterior_variance_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    return muygps.fused_regress(indices, nn_indices, test, train, targets, want_mean=False)


def regress_from_indices(muygps, indices, nn_indices, test, train, targets, **kwargs):
    return muygps.fused_regress(indices, nn_indices, test, train, targets)


def fast_posterior_mean_from_indices(muygps, indices, nn_indices, test_features,
                                     train_features, closest_index, coeffs_tensor):
    deformation = muygps.kernel.deformation
    ls = deformation.length_scales()
    out = ops.fast_mean(
        fdev(train_features), fdev(test_features), idev(indices), idev(nn_indices),
        idev(closest_index), fdev(coeffs_tensor), kernel_id=muygps.kernel.kernel_id,
        metric_id=deformation.metric.metric_id,
        length_scale=ls if deformation.anisotropic else ls[0])
    if fdev(coeffs_tensor).dim() == 2:
        out = out[:, 0]
    return like_input(out, indices, nn_indices, test_features, train_features, coeffs_tensor)


def optimize_from_indices(muygps, batch_indices, batch_nn_indices, train_features,
                          train_targets, loss_fn: LossFn = lool_fn,
                          opt_fn: OptimizeFn = L_BFGS_B_optimize, verbose: bool = False,
                          **kwargs):
    return opt_fn.from_indices(muygps, batch_indices, batch_nn_indices, train_features,
                               train_targets, loss_fn=loss_fn, verbose=verbose, **kwargs)
