"""Leave-one-out objective functions (S/optimize/objective.py:20-118, loss.py:26-178).

`make_loo_crossval_fn` keeps the reference's signature and works on materialised
difference/distance tensors through the staged kernels.  `make_fused_loo_crossval_fn`
is the fast path: it captures (features, indices, targets) instead and every
`obj_fn(**theta)` call is one K1 launch + one loss reduction, ending in a single
8-double all-reduce when several ranks share the batch.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

from . import _lib as L
from . import ops
from ._arrays import fdev, idev
from .distributed import allreduce_partials
from .losses import LossFn


def _finish_loss(loss_fn: LossFn, rec) -> float:
    """Turn an (all-reduced) partials record of a variance-free loss into its value."""
    if loss_fn.loss_id == L.LOSS_MSE:
        return rec[L.P_SQERR] / rec[L.P_COUNT]
    return rec[L.P_AUX]


def make_loo_crossval_fn(loss_fn: LossFn, kernel_fn: Callable, mean_fn: Callable,
                         var_fn: Callable, scale_fn: Callable, pairwise_diffs, crosswise_diffs,
                         batch_nn_targets, batch_targets, batch_features=None,
                         target_mask=None, loss_kwargs: Dict = dict()) -> Callable:
    """Staged objective with the reference's argument list; returns `obj_fn(**theta) -> -loss`."""
    pw, cw = fdev(pairwise_diffs), fdev(crosswise_diffs)
    y_nn, y_b = fdev(batch_nn_targets), fdev(batch_targets)

    def obj_fn(*args, **kwargs):
        Kin = kernel_fn(pw, **kwargs)
        Kcross = kernel_fn(cw, **kwargs)
        predictions = mean_fn(Kin, Kcross, y_nn, **kwargs)
        if loss_fn.needs_variance:
            scale = scale_fn(Kin, y_nn, **kwargs)
            variances = var_fn(Kin, Kcross, **kwargs)
            if target_mask is not None:
                predictions = predictions[:, target_mask]
                variances = variances[:, target_mask, target_mask]
            return -loss_fn(predictions, y_b, variances, scale, **loss_kwargs)
        if target_mask is not None:
            predictions = predictions[:, target_mask]
        return -loss_fn(predictions, y_b, **loss_kwargs)

    return obj_fn


def make_fused_loo_crossval_fn(muygps, loss_fn: LossFn, batch_indices, batch_nn_indices,
                               train_features, train_targets, target_mask=None,
                               loss_kwargs: Optional[Dict] = None, group=None,
                               distributed: bool = False) -> Callable:
    """Fused objective: nothing of size (b,k,k) is ever materialised.

    With `distributed=True` the given batch rows are this rank's shard; partial
    sums are combined across `group` so every rank returns the global objective.
    Accepts the keyword names the reference's optimiser uses: `length_scale` or
    `length_scale0..`, `noise`; anything else (e.g. `smoothness`) is ignored.
    """
    loss_kwargs = dict(loss_kwargs or {})
    x = fdev(train_features)
    y = fdev(train_targets)
    bi, bnn = idev(batch_indices), idev(batch_nn_indices)
    y_b = y[bi].contiguous()
    if target_mask is not None:
        y_b = y_b[:, target_mask].contiguous()
    k = bnn.shape[1]
    needs_var = loss_fn.needs_variance
    analytic = needs_var and muygps.scale.analytic
    delta = float(loss_kwargs.get("boundary_scale", loss_fn.default_boundary()))
    model_noise = muygps.noise.value(None)
    reduce = (lambda rec: allreduce_partials(rec, group)) if distributed else (lambda rec: rec)
    rec_pin = torch.empty((L.MGP_PARTIALS,), dtype=torch.float64).pin_memory()

    def to_host(rec):
        """The 8-double record through a page-locked buffer (cheaper than Tensor.cpu())."""
        rec_pin.copy_(rec, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return rec_pin.numpy().copy()

    def obj_fn(*args, **theta):
        noise_kw = theta.get("noise")
        same_noise = noise_kw is None or muygps.noise.heteroscedastic or \
            float(noise_kw) == float(model_noise)
        out = muygps._fused(bi, bnn, x, x, y, theta=theta, scale=1.0, want_mean=True,
                            want_var=needs_var, want_yky=analytic and same_noise)
        yky = out.get("yky")
        if analytic and not same_noise:
            # reference quirk: the scale ignores the optimiser's nugget (scale.py:206-208)
            th2 = {kk: v for kk, v in theta.items() if kk != "noise"}
            yky = muygps._fused(bi, bnn, x, x, y, theta=th2, scale=1.0, want_mean=False,
                                want_var=False, want_yky=True)["yky"]
        mean = out["mean"]
        if y.dim() == 1:
            mean = mean[:, 0]
        elif target_mask is not None:
            mean = mean[:, target_mask].contiguous()
        if not needs_var:
            rec = ops.loss_partials(loss_fn.loss_id, mean, y_b, boundary_scale=delta)
            rec = to_host(reduce(rec))
            return -float(_finish_loss(loss_fn, rec))
        var = out["var"]
        if analytic:
            if loss_fn.loss_id == L.LOSS_LOOL:
                # lool is affine in 1/sigma^2 and log sigma^2, so the per-rank sums of
                # e^2/v, log v and y^T K^-1 y finish it after ONE all-reduce
                rec = ops.loss_partials(L.LOSS_LOOL, mean, y_b, var=var, yky=yky)
                rec = to_host(reduce(rec))
                rows = rec[L.P_ROWS]
                sigma2 = muygps.scale.from_mean_quadratic_form(rec[L.P_YKY] / (rows * k))
                loss = rec[L.P_SQERR_V] / sigma2 + rec[L.P_LOGV] + rows * math.log(sigma2)
                return -float(loss)
            # looph is nonlinear in sigma^2: reduce the scale first, then the loss
            rec = reduce(ops.loss_partials(L.LOSS_NONE, mean, y_b, var=var, yky=yky))
            sigma0 = rec[L.P_YKY] / (rec[L.P_ROWS] * k)
            sigma2_dev = _iterate_scale(sigma0, muygps.scale.iteration_count).reshape(1)
        else:
            sigma2_dev = torch.full((1,), float(muygps.scale()), dtype=torch.float64,
                                    device=mean.device)
        rec2 = ops.loss_partials(loss_fn.loss_id, mean, y_b, var=var, scale_dev=sigma2_dev,
                                 boundary_scale=delta)
        rec2 = to_host(reduce(rec2))
        return -float(rec2[L.P_AUX])

    return obj_fn


def _iterate_scale(sigma0: torch.Tensor, iteration_count: int) -> torch.Tensor:
    """Device-side twin of AnalyticScale.from_mean_quadratic_form (no host sync)."""
    s = sigma0
    for _ in range(1, iteration_count):
        s = 0.5 * (s + sigma0 / s)
    return s
