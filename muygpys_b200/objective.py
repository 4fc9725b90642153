"""Leave-one-out objective functions (S/optimize/objective.py:20-118, loss.py:26-178).

`make_loo_crossval_fn` keeps the reference's signature and works on materialised
difference/distance tensors through the staged kernels.  `make_fused_loo_crossval_fn` is the
fast path: it captures (features, indices, targets) instead and, for one response with the mse,
lool, pseudo-Huber or looph loss, every `obj_fn(**theta)` call is ONE kernel launch
(`mgp_fused_loo`: K1 with the loss / scale partials folded into its epilogue; looph with the
analytic scale: one launch for sigma^2, one for the loss) followed by a 64-byte read -- with several ranks, the per-rank records are summed by one all-reduce
(`distributed.PartialsReducer`).  Other shapes take K1 + the loss kernels.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._arrays import fdev, idev
from .adapt import ModelSpec
from .distributed import PartialsReducer
from .fused import fused_call
from .losses import LossFn, as_loss


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


def make_fused_loo_crossval_fn(muygps, loss_fn, batch_indices, batch_nn_indices,
                               train_features, train_targets, target_mask=None,
                               loss_kwargs: Optional[Dict] = None, group=None,
                               distributed: bool = False) -> Callable:
    """Fused objective: nothing of size (b,k,k) is ever materialised.

    `muygps` is a `MuyGPyS.gp.MuyGPS` or this package's mirror object, `loss_fn` a loss object
    of either family.  With `distributed=True` the given batch rows are this rank's shard;
    partial sums are combined across `group` so every rank returns the global objective.
    Accepts the keyword names the reference's optimiser uses: `length_scale` or
    `length_scale0..`, `noise`; anything else (e.g. `smoothness`) is ignored.
    """
    spec = ModelSpec.of(muygps)
    loss_fn = as_loss(loss_fn)
    loss_kwargs = dict(loss_kwargs or {})
    x = fdev(train_features)
    y = fdev(train_targets)
    bi, bnn = idev(batch_indices), idev(batch_nn_indices)
    k = bnn.shape[1]
    needs_var = loss_fn.needs_variance
    analytic = needs_var and spec.analytic
    delta = float(loss_kwargs.get("boundary_scale", loss_fn.default_boundary()))
    model_noise = spec.noise(None)
    d = 1 if x.dim() == 1 else x.shape[1]
    r = 1 if y.dim() == 1 else y.shape[1]
    reducer = PartialsReducer(x.device, group) if distributed else None

    def same_noise(theta) -> bool:
        noise_kw = theta.get("noise")
        return noise_kw is None or spec.heteroscedastic or float(noise_kw) == float(model_noise)

    def finish(rec, rec_scale=None) -> float:
        """Loss value from the (summed) host record(s)."""
        if not needs_var:
            return float(_finish_loss(loss_fn, rec))
        rows = rec[L.P_ROWS]
        if analytic:
            src = rec if rec_scale is None else rec_scale
            sigma2 = spec.sigma_from_mean_quadratic_form(src[L.P_YKY] / (src[L.P_ROWS] * k))
        else:
            sigma2 = spec.scale()
        # lool is affine in 1/sigma^2 and log sigma^2, so the sums of e^2/v, log v and
        # y^T K^-1 y finish it after ONE reduction
        return float(rec[L.P_SQERR_V] / sigma2 + rec[L.P_LOGV] + rows * math.log(sigma2))

    looph = loss_fn.loss_id == L.LOSS_LOOPH
    one_launch = (target_mask is None
                  and loss_fn.loss_id in (L.LOSS_MSE, L.LOSS_LOOL, L.LOSS_PSEUDO_HUBER,
                                          L.LOSS_LOOPH)
                  and ops.fused_loo_supported(d, k, r, spec.kernel_id, spec.metric_id,
                                              spec.heteroscedastic)
                  and x.data_ptr() % 16 == 0)
    if one_launch:
        # several GPUs: the kernel's own epilogue sums the records over NVLink peer memory
        # (then, as on one GPU, its last block writes the result straight into pinned host
        # memory); without peer access the record stays on the device for an all-reduce
        chan = None if reducer is None else reducer.channel()
        chan_scale = None if reducer is None else reducer.channel()
        on_device = reducer is not None and chan is None
        loo = ops.FusedLoo(x, y, bi, bnn, kernel_id=spec.kernel_id, metric_id=spec.metric_id,
                           loss_id=loss_fn.loss_id, boundary_scale=delta,
                           partials=reducer.slot() if on_device else None)
        loo_scale = []  # lazily: a second evaluator for the noise quirk below

        def read(ev, dev_rec, fused_sum):
            if reducer is None or fused_sum:
                return ev.record(dev_rec)
            return reducer.sum_to_host(dev_rec)

        def obj_fn(*args, **theta):
            ls = spec.length_scale_arg(**theta)
            rec_scale = None
            # (looph is nonlinear in sigma^2: the scale launch always comes first, and its
            #  sigma^2 goes into the loss launch -- two launches, S/optimize/objective.py:94-105)
            if analytic and (looph or not same_noise(theta)):
                # reference quirk: the analytic scale perturbs with the MODEL's nugget, ignoring
                # the optimiser's `noise=` (S/gp/hyperparameter/scale.py:206-208)
                if not loo_scale:
                    loo_scale.append(ops.FusedLoo(
                        x, y, bi, bnn, kernel_id=spec.kernel_id, metric_id=spec.metric_id,
                        loss_id=L.LOSS_NONE, partials=reducer.slot() if on_device else None))
                rec_scale = read(loo_scale[0], loo_scale[0].launch(ls, model_noise, chan_scale),
                                 chan_scale is not None)
            if looph:
                if analytic:
                    sigma2 = spec.sigma_from_mean_quadratic_form(
                        rec_scale[L.P_YKY] / (rec_scale[L.P_ROWS] * k))
                else:
                    sigma2 = spec.scale()
                rec = read(loo, loo.launch(ls, spec.noise(theta.get("noise")), chan,
                                           scale=sigma2), chan is not None)
                return -float(rec[L.P_AUX] + rec[L.P_LOGV]
                              + rec[L.P_ROWS] * math.log(sigma2))
            rec = read(loo, loo.launch(ls, spec.noise(theta.get("noise")), chan),
                       chan is not None)
            return -finish(rec, rec_scale)

        return obj_fn

    # ---- general shapes: K1, then the loss kernels ------------------------------------------
    y_b = y[bi].contiguous()
    if target_mask is not None:
        y_b = y_b[:, target_mask].contiguous()
    rec_pin = torch.empty((L.MGP_PARTIALS,), dtype=torch.float64).pin_memory()

    def to_host(rec):
        if reducer is not None:
            return reducer.sum_to_host(rec)
        rec_pin.copy_(rec, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return rec_pin.numpy().copy()

    def obj_fn(*args, **theta):
        same = same_noise(theta)
        out = fused_call(spec, bi, bnn, x, x, y, theta=theta, scale=1.0, want_mean=True,
                         want_var=needs_var, want_yky=analytic and same)
        yky = out.get("yky")
        if analytic and not same:
            th2 = {kk: v for kk, v in theta.items() if kk != "noise"}
            yky = fused_call(spec, bi, bnn, x, x, y, theta=th2, scale=1.0, want_mean=False,
                             want_var=False, want_yky=True)["yky"]
        mean = out["mean"]
        if y.dim() == 1:
            mean = mean[:, 0]
        elif target_mask is not None:
            mean = mean[:, target_mask].contiguous()
        if not needs_var:
            rec = ops.loss_partials(loss_fn.loss_id, mean, y_b, boundary_scale=delta)
            return -finish(to_host(rec))
        var = out["var"]
        if loss_fn.loss_id == L.LOSS_LOOL:
            rec = ops.loss_partials(L.LOSS_LOOL, mean, y_b, var=var, yky=yky)
            return -finish(to_host(rec))
        # looph is nonlinear in sigma^2: reduce the scale first, then the loss
        if analytic:
            rec = to_host(ops.loss_partials(L.LOSS_NONE, mean, y_b, var=var, yky=yky))
            sigma2 = spec.sigma_from_mean_quadratic_form(rec[L.P_YKY] / (rec[L.P_ROWS] * k))
        else:
            sigma2 = spec.scale()
        sigma2_dev = torch.full((1,), float(sigma2), dtype=torch.float64, device=mean.device)
        rec2 = ops.loss_partials(loss_fn.loss_id, mean, y_b, var=var, scale_dev=sigma2_dev,
                                 boundary_scale=delta)
        return -float(to_host(rec2)[L.P_AUX])

    return obj_fn


def finish_value_and_grad(rec, g, *, loss_id, k, d, anisotropic, analytic, sigma2=None,
                          fixed_scale=None, g_sigma=None):
    """Objective (negated loss) and its gradient from the summed records of `mgp_fused_loo_grad`:
    `rec` the MGP_P_* partials, `g[t]` the five gradient sums of parameter slot t (length scale
    of feature 0..2, nugget in slot 3; include/muygpys_b200.h).  Pure host arithmetic (tested on
    the CPU against finite differences of the oracle, tests/test_grad_finish_cpu.py).

    mse: sum e^2 / count.  lool: S / sigma^2 + sum log v + n log sigma^2 with S = sum e^2 / v.
    looph: the kernel was handed sigma^2 and weighted S and the sums g[t][1], g[t][2] with the
    Huber weight 1 / sqrt(1 + u), u = e^2 / (b^2 sigma^2 v); AUX = sum 2 b^2 (sqrt(1 + u) - 1)
    replaces S / sigma^2 and the derivative keeps lool's form.  Analytic scale: sigma^2 =
    sum yky / (n k) and d sigma^2 = sum d yky / (n k), taken from `g_sigma` when the scale was
    evaluated in a separate launch at the MODEL's nugget (reference quirk,
    S/gp/hyperparameter/scale.py:206-208); the optimiser's nugget never moves sigma^2."""
    rows = rec[L.P_ROWS]
    looph = loss_id == L.LOSS_LOOPH
    if loss_id in (L.LOSS_LOOL, L.LOSS_LOOPH):
        S = rec[L.P_SQERR_V]
        if sigma2 is None:
            sigma2 = rec[L.P_YKY] / (rows * k) if analytic else fixed_scale
        head = rec[L.P_AUX] if looph else S / sigma2
        value = head + rec[L.P_LOGV] + rows * math.log(sigma2)
        gs = g if g_sigma is None else g_sigma  # where d sum yky comes from

        def dloss(t, slot=None):
            if not analytic or slot == 3:
                dsig = 0.0  # fixed scale; or the nugget, which sigma^2 ignores (quirk)
            elif slot is None:
                dsig = gs[:d, 4].sum() / (rows * k)
            else:
                dsig = gs[slot, 4] / (rows * k)
            return ((t[1] - t[2]) / sigma2 + t[3]
                    + dsig * (-S / (sigma2 * sigma2) + rows / sigma2))
    else:
        value = rec[L.P_SQERR] / rec[L.P_COUNT]

        def dloss(t, slot=None):
            return t[0] / rec[L.P_COUNT]

    grads = {}
    if anisotropic:
        for f in range(d):
            grads[f"length_scale{f}"] = -float(dloss(g[f], f))
    else:
        grads["length_scale"] = -float(dloss(g[:d].sum(axis=0)))
    grads["noise"] = -float(dloss(g[3], 3))
    return -float(value), grads


def make_fused_loo_value_and_grad_fn(muygps, loss_fn, batch_indices, batch_nn_indices,
                                     train_features, train_targets, group=None,
                                     distributed: bool = False,
                                     loss_kwargs: Optional[Dict] = None) -> Callable:
    """`fn(**theta) -> (objective, {name: d objective / d name})` with the ANALYTIC gradient
    (SURVEY.md 8f-2): one launch of `mgp_fused_loo_grad` per call, where the reference's
    optimiser needs 1 + p objective evaluations for its finite differences
    (S/_src/optimize/chassis/numpy.py:68-74).  The objective is the negated loss, as
    `make_loo_crossval_fn` returns it.  Supported: mse, lool and looph (fixed scale, or analytic
    scale with iteration_count == 1), one response, the shapes of
    `mgp_fused_loo`, including the nugget under the analytic scale (the reference's quirk: sigma^2
    is taken at the MODEL's nugget; when the optimiser's differs, a second gradient launch at the
    model's nugget supplies sigma^2 and its derivatives).  looph is nonlinear in the scale: with the analytic scale a plain launch
    (y^T K^-1 y only) fixes sigma^2 first, then the gradient launch is handed that value -- two
    launches, where finite differences take 2 (1 + p).
    Gradient names follow the optimiser's keywords: `length_scale` | `length_scale0..`,
    `noise`."""
    spec = ModelSpec.of(muygps)
    loss_fn = as_loss(loss_fn)
    loss_kwargs = dict(loss_kwargs or {})
    x = fdev(train_features)
    y = fdev(train_targets)
    bi, bnn = idev(batch_indices), idev(batch_nn_indices)
    k = bnn.shape[1]
    d = 1 if x.dim() == 1 else x.shape[1]
    r = 1 if y.dim() == 1 else y.shape[1]
    if loss_fn.loss_id not in (L.LOSS_MSE, L.LOSS_LOOL, L.LOSS_LOOPH):
        raise NotImplementedError(
            f"analytic gradient: loss {loss_fn.name} (mse, lool and looph only)")
    if not ops.fused_loo_supported(d, k, r, spec.kernel_id, spec.metric_id, spec.heteroscedastic,
                                   grad=True):
        raise NotImplementedError("analytic gradient: shape not supported by mgp_fused_loo_grad")
    looph = loss_fn.loss_id == L.LOSS_LOOPH
    lool = loss_fn.loss_id == L.LOSS_LOOL or looph  # (looph finishes like lool, weighted sums)
    analytic = lool and spec.analytic
    if analytic and spec.iteration_count != 1:
        raise NotImplementedError("analytic gradient with AnalyticScale(iteration_count > 1)")
    delta = float(loss_kwargs.get("boundary_scale", loss_fn.default_boundary()))
    loo = ops.FusedLoo(x, y, bi, bnn, kernel_id=spec.kernel_id, metric_id=spec.metric_id,
                       loss_id=loss_fn.loss_id, boundary_scale=delta, want_grad=True)
    # looph with the analytic scale: sigma^2 from a plain launch before the gradient launch
    loo_scale = (ops.FusedLoo(x, y, bi, bnn, kernel_id=spec.kernel_id, metric_id=spec.metric_id,
                              loss_id=L.LOSS_NONE) if looph and analytic else None)
    model_noise = spec.noise(None)
    # Reference quirk (S/gp/hyperparameter/scale.py:206-208): the analytic scale perturbs with
    # the MODEL's nugget whatever `noise=` the optimiser passes.  So sigma^2 never depends on the
    # optimiser's nugget (d sigma^2 / d noise = 0), and when the two differ sigma^2 and its
    # length-scale derivatives come from a second gradient launch at the model's nugget.
    loo_model = []

    def sum_ranks(rec, g):
        from .distributed import allreduce_partials

        both = torch.as_tensor(np.concatenate((rec, g.ravel()))).to(x.device)
        allreduce_partials(both, group)  # per-rank sums of the record and the gradient
        both = both.cpu().numpy()
        return both[:8], both[8:].reshape(L.MGP_GRAD_PARAMS, 5)

    def fn(*args, **theta):
        ls, nz = spec.length_scale_arg(**theta), spec.noise(theta.get("noise"))
        other_noise = (analytic and not spec.heteroscedastic and "noise" in theta
                       and float(theta["noise"]) != float(model_noise))
        sigma2, rec_s, g_s = None, None, None
        if other_noise:
            if not loo_model:
                loo_model.append(ops.FusedLoo(x, y, bi, bnn, kernel_id=spec.kernel_id,
                                              metric_id=spec.metric_id, loss_id=L.LOSS_NONE,
                                              want_grad=True))
            ev = loo_model[0]
            rec_s = ev.record(ev.launch(ls, model_noise))
            g_s = ev.grad.numpy().reshape(L.MGP_GRAD_PARAMS, 5).copy()
            if distributed:
                rec_s, g_s = sum_ranks(rec_s, g_s)
            sigma2 = spec.sigma_from_mean_quadratic_form(
                rec_s[L.P_YKY] / (rec_s[L.P_ROWS] * k))
        elif looph:
            if analytic:
                rs = loo_scale.record(loo_scale.launch(ls, nz))
                if distributed:
                    from .distributed import allreduce_partials

                    rs_dev = torch.as_tensor(rs).to(x.device)
                    allreduce_partials(rs_dev, group)
                    rs = rs_dev.cpu().numpy()
                sigma2 = spec.sigma_from_mean_quadratic_form(rs[L.P_YKY] / (rs[L.P_ROWS] * k))
            else:
                sigma2 = spec.scale()
        rec = loo.record(loo.launch(ls, nz, scale=sigma2 if looph else None))
        g = loo.grad.numpy().reshape(L.MGP_GRAD_PARAMS, 5).copy()
        if distributed:
            rec, g = sum_ranks(rec, g)
        return finish_value_and_grad(
            rec, g, loss_id=loss_fn.loss_id, k=k, d=d, anisotropic=spec.anisotropic,
            analytic=analytic, sigma2=sigma2,
            fixed_scale=spec.scale() if (lool and not analytic) else None,
            g_sigma=g_s)

    return fn
