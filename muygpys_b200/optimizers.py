"""Outer optimisers (S/optimize/chassis.py:23-194, S/_src/optimize/chassis/numpy.py).

The optimiser itself stays on the host (scipy L-BFGS-B with finite differences,
exactly what the reference uses); what changes is that each objective evaluation
is a GPU launch.  `OptimizeFn.__call__` keeps the reference's staged signature;
`.from_indices` is the fused path used by `optimize_from_indices`.
"""

from __future__ import annotations

from copy import deepcopy
from typing import Callable, Dict, Optional

from scipy import optimize as _sciopt

from .losses import LossFn, lool_fn
from .objective import make_fused_loo_crossval_fn, make_loo_crossval_fn


def _updated_copy(muygps, names, bounds, values: Dict[str, float]):
    new = deepcopy(muygps)
    for i, key in enumerate(names):
        lo, hi = bounds[i]
        val = min(max(float(values[key]), lo), hi)
        if key == "noise":
            new.noise._set_val(val)
        else:
            new.kernel._hyperparameters[key]._set_val(val)
    new._make()
    return new


def _scipy_optimize(muygps, obj_fn: Callable, verbose: bool = False, **kwargs):
    names, x0, bounds = muygps.get_opt_params()
    if verbose:
        print(f"parameters to be optimized: {names}")
        print(f"bounds: {bounds}")
        print(f"initial x0: {x0}")

    def negated(x, *args):
        return -obj_fn(*args, **{name: x[i] for i, name in enumerate(names)})

    res = _sciopt.minimize(negated, x0, method="L-BFGS-B", bounds=bounds, **kwargs)
    if verbose:
        print(f"optimizer results: \n{res}")
    return _updated_copy(muygps, names, bounds, {n: res.x[i] for i, n in enumerate(names)})


def _bayes_opt_optimize(muygps, obj_fn: Callable, verbose: bool = False, **kwargs):
    try:
        from bayes_opt import BayesianOptimization
    except Exception as exc:  # pragma: no cover - optional dependency
        raise ModuleNotFoundError("bayes_opt is not installed") from exc
    names, x0, bounds = muygps.get_opt_params()
    opt_kw = {k: kwargs[k] for k in kwargs
              if k in {"random_state", "verbose", "bounds_transformer", "allow_duplicate_points"}}
    opt_kw.setdefault("verbose", 2 if verbose else 0)
    opt_kw.setdefault("allow_duplicate_points", True)
    max_kw = {k: kwargs[k] for k in kwargs
              if k in {"init_points", "n_iter", "acq", "kappa", "kappa_decay",
                       "kappa_decay_delay", "xi"}}
    max_kw.setdefault("init_points", 5)
    max_kw.setdefault("n_iter", 20)
    optimizer = BayesianOptimization(
        f=obj_fn, pbounds={n: tuple(bounds[i]) for i, n in enumerate(names)}, **opt_kw)
    optimizer.probe({n: x0[i] for i, n in enumerate(names)}, lazy=True)
    optimizer.maximize(**max_kw)
    return _updated_copy(muygps, names, bounds, optimizer.max["params"])


class OptimizeFn:
    def __init__(self, optimize_fn: Callable, make_obj_fn: Callable = make_loo_crossval_fn):
        self._fn = optimize_fn
        self._make_obj_fn = make_obj_fn

    def make_obj_fn(self, muygps, batch_targets, batch_nn_targets, crosswise_diffs,
                    pairwise_diffs, batch_features=None, target_mask=None,
                    loss_fn: LossFn = lool_fn, loss_kwargs: Dict = dict(), **kwargs) -> Callable:
        return self._make_obj_fn(
            loss_fn, muygps.kernel.get_opt_fn(), muygps.get_opt_mean_fn(),
            muygps.get_opt_var_fn(), muygps.get_opt_scale_fn(), pairwise_diffs, crosswise_diffs,
            batch_nn_targets, batch_targets, batch_features=batch_features,
            target_mask=target_mask, loss_kwargs=loss_kwargs)

    def __call__(self, muygps, batch_targets, batch_nn_targets, crosswise_diffs, pairwise_diffs,
                 batch_features=None, loss_fn: LossFn = lool_fn, loss_kwargs: Dict = dict(),
                 target_mask=None, verbose: bool = False, **kwargs):
        obj_fn = self.make_obj_fn(muygps, batch_targets, batch_nn_targets, crosswise_diffs,
                                  pairwise_diffs, batch_features=batch_features,
                                  target_mask=target_mask, loss_fn=loss_fn,
                                  loss_kwargs=loss_kwargs)
        return self._fn(muygps, obj_fn, verbose=verbose, **kwargs)

    def from_indices(self, muygps, batch_indices, batch_nn_indices, train_features,
                     train_targets, loss_fn: LossFn = lool_fn,
                     loss_kwargs: Optional[Dict] = None, target_mask=None, verbose: bool = False,
                     group=None, distributed: bool = False, **kwargs):
        obj_fn = make_fused_loo_crossval_fn(
            muygps, loss_fn, batch_indices, batch_nn_indices, train_features, train_targets,
            target_mask=target_mask, loss_kwargs=loss_kwargs, group=group,
            distributed=distributed)
        return self._fn(muygps, obj_fn, verbose=verbose, **kwargs)


L_BFGS_B_optimize = OptimizeFn(_scipy_optimize)
Bayes_optimize = OptimizeFn(_bayes_opt_optimize)
