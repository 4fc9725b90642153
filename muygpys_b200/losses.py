"""Loss functors (S/optimize/loss.py:181-396) evaluated by the K-loss CUDA kernels.

Each `LossFn` is callable like the reference's (`loss(predictions, targets[,
variances, scale], **kw) -> float`) and carries the id the fused objective uses.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._arrays import fdev


class LossFn:
    def __init__(self, name: str, loss_id: int, needs_variance: bool, finish):
        self.name = name
        self.loss_id = loss_id
        self.needs_variance = needs_variance
        self._finish = finish

    def partials(self, predictions, targets, variances=None, scale=None, yky=None, **kwargs):
        scale_dev = None
        if scale is not None:
            scale_dev = scale if isinstance(scale, torch.Tensor) else torch.tensor(
                [float(scale)], dtype=torch.float64, device=fdev(predictions).device)
        return ops.loss_partials(
            self.loss_id, fdev(predictions), fdev(targets),
            None if variances is None else fdev(variances), yky, scale_dev,
            boundary_scale=float(kwargs.get("boundary_scale", self.default_boundary())))

    def default_boundary(self) -> float:
        return {L.LOSS_LOOPH: 3.0, L.LOSS_PSEUDO_HUBER: 1.5}.get(self.loss_id, 1.0)

    def finish(self, record: np.ndarray) -> float:
        return float(self._finish(record))

    def __call__(self, predictions, targets, *args, **kwargs) -> float:
        if self.needs_variance:
            if len(args) < 2:
                raise TypeError(f"{self.name} expects (predictions, targets, variances, scale)")
            rec = self.partials(predictions, targets, args[0], args[1], **kwargs)
        else:
            rec = self.partials(predictions, targets, **kwargs)
        return self.finish(rec.cpu().numpy())

    def __str__(self) -> str:
        return self.name


mse_fn = LossFn("mse", L.LOSS_MSE, False, lambda p: p[L.P_SQERR] / p[L.P_COUNT])
cross_entropy_fn = LossFn("cross_entropy", L.LOSS_CROSS_ENTROPY, False, lambda p: p[L.P_AUX])
pseudo_huber_fn = LossFn("pseudo_huber", L.LOSS_PSEUDO_HUBER, False, lambda p: p[L.P_AUX])
lool_fn = LossFn("lool", L.LOSS_LOOL, True, lambda p: p[L.P_AUX])
looph_fn = LossFn("looph", L.LOSS_LOOPH, True, lambda p: p[L.P_AUX])


_REFERENCE_LOSS_NAMES = {"_mse_fn": mse_fn, "_cross_entropy_fn": cross_entropy_fn,
                         "_pseudo_huber_fn": pseudo_huber_fn, "_lool_fn": lool_fn,
                         "_looph_fn": looph_fn}


def as_loss(loss_fn) -> LossFn:
    """Our LossFn for either family of loss objects: a `MuyGPyS.optimize.loss.LossFn` wraps the
    numpy-backend function in `._fn` (S/optimize/loss.py:205-213), whose name picks ours."""
    if isinstance(loss_fn, LossFn):
        return loss_fn
    inner = getattr(loss_fn, "_fn", loss_fn)
    ours = _REFERENCE_LOSS_NAMES.get(getattr(inner, "__name__", ""))
    if ours is None:
        raise NotImplementedError(f"loss function {loss_fn!r} is not on the fused path")
    return ours
