"""Noise models (S/gp/noise/{homoscedastic,heteroscedastic,null}.py)."""

from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from . import ops
from ._arrays import fdev, like_input
from .hyperparameter import Parameter, TensorParam


class NoiseFn:
    heteroscedastic = False

    def fixed(self) -> bool:
        return True

    def append_lists(self, names, params, bounds) -> None:
        return None


class HomoscedasticNoise(Parameter, NoiseFn):
    """A single nugget tau^2 added to every diagonal; may be optimised as `noise=`."""

    def __init__(self, val: Union[str, float], bounds: Union[str, Tuple[float, float]] = "fixed"):
        Parameter.__init__(self, val, bounds)
        if not self.fixed() and (self._bounds[0] < 0.0 or self._bounds[1] < 0.0):
            raise ValueError(
                f"Homoscedastic noise optimization bounds {self._bounds} are not strictly "
                "positive!"
            )

    def value(self, noise: Optional[float] = None) -> float:
        return self._val if noise is None else float(noise)

    def perturb(self, Kin, noise: Optional[float] = None, **kwargs):
        return like_input(ops.perturb(fdev(Kin), self.value(noise)), Kin)

    def append_lists(self, names, params, bounds) -> None:
        if not self.fixed():
            names.append("noise")
            params.append(self())
            bounds.append(self.get_bounds())


class HeteroscedasticNoise(TensorParam, NoiseFn):
    """Per-neighbour nugget of shape (batch_count, nn_count); never optimised."""

    heteroscedastic = True

    def __init__(self, val):
        TensorParam.__init__(self, val)
        t = fdev(val)
        if bool((t < 0).any()):
            raise ValueError("Heteroscedastic noise values are not strictly non-negative!")
        self._dev = t

    def value(self, noise=None) -> torch.Tensor:
        return self._dev

    def perturb(self, Kin, **kwargs):
        return like_input(ops.perturb(fdev(Kin), self._dev), Kin)


class NullNoise(NoiseFn):
    def __call__(self, *args, **kwargs) -> float:
        return 0.0

    def value(self, noise=None) -> float:
        return 0.0

    def perturb(self, Kin, **kwargs):
        return Kin
