"""Hyperparameter containers with the reference's names and calling conventions.

Mirrors the user-facing behaviour of S/gp/hyperparameter/{scalar,vector,tensor,
scale}.py: `Parameter(val, bounds)`, `VectorParameter(*params)`, `FixedScale`,
`AnalyticScale`.  The reference threads optimiser values through nested
closures; here parameters only resolve `name=value` keyword overrides, which is
all the hot path needs.
"""

from __future__ import annotations

from collections.abc import Sequence
from numbers import Number
from typing import Dict, List, Tuple, Union

import numpy as np


class Parameter:
    """A scalar hyperparameter with optional optimisation bounds.

    `bounds` is "fixed" or a `(lower, upper)` pair; `val` may be a number or,
    for bounded parameters, "sample" / "log_sample" (S/gp/hyperparameter/scalar.py:29-200).
    """

    def __init__(self, val: Union[str, float], bounds: Union[str, Tuple[float, float]] = "fixed"):
        self._set_bounds(bounds)
        self._set_val(val)

    def _set_bounds(self, bounds) -> None:
        if isinstance(bounds, str):
            if bounds != "fixed":
                raise ValueError(f"Unknown bound option {bounds}.")
            self._bounds, self._fixed = (0.0, 0.0), True
            return
        if not hasattr(bounds, "__iter__"):
            raise ValueError(
                f"Unknown bound optiom {bounds} of a non-iterable type {type(bounds)}."
            )
        if len(bounds) != 2:
            raise ValueError(
                f"Provided hyperparameter optimization bounds have unsupported length "
                f"{len(bounds)}."
            )
        for edge in bounds:
            if not isinstance(edge, Number):
                raise ValueError(
                    f"Nonscalar {edge} of type {type(edge)} is not a supported "
                    "hyperparameter bound type."
                )
        lo, hi = float(bounds[0]), float(bounds[1])
        if lo > hi:
            raise ValueError(f"Lower bound {lo} is not lesser than upper bound {hi}.")
        self._bounds, self._fixed = (lo, hi), False

    def _set_val(self, val) -> None:
        if isinstance(val, str):
            if self._fixed:
                raise ValueError(f"Fixed bounds do not support string value ({val}) prompts.")
            lo, hi = self._bounds
            if val == "sample":
                val = float(np.random.uniform(low=lo, high=hi))
            elif val == "log_sample":
                val = float(np.exp(np.random.uniform(low=np.log(lo), high=np.log(hi))))
            else:
                raise ValueError(f"Unsupported string hyperparameter value {val}.")
        if isinstance(val, Sequence) or hasattr(val, "__len__"):
            raise ValueError(f"Nonscalar hyperparameter value {val} is not allowed.")
        val = float(val)
        if not self._fixed:
            lo, hi = self._bounds
            if val < lo - 1e-5:
                raise ValueError(
                    f"Hyperparameter value {val} is lesser than the optimization lower bound {lo}"
                )
            if val > hi + 1e-5:
                raise ValueError(
                    f"Hyperparameter value {val} is greater than the optimization upper bound {hi}"
                )
        self._val = val

    def __call__(self, **kwargs) -> float:
        return self._val

    def get_bounds(self) -> Tuple[float, float]:
        return self._bounds

    def fixed(self) -> bool:
        return self._fixed

    def __str__(self) -> str:
        return f"{type(self).__name__}({self._val}, {'fixed' if self._fixed else self._bounds})"

    __repr__ = __str__


ScalarParam = Parameter


class VectorParameter:
    """An ordered tuple of scalar parameters (anisotropic length scales)."""

    def __init__(self, *params: Parameter):
        for p in params:
            if not isinstance(p, Parameter):
                raise ValueError(f"VectorParameter expects Parameter entries, not {type(p)}")
        self._params = list(params)

    def __len__(self) -> int:
        return len(self._params)

    def __call__(self, **kwargs) -> np.ndarray:
        return np.array([p() for p in self._params], dtype=np.float64)

    def fixed(self) -> bool:
        return all(p.fixed() for p in self._params)

    def __str__(self) -> str:
        return f"{type(self).__name__}({', '.join(str(p) for p in self._params)})"

    __repr__ = __str__


VectorParam = VectorParameter


class _Named:
    """Binds parameter(s) to the keyword name(s) the optimiser uses.

    A scalar named `length_scale` answers to `length_scale=`; a vector answers to
    `length_scale0=`, `length_scale1=`, ... (S/gp/hyperparameter/vector.py:92-127).
    """

    def __init__(self, name: str, param):
        self.name = name
        self.param = param
        if isinstance(param, VectorParameter):
            self.entries = [(f"{name}{i}", p) for i, p in enumerate(param._params)]
        else:
            self.entries = [(name, param)]

    def resolve(self, kwargs: Dict) -> List[float]:
        return [float(kwargs.get(key, p())) for key, p in self.entries]

    def populate(self, table: Dict) -> None:
        for key, p in self.entries:
            table[key] = p

    def append_lists(self, names, params, bounds) -> None:
        for key, p in self.entries:
            if not p.fixed():
                names.append(key)
                params.append(p())
                bounds.append(p.get_bounds())


class TensorParam:
    """A fixed tensor-valued parameter (heteroscedastic noise)."""

    def __init__(self, val):
        if isinstance(val, str):
            raise ValueError("TensorParam class does not support strings.")
        self._val = val

    def __call__(self):
        return self._val

    def fixed(self) -> bool:
        return True

    def get_bounds(self):
        raise NotImplementedError("TensorParam does not support optimization bounds!")

    def append_lists(self, names, params, bounds) -> None:
        return None


class ScaleFn:
    """Variance scale sigma^2 (S/gp/hyperparameter/scale.py:21-109)."""

    def __init__(self, val: float = 1.0, **kwargs):
        self.val = self._check(val)
        self._trained = False

    @staticmethod
    def _check(val) -> float:
        if isinstance(val, Sequence) or (hasattr(val, "__len__") and len(val) != 1):
            raise ValueError(f"Scale parameter must be scalar, not {val}.")
        val = float(val)
        if val <= 0.0:
            raise ValueError(f"Scale parameter must be positive, not {val}.")
        return val

    def _set(self, val) -> None:
        self.val = self._check(val)
        self._trained = True

    def __call__(self) -> float:
        return self.val

    @property
    def trained(self) -> bool:
        return self._trained

    analytic = False
    iteration_count = 1

    def __str__(self) -> str:
        return f"{type(self).__name__}({self.val})"


class FixedScale(ScaleFn):
    """sigma^2 that optimisation leaves alone (scale.py:112-145)."""


class AnalyticScale(ScaleFn):
    """sigma^2 = mean_b y^T (K+eps)^-1 y / k, optionally iterated (scale.py:148-219)."""

    analytic = True

    def __init__(self, iteration_count: int = 1, **kwargs):
        super().__init__(**kwargs)
        self.iteration_count = int(iteration_count)

    def from_mean_quadratic_form(self, sigma0_sq: float) -> float:
        """Finish the reference's fixed-point loop from the one quantity it needs.

        The loop `s <- (s + f(s*K)) / 2` (scale.py:210-216) has f(s*K) = f(K)/s, so
        it only ever needs f(K) = sum y^T K^-1 y / (b k), which the fused kernel
        already produced.
        """
        s = sigma0_sq
        for _ in range(1, self.iteration_count):
            s = 0.5 * (s + sigma0_sq / s)
        return s
