"""Metric and deformation functors with the reference's interface, CUDA-backed.

`l2` / `F2` mirror S/gp/deformation/metric.py:237-265, `Isotropy` mirrors
S/gp/deformation/isotropy.py:22-161 (tensors are DISTANCES) and `Anisotropy`
mirrors S/gp/deformation/anisotropy.py:15-143 (tensors are DIFFERENCES with a
trailing feature axis).  All tensor makers run the staged K5 kernels.
"""

from __future__ import annotations

from typing import Optional

from . import _lib as L
from . import ops
from ._arrays import fdev, idev, like_input
from .hyperparameter import Parameter, VectorParameter, _Named


class MetricFn:
    def __init__(self, name: str, metric_id: int):
        self.name = name
        self.metric_id = metric_id

    def __call__(self, diffs, length_scale=None):
        """`(..., d)` differences -> `(...)` distances (S/_src/gp/tensors/numpy.py:89-94)."""
        out = ops.metric_reduce(self.metric_id, fdev(diffs), length_scale=length_scale)
        return like_input(out, diffs)

    def crosswise_differences(self, data, nn_data, data_indices, nn_indices, **kwargs):
        out = ops.crosswise_diffs(fdev(data), fdev(nn_data), idev(data_indices), idev(nn_indices))
        return like_input(out, data, nn_data)

    def pairwise_differences(self, data, nn_indices, **kwargs):
        return like_input(ops.pairwise_diffs(fdev(data), idev(nn_indices)), data)

    def crosswise_distances(self, data, nn_data, data_indices, nn_indices, **kwargs):
        out = ops.crosswise_dists(self.metric_id, fdev(data), fdev(nn_data), idev(data_indices),
                                  idev(nn_indices))
        return like_input(out, data, nn_data)

    def pairwise_distances(self, data, nn_indices, **kwargs):
        return like_input(ops.pairwise_dists(self.metric_id, fdev(data), idev(nn_indices)), data)

    def length_scale_factor(self, length_scale: float) -> float:
        """l2: x / l ; F2: x / l**2 (metric.py:241,264), as a multiplier."""
        if self.metric_id == L.METRIC_L2:
            return 1.0 / length_scale
        return 1.0 / (length_scale * length_scale)

    def apply_length_scale(self, dists, length_scale: float):
        return dists * self.length_scale_factor(float(length_scale))

    def __str__(self) -> str:
        return self.name

    __repr__ = __str__


l2 = MetricFn("l2", L.METRIC_L2)
F2 = MetricFn("F2", L.METRIC_F2)


class DeformationFn:
    anisotropic = False


class Isotropy(DeformationFn):
    """One length scale applied to a scalar distance."""

    def __init__(self, metric: MetricFn, length_scale: Parameter):
        if not isinstance(length_scale, Parameter):
            raise ValueError(
                f"Expected ScalarParam type for length_scale, not {type(length_scale)}"
            )
        self.metric = metric
        self.length_scale = _Named("length_scale", length_scale)

    def length_scales(self, **kwargs):
        return self.length_scale.resolve(kwargs)

    def __call__(self, dists, length_scale: Optional[float] = None, **kwargs):
        if length_scale is None:
            length_scale = self.length_scale.param()
        return self.metric.apply_length_scale(dists, length_scale)

    def pairwise_tensor(self, data, nn_indices, **kwargs):
        return self.metric.pairwise_distances(data, nn_indices)

    def crosswise_tensor(self, data, nn_data, data_indices, nn_indices, **kwargs):
        return self.metric.crosswise_distances(data, nn_data, data_indices, nn_indices)

    def __str__(self) -> str:
        return f"Isotropy({self.metric}, {self.length_scale.param})"


class Anisotropy(DeformationFn):
    """Per-feature length scales applied to difference tensors."""

    anisotropic = True

    def __init__(self, metric: MetricFn, length_scale: VectorParameter):
        if not isinstance(length_scale, VectorParameter):
            raise ValueError(
                f"Expected VectorParam type for length_scale, not {type(length_scale)}"
            )
        self.metric = metric
        self.length_scale = _Named("length_scale", length_scale)

    def length_scales(self, **kwargs):
        return self.length_scale.resolve(kwargs)

    def __call__(self, diffs, **length_scales):
        ls = self.length_scales(**length_scales)
        if diffs.shape[-1] != len(ls):
            raise ValueError(
                f"Difference tensor of shape {tuple(diffs.shape)} must have final dimension "
                f"size of {len(ls)}"
            )
        return self.metric(diffs, length_scale=ls)

    def pairwise_tensor(self, data, nn_indices, **kwargs):
        return self.metric.pairwise_differences(data, nn_indices)

    def crosswise_tensor(self, data, nn_data, data_indices, nn_indices, **kwargs):
        return self.metric.crosswise_differences(data, nn_data, data_indices, nn_indices)

    def __str__(self) -> str:
        return f"Anisotropy({self.metric}, {self.length_scale.param})"
