"""Same names as `MuyGPyS.gp.deformation`."""

from ..deformation import F2, Anisotropy, DeformationFn, Isotropy, MetricFn, l2  # noqa: F401
