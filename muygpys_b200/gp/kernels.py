"""Same names as `MuyGPyS.gp.kernels`."""

from ..covariance import RBF, KernelFn, Matern  # noqa: F401
