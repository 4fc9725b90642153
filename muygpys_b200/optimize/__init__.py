"""`muygpys_b200.optimize` mirrors `MuyGPyS.optimize` (S/optimize/__init__.py)."""

from ..optimizers import Bayes_optimize, L_BFGS_B_optimize, OptimizeFn  # noqa: F401
