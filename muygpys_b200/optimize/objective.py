"""Same names as `MuyGPyS.optimize.objective`, plus the fused factory."""

from ..objective import make_fused_loo_crossval_fn, make_loo_crossval_fn  # noqa: F401
