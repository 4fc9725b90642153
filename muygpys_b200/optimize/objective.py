"""Same names as `MuyGPyS.optimize.objective`, plus the fused factory."""

from ..objective import (  # noqa: F401
    make_fused_loo_crossval_fn,
    make_fused_loo_value_and_grad_fn,
    make_loo_crossval_fn,
)
