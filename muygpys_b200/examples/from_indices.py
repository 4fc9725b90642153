"""Same names as `MuyGPyS.examples.from_indices`; every entry point is fused."""

from ..from_indices import (  # noqa: F401
    fast_posterior_mean_from_indices,
    optimize_from_indices,
    posterior_mean_from_indices,
    posterior_variance_from_indices,
    regress_from_indices,
    tensors_from_indices,
)
