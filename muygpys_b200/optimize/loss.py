"""Same names as `MuyGPyS.optimize.loss`."""

from ..losses import (  # noqa: F401
    LossFn,
    cross_entropy_fn,
    lool_fn,
    looph_fn,
    mse_fn,
    pseudo_huber_fn,
)
