"""Same names as `MuyGPyS.gp.hyperparameter`."""

from ..hyperparameter import (  # noqa: F401
    AnalyticScale,
    FixedScale,
    Parameter,
    ScalarParam,
    ScaleFn,
    TensorParam,
    VectorParam,
    VectorParameter,
)
