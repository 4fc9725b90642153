"""Same names as `MuyGPyS.optimize.batch`."""

from ..batch import (  # noqa: F401
    full_filtered_batch,
    get_balanced_batch,
    sample_balanced_batch,
    sample_batch,
)
