"""Same names as `MuyGPyS.optimize.batch`."""

from ..batch import sample_batch  # noqa: F401
