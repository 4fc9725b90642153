"""`muygpys_b200.gp` mirrors the import layout of `MuyGPyS.gp` (S/gp/__init__.py)."""

from ..model import MultivariateMuyGPS, MuyGPS  # noqa: F401
