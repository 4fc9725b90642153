"""Same names as `MuyGPyS.gp.noise`."""

from ..noise import HeteroscedasticNoise, HomoscedasticNoise, NoiseFn, NullNoise  # noqa: F401
