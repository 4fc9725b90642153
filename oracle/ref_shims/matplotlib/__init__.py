"""Import shim (build container only): MuyGPyS/_test/sampler.py:6 imports pyplot."""
