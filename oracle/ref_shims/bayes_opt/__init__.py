"""Import shim used ONLY by oracle/make_golden.py in the build container.

The reference imports ``bayes_opt.BayesianOptimization`` at module import time
(/root/reference/src/MuyGPyS/_src/optimize/chassis/numpy.py:9) and the package is
not installed here.  Nothing on the hot path uses it.
"""


class BayesianOptimization:  # pragma: no cover - shim
    def __init__(self, *args, **kwargs):
        raise ModuleNotFoundError("bayes_opt is not installed (shim)")
