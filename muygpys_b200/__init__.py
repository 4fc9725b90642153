"""muygpys_b200: the MuyGPyS per-neighbourhood GP hot path as hand-written sm_100a CUDA.

Host code is Python over torch CUDA tensors calling a C-ABI shared library
(include/muygpys_b200.h) through ctypes.  There is no CPU fallback.
"""

__version__ = "0.1.0"
