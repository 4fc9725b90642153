"""Array plumbing between the numpy-facing reference API and CUDA tensors."""

from __future__ import annotations

import numpy as np
import torch

f64 = torch.float64
i64 = torch.int64


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "muygpys_b200 needs a CUDA device (B200); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def is_host(x) -> bool:
    return isinstance(x, np.ndarray) or (isinstance(x, torch.Tensor) and not x.is_cuda)


def to_dev(x, dtype=f64):
    """numpy / CPU tensor / CUDA tensor -> contiguous CUDA tensor of `dtype`."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.ascontiguousarray(x))
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def fdev(x):
    return to_dev(x, f64)


def idev(x):
    return to_dev(x, i64)


def _is_f32(a) -> bool:
    return (isinstance(a, np.ndarray) and a.dtype == np.float32) or (
        isinstance(a, torch.Tensor) and a.dtype == torch.float32)


def like_input(result: torch.Tensor, *inputs):
    """Return `result` as numpy when the caller handed us host arrays.

    fp32 I/O mode (the counterpart of the reference's MUYGPYS_FTYPE=32,
    S/_src/math/numpy.py:92): when the caller's FEATURE / TARGET arrays are float32, floating
    results come back as float32.  The kernels always compute in fp64 on the (exactly) widened
    values, so the mode costs nothing in accuracy beyond the rounding of the inputs themselves
    and halves the bytes that cross the host link."""
    floats = [a for a in inputs if isinstance(a, (np.ndarray, torch.Tensor))
              and (a.dtype in (np.float32, np.float64) if isinstance(a, np.ndarray)
                   else a.dtype in (torch.float32, torch.float64))]
    if result.dtype == f64 and floats and all(_is_f32(a) for a in floats):
        result = result.to(torch.float32)
    if any(isinstance(a, np.ndarray) for a in inputs):
        return result.cpu().numpy()
    return result
