"""pytest configuration: registers the `gpu` marker and shared helpers."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)"
    )


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def assert_close(actual, desired, rtol=1e-10, what=""):
    """north_star tolerance: |a-b| <= rtol * max(|b|, ||b||_inf)  (SURVEY.md section 7)."""
    a = np.asarray(actual, dtype=np.float64)
    b = np.asarray(desired, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    if b.size == 0:
        return
    scale = np.maximum(np.abs(b), np.max(np.abs(b)))
    err = np.abs(a - b)
    bad = err > rtol * scale
    assert not np.any(bad) and np.all(np.isfinite(a)), (
        f"{what}: max rel err {np.max(err / np.maximum(scale, 1e-300)):.3e} > {rtol:g}"
    )
