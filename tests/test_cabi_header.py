"""CPU: the shared library loads and exports every symbol include/muygpys_b200.h
declares; enum values agree between header, ctypes layer and oracle."""

import os
import re

import pytest

from muygpys_b200 import _lib as L
from oracle import numpy_oracle as O


def _header():
    with open(L.HEADER_PATH) as f:
        return f.read()


def test_header_symbols_are_bound_and_exported():
    src = _header()
    declared = set(re.findall(r"\b(mgp_[a-z0-9_]+)\s*\(", src))
    declared -= {"mgp_problem"}
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    if not os.path.exists(L.LIB_PATH):
        L.build()
    lib = L.lib()  # raises if any symbol is missing
    assert lib.mgp_version() == int(re.search(r"#define MGP_VERSION (\d+)", src).group(1))


def test_enums_in_sync():
    src = _header()

    def val(name):
        return int(re.search(rf"\b{name}\s*=\s*(-?\d+)", src).group(1))

    for name, py in [("MGP_KERNEL_RBF", L.KERNEL_RBF), ("MGP_KERNEL_MATERN_05", L.KERNEL_MATERN_05),
                     ("MGP_KERNEL_MATERN_15", L.KERNEL_MATERN_15),
                     ("MGP_KERNEL_MATERN_25", L.KERNEL_MATERN_25),
                     ("MGP_KERNEL_MATERN_INF", L.KERNEL_MATERN_INF),
                     ("MGP_METRIC_L2", L.METRIC_L2), ("MGP_METRIC_F2", L.METRIC_F2),
                     ("MGP_LOSS_MSE", L.LOSS_MSE), ("MGP_LOSS_LOOL", L.LOSS_LOOL),
                     ("MGP_LOSS_LOOPH", L.LOSS_LOOPH),
                     ("MGP_LOSS_PSEUDO_HUBER", L.LOSS_PSEUDO_HUBER),
                     ("MGP_LOSS_CROSS_ENTROPY", L.LOSS_CROSS_ENTROPY),
                     ("MGP_P_SQERR", L.P_SQERR), ("MGP_P_COUNT", L.P_COUNT),
                     ("MGP_P_YKY", L.P_YKY), ("MGP_P_ROWS", L.P_ROWS),
                     ("MGP_P_SQERR_V", L.P_SQERR_V), ("MGP_P_LOGV", L.P_LOGV),
                     ("MGP_P_AUX", L.P_AUX), ("MGP_P_BAD", L.P_BAD)]:
        assert val(name) == py, name
    assert (O.KERNEL_RBF, O.KERNEL_MATERN_05, O.KERNEL_MATERN_15, O.KERNEL_MATERN_25,
            O.KERNEL_MATERN_INF) == (L.KERNEL_RBF, L.KERNEL_MATERN_05, L.KERNEL_MATERN_15,
                                     L.KERNEL_MATERN_25, L.KERNEL_MATERN_INF)
    assert (O.METRIC_L2, O.METRIC_F2) == (L.METRIC_L2, L.METRIC_F2)
    assert (O.LOSS_MSE, O.LOSS_LOOL, O.LOSS_LOOPH, O.LOSS_PSEUDO_HUBER,
            O.LOSS_CROSS_ENTROPY) == (L.LOSS_MSE, L.LOSS_LOOL, L.LOSS_LOOPH,
                                      L.LOSS_PSEUDO_HUBER, L.LOSS_CROSS_ENTROPY)


def test_struct_layout_matches_header_order():
    src = _header()
    body = re.search(r"typedef struct mgp_problem \{(.*?)\} mgp_problem;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        decl = stmt.split(",")
        for part in decl:
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert names == [f[0] for f in L.MgpProblem._fields_]


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muygpys_b200 import ops

    with pytest.raises(Exception):
        ops.kernel_apply(0, torch.zeros(4, dtype=torch.float64), 1.0)
