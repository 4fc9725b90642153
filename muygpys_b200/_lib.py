"""ctypes binding of libmuygpys_b200.so (the C ABI declared in include/muygpys_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, the caller gets an exception.  `build()` (re)compiles it in-tree with nvcc
for sm_100a; the driver calls it through `__graft_entry__.build()`.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmuygpys_b200.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "muygpys_b200.h")

MGP_OK = 0
MGP_ERR_BAD_ARG = -1
MGP_ERR_UNSUPPORTED = -2
MGP_ERR_CUDA = -3
MGP_ERR_WORKSPACE = -4
MGP_PARTIALS = 8
MGP_MAX_PEERS = 8
MGP_GRAD_PARAMS = 4
MGP_GRAD_DOUBLES = 20
MGP_MAX_ANISO_DIM = 32

# enums (mgp_kernel_id, mgp_metric_id, mgp_loss_id, partial slots)
KERNEL_RBF, KERNEL_MATERN_05, KERNEL_MATERN_15, KERNEL_MATERN_25, KERNEL_MATERN_INF = range(5)
METRIC_L2, METRIC_F2 = 0, 1
(LOSS_NONE, LOSS_MSE, LOSS_LOOL, LOSS_LOOPH, LOSS_PSEUDO_HUBER, LOSS_CROSS_ENTROPY) = range(6)
(P_SQERR, P_COUNT, P_YKY, P_ROWS, P_SQERR_V, P_LOGV, P_AUX, P_BAD) = range(8)

_dp = C.c_void_p  # device pointers travel as integers
_i64, _i32, _f64, _sz = C.c_int64, C.c_int32, C.c_double, C.c_size_t


class MgpProblem(C.Structure):
    """Mirror of `struct mgp_problem` (include/muygpys_b200.h)."""

    _fields_ = [
        ("train_x", _dp), ("query_x", _dp), ("query_idx", _dp), ("nn_idx", _dp),
        ("train_y", _dp),
        ("n", _i64), ("t", _i64), ("b", _i64),
        ("k", _i32), ("d", _i32), ("r", _i32),
        ("kernel_id", _i32), ("metric_id", _i32), ("length_scale_count", _i32),
        ("length_scale", C.POINTER(C.c_double)),
        ("noise", _f64), ("noise_bk", _dp), ("scale", _f64),
        ("mean", _dp), ("var", _dp), ("yky", _dp), ("coeffs", _dp), ("status", _dp),
    ]


class MgpPeerGroup(C.Structure):
    """Mirror of `struct mgp_peer_group`."""

    _fields_ = [("rank", _i32), ("world", _i32), ("epoch", C.c_uint64),
                ("peer_buf", C.c_void_p * 8)]


_PP = C.POINTER(MgpProblem)
_PG = C.POINTER(MgpPeerGroup)
_HOSTD = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    "mgp_version": (C.c_int, []),
    "mgp_last_error": (C.c_char_p, []),
    "mgp_fused_workspace_bytes": (_sz, [_PP]),
    "mgp_fused_posterior": (C.c_int, [_PP, _dp, _sz, _dp]),
    "mgp_fused_posterior_host": (C.c_int, [_PP, _dp, _dp, _dp, _dp, _dp, _sz, _dp]),
    "mgp_fused_posterior_host32": (C.c_int, [_PP, _dp, _dp, _dp, _dp, _dp, _dp, _sz, _dp]),
    "mgp_set_fused_variant": (C.c_int, [_i32]),
    "mgp_fused_loo_workspace_bytes": (_sz, [_PP]),
    "mgp_fused_loo": (C.c_int, [_PP, _i32, _f64, _dp, _dp, _sz, _dp]),
    "mgp_peer_buffer_bytes": (_sz, []),
    "mgp_peer_sum8": (C.c_int, [_dp, _PG, _dp]),
    "mgp_fused_loo_peers": (C.c_int, [_PP, _i32, _f64, _dp, _dp, _sz, _PG, _dp]),
    "mgp_fused_loo_grad": (C.c_int, [_PP, _i32, _f64, _dp, _dp, _dp, _sz, _PG, _dp]),
    "mgp_loss_workspace_bytes": (_sz, [_i64, _i32]),
    "mgp_loss_partials": (C.c_int, [_i32, _dp, _dp, _dp, _dp, _dp, _f64, _i64, _i32, _dp, _dp,
                                    _sz, _dp]),
    "mgp_knn_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "mgp_knn": (C.c_int, [_dp, _i64, _dp, _i64, _i32, _i32, _i32, _dp, _dp, _dp, _dp, _sz, _dp]),
    "mgp_knn_grid_cells": (C.c_int, [_dp, _i64, _i32, C.POINTER(C.c_int32), _HOSTD, _f64, _dp, _dp]),
    "mgp_knn_grid_query": (C.c_int, [_dp, _dp, _dp, _i64, _i32, C.POINTER(C.c_int32), _HOSTD, _f64,
                                     _dp, _dp, _i64, _i32, _dp, _dp, _dp, _dp]),
    "mgp_fast_mean": (C.c_int, [_PP, _dp, _dp, _dp]),
    "mgp_crosswise_diffs": (C.c_int, [_dp, _dp, _dp, _dp, _i64, _i32, _i32, _dp, _dp]),
    "mgp_pairwise_diffs": (C.c_int, [_dp, _dp, _i64, _i32, _i32, _dp, _dp]),
    "mgp_metric_reduce": (C.c_int, [_i32, _dp, _i64, _i32, _HOSTD, _dp, _dp]),
    "mgp_crosswise_dists": (C.c_int, [_i32, _dp, _dp, _dp, _dp, _i64, _i32, _i32, _dp, _dp]),
    "mgp_pairwise_dists": (C.c_int, [_i32, _dp, _dp, _i64, _i32, _i32, _dp, _dp]),
    "mgp_kernel_apply": (C.c_int, [_i32, _dp, _f64, _i64, _dp, _dp]),
    "mgp_perturb": (C.c_int, [_dp, _i64, _i32, _f64, _dp, _dp, _dp]),
    "mgp_solve_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "mgp_solve": (C.c_int, [_dp, _dp, _dp, _i64, _i32, _i32, _f64, _dp, _dp, _dp, _dp, _dp, _dp,
                            _sz, _dp]),
    "mgp_nn_label_mask": (C.c_int, [_dp, _i64, _dp, _i64, _i32, _dp, _dp]),
    "mgp_rowdot": (C.c_int, [_dp, _dp, _i64, _i32, _i32, _dp, _dp]),
    "mgp_fp64_probe": (C.c_int, [_i32, _i32, _i32, _i32, _dp, _dp]),
}

_lib = None


class MgpError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Returns the .so path.

    `make` only recompiles what changed, so an up-to-date tree proves nothing about the
    toolchain: every call therefore also compiles one small unit (probe.cu) from scratch into a
    temporary directory, and the outcome of both goes into profiles/build_record.json
    (`build_mode`, `build_exercised`, objects recompiled, nvcc version, size / hash of the .so)."""
    import hashlib
    import json
    import tempfile
    import time

    t0 = time.time()
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise MgpError(f"building {LIB_PATH} failed (exit {res.returncode})")
    recompiled = sum(1 for ln in res.stdout.splitlines() if " -c " in ln and "nvcc" in ln)
    canary_ok = False
    with tempfile.TemporaryDirectory() as tmp:
        canary = subprocess.run(
            ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode",
             "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
             "-c", os.path.join(CSRC_DIR, "probe.cu"), "-o", os.path.join(tmp, "probe.o")],
            capture_output=True, text=True)
        canary_ok = canary.returncode == 0 and os.path.getsize(os.path.join(tmp, "probe.o")) > 0
    if not canary_ok:
        raise MgpError("nvcc could not compile csrc/probe.cu for sm_100a:\n" + canary.stderr)
    try:
        ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
        ver = ver.strip().splitlines()[-2:]
        with open(LIB_PATH, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        record = {
            "build_mode": "make + nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo "
                          "(muygpys_b200/csrc/Makefile), in-tree .so",
            "build_exercised": True,
            "objects_recompiled_by_this_call": recompiled,
            "library_was_up_to_date": recompiled == 0,
            "canary": "csrc/probe.cu compiled from scratch for sm_100a by this call: ok",
            "nvcc": ver, "so_bytes": os.path.getsize(LIB_PATH), "so_sha256": digest,
            "seconds": round(time.time() - t0, 1),
            "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        }
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        with open(os.path.join(root, "profiles", "build_record.json"), "w") as f:
            json.dump(record, f, indent=1)
            f.write("\n")
    except OSError:
        pass  # (read-only checkout: the record is a courtesy, the build itself succeeded)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgpError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {CSRC_DIR}`).  muygpys_b200 has no CPU fallback."
        )
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if handle.mgp_version() < 100:
        raise MgpError("libmuygpys_b200.so is older than this Python package")
    _lib = handle
    return _lib


def check(rc: int) -> None:
    """Map a negative mgp_status to the exception the reference would raise."""
    if rc == MGP_OK:
        return
    msg = lib().mgp_last_error().decode("utf-8", "replace")
    if rc in (MGP_ERR_BAD_ARG,):
        raise ValueError(msg)
    if rc == MGP_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MgpError(f"muygpys_b200 error {rc}: {msg}")
