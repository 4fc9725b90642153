"""`regress_any` with the signature of `MuyGPyS.examples.regress.regress_any`
(S/examples/regress.py:602-668): test FEATURES in, posterior mean and variance out.

The reference finds the neighbours with sklearn on the host and hands the `(t,k)` int64 index
array to `regress_from_indices`.  Here the search runs on the GPU against the resident training
set (`muygpys_b200.neighbors.NN_Wrapper`) and its result feeds the fused kernel directly: the
neighbour indices never exist in host memory, so a step moves `8 t d` bytes up and `16 t` bytes
down instead of `8 t (k + d + 1)` up.
"""

from __future__ import annotations

from time import perf_counter
from typing import Dict, Tuple

import torch

from .. import fused
from .._arrays import fdev, like_input
from ..adapt import ModelSpec
from ..neighbors import NN_Wrapper


def regress_any(regressor, test_features, train_features, train_nbrs_lookup, train_targets,
                sync_timing: bool = True) -> Tuple[object, object, Dict[str, float]]:
    """(means, variances, timing).  `train_nbrs_lookup` is a `muygpys_b200.neighbors.NN_Wrapper`
    (any other lookup object with a `.train` array and `.nn_count` is re-indexed on the device
    once and cached on the object).  `sync_timing=False` skips the stream synchronisations that
    make the "nn" / "pred" wall-clock entries meaningful."""
    spec = ModelSpec.of(regressor)
    nbrs = train_nbrs_lookup
    if not isinstance(nbrs, NN_Wrapper):
        cached = getattr(nbrs, "_mgp_device_index", None)
        if cached is None:
            cached = NN_Wrapper(fdev(nbrs.train), nbrs.nn_count)
            try:
                nbrs._mgp_device_index = cached
            except AttributeError:
                pass
        nbrs = cached
    x = nbrs.train  # device-resident copy the index was built on
    y = fdev(train_targets)
    q = fdev(test_features)
    if q.dim() == 1:
        q = q[:, None]

    def sync():
        if sync_timing:
            torch.cuda.current_stream().synchronize()

    t0 = perf_counter()
    nn_idx, _ = nbrs._query(q, nbrs.nn_count)  # stays on the device
    sync()
    t1 = perf_counter()
    out = fused.fused_call(spec, None, nn_idx, q, x, y, want_mean=True, want_var=True)
    mean = out["mean"][:, 0] if y.dim() == 1 else out["mean"]
    host = (test_features, train_features, train_targets)
    mean, var = like_input(mean, *host), like_input(out["var"], *host)
    sync()
    t2 = perf_counter()
    return mean, var, {"nn": t1 - t0, "agree": 0.0, "pred": t2 - t1}
