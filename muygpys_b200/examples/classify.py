"""`classify_any` with the signature of `MuyGPyS.examples.classify.classify_any`
(S/examples/classify.py:536-608), end to end on the device: exact KNN against the resident
training set, the label filter (`mgp_nn_label_mask`: neighbourhoods whose labels all agree take
that label without a solve), and the fused posterior mean of the surrogate on the remaining
rows.  The `(t,k)` neighbour indices and the gathered `(t,k,c)` labels of the reference never
exist in host memory."""

from __future__ import annotations

from time import perf_counter
from typing import Dict, Tuple

import torch

from .. import fused, ops
from .._arrays import fdev, like_input
from ..adapt import ModelSpec
from ..neighbors import NN_Wrapper


def classify_any(surrogate, test_features, train_features, train_nbrs_lookup, train_labels,
                 sync_timing: bool = True) -> Tuple[object, Dict[str, float]]:
    spec = ModelSpec.of(surrogate)
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
    x = nbrs.train
    labels = fdev(train_labels)
    q = fdev(test_features)
    if q.dim() == 1:
        q = q[:, None]

    def sync():
        if sync_timing:
            torch.cuda.current_stream().synchronize()

    t0 = perf_counter()
    nn_idx, _ = nbrs._query(q, nbrs.nn_count)
    sync()
    t1 = perf_counter()
    # rows whose neighbours all carry the same label take that label (one-hot row of the
    # nearest neighbour), classify.py:575-585
    mask = ops.nn_label_mask(labels, nn_idx)
    predictions = labels[nn_idx[:, 0]].clone()
    sync()
    t2 = perf_counter()
    rows = torch.nonzero(mask)[:, 0]
    if rows.numel() > 0:
        out = fused.fused_call(spec, rows, nn_idx[rows].contiguous(), q, x, labels,
                               want_mean=True, want_var=False)
        predictions[rows] = out["mean"]
    sync()
    t3 = perf_counter()
    host = (test_features, train_features, train_labels)
    return like_input(predictions, *host), {"nn": t1 - t0, "agree": t2 - t1, "pred": t3 - t2}
