"""Batch sampling on the device (S/optimize/batch.py:25-228).

The reference draws indices with numpy on the host, asks sklearn for the neighbours and
filters / balances with fancy indexing on `(n,k)` host arrays.  Here the draw is a device
permutation, the neighbours come from the resident KNN index, the "nonconstant neighbourhood"
filter is one kernel (`mgp_nn_label_mask`, which never materialises `labels[nn_indices]`) and
the compaction / per-class selection are device index operations: the `(n,k)` neighbour array
never visits host memory.  Results are device tensors when the lookup was built from a device
tensor and numpy arrays when it was built from numpy (as every other entry point).

The random stream is torch's device generator, not numpy's: the SETS drawn differ from the
reference's for the same seed; their distribution and every deterministic property (sizes,
uniqueness, class balance, the filter) are the reference's -- tests/test_gpu_batch.py.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._arrays import fdev, idev, like_input
from .neighbors import NN_Wrapper


def _out(nbrs: NN_Wrapper, *tensors):
    if getattr(nbrs, "_host_api", False):
        return tuple(t.cpu().numpy() for t in tensors)
    return tensors


def _batch_nns_dev(nbrs: NN_Wrapper, batch_indices: torch.Tensor) -> torch.Tensor:
    idx, _ = nbrs._query(nbrs.train[batch_indices], nbrs.nn_count + 1)
    return idx[:, 1:].contiguous()  # drop the self-match, as get_batch_nns does


def sample_batch(nbrs_lookup: NN_Wrapper, batch_count: int, train_count: int,
                 generator: Optional[torch.Generator] = None) -> Tuple:
    """Uniform batch without replacement + its neighbours (S/optimize/batch.py:197-228)."""
    dev = nbrs_lookup.train.device
    if train_count > batch_count:
        batch_indices = torch.randperm(train_count, device=dev, generator=generator)[:batch_count]
    else:
        batch_indices = torch.arange(train_count, device=dev)
    batch_indices = batch_indices.to(torch.int64).contiguous()
    return _out(nbrs_lookup, batch_indices, _batch_nns_dev(nbrs_lookup, batch_indices))


def _nonconstant(nbrs: NN_Wrapper, labels):
    lab = fdev(labels)
    indices = torch.arange(lab.shape[0], device=lab.device)
    nn_indices = _batch_nns_dev(nbrs, indices)
    return lab, indices, nn_indices, ops.nn_label_mask(lab, nn_indices)


def full_filtered_batch(nbrs_lookup: NN_Wrapper, labels) -> Tuple:
    """Every training point whose neighbourhood holds more than one label
    (S/optimize/batch.py:67-113)."""
    _, indices, nn_indices, mask = _nonconstant(nbrs_lookup, labels)
    return _out(nbrs_lookup, indices[mask].contiguous(), nn_indices[mask].contiguous())


def sample_balanced_batch(nbrs_lookup: NN_Wrapper, labels, batch_count: int,
                          generator: Optional[torch.Generator] = None) -> Tuple:
    """Up to batch_count / class_count nonconstant-neighbourhood points per class, classes in
    ascending order (S/optimize/batch.py:116-194)."""
    lab, indices, nn_indices, mask = _nonconstant(nbrs_lookup, labels)
    classes = torch.unique(lab)
    each = int(batch_count / classes.numel())
    chosen = []
    for c in classes.tolist():
        cand = indices[mask & (lab == c)]
        take = min(int(cand.numel()), each)
        perm = torch.randperm(cand.numel(), device=lab.device, generator=generator)[:take]
        chosen.append(cand[perm])
    batch_indices = torch.cat(chosen).to(torch.int64).contiguous()
    return _out(nbrs_lookup, batch_indices, nn_indices[batch_indices].contiguous())


def get_balanced_batch(nbrs_lookup: NN_Wrapper, labels, batch_count: int,
                       generator: Optional[torch.Generator] = None) -> Tuple:
    """S/optimize/batch.py:25-64."""
    if len(labels) > batch_count:
        return sample_balanced_batch(nbrs_lookup, labels, batch_count, generator)
    return full_filtered_batch(nbrs_lookup, labels)
