"""Batch sampling (S/optimize/batch.py:197-228): host RNG, neighbours on the GPU."""

from __future__ import annotations

import numpy as np

from ._arrays import idev, like_input
from .neighbors import NN_Wrapper


def sample_batch(nbrs_lookup: NN_Wrapper, batch_count: int, train_count: int):
    if train_count > batch_count:
        batch_indices = np.random.choice(train_count, batch_count, replace=False).astype(np.int64)
    else:
        batch_indices = np.arange(train_count, dtype=np.int64)
    batch_nn_indices, _ = nbrs_lookup.get_batch_nns(batch_indices)
    return batch_indices, batch_nn_indices
