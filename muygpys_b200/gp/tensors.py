"""Tensor helpers with the names of `MuyGPyS.gp.tensors` (S/gp/tensors.py:23-160)."""

from __future__ import annotations

import torch

from .. import ops
from .._arrays import fdev, idev, like_input


def make_heteroscedastic_tensor(measurement_noise, batch_nn_indices):
    """measurement_noise[batch_nn_indices]   (S/_src/gp/tensors/numpy.py:11-15)."""
    out = fdev(measurement_noise)[idev(batch_nn_indices)]
    return like_input(out, measurement_noise, batch_nn_indices)


def fast_nn_update(train_nn_indices):
    """[i | nn_0 .. nn_{k-2}] per training point   (S/_src/gp/tensors/numpy.py:97-108)."""
    nn = idev(train_nn_indices)
    own = torch.arange(nn.shape[0], dtype=nn.dtype, device=nn.device)[:, None]
    return like_input(torch.cat((own, nn[:, :-1]), dim=1).contiguous(), train_nn_indices)


def make_fast_predict_tensors(batch_nn_indices, train_features, train_targets):
    """(pairwise DIFFERENCES (n,k,k,d), nn targets) of the updated neighbour sets
    (S/_src/gp/tensors/numpy.py:18-37); the caller applies its metric."""
    nn_fast = idev(fast_nn_update(idev(batch_nn_indices)))
    pairwise = ops.pairwise_diffs(fdev(train_features), nn_fast)
    targets = fdev(train_targets)[nn_fast]
    host = (batch_nn_indices, train_features, train_targets)
    return like_input(pairwise, *host), like_input(targets, *host)


def batch_features_tensor(features, batch_indices):
    return like_input(fdev(features)[idev(batch_indices)], features, batch_indices)
