"""CPU (gloo, world_size 2): the data-parallel host logic -- chunk rule, row
sharding, the single SUM all-reduce of partial records and how a loss is finished
from them -- reproduces the full-batch reference values.  The per-rank partial
sums are computed with the numpy oracle here (no GPU in this container); on the
GPU box the same records come from the CUDA loss kernel (tests/test_gpu_ops.py).
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muygpys_b200 import _lib as L
from muygpys_b200 import distributed as D
from oracle import numpy_oracle as O


def test_chunk_rule_matches_reference():
    for count in (0, 1, 7, 10, 1000, 100_003):
        for size in (1, 2, 3, 4, 8):
            assert D.get_chunk_sizes(count, size) == O.chunk_sizes(count, size)
    assert D.get_chunk_sizes(10, 4) == [2, 2, 3, 3]  # larger chunks on the LAST ranks
    ranges = [D.local_range(10, rank=r, size=4) for r in range(4)]
    assert ranges == [(0, 2), (2, 4), (4, 7), (7, 10)]


def _record(pred, targ, var, yky):
    rec = np.zeros(L.MGP_PARTIALS)
    rec[L.P_SQERR] = np.sum((pred - targ) ** 2)
    rec[L.P_COUNT] = pred.size
    rec[L.P_ROWS] = pred.shape[0]
    rec[L.P_YKY] = np.sum(yky)
    rec[L.P_SQERR_V] = np.sum((pred - targ) ** 2 / var)
    rec[L.P_LOGV] = np.sum(np.log(var))
    return rec


def _worker(rank, size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        rng = np.random.default_rng(5)
        b, k = 101, 7
        pred, targ = rng.normal(size=b), rng.normal(size=b)
        var, yky = rng.uniform(0.1, 2.0, size=b), rng.uniform(1.0, 9.0, size=b)
        lo, hi = D.local_range(b)
        assert (lo, hi) == (sum(O.chunk_sizes(b, size)[:rank]),
                            sum(O.chunk_sizes(b, size)[: rank + 1]))
        shard = D.shard_rows(torch.as_tensor(pred), torch.as_tensor(targ))
        assert shard[0].shape[0] == hi - lo
        rec = torch.as_tensor(_record(pred[lo:hi], targ[lo:hi], var[lo:hi], yky[lo:hi]))
        rec = D.allreduce_partials(rec).numpy()
        # mse: two reference allreduces (S/_src/optimize/loss/mpi.py:21-25) in one record
        mse = rec[L.P_SQERR] / rec[L.P_COUNT]
        # analytic scale: S/_src/optimize/scale/mpi.py:19-37
        sigma2 = rec[L.P_YKY] / (rec[L.P_ROWS] * k)
        # lool from linear partials (one all-reduce instead of scale-then-loss)
        lool = rec[L.P_SQERR_V] / sigma2 + rec[L.P_LOGV] + rec[L.P_ROWS] * np.log(sigma2)
        gathered = D.allgather_rows(torch.as_tensor(pred[lo:hi]), b).numpy()
        if rank == 0:
            out.put((mse, sigma2, lool, gathered, pred, targ, var, yky, k))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_two_rank_reduction_equals_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    mse, sigma2, lool, gathered, pred, targ, var, yky, k = out.get()
    np.testing.assert_allclose(mse, O.mse(pred, targ), rtol=1e-13)
    np.testing.assert_allclose(sigma2, yky.sum() / (len(pred) * k), rtol=1e-13)
    np.testing.assert_allclose(lool, O.lool(pred, targ, var, sigma2), rtol=1e-12)
    np.testing.assert_array_equal(gathered, pred)


def test_single_process_paths_are_noops():
    rec = torch.arange(8, dtype=torch.float64)
    assert D.allreduce_partials(rec.clone()).equal(rec)
    assert D.rank_and_size() == (0, 1)
    assert D.local_range(17) == (0, 17)
    x = torch.arange(10)
    assert D.allgather_rows(x, 10).equal(x)
