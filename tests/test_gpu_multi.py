"""GPU, >= 2 devices: the data-parallel objective.  Two processes, one
GPU each, NCCL rendezvous on 127.0.0.1: the batch rows are split with the reference's chunk rule,
the per-rank partial records are summed over NVLink peer memory inside the objective kernel
(`mgp_fused_loo_peers`) or by `mgp_peer_sum8`, and every rank must return the single-GPU
objective of the whole batch."""

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

needs_two = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, no_peer, out_dir):
    import torch.distributed as dist

    if no_peer:
        os.environ["MGP_NO_PEER"] = "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world, device_id=torch.device("cuda", rank))
    from muygpys_b200 import _lib as L
    from muygpys_b200 import ops
    from muygpys_b200.distributed import PartialsReducer, local_range
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import Isotropy, l2
    from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter
    from muygpys_b200.gp.kernels import Matern
    from muygpys_b200.gp.noise import HomoscedasticNoise
    from muygpys_b200.optimize.loss import lool_fn, looph_fn, mse_fn
    from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn

    dev = torch.device("cuda", rank)
    res = {}
    # ---- the bare exchange: many epochs, both parities ---------------------------------
    red = PartialsReducer(dev)
    res["peer_path"] = red.peers is not None
    for e in range(1, 40):
        rec = torch.arange(8, dtype=torch.float64, device=dev) * (rank + 1) + e
        got = red.sum_to_host(rec)
        want = sum(np.arange(8.0) * (r + 1) + e for r in range(world))
        np.testing.assert_array_equal(got, want)
    # ---- the objective: sharded rows == whole batch on one GPU -------------------------
    rng = np.random.default_rng(11)
    n, b, k = 20_000, 3001, 50
    x = torch.as_tensor(rng.uniform(size=(n, 2))).to(dev)
    y = torch.as_tensor(np.sin(5 * rng.uniform(size=n)) + 0.1 * rng.normal(size=n)).to(dev)
    bi = torch.as_tensor(np.sort(rng.choice(n, b, replace=False))).to(dev)
    nn, _ = ops.knn(x, x[bi], k + 1)
    nn = nn[:, 1:].contiguous()
    model = MuyGPS(kernel=Matern(smoothness=Parameter(1.5),
                                 deformation=Isotropy(l2, Parameter(0.1, (0.01, 1.0)))),
                   noise=HomoscedasticNoise(1e-3, (1e-5, 1e-1)), scale=AnalyticScale())
    lo, hi = local_range(b)
    for name, lf in (("mse", mse_fn), ("lool", lool_fn), ("looph", looph_fn)):
        whole = make_fused_loo_crossval_fn(model, lf, bi, nn, x, y)
        shard = make_fused_loo_crossval_fn(model, lf, bi[lo:hi], nn[lo:hi], x, y,
                                           distributed=True)
        vals = []
        for theta in ({"length_scale": 0.07}, {"length_scale": 0.2},
                      {"length_scale": 0.1, "noise": 3e-3}):
            for _ in range(3):
                vals.append((whole(**theta), shard(**theta)))
        res[name] = vals
    # k = 100 (tile kernel + loss kernels + mgp_peer_sum8 / all-reduce)
    nn2, _ = ops.knn(x, x[bi], 101)
    nn2 = nn2[:, 1:].contiguous()
    whole = make_fused_loo_crossval_fn(model, lool_fn, bi, nn2, x, y)
    shard = make_fused_loo_crossval_fn(model, lool_fn, bi[lo:hi], nn2[lo:hi], x, y,
                                       distributed=True)
    res["lool_k100"] = [(whole(length_scale=0.1), shard(length_scale=0.1))]
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@needs_two
@pytest.mark.parametrize("no_peer", [False, True], ids=["peer-memory", "nccl-allreduce"])
def test_sharded_objective_equals_single_gpu(tmp_path, no_peer):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), no_peer, str(tmp_path)), nprocs=world, join=True)
    results = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    assert results[0]["peer_path"] == (not no_peer)
    for name in ("mse", "lool", "looph", "lool_k100"):
        for r in range(world):
            for whole, shard in results[r][name]:
                assert abs(shard - whole) <= 1e-11 * abs(whole), (name, r, whole, shard)
        # every rank returns the same bits (fixed summation order on every GPU)
        assert [s for _, s in results[0][name]] == [s for _, s in results[1][name]]
