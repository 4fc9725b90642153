"""GPU, BASELINE.json sizes: size-independent properties of the CUDA path where the
oracle cannot run in full -- agreement of the two independently written fused kernels,
oracle agreement on sampled rows, linearity in the targets, permutation invariance,
variance bounds, sortedness and exactness of the neighbour lists, run-to-run bitwise
determinism, and shard-and-combine equality of the loss partials."""

import numpy as np
import pytest
import torch

from oracle import numpy_oracle as O

from conftest import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def c2_targets(x, rng):
    return (np.sin(4 * x[:, 0]) + np.cos(3 * x[:, 1]) + 0.3 * np.sin(11 * x[:, 0] * x[:, 1])
            + 0.05 * rng.normal(size=x.shape[0]))


@pytest.fixture(scope="module")
def c2():
    """C2 at full size: 1 M train, 100 k test, d=2, k=50."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(2)
    x = rng.uniform(size=(1_000_000, 2))
    q = rng.uniform(size=(100_000, 2))
    y = c2_targets(x, rng)
    xd, qd, yd = dev(x), dev(q), dev(y)
    nn, d2 = ops.knn(xd, qd, 50)
    return dict(x=x, q=q, y=y, xd=xd, qd=qd, yd=yd, nn=nn, d2=d2, rng=rng)


def test_c2_grid_knn_equals_brute_force(c2):
    """1 M train / 100 k queries: the grid search returns the brute-force result exactly."""
    import time

    from muygpys_b200 import ops

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    grid = ops.KnnGrid(c2["xd"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gi, gd = grid.query(c2["qd"], 50)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"grid build {t1 - t0:.3f} s, 100k queries {t2 - t1:.4f} s")
    assert torch.equal(gi, c2["nn"]) and torch.equal(gd, c2["d2"])


def test_c2_knn_properties(c2):
    nn, d2 = c2["nn"].cpu().numpy(), c2["d2"].cpu().numpy()
    assert nn.shape == (100_000, 50) and nn.dtype == np.int64
    assert np.all(np.diff(d2, axis=1) >= 0.0), "distances must ascend"
    assert np.all((nn >= 0) & (nn < 1_000_000))
    assert np.all(np.sort(nn, axis=1)[:, 1:] != np.sort(nn, axis=1)[:, :-1]), "duplicates"
    # returned squared distances are the direct-difference ones, bit for bit
    rows = c2["rng"].choice(100_000, 3000, replace=False)
    diff = c2["q"][rows, None, :] - c2["x"][nn[rows]]
    direct = diff[..., 0] ** 2
    direct = direct + diff[..., 1] ** 2
    np.testing.assert_array_equal(d2[rows], direct)
    # exactness against a brute-force scan of the whole training set
    sample = rows[:150]
    want_idx, want_d2 = O.knn_exact(c2["x"], c2["q"][sample], 50, chunk=25)
    np.testing.assert_array_equal(nn[sample], want_idx)
    np.testing.assert_array_equal(d2[sample], want_d2)


def test_c2_fused_kernels_agree_and_match_oracle(c2):
    from muygpys_b200 import ops

    kw = dict(kernel_id=O.KERNEL_MATERN_15, metric_id=O.METRIC_L2, length_scale=0.1, noise=1e-3,
              want_yky=True, want_status=True)
    ops.set_fused_variant(3)  # column-direct kernel (the one bench.py times)
    tile = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"], **kw)
    again = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"], **kw)
    ops.set_fused_variant(2)  # plain tile kernel
    plain = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"], **kw)
    ops.set_fused_variant(1)  # generic shared-memory kernel
    gen = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"], **kw)
    ops.set_fused_variant(0)
    assert int(tile["status"].sum()) == 0 and int(gen["status"].sum()) == 0
    for key in ("mean", "var", "yky"):
        assert torch.equal(tile[key], again[key]), f"{key} not bitwise reproducible"
        assert_close(tile[key].cpu().numpy(), gen[key].cpu().numpy(), RTOL, f"column vs generic {key}")
        assert_close(tile[key].cpu().numpy(), plain[key].cpu().numpy(), RTOL, f"column vs tile {key}")
    var = tile["var"].cpu().numpy()
    assert np.all(var > 0.0) and np.all(var <= 1.0 + 1e-12), "unscaled variance must lie in (0,1]"
    rows = np.random.default_rng(7).choice(100_000, 400, replace=False)
    mean, v = O.predict(O.KERNEL_MATERN_15, O.METRIC_L2, 0.1, 1e-3, 1.0, c2["x"], c2["y"],
                        c2["q"], rows, c2["nn"].cpu().numpy()[rows])
    assert_close(tile["mean"].cpu().numpy()[rows, 0], mean, RTOL, "mean vs oracle")
    assert_close(var[rows], v, RTOL, "var vs oracle")


def test_c2_linearity_and_permutation_invariance(c2):
    from muygpys_b200 import ops

    kw = dict(kernel_id=O.KERNEL_MATERN_15, metric_id=O.METRIC_L2, length_scale=0.1, noise=1e-3)
    rng = np.random.default_rng(11)
    y2 = dev(rng.normal(size=1_000_000))
    m1 = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"], **kw)
    m2 = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], y2, **kw)
    m12 = ops.fused_posterior(c2["xd"], c2["qd"], None, c2["nn"], c2["yd"] + 2.0 * y2, **kw)
    assert_close((m1["mean"] + 2.0 * m2["mean"]).cpu().numpy(), m12["mean"].cpu().numpy(), 1e-11,
                 "posterior mean is linear in the targets")
    assert torch.equal(m1["var"], m2["var"]), "variance does not depend on the targets"
    perm = torch.argsort(torch.rand(c2["nn"].shape, device="cuda"), dim=1)
    shuffled = torch.gather(c2["nn"], 1, perm)
    mp = ops.fused_posterior(c2["xd"], c2["qd"], None, shuffled, c2["yd"], **kw)
    assert_close(mp["mean"].cpu().numpy(), m1["mean"].cpu().numpy(), RTOL, "neighbour order")
    assert_close(mp["var"].cpu().numpy(), m1["var"].cpu().numpy(), RTOL, "neighbour order")
    # two response columns at once == one at a time
    both = ops.fused_posterior(c2["xd"], c2["qd"][:20000], None, c2["nn"][:20000],
                               torch.stack((c2["yd"], y2), dim=1), **kw)
    assert_close(both["mean"][:, 0].cpu().numpy(), m1["mean"][:20000, 0].cpu().numpy(), RTOL)
    assert_close(both["mean"][:, 1].cpu().numpy(), m2["mean"][:20000, 0].cpu().numpy(), RTOL)


def test_c2_objective_shards_combine(c2):
    """mse / lool objectives on a 10 k batch equal the combination of 8 shard records
    (the 8-GPU data-parallel evaluation) and the oracle on a subsample."""
    from muygpys_b200 import _lib as L
    from muygpys_b200 import distributed as D
    from muygpys_b200 import ops

    rng = np.random.default_rng(3)
    bi = np.sort(rng.choice(1_000_000, 10_000, replace=False))
    bid = dev(bi)
    bnn, _ = ops.knn(c2["xd"], c2["xd"][bid], 51)
    bnn = bnn[:, 1:].contiguous()
    kw = dict(kernel_id=O.KERNEL_MATERN_15, metric_id=O.METRIC_L2, length_scale=0.07, noise=1e-3,
              want_yky=True)
    out = ops.fused_posterior(c2["xd"], c2["xd"], bid, bnn, c2["yd"], **kw)
    yb = c2["yd"][bid]
    full = ops.loss_partials(L.LOSS_LOOL, out["mean"][:, 0], yb, var=out["var"],
                             yky=out["yky"]).cpu().numpy()
    parts = np.zeros(L.MGP_PARTIALS)
    for r in range(8):
        lo, hi = D.local_range(10_000, rank=r, size=8)
        o = ops.fused_posterior(c2["xd"], c2["xd"], bid[lo:hi], bnn[lo:hi], c2["yd"], **kw)
        parts += ops.loss_partials(L.LOSS_LOOL, o["mean"][:, 0], yb[lo:hi], var=o["var"],
                                   yky=o["yky"]).cpu().numpy()
    assert_close(parts, full, 1e-12, "shard records sum to the full-batch record")
    sub = np.arange(0, 10_000, 25)
    want_mse, _ = O.loo_objective(O.LOSS_MSE, O.KERNEL_MATERN_15, O.METRIC_L2, 0.07, 1e-3,
                                  c2["x"], c2["y"], bi[sub], bnn.cpu().numpy()[sub])
    got = ops.loss_partials(L.LOSS_MSE, out["mean"][sub, 0], yb[sub]).cpu().numpy()
    assert_close(-got[L.P_SQERR] / got[L.P_COUNT], want_mse, RTOL, "mse objective vs oracle")


def test_c4_shape_anisotropic_k100_lool():
    """C4 at full size: anisotropic Matern 5/2, 10 M training points, k=100, lool, 10 k batch
    rows (one GPU's share of the 80 k-row batch); oracle check on a subsample."""
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import Anisotropy, l2
    from muygpys_b200.gp.hyperparameter import AnalyticScale, Parameter, VectorParameter
    from muygpys_b200.gp.kernels import Matern
    from muygpys_b200.gp.noise import HomoscedasticNoise
    from muygpys_b200.neighbors import NN_Wrapper
    from muygpys_b200.optimize.loss import lool_fn
    from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn

    rng = np.random.default_rng(4)
    n, b, k = 10_000_000, 10_000, 100
    x = rng.uniform(size=(n, 2))
    y = np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1] / 5) + 0.05 * rng.normal(size=n)
    xd, yd = dev(x), dev(y)
    bi = np.sort(rng.choice(n, b, replace=False))
    bnn, _ = NN_Wrapper(xd, k).get_batch_nns(dev(bi))
    model = MuyGPS(kernel=Matern(smoothness=Parameter(2.5), deformation=Anisotropy(
        l2, VectorParameter(Parameter(0.1, (0.01, 1)), Parameter(0.5, (0.05, 5))))),
        noise=HomoscedasticNoise(1e-3), scale=AnalyticScale())
    obj = make_fused_loo_crossval_fn(model, lool_fn, dev(bi), bnn, xd, yd)
    got = obj(length_scale0=0.12, length_scale1=0.4)
    assert np.isfinite(got) and got == obj(length_scale0=0.12, length_scale1=0.4)
    sub = np.arange(0, b, 25)
    sub_obj = make_fused_loo_crossval_fn(model, lool_fn, dev(bi[sub]), bnn[dev(sub)], xd, yd)
    want, _ = O.loo_objective(O.LOSS_LOOL, O.KERNEL_MATERN_25, O.METRIC_L2,
                              np.array([0.12, 0.4]), 1e-3, x, y, bi[sub],
                              bnn.cpu().numpy()[sub])
    assert_close(sub_obj(length_scale0=0.12, length_scale1=0.4), want, RTOL, "C4 lool")


def test_c3_shape_784d_classification():
    """C3 shape: 784 features, r=10 one-hot targets, RBF/F2, k=30, cross-entropy."""
    from muygpys_b200 import _lib as L
    from muygpys_b200 import ops

    rng = np.random.default_rng(3)
    n, t, d, r, k = 60_000, 512, 784, 10, 30
    cent = rng.normal(0, 0.5, size=(r, d))
    lab = rng.integers(0, r, size=n)
    x = cent[lab] + rng.normal(size=(n, d))
    qlab = rng.integers(0, r, size=t)
    q = cent[qlab] + rng.normal(size=(t, d))
    y = -0.1 * np.ones((n, r))
    y[np.arange(n), lab] = 0.9
    xd, qd, yd = dev(x), dev(q), dev(y)
    nn, d2 = ops.knn(xd, qd, k)
    want_idx, want_d2 = O.knn_exact(x, q[:24], k, chunk=4)
    np.testing.assert_array_equal(nn[:24].cpu().numpy(), want_idx)
    assert_close(d2[:24].cpu().numpy(), want_d2, 1e-12)
    out = ops.fused_posterior(xd, qd, None, nn, yd, kernel_id=O.KERNEL_RBF, metric_id=O.METRIC_F2,
                              length_scale=28.0, noise=1e-3)
    rows = np.arange(0, t, 8)
    mean, var = O.predict(O.KERNEL_RBF, O.METRIC_F2, 28.0, 1e-3, 1.0, x, y, q, rows,
                          nn.cpu().numpy()[rows])
    assert_close(out["mean"].cpu().numpy()[rows], mean, RTOL, "C3 mean")
    assert_close(out["var"].cpu().numpy()[rows], var, RTOL, "C3 var")
    onehot = -0.1 * np.ones((t, r))
    onehot[np.arange(t), qlab] = 0.9
    ce = ops.loss_partials(L.LOSS_CROSS_ENTROPY, out["mean"], dev(onehot)).cpu().numpy()[L.P_AUX]
    assert_close(ce, O.cross_entropy(out["mean"].cpu().numpy(), onehot), 1e-12, "C3 CE")
    acc = float((out["mean"].argmax(dim=1).cpu().numpy() == qlab).mean())
    assert acc > 0.9, acc


def test_c5_shape_10m_train_mean_var_and_fast_mean():
    """C5 shape (Matern 1/2, k=50) on 10 M training points: mean+variance vs the oracle on a
    sample, and the fast posterior mean pipeline vs its staged definition."""
    from muygpys_b200 import ops
    from muygpys_b200.gp.tensors import fast_nn_update

    rng = np.random.default_rng(5)
    n, t, k = 10_000_000, 50_000, 50
    x = rng.uniform(size=(n, 2))
    y = c2_targets(x, rng)
    q = rng.uniform(size=(t, 2))
    xd, yd, qd = dev(x), dev(y), dev(q)
    nn, _ = ops.knn(xd, qd, k)
    kw = dict(kernel_id=O.KERNEL_MATERN_05, metric_id=O.METRIC_L2, length_scale=0.1, noise=1e-3)
    out = ops.fused_posterior(xd, qd, None, nn, yd, **kw)
    rows = np.arange(0, t, 250)
    mean, var = O.predict(O.KERNEL_MATERN_05, O.METRIC_L2, 0.1, 1e-3, 1.0, x, y, q, rows,
                          nn.cpu().numpy()[rows])
    assert_close(out["mean"].cpu().numpy()[rows, 0], mean, RTOL, "C5 mean")
    assert_close(out["var"].cpu().numpy()[rows], var, RTOL, "C5 var")
    # fast mean: coefficients are only needed for the training points that are some test
    # point's nearest neighbour
    closest = torch.unique(nn[:, 0])
    cnn, _ = ops.knn(xd, xd[closest], k)           # includes the point itself first
    assert torch.equal(cnn[:, 0], closest)
    nn_fast = cnn                                   # == fast_nn_update of the k-1 others
    coeffs = ops.fused_posterior(xd, xd, closest, nn_fast, yd, want_mean=False, want_var=False,
                                 want_coeffs=True, **kw)["coeffs"]
    slot = torch.searchsorted(closest, nn[:, 0])
    fast = ops.fast_mean(xd, qd, None, nn_fast[slot], slot, coeffs, kernel_id=O.KERNEL_MATERN_05,
                         metric_id=O.METRIC_L2, length_scale=0.1)
    srows = rows[:40]
    nf = nn_fast[slot].cpu().numpy()[srows]
    Kin, Kcross = O.kernel_tensors(O.KERNEL_MATERN_05, O.METRIC_L2, 0.1, x, q, srows, nf)
    want = np.einsum("bk,bk->b", Kcross, np.linalg.solve(
        O.homoscedastic_perturb(Kin, 1e-3), y[nf][..., None])[..., 0])
    assert_close(fast.cpu().numpy()[srows, 0], want, 1e-9, "C5 fast mean")
    assert fast_nn_update(cnn[:, 1:]).shape == cnn[:, 1:].shape


def test_c5_full_size_100m_train_10m_test():
    """C5 at FULL size: 100 M training points (the grid index's largest exercised size), 10 M
    test points, Matern 1/2, k = 50: exact KNN, posterior mean + variance, and the fast
    posterior mean; spot checks against the numpy oracle with brute-force neighbours on the
    host.  Data are generated on the device (2.4 GB of host random numbers per run would
    dominate the test) and only the rows the oracle needs are copied back."""
    from muygpys_b200 import ops
    from muygpys_b200.neighbors import NN_Wrapper

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~60 GB of free device memory")
    n, t, k = 100_000_000, 10_000_000, 50
    gen = torch.Generator(device="cuda").manual_seed(5)
    xd = torch.rand((n, 2), generator=gen, device="cuda", dtype=torch.float64)
    yd = (torch.sin(4 * xd[:, 0]) + torch.cos(3 * xd[:, 1]) + 0.3 * torch.sin(11 * xd[:, 0] * xd[:, 1])
          + 0.05 * torch.randn(n, generator=gen, device="cuda", dtype=torch.float64))
    qd = torch.rand((t, 2), generator=gen, device="cuda", dtype=torch.float64)
    nbrs = NN_Wrapper(xd, k)
    assert nbrs._grid is not None and nbrs._grid.n == n
    nn, d2 = nbrs._query(qd, k)
    assert nn.shape == (t, k) and int(nn.min()) >= 0 and int(nn.max()) < n
    assert bool((d2[:, 1:] >= d2[:, :-1]).all()), "neighbours sorted by distance"
    # brute-force neighbours of a few queries on the host (100 M distances each)
    x = xd.cpu().numpy()
    rows = np.array([0, 1, 4_999_999, 9_999_999, 123_456, 7_654_321])
    q_rows = qd[torch.as_tensor(rows).cuda()].cpu().numpy()
    for i, r in enumerate(rows):
        diff = x - q_rows[i]
        dist = diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]
        order = np.argpartition(dist, k)[:k]
        order = order[np.lexsort((order, dist[order]))]
        np.testing.assert_array_equal(nn[r].cpu().numpy(), order)
    kw = dict(kernel_id=O.KERNEL_MATERN_05, metric_id=O.METRIC_L2, length_scale=0.1, noise=1e-3)
    out = ops.fused_posterior(xd, qd, None, nn, yd, want_status=True, **kw)
    assert int(out["status"].sum()) == 0 and bool(torch.isfinite(out["mean"]).all())
    assert bool((out["var"] > 0).all()) and bool((out["var"] < 1.0).all())
    # oracle on the checked rows: gather their neighbourhoods into a small problem
    sel = nn[torch.as_tensor(rows).cuda()].cpu().numpy()
    uniq, inv = np.unique(sel, return_inverse=True)
    ud = torch.as_tensor(uniq).cuda()
    mean, var = O.predict(O.KERNEL_MATERN_05, O.METRIC_L2, 0.1, 1e-3, 1.0, xd[ud].cpu().numpy(),
                          yd[ud].cpu().numpy(), q_rows, np.arange(len(rows)),
                          inv.reshape(sel.shape))
    assert_close(out["mean"].cpu().numpy()[rows, 0], mean, RTOL, "C5 full-size mean")
    assert_close(out["var"].cpu().numpy()[rows], var, RTOL, "C5 full-size variance")
    # fast posterior mean: coefficients of the training points that are some test point's
    # nearest neighbour, then a k-dot per test point; equals the definition on the checked rows
    closest = torch.unique(nn[:, 0])
    cnn, _ = nbrs._query(xd[closest], k)
    assert torch.equal(cnn[:, 0], closest)
    coeffs = ops.fused_posterior(xd, xd, closest, cnn, yd, want_mean=False, want_var=False,
                                 want_coeffs=True, **kw)["coeffs"]
    slot = torch.searchsorted(closest, nn[:, 0].contiguous())
    nn_fast = cnn[slot]
    fast = ops.fast_mean(xd, qd, None, nn_fast, slot, coeffs, kernel_id=O.KERNEL_MATERN_05,
                         metric_id=O.METRIC_L2, length_scale=0.1)
    nf = nn_fast[torch.as_tensor(rows).cuda()].cpu().numpy()
    uniq, inv = np.unique(nf, return_inverse=True)
    ud = torch.as_tensor(uniq).cuda()
    xs, ys = xd[ud].cpu().numpy(), yd[ud].cpu().numpy()
    nfl = inv.reshape(nf.shape)
    Kin, Kcross = O.kernel_tensors(O.KERNEL_MATERN_05, O.METRIC_L2, 0.1, xs, q_rows,
                                   np.arange(len(rows)), nfl)
    want = np.einsum("bk,bk->b", Kcross, np.linalg.solve(
        O.homoscedastic_perturb(Kin, 1e-3), ys[nfl][..., None])[..., 0])
    assert_close(fast.cpu().numpy()[rows, 0], want, 1e-9, "C5 full-size fast mean")
