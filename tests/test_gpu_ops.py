"""GPU parity: every C-ABI entry point against the numpy oracle (and through it
the reference's recorded outputs) on the seeded cases of oracle/cases.py.

Tolerance (north_star): |a-b| <= 1e-10 * max(|b|, ||b||_inf) in fp64; index
arrays bit-exact."""

import numpy as np
import pytest
import torch

from oracle import numpy_oracle as O
from oracle.cases import CASES, make_data

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def _ls(case, f=1.0):
    return (np.asarray(case.length_scale) * f) if case.anisotropic else case.length_scale * f


def _2d(x):
    return x[:, None] if x.ndim == 1 else x


@pytest.fixture(scope="module", params=["auto", "tile", "generic"])
def ops(request):
    """Every test runs three times: through the automatically chosen fused kernel (the
    column-direct kernel where the shape allows, else the tile kernel), through the tile kernel
    and through the generic shared-memory kernel -- independently written implementations of
    the same contract."""
    from muygpys_b200 import ops as _ops

    _ops.set_fused_variant({"auto": 0, "tile": 2, "generic": 1}[request.param])
    yield _ops
    _ops.set_fused_variant(0)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_fused_predict_matches_reference(ops, case):
    g = load_golden(case.name)
    data = make_data(case)
    nn = g["test_nn_idx"]
    y = data["train_y"]
    noise = dev(data["hetero_train_noise"][nn]) if case.hetero else case.noise
    out = ops.fused_posterior(dev(data["train_x"]), dev(data["test_x"]), None, dev(nn), dev(y),
                              kernel_id=case.kernel_id, metric_id=case.metric_id,
                              length_scale=_ls(case), noise=noise, scale=float(g["scale_val"]),
                              want_status=True)
    mean = out["mean"].cpu().numpy()
    if case.r == 1:
        mean = mean[:, 0]
    assert int(out["status"].sum()) == 0
    assert_close(mean, g["mean"], RTOL, "mean vs reference")
    assert_close(out["var"].cpu().numpy(), g["var"], RTOL, "var vs reference")


@pytest.mark.parametrize("case", [c for c in CASES if c.batch and c.r == 1],
                         ids=lambda c: c.name)
def test_fused_train_batch_yky_and_query_idx(ops, case):
    g = load_golden(case.name)
    data = make_data(case)
    x = _2d(data["train_x"])
    y = data["train_y"][:, 0]
    bi, bnn = data["batch_idx"], g["batch_nn_idx"]
    out = ops.fused_posterior(dev(x), dev(x), dev(bi), dev(bnn), dev(y),
                              kernel_id=case.kernel_id, metric_id=case.metric_id,
                              length_scale=_ls(case), noise=case.noise, want_yky=True)
    Kin, Kcross = O.kernel_tensors(case.kernel_id, case.metric_id, _ls(case), x, x, bi, bnn)
    pK = O.homoscedastic_perturb(Kin, case.noise)
    assert_close(out["mean"].cpu().numpy()[:, 0], O.posterior_mean(pK, Kcross, y[bnn]), RTOL)
    assert_close(out["var"].cpu().numpy(), O.diagonal_variance(pK, Kcross), RTOL)
    yky = out["yky"].cpu().numpy()
    want = np.einsum("bi,bi->b", y[bnn], np.linalg.solve(pK, y[bnn][..., None])[..., 0])
    assert_close(yky, want, RTOL, "yky")
    assert_close(yky.sum() / (len(bi) * case.k), g["analytic_scale"], RTOL, "analytic scale")


@pytest.mark.parametrize("case", [c for c in CASES if c.fast], ids=lambda c: c.name)
def test_fused_coeffs_and_fast_mean(ops, case):
    g = load_golden(case.name)
    data = make_data(case)
    x, tx = _2d(data["train_x"]), _2d(data["test_x"])
    y = data["train_y"]
    tr_nn, _ = O.knn_exact(x, x, case.k)
    fast_nn = O.fast_nn_update(tr_nn)
    closest = g["test_nn_idx"][:, 0]
    # precompute for every training point, like the reference workflow
    out = ops.fused_posterior(dev(x), dev(x), None, dev(fast_nn), dev(y),
                              kernel_id=case.kernel_id, metric_id=case.metric_id,
                              length_scale=_ls(case), noise=case.noise, want_mean=False,
                              want_var=False, want_coeffs=True)
    coeffs = out["coeffs"]
    got = coeffs.cpu().numpy()
    if case.r == 1:
        got = got[:, :, 0]
    assert_close(got[:64], g["fast_coeffs_head"], 1e-9, "coeffs head")
    assert_close(got[closest], g["fast_coeffs_closest"], 1e-9, "coeffs[closest]")
    fm = ops.fast_mean(dev(x), dev(tx), None, dev(fast_nn[closest]), dev(closest), coeffs,
                       kernel_id=case.kernel_id, metric_id=case.metric_id,
                       length_scale=_ls(case)).cpu().numpy()
    if case.r == 1:
        fm = fm[:, 0]
    assert_close(fm, g["fast_mean"], RTOL, "fast mean")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_knn_bit_exact(ops, case):
    g = load_golden(case.name)
    data = make_data(case)
    idx, d2 = ops.knn(dev(data["train_x"]), dev(data["test_x"]), case.k)
    assert idx.dtype == torch.int64
    np.testing.assert_array_equal(idx.cpu().numpy(), g["test_nn_idx"])
    assert_close(d2.cpu().numpy(), g["test_nn_d2"], 1e-12, "dist2")
    if case.batch:
        x = dev(_2d(data["train_x"]))
        bi = dev(data["batch_idx"])
        # reference semantics: query k+1, drop column 0 (S/neighbors.py:207-211)
        idx1, d21 = ops.knn(x, x[bi], case.k + 1)
        np.testing.assert_array_equal(idx1[:, 1:].cpu().numpy(), g["batch_nn_idx"])
        # explicit self exclusion gives the same rows when points are distinct
        idx2, d22 = ops.knn(x, x[bi], case.k, self_idx=bi)
        np.testing.assert_array_equal(idx2.cpu().numpy(), g["batch_nn_idx"])
        assert_close(d22.cpu().numpy(), g["batch_nn_d2"], 1e-12)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_staged_ops(ops, case):
    g = load_golden(case.name)
    data = make_data(case)
    rows = g["stage_Kin"].shape[0]
    nn = g["test_nn_idx"][:rows]
    x, tx = dev(data["train_x"]), dev(data["test_x"])
    ar = dev(np.arange(rows))
    cd = ops.crosswise_diffs(tx, x, ar, dev(nn))
    pd = ops.pairwise_diffs(x, dev(nn))
    np.testing.assert_array_equal(
        cd.cpu().numpy(), O.crosswise_tensor(data["test_x"], data["train_x"], np.arange(rows), nn))
    np.testing.assert_array_equal(pd.cpu().numpy(), O.pairwise_tensor(data["train_x"], nn))
    if case.anisotropic:
        xc = ops.metric_reduce(case.metric_id, cd, length_scale=_ls(case))
        xp = ops.metric_reduce(case.metric_id, pd, length_scale=_ls(case))
        pre = 1.0
    else:
        xc = ops.crosswise_dists(case.metric_id, tx, x, ar, dev(nn))
        xp = ops.pairwise_dists(case.metric_id, x, dev(nn))
        assert_close(xc.cpu().numpy(), g["stage_crosswise"], 1e-14)
        assert_close(xp.cpu().numpy(), g["stage_pairwise"], 1e-14)
        assert_close(ops.metric_reduce(case.metric_id, cd).cpu().numpy(), g["stage_crosswise"],
                     1e-14)
        ls = case.length_scale
        pre = 1.0 / ls if case.metric_id == O.METRIC_L2 else 1.0 / ls**2
    Kin = ops.kernel_apply(case.kernel_id, xp, pre)
    Kcross = ops.kernel_apply(case.kernel_id, xc, pre)
    assert_close(Kin.cpu().numpy(), g["stage_Kin"], 1e-13, "Kin")
    assert_close(Kcross.cpu().numpy(), g["stage_Kcross"], 1e-13, "Kcross")
    # perturb + solve on the reference's own tensors
    Kin_ref, Kcross_ref, Y = g["stage_Kin"], g["stage_Kcross"], g["stage_nn_targets"]
    if case.hetero:
        nz = data["hetero_train_noise"][nn]
        pK = ops.perturb(dev(Kin_ref), dev(nz))
        want_pK = O.heteroscedastic_perturb(Kin_ref, nz)
    else:
        pK = ops.perturb(dev(Kin_ref), case.noise)
        want_pK = O.homoscedastic_perturb(Kin_ref, case.noise)
    np.testing.assert_array_equal(pK.cpu().numpy(), want_pK)
    out = ops.solve(pK, dev(Kcross_ref), dev(Y), 1.0, want_mean=True, want_var=True,
                    want_yky=True, want_coeffs=True)
    mean = out["mean"].cpu().numpy().reshape(Y.shape[:1] + Y.shape[2:])
    assert_close(mean, O.posterior_mean(want_pK, Kcross_ref, Y), RTOL, "staged mean")
    assert_close(out["var"].cpu().numpy(), O.diagonal_variance(want_pK, Kcross_ref), RTOL)
    assert_close(out["yky"].cpu().numpy().sum(), O.analytic_scale_unnormalized(want_pK, Y), RTOL)
    coeffs = O.fast_precompute(want_pK, Y)
    assert_close(out["coeffs"].cpu().numpy().reshape(coeffs.shape), coeffs, 1e-9, "coeffs")
    fm = ops.rowdot(dev(Kcross_ref), out["coeffs"]).cpu().numpy()
    assert_close(fm.reshape(mean.shape), O.fast_posterior_mean(Kcross_ref, coeffs), 1e-9)


def test_losses(ops):
    g = load_golden("losses")
    from muygpys_b200 import _lib as L

    for r in (1, 2, 10):
        p, t, v = dev(g[f"pred_r{r}"]), dev(g[f"targ_r{r}"]), dev(g[f"var_r{r}"])
        P = ops.loss_partials(L.LOSS_MSE, p, t).cpu().numpy()
        assert_close(P[L.P_SQERR] / P[L.P_COUNT], g[f"mse_r{r}"], 1e-13)
        assert P[L.P_ROWS] == p.shape[0]
        P = ops.loss_partials(L.LOSS_PSEUDO_HUBER, p, t, boundary_scale=1.5).cpu().numpy()
        assert_close(P[L.P_AUX], g[f"phuber_r{r}"], 1e-13)
        P = ops.loss_partials(L.LOSS_PSEUDO_HUBER, p, t, boundary_scale=2.5).cpu().numpy()
        assert_close(P[L.P_AUX], g[f"phuber25_r{r}"], 1e-13)
        if r == 1:
            s = dev(np.array([1.3]))
            P = ops.loss_partials(L.LOSS_LOOL, p, t, var=v, scale_dev=s).cpu().numpy()
            assert_close(P[L.P_AUX], g["lool_r1"], 1e-13)
            # single-allreduce form: sum e^2/v / s + sum log v + b log s
            alt = P[L.P_SQERR_V] / 1.3 + P[L.P_LOGV] + P[L.P_ROWS] * np.log(1.3)
            assert_close(alt, g["lool_r1"], 1e-13)
            P = ops.loss_partials(L.LOSS_LOOPH, p, t, var=v, scale_dev=s,
                                  boundary_scale=3.0).cpu().numpy()
            assert_close(P[L.P_AUX], g["looph_r1"], 1e-13)
            P = ops.loss_partials(L.LOSS_LOOPH, p, t, var=v, scale_dev=dev(np.array([0.7])),
                                  boundary_scale=2.0).cpu().numpy()
            assert_close(P[L.P_AUX], g["looph2_r1"], 1e-13)
        else:
            oh = dev(g[f"onehot_r{r}"])
            P = ops.loss_partials(L.LOSS_CROSS_ENTROPY, p, oh).cpu().numpy()
            assert_close(P[L.P_AUX], g[f"ce_r{r}"], 1e-13)
            P = ops.loss_partials(L.LOSS_CROSS_ENTROPY, p * 60.0, oh).cpu().numpy()
            assert_close(P[L.P_AUX], g[f"ce_big_r{r}"], 1e-13)


def test_error_behaviour(ops):
    x = torch.rand(50, 2, dtype=torch.float64, device="cuda")
    y = torch.rand(50, dtype=torch.float64, device="cuda")
    nn = torch.randint(0, 50, (4, 5), device="cuda")
    with pytest.raises(ValueError):  # anisotropic length-scale count mismatch
        ops.fused_posterior(x, x, None, nn, y, kernel_id=2, metric_id=0,
                            length_scale=[0.1, 0.2, 0.3])
    with pytest.raises(ValueError):
        ops.fused_posterior(x, x, None, nn, y, kernel_id=9, metric_id=0, length_scale=0.1)
    with pytest.raises(Exception):
        ops.fused_posterior(x.cpu(), x, None, nn, y, kernel_id=2, metric_id=0, length_scale=0.1)
    with pytest.raises(ValueError):
        ops.perturb(torch.rand(3, 4, dtype=torch.float64, device="cuda"), 1e-3)
    # empty batch is a no-op
    out = ops.fused_posterior(x, x, None, nn[:0], y, kernel_id=2, metric_id=0, length_scale=0.1)
    assert out["mean"].shape == (0, 1) and out["var"].shape == (0,)
    # non-SPD neighbourhood (duplicate rows, zero nugget) is flagged, not fatal
    dup = torch.zeros((1, 5), dtype=torch.int64, device="cuda")
    out = ops.fused_posterior(x, x, None, dup, y, kernel_id=2, metric_id=0, length_scale=0.1,
                              noise=0.0, want_status=True)
    assert int(out["status"][0]) == 1 and torch.isnan(out["var"][0])


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("layout", ["uniform", "clustered", "lattice"])
def test_grid_knn_equals_brute_force_bit_for_bit(d, layout):
    """The uniform-grid search (d <= 3) must return exactly what the brute-force kernel
    returns: same indices (ties resolved to the lower train row) and the same bits of
    squared distance -- on uniform data, strongly clustered data and an integer lattice
    where almost every distance is tied."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(100 * d + len(layout))
    n, q, k = 20_000, 3_000, 37
    if layout == "uniform":
        x = rng.uniform(size=(n, d))
        qs = rng.uniform(-0.1, 1.1, size=(q, d))  # some queries outside the bounding box
    elif layout == "clustered":
        centers = rng.uniform(size=(5, d))
        x = centers[rng.integers(0, 5, n)] + 1e-3 * rng.normal(size=(n, d))
        x[:50] = rng.uniform(-3, 3, size=(50, d))  # a few far outliers stretch the grid
        qs = centers[rng.integers(0, 5, q)] + 2e-3 * rng.normal(size=(q, d))
    else:
        side = int(round(n ** (1.0 / d))) + 1
        x = rng.integers(0, side, size=(n, d)).astype(np.float64)  # duplicates and ties
        qs = rng.integers(0, side, size=(q, d)).astype(np.float64)
    xd, qd = dev(x), dev(qs)
    grid = ops.KnnGrid(xd)
    gi, gd = grid.query(qd, k)
    bi, bd = ops.knn(xd, qd, k)
    assert torch.equal(gi, bi) and torch.equal(gd, bd)
    # self-exclusion and the reference's k+1-and-drop both agree with brute force too
    rows = dev(rng.choice(n, 500, replace=False))
    gi2, gd2 = grid.query(xd[rows], k, self_idx=rows)
    bi2, bd2 = ops.knn(xd, xd[rows], k, self_idx=rows)
    assert torch.equal(gi2, bi2) and torch.equal(gd2, bd2)
    gi3, _ = grid.query(xd[rows], k + 1)
    bi3, _ = ops.knn(xd, xd[rows], k + 1)
    assert torch.equal(gi3, bi3)
    # a single query and k = 1
    g1, _ = grid.query(qd[:1], 1)
    b1, _ = ops.knn(xd, qd[:1], 1)
    assert torch.equal(g1, b1)


@pytest.mark.parametrize("k", [5, 50, 64, 65, 100, 128, 129])
def test_grid_knn_kernel_selection_and_overflow(k):
    """The warp-per-query grid kernel (k <= 128: candidate list of 256 entries up to k = 64, 512
    above, exact radix selection, serial fallback when the list overflows or one distance is
    shared by more than 32 rows) and the thread-per-query kernel (k > 128) against brute force,
    bit for bit: uniform data with one cell holding thousands of points (overflow), a block of
    exact duplicates (ties at distance 0) and queries on both."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(900 + k)
    n = 30_000
    x = rng.uniform(size=(n, 2))
    x[:4000] = 0.5 + 1e-4 * rng.normal(size=(4000, 2))   # one very dense cell
    x[4000:4100] = np.array([0.25, 0.75])                # 100 copies of one point
    q = np.concatenate([rng.uniform(size=(1500, 2)), 0.5 + 2e-4 * rng.normal(size=(300, 2)),
                        np.tile([[0.25, 0.75]], (10, 1)), x[:200]])
    xd, qd = dev(x), dev(q)
    grid = ops.KnnGrid(xd)
    gi, gd = grid.query(qd, k)
    bi, bd = ops.knn(xd, qd, k)
    assert torch.equal(gi, bi) and torch.equal(gd, bd)
    rows = dev(np.arange(3900, 4150))
    gi2, gd2 = grid.query(xd[rows], k, self_idx=rows)
    bi2, bd2 = ops.knn(xd, xd[rows], k, self_idx=rows)
    assert torch.equal(gi2, bi2) and torch.equal(gd2, bd2)


def test_fast_coefficients_reproducible_over_many_launches(ops):
    """Regression: stale NaNs in never-assembled shared-memory cells once leaked into the
    coefficient back substitution on some launches.  Interleave other launches (which leave
    different garbage behind) and require bit-identical, finite coefficients every time."""
    case = next(c for c in CASES if c.name == "c5_m05_2d")
    data = make_data(case)
    x, y = dev(data["train_x"]), dev(data["train_y"][:, 0])
    nn, _ = ops.knn(x, x, case.k)
    other = torch.rand(5000, 2, dtype=torch.float64, device="cuda") * 1e3
    kw = dict(kernel_id=case.kernel_id, metric_id=case.metric_id, length_scale=0.1, noise=1e-3)
    first = None
    for it in range(40):
        c = ops.fused_posterior(x, x, None, nn, y, want_mean=False, want_var=False,
                                want_coeffs=True, **kw)["coeffs"]
        assert bool(torch.isfinite(c).all()), f"non-finite coefficients at launch {it}"
        first = c.clone() if first is None else first
        assert torch.equal(c, first), f"coefficients changed at launch {it}"
        # a launch on badly scaled data (zero nugget: non-SPD rows produce NaN/inf internally)
        onn = torch.randint(0, 5000, (512, case.k), device="cuda")
        ops.fused_posterior(other, other, None, onn, other[:, 0].contiguous(), kernel_id=0,
                            metric_id=1, length_scale=1e-3, noise=0.0)


def test_high_d_knn_tiled_matches_warp_kernel_and_oracle():
    """d = 784: the register-tiled sweep (q >= 8, training set split over CTAs and merged) and
    the warp-per-query kernel (q < 8) agree, and the tiled kernel reproduces the oracle's
    direct-difference distances bit for bit (same feature order, no FMA)."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(784)
    n, q, d, k = 9_000, 70, 784, 30
    x = rng.normal(size=(n, d))
    qs = rng.normal(size=(q, d))
    x[17] = x[4000]  # an exact duplicate: tie must resolve to the lower row
    xd, qd = dev(x), dev(qs)
    ti, td = ops.knn(xd, qd, k)
    want_i, want_d = O.knn_exact(x, qs, k, chunk=8)
    np.testing.assert_array_equal(ti.cpu().numpy(), want_i)
    np.testing.assert_array_equal(td.cpu().numpy(), want_d)
    wi = torch.cat([ops.knn(xd, qd[i:i + 5], k)[0] for i in range(0, 20, 5)])
    assert torch.equal(wi, ti[:20])
    rows = dev(np.array([17, 4000, 5, 8999, 123, 77, 4001, 16]))
    si, _ = ops.knn(xd, xd[rows], k, self_idx=rows)
    s1, _ = ops.knn(xd, xd[rows], k + 1)
    assert int(si[0, 0]) == 4000 and int(si[1, 0]) == 17  # the duplicate is the nearest other
    assert torch.equal(s1[2:, 1:], si[2:])


SHAPES = [(1, 1, 2), (2, 1, 1), (3, 2, 2), (7, 1, 3), (8, 3, 2), (12, 5, 2), (30, 10, 3),
          (33, 1, 2), (49, 1, 2), (50, 10, 2), (51, 1, 2), (52, 1, 2), (60, 4, 5), (64, 1, 8),
          (99, 2, 2), (100, 1, 2), (100, 3, 2), (101, 1, 2), (104, 1, 2), (108, 2, 2),
          (120, 1, 2), (124, 1, 2), (125, 1, 2)]


@pytest.mark.parametrize("k,r,d", SHAPES, ids=lambda v: str(v))
def test_shape_sweep_tile_vs_generic_vs_oracle(k, r, d):
    """Every (k, r, d) corner of the tile kernels -- one tile, partial last panel, Schur block
    spanning two tile rows (r = 10), the register / shared-memory factor split at T = 7/8, the
    pipelined specialisation at k in 49..52, k beyond the tile kernels -- against the generic
    kernel and the oracle, for mean, variance, y^T K^-1 y and the fast-mean coefficients."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(1000 * k + 10 * r + d)
    n, b = 600, 37
    x = rng.uniform(size=(n, d))
    y = rng.normal(size=(n, r))
    q = rng.uniform(size=(b, d))
    nn, _ = O.knn_exact(x, q, k)
    kid = int(rng.integers(0, 5))
    metric = O.METRIC_F2 if kid == O.KERNEL_RBF else O.METRIC_L2
    ls = 0.3 if metric == O.METRIC_L2 else 0.5
    kw = dict(kernel_id=kid, metric_id=metric, length_scale=ls, noise=1e-3, scale=1.7,
              want_yky=True, want_coeffs=True, want_status=True)
    outs = {}
    for name, variant in (("auto", 0), ("tile", 2), ("generic", 1)):
        ops.set_fused_variant(variant)
        outs[name] = ops.fused_posterior(dev(x), dev(q), None, dev(nn), dev(y), **kw)
    ops.set_fused_variant(0)
    Kin, Kcross = O.kernel_tensors(kid, metric, ls, x, q, np.arange(b), nn)
    pK = O.homoscedastic_perturb(Kin, 1e-3)
    want = dict(mean=O.posterior_mean(pK, Kcross, y[nn]),
                var=1.7 * O.diagonal_variance(pK, Kcross),
                yky=np.einsum("bkr,bkr->b", y[nn], np.linalg.solve(pK, y[nn])),
                coeffs=np.linalg.solve(pK, y[nn]))
    for name, out in outs.items():
        assert int(out["status"].sum()) == 0
        for key, tol in (("mean", RTOL), ("var", RTOL), ("yky", RTOL), ("coeffs", 1e-9)):
            assert_close(out[key].cpu().numpy(), want[key], tol, f"{name} {key} k={k} r={r} d={d}")


HIGH_D_SHAPES = [(30, 10, 784, False), (5, 1, 9, False), (30, 1, 10, True), (50, 1, 16, False),
                 (63, 2, 33, False), (64, 1, 12, True), (100, 1, 20, False), (31, 3, 11, True),
                 (8, 2, 100, False), (40, 1, 32, True)]


@pytest.mark.parametrize("k,r,d,aniso", HIGH_D_SHAPES, ids=lambda v: str(v))
def test_high_dimensional_gram_assembly(k, r, d, aniso):
    """d > 8: the tile kernel takes its squared distances from query-centred DMMA Gram tiles
    (csrc/gram.cuh) -- even / odd d (vector / scalar row loads), anisotropic scaling, one to four
    row blocks (k = 100 exercises the off-diagonal blocks) -- against the generic kernel and the
    oracle."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(77 * k + 3 * r + d)
    n, b = 500, 41
    x = rng.uniform(size=(n, d)) + 5.0          # far from the origin: centring must matter
    y = rng.normal(size=(n, r))
    q = rng.uniform(size=(b, d)) + 5.0
    nn, _ = O.knn_exact(x, q, k)
    kid = int(rng.integers(0, 5))
    metric = O.METRIC_F2 if kid == O.KERNEL_RBF else O.METRIC_L2
    base = (0.5 if metric == O.METRIC_L2 else 0.7) * np.sqrt(d)
    ls = base * rng.uniform(0.7, 1.4, size=d) if aniso else base
    kw = dict(kernel_id=kid, metric_id=metric, length_scale=ls, noise=1e-3, scale=0.9,
              want_yky=True, want_coeffs=True, want_status=True)
    outs = {}
    for name, variant in (("auto", 0), ("generic", 1)):
        ops.set_fused_variant(variant)
        outs[name] = ops.fused_posterior(dev(x), dev(q), None, dev(nn), dev(y), **kw)
    ops.set_fused_variant(0)
    Kin, Kcross = O.kernel_tensors(kid, metric, ls, x, q, np.arange(b), nn)
    pK = O.homoscedastic_perturb(Kin, 1e-3)
    want = dict(mean=O.posterior_mean(pK, Kcross, y[nn]),
                var=0.9 * O.diagonal_variance(pK, Kcross),
                yky=np.einsum("bkr,bkr->b", y[nn], np.linalg.solve(pK, y[nn])),
                coeffs=np.linalg.solve(pK, y[nn]))
    for name, out in outs.items():
        assert int(out["status"].sum()) == 0
        for key, tol in (("mean", RTOL), ("var", RTOL), ("yky", RTOL), ("coeffs", 1e-9)):
            assert_close(out[key].cpu().numpy(), want[key], tol, f"{name} {key} k={k} r={r} d={d}")


@pytest.mark.parametrize("kid", [O.KERNEL_MATERN_05, O.KERNEL_MATERN_25, O.KERNEL_RBF])
def test_high_dimensional_gram_cancellation_fixup(kid):
    """Gram-identity distances cancel when two neighbours are much closer to each other than to
    the query: exact duplicates, near-duplicates (1e-7 apart) and queries far outside their
    neighbourhood must still match the oracle's direct differences (Matern 1/2 is not smooth at
    0, so a distance error of 1e-14 would show up as 1e-7 in the covariance)."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(4242 + kid)
    n, b, k, d, r = 300, 23, 20, 24, 2
    x = rng.normal(size=(n, d))
    x[1] = x[0]                                   # exact duplicate
    x[3] = x[2] + 1e-7 * rng.normal(size=d)       # near duplicate
    x[5] = x[4] + 1e-3 * rng.normal(size=d)
    y = rng.normal(size=(n, r))
    q = rng.normal(size=(b, d))
    q[:8] = x[[0, 2, 4, 0, 2, 4, 0, 2]] + 0.05 * rng.normal(size=(8, d))  # near the duplicates
    q[8:12] += 40.0                               # far outside the data: everything cancels
    nn, _ = O.knn_exact(x, q, k)
    metric = O.METRIC_F2 if kid == O.KERNEL_RBF else O.METRIC_L2
    ls = 3.0 if metric == O.METRIC_L2 else 4.0
    noise = 1e-2
    ops.set_fused_variant(0)
    out = ops.fused_posterior(dev(x), dev(q), None, dev(nn), dev(y), kernel_id=kid,
                              metric_id=metric, length_scale=ls, noise=noise, want_status=True)
    Kin, Kcross = O.kernel_tensors(kid, metric, ls, x, q, np.arange(b), nn)
    pK = O.homoscedastic_perturb(Kin, noise)
    assert int(out["status"].sum()) == 0
    # rows 8..11 have Kcross ~ 0; compare them on the absolute scale of the others
    assert_close(out["mean"].cpu().numpy(), O.posterior_mean(pK, Kcross, y[nn]), 1e-9, "mean")
    assert_close(out["var"].cpu().numpy(), O.diagonal_variance(pK, Kcross), 1e-9, "var")


@pytest.mark.parametrize("layout,d", [("random", 33), ("offset", 33), ("duplicates", 33),
                                      ("lattice", 33), ("random", 12), ("lattice", 9),
                                      ("duplicates", 16), ("random", 4), ("lattice", 5),
                                      ("duplicates", 8), ("offset", 2)])
def test_high_d_knn_gram_prefilter_is_exact(layout, d):
    """Any d (n >= 2048): the DMMA Gram pre-filter + certified exact re-rank (csrc/knn_gram.cu) must return
    bit for bit what the exact sweep returns -- on random data, on data far from the origin
    (large norms -> large cancellation bound), with every point repeated 12 times (ties straddle
    the candidate boundary: certification fails and the exact sweep re-runs those queries) and
    on an integer lattice where most distances are tied."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(len(layout) + d)
    n, q, k = 4_000, 150, 41
    if layout == "random":
        x = rng.normal(size=(n, d))
        qs = rng.normal(size=(q, d))
    elif layout == "offset":
        x = rng.normal(size=(n, d)) + 1e4
        qs = rng.normal(size=(q, d)) + 1e4
    elif layout == "duplicates":
        base = rng.normal(size=(n // 12 + 1, d))
        x = np.repeat(base, 12, axis=0)[:n]
        qs = base[rng.integers(0, len(base), q)] + 1e-3 * rng.normal(size=(q, d))
    else:
        x = rng.integers(0, 2, size=(n, d)).astype(np.float64)
        qs = rng.integers(0, 2, size=(q, d)).astype(np.float64)
    xd, qd = dev(x), dev(qs)
    for kk in (k, 88, 100):  # 100 > 88: beyond the candidate lists, i.e. the exact sweep
        gi, gd = ops.knn(xd, qd, kk)
        want_i, want_d = O.knn_exact(x, qs, kk, chunk=16)
        np.testing.assert_array_equal(gi.cpu().numpy(), want_i)
        np.testing.assert_array_equal(gd.cpu().numpy(), want_d)
    rows = dev(rng.integers(0, n, 64))
    si, sd = ops.knn(xd, xd[rows], k, self_idx=rows)
    rn = rows.cpu().numpy()
    dist = np.zeros((len(rn), n))
    for f in range(d):  # the oracle's arithmetic: direct differences in feature order
        dist += (x[rn, f:f + 1] - x[None, :, f]) ** 2
    dist[np.arange(len(rn)), rn] = np.inf  # explicit self exclusion
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]
    np.testing.assert_array_equal(si.cpu().numpy(), order)
    np.testing.assert_array_equal(sd.cpu().numpy(), np.take_along_axis(dist, order, axis=1))


@pytest.mark.parametrize("d", [12, 40])
def test_high_dimensional_gram_training_batch_heteroscedastic(d):
    """d > 8 with everything a training batch uses: queries taken from the training set through
    `query_idx` (the query row is then a training row, the neighbourhood excludes it), a
    heteroscedastic nugget tensor and the y^T K^-1 y output -- against the generic kernel and
    the oracle."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(9000 + d)
    n, b, k, r = 2_500, 57, 22, 2
    x = rng.normal(size=(n, d))
    y = rng.normal(size=(n, r))
    bi = rng.choice(n, b, replace=False)
    nn, _ = ops.knn(dev(x), dev(x[bi]), k, self_idx=dev(bi))
    nn_h = nn.cpu().numpy()
    assert not (nn_h == bi[:, None]).any()
    noise = rng.uniform(1e-3, 5e-2, size=(b, k))
    ls = 0.8 * np.sqrt(d)
    kw = dict(kernel_id=O.KERNEL_MATERN_25, metric_id=O.METRIC_L2, length_scale=ls,
              noise=dev(noise), scale=1.0, want_yky=True, want_status=True)
    outs = {}
    for name, variant in (("auto", 0), ("generic", 1)):
        ops.set_fused_variant(variant)
        outs[name] = ops.fused_posterior(dev(x), dev(x), dev(bi), nn, dev(y), **kw)
    ops.set_fused_variant(0)
    Kin, Kcross = O.kernel_tensors(O.KERNEL_MATERN_25, O.METRIC_L2, ls, x, x, bi, nn_h)
    pK = O.heteroscedastic_perturb(Kin, noise)
    want_mean = O.posterior_mean(pK, Kcross, y[nn_h])
    want_var = O.diagonal_variance(pK, Kcross)
    want_yky = np.einsum("bkr,bkr->b", y[nn_h], np.linalg.solve(pK, y[nn_h]))
    for name, out in outs.items():
        assert int(out["status"].sum()) == 0
        assert_close(out["mean"].cpu().numpy(), want_mean, RTOL, f"{name} mean d={d}")
        assert_close(out["var"].cpu().numpy(), want_var, RTOL, f"{name} var d={d}")
        assert_close(out["yky"].cpu().numpy(), want_yky, RTOL, f"{name} yky d={d}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_kernels_follow_their_tensors_to_a_second_device():
    """One process, two GPUs: launches, the exp table (kernel parameters) and the cached device
    attributes must follow the tensors' device, not the current one (r1 advisor finding: a
    per-process constant table left device 1 with zeros -> mean 0, var = scale, no error)."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(3)
    n, b, k = 3000, 500, 50
    xh, yh, qh = rng.uniform(size=(n, 2)), rng.normal(size=n), rng.uniform(size=(b, 2))
    outs = []
    for dev in (0, 1):
        x, y, q = (torch.as_tensor(a).to(f"cuda:{dev}") for a in (xh, yh, qh))
        torch.cuda.set_device(0)  # the CURRENT device stays 0 throughout
        nn, _ = ops.knn(x, q, k)
        assert nn.device.index == dev
        for variant in (3, 2, 1):
            ops.set_fused_variant(variant)
            out = ops.fused_posterior(x, q, None, nn, y, kernel_id=2, metric_id=0,
                                      length_scale=0.1, noise=1e-3)
            assert out["mean"].device.index == dev
            outs.append((dev, variant, out["mean"].cpu().numpy(), out["var"].cpu().numpy()))
    ops.set_fused_variant(0)
    for dev, variant, mean, var in outs[1:]:
        np.testing.assert_allclose(mean, outs[0][2], rtol=0, atol=1e-11)
        np.testing.assert_allclose(var, outs[0][3], rtol=0, atol=1e-11)
        assert np.abs(mean).max() > 1e-3


@pytest.mark.parametrize("col_variant", [3, 4])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_column_kernel_shape_sweep(d, col_variant):
    """The column-direct kernels (forced: an unsupported shape would raise) -- variant 3 = the
    thread-per-tile kernel (factor warp + update warps, csrc/fused_tp.cuh), variant 4 = the
    kernel with lane-parallel column steps (csrc/fused_col.cuh) -- against the numpy oracle for
    every neighbour count around the tile boundaries -- the augmented rows (cross-covariance,
    targets) move through the last two tile rows as k changes -- and every covariance function
    they build, prediction and training-batch (query_idx) forms.  64 rows do not fill the last
    CTA of the thread-per-tile kernel (15 neighbourhoods per CTA): the padding rows are covered."""
    from muygpys_b200 import ops

    rng = np.random.default_rng(40 + d)
    n, b = 1500, 64
    x = rng.uniform(size=(n, d))
    y = rng.normal(size=(n, 1))
    q = rng.uniform(size=(b, d))
    x[7] = x[3]  # an exact duplicate: zero distance inside neighbourhoods
    ks = [7, 8, 9, 14, 15, 16, 22, 23, 30, 31, 38, 39, 46, 47, 50, 54, 55, 62]
    if col_variant == 3:  # the thread-per-tile kernel goes on to 13 tile rows (k = 102)
        ks += [63, 64, 70, 71, 78, 79, 86, 87, 94, 95, 100, 102]
    try:
        for k in ks:
            nn, _ = O.knn_exact(x, q, k)
            kid = int(rng.integers(0, 5))
            metric = O.METRIC_F2 if kid == O.KERNEL_RBF else O.METRIC_L2
            ls = rng.uniform(0.2, 0.5, size=d) if (d > 1 and k % 2) else 0.3
            kw = dict(kernel_id=kid, metric_id=metric, length_scale=ls, noise=1e-3, scale=1.3,
                      want_yky=True, want_status=True)
            ops.set_fused_variant(col_variant)
            got = ops.fused_posterior(dev(x), dev(q), None, dev(nn), dev(y), **kw)
            ops.set_fused_variant(1)
            ref = ops.fused_posterior(dev(x), dev(q), None, dev(nn), dev(y), **kw)
            assert int(got["status"].sum()) == 0
            want_mean, want_var = O.predict(kid, metric, ls, 1e-3, 1.3, x, y[:, 0], q,
                                            np.arange(b), nn)
            assert_close(got["mean"].cpu().numpy()[:, 0], want_mean, RTOL, f"mean k={k} d={d}")
            assert_close(got["var"].cpu().numpy(), want_var, RTOL, f"var k={k} d={d}")
            assert_close(got["yky"].cpu().numpy(), ref["yky"].cpu().numpy(), RTOL, f"yky k={k}")
        # training-batch form: queries are rows of the training set
        k = 50
        bi = np.sort(rng.choice(n, b, replace=False))
        bnn, _ = O.knn_exact(x, x[bi], k + 1)
        bnn = np.ascontiguousarray(bnn[:, 1:])
        ops.set_fused_variant(col_variant)
        got = ops.fused_posterior(dev(x), dev(x), dev(bi), dev(bnn), dev(y), kernel_id=2,
                                  metric_id=0, length_scale=0.3, noise=1e-3)
        want_mean, want_var = O.predict(2, 0, 0.3, 1e-3, 1.0, x, y[:, 0], x, bi, bnn)
        assert_close(got["mean"].cpu().numpy()[:, 0], want_mean, RTOL, "batch mean")
        assert_close(got["var"].cpu().numpy(), want_var, RTOL, "batch var")
        # a non-positive pivot is reported, not hidden: negative nugget on duplicated points
        nn, _ = O.knn_exact(x, x[[3]], 10)
        ops.set_fused_variant(col_variant)
        bad = ops.fused_posterior(dev(x), dev(x[[3]]), None, dev(nn), dev(y), kernel_id=2,
                                  metric_id=0, length_scale=0.3, noise=-1e-3, want_status=True)
        assert int(bad["status"][0]) == 1 and bool(torch.isnan(bad["mean"]).all())
    finally:
        ops.set_fused_variant(0)


@pytest.mark.parametrize("loss_id,k,d", [(1, 50, 2), (2, 50, 2), (4, 30, 1), (2, 23, 3), (0, 62, 2),
                                         (2, 100, 2), (1, 71, 3)])
def test_fused_loo_record_matches_two_pass_path(loss_id, k, d):
    """mgp_fused_loo (K1 with the loss / scale partials in its epilogue, one launch) against
    K1 + mgp_loss_partials on the same batch, slot by slot; repeated launches reuse the
    self-resetting workspace and are bit-reproducible."""
    from muygpys_b200 import _lib as L
    from muygpys_b200 import ops

    rng = np.random.default_rng(loss_id * 100 + k)
    n, b = 5000, 1237
    x = dev(rng.uniform(size=(n, d)))
    y = dev(np.sin(3 * rng.uniform(size=n)) + 0.1 * rng.normal(size=n))
    bi = dev(np.sort(rng.choice(n, b, replace=False)))
    nn, _ = ops.knn(x, x[bi], k + 1)
    nn = nn[:, 1:].contiguous()
    kid, ls, noise = (2, 0.2, 1e-3) if d > 1 else (0, 0.05, 1e-3)
    mid = 1 if kid == 0 else 0
    loo = ops.FusedLoo(x, y, bi, nn, kernel_id=kid, metric_id=mid, loss_id=loss_id,
                       boundary_scale=1.5)
    rec = loo.record(loo.launch(ls, noise))
    out = ops.fused_posterior(x, x, bi, nn, y, kernel_id=kid, metric_id=mid, length_scale=ls,
                              noise=noise, want_yky=True)
    two_pass = ops.loss_partials(loss_id, out["mean"][:, 0].contiguous(), y[bi].contiguous(),
                                 var=out["var"], yky=out["yky"], boundary_scale=1.5)
    want = two_pass.cpu().numpy()
    assert rec[L.P_ROWS] == b and rec[L.P_COUNT] == b and rec[L.P_BAD] == 0
    slots = [L.P_SQERR, L.P_YKY]
    slots += [L.P_SQERR_V, L.P_LOGV] if loss_id == 2 else []
    slots += [L.P_AUX] if loss_id == 4 else []
    for s_ in slots:
        assert abs(rec[s_] - want[s_]) <= 1e-11 * abs(want[s_]), (s_, rec[s_], want[s_])
    for _ in range(3):  # same launch again: identical bits (fixed summation order)
        again = loo.record(loo.launch(ls, noise))
        np.testing.assert_array_equal(again, rec)
    other = loo.record(loo.launch(ls * 1.5, noise * 2))
    assert other[L.P_SQERR] != rec[L.P_SQERR]


def test_grid_knn_with_a_nearly_degenerate_axis_and_bad_indices():
    """r1 advisor findings: (1) a feature that is constant up to rounding must not blow up the
    cell grid (it collapses to one layer) and the search stays bit-identical to brute force;
    (2) out-of-range neighbour indices are a ValueError where the objective is set up (and in
    every fused call under MGP_CHECK_INDICES=1), not an out-of-bounds read."""
    from muygpys_b200 import _lib as L
    from muygpys_b200 import ops

    rng = np.random.default_rng(12)
    n, q, k = 20000, 500, 20
    x = rng.uniform(size=(n, 3))
    x[:, 1] = 0.5 + 1e-13 * rng.normal(size=n)      # degenerate up to rounding
    x[:, 2] *= 1e-9                                  # thin, but genuinely spread
    qs = x[rng.choice(n, q, replace=False)] + 1e-12
    grid = ops.KnnGrid(dev(x))
    ncells = int(np.prod(grid.dims))
    assert ncells <= ops.KnnGrid.MAX_CELLS and max(grid.dims) < 2 ** 31
    idx, d2 = grid.query(dev(qs), k)
    want_idx, want_d2 = O.knn_exact(x, qs, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(d2.cpu().numpy(), want_d2)
    # duplicate-heavy data far from the origin, points sitting on cell edges
    base = 1.0e6 + np.round(rng.uniform(size=(400, 2)) * 64) / 64
    x2 = np.repeat(base, 25, axis=0)
    q2 = base[:200]
    idx2, d22 = ops.KnnGrid(dev(x2)).query(dev(q2), 30)
    want_idx2, want_d22 = O.knn_exact(x2, q2, 30)
    np.testing.assert_array_equal(idx2.cpu().numpy(), want_idx2)
    np.testing.assert_array_equal(d22.cpu().numpy(), want_d22)
    # index validation
    xs, ys = dev(rng.uniform(size=(500, 2))), dev(rng.normal(size=500))
    bi = dev(np.arange(40))
    nn = torch.randint(0, 500, (40, 20), device="cuda")
    nn[3, 4] = 500
    with pytest.raises(ValueError, match="batch_nn_indices"):
        ops.FusedLoo(xs, ys, bi, nn, kernel_id=2, metric_id=0, loss_id=L.LOSS_MSE)
    nn[3, 4] = -1
    with pytest.raises(ValueError, match="batch_nn_indices"):
        ops.FusedLoo(xs, ys, bi, nn, kernel_id=2, metric_id=0, loss_id=L.LOSS_MSE)
    old = ops._CHECK_INDICES
    ops._CHECK_INDICES = True
    try:
        with pytest.raises(ValueError, match="nn_indices"):
            ops.fused_posterior(xs, xs, bi, nn, ys, kernel_id=2, metric_id=0, length_scale=0.2,
                                noise=1e-3)
    finally:
        ops._CHECK_INDICES = old
