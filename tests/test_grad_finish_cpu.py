"""CPU: the host arithmetic that finishes the analytic gradient (`objective.finish_value_and_grad`)
from the sums `mgp_fused_loo_grad` delivers.  The sums are rebuilt here in numpy from the
oracle's per-row posterior mean / variance / y^T K^-1 y and their per-row central differences;
the finished objective and gradient must equal the oracle's objective
(S/optimize/objective.py:20-118) and its central differences -- for mse, lool and looph, fixed
and analytic scale, isotropic and anisotropic length scales, and the reference's nugget quirk
(sigma^2 at the MODEL's nugget, S/gp/hyperparameter/scale.py:206-208)."""

import numpy as np
import pytest

from muygpys_b200 import _lib as L
from muygpys_b200.objective import finish_value_and_grad
from oracle import numpy_oracle as O

K, D, B, N = 12, 2, 40, 400
BS = 2.5  # looph boundary scale


def _data():
    rng = np.random.default_rng(3)
    x = rng.uniform(size=(N, D))
    y = np.sin(4 * x[:, 0]) + np.cos(3 * x[:, 1]) + 0.05 * rng.normal(size=N)
    y[::37] += 0.6  # outliers: Huber weights well below 1
    bi = np.sort(rng.choice(N, B, replace=False))
    bnn, _ = O.knn_batch(x, bi, K)
    return x, y, bi, bnn


def _rows(x, y, bi, bnn, ls, noise):
    """Per-row mean, unscaled variance and y^T K^-1 y (length scales as a vector)."""
    Kin, Kcross = O.kernel_tensors(O.KERNEL_MATERN_15, O.METRIC_L2, np.asarray(ls, float), x, x,
                                   bi, bnn)
    pK = O.homoscedastic_perturb(Kin, noise)
    ynn = y[bnn]
    m = np.asarray(O.posterior_mean(pK, Kcross, ynn)).reshape(-1)
    v = np.asarray(O.diagonal_variance(pK, Kcross, 1.0)).reshape(-1)
    yky = np.einsum("bk,bk->b", ynn, np.linalg.solve(pK, ynn[:, :, None])[:, :, 0])
    return m, v, yky


def _row_derivatives(x, y, bi, bnn, ls, noise):
    """Slots 0..D-1: d/d l_f, slot 3: d/d noise, each a (dm, dv, dyky) triple per row."""
    out = {}
    for slot in list(range(D)) + [3]:
        base = noise if slot == 3 else ls[slot]
        h = 1e-5 * base

        def at(val):
            l2, nz = list(ls), noise
            if slot == 3:
                nz = val
            else:
                l2[slot] = val
            return _rows(x, y, bi, bnn, l2, nz)

        up, dn = at(base + h), at(base - h)
        out[slot] = tuple((a - b) / (2 * h) for a, b in zip(up, dn))
    return out


def _kernel_sums(x, y, bi, bnn, ls, noise, loss_id, sigma2):
    """What the kernel's epilogue accumulates (include/muygpys_b200.h, mgp_fused_loo_grad)."""
    m, v, yky = _rows(x, y, bi, bnn, ls, noise)
    e = m - y[bi]
    w = np.ones_like(e)
    rec = np.zeros(L.MGP_PARTIALS)
    if loss_id == L.LOSS_LOOPH:
        u = e * e / (BS * BS * sigma2 * v)
        rec[L.P_AUX] = np.sum(2 * BS * BS * (np.sqrt(1 + u) - 1))
        w = 1 / np.sqrt(1 + u)
    rec[L.P_SQERR], rec[L.P_COUNT], rec[L.P_ROWS] = np.sum(e * e), len(e), len(e)
    rec[L.P_YKY], rec[L.P_SQERR_V], rec[L.P_LOGV] = yky.sum(), np.sum(w * e * e / v), np.log(v).sum()
    g = np.zeros((L.MGP_GRAD_PARAMS, 5))
    for slot, (dm, dv, dy) in _row_derivatives(x, y, bi, bnn, ls, noise).items():
        g[slot] = [np.sum(2 * e * dm), np.sum(w * 2 * e * dm / v), np.sum(w * e * e * dv / v**2),
                   np.sum(dv / v), dy.sum()]
    return rec, g


COMBOS = [(loss, analytic, aniso, other)
          for loss in ("mse", "lool", "looph") for analytic in (False, True)
          for aniso in (False, True) for other in (False, True)
          # (the nugget quirk only exists under the analytic scale)
          if not other or (analytic and loss != "mse")]


@pytest.mark.parametrize("loss,analytic,aniso,other_noise", COMBOS)
def test_finish_value_and_grad_matches_oracle_finite_differences(loss, analytic, aniso,
                                                                 other_noise):
    x, y, bi, bnn = _data()
    loss_id = {"mse": L.LOSS_MSE, "lool": L.LOSS_LOOL, "looph": L.LOSS_LOOPH}[loss]
    oid = {"mse": O.LOSS_MSE, "lool": O.LOSS_LOOL, "looph": O.LOSS_LOOPH}[loss]
    ls = [0.3, 0.45] if aniso else [0.3, 0.3]
    model_noise = 2e-3
    noise = 5e-3 if other_noise else model_noise
    fixed = 1.3
    needs_var = loss != "mse"
    is_analytic = analytic and needs_var

    def oracle(ls_now, nz):
        arg = np.asarray(ls_now, float) if aniso else float(ls_now[0])
        kw = {"loss_kwargs": {"boundary_scale": BS}} if loss == "looph" else {}
        return O.loo_objective(oid, O.KERNEL_MATERN_15, O.METRIC_L2, arg, nz, x, y, bi, bnn,
                               analytic=is_analytic, fixed_scale=fixed, model_noise=model_noise,
                               **kw)[0]

    # sigma^2 as the launches would have it: analytic -> from y^T K^-1 y at the MODEL's nugget
    g_sigma = None
    if is_analytic:
        _, _, yky_model = _rows(x, y, bi, bnn, ls, model_noise)
        sigma2 = yky_model.sum() / (B * K)
        if other_noise:
            _, g_sigma = _kernel_sums(x, y, bi, bnn, ls, model_noise, L.LOSS_NONE, sigma2)
    else:
        sigma2 = fixed
    rec, g = _kernel_sums(x, y, bi, bnn, ls, noise, loss_id, sigma2)
    value, grads = finish_value_and_grad(
        rec, g, loss_id=loss_id, k=K, d=D, anisotropic=aniso, analytic=is_analytic,
        sigma2=sigma2 if (loss == "looph" or other_noise) else None,
        fixed_scale=None if is_analytic else fixed, g_sigma=g_sigma)
    want = oracle(ls, noise)
    assert abs(value - want) <= 1e-10 * abs(want)

    def fd(fn, base):
        h = 1e-5 * base
        return (fn(base + h) - fn(base - h)) / (2 * h)

    checks = {}
    if aniso:
        for f in range(D):
            checks[f"length_scale{f}"] = fd(
                lambda v, f=f: oracle([v if i == f else ls[i] for i in range(D)], noise), ls[f])
    else:
        checks["length_scale"] = fd(lambda v: oracle([v] * D, noise), ls[0])
    # the oracle's objective moves only the optimiser's nugget, like scipy's finite differences
    checks["noise"] = fd(lambda v: oracle(ls, v), noise)
    for name, want_d in checks.items():
        assert abs(grads[name] - want_d) <= 2e-5 * max(abs(want_d), 1e-6 * abs(want)), (
            name, grads[name], want_d)


@pytest.mark.parametrize("loss", ["lool", "looph"])
def test_rank_shards_sum_to_the_full_batch(loss):
    """N > 1: every rank's launch covers its shard of the batch rows, the records and gradient
    sums are added across ranks (peer exchange / all-reduce) and finished once -- the result
    must be the full batch's (the sums are linear in the rows; sigma^2 is global)."""
    x, y, bi, bnn = _data()
    loss_id = {"lool": L.LOSS_LOOL, "looph": L.LOSS_LOOPH}[loss]
    ls, noise = [0.3, 0.45], 2e-3
    _, _, yky = _rows(x, y, bi, bnn, ls, noise)
    sigma2 = yky.sum() / (B * K)  # (looph: fixed by the global scale launch before the loss launch)
    full = _kernel_sums(x, y, bi, bnn, ls, noise, loss_id, sigma2)
    half = B // 2
    parts = [_kernel_sums(x, y, bi[s], bnn[s], ls, noise, loss_id, sigma2)
             for s in (slice(0, half), slice(half, B))]
    rec = parts[0][0] + parts[1][0]
    g = parts[0][1] + parts[1][1]
    np.testing.assert_allclose(rec, full[0], rtol=1e-12)
    kw = dict(loss_id=loss_id, k=K, d=D, anisotropic=True, analytic=True,
              sigma2=sigma2 if loss == "looph" else None)
    v_sum, g_sum = finish_value_and_grad(rec, g, **kw)
    v_full, g_full = finish_value_and_grad(*full, **kw)
    assert abs(v_sum - v_full) <= 1e-12 * abs(v_full)
    for name in g_full:
        assert abs(g_sum[name] - g_full[name]) <= 1e-9 * max(abs(g_full[name]), 1e-9)
