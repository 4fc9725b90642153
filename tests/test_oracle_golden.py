"""CPU: the numpy oracle reproduces the real reference's recorded outputs.

The .npz files under tests/golden were produced by oracle/make_golden.py, which
drives the unmodified reference (numpy backend) through its public objects.
"""

import numpy as np
import pytest

from oracle import numpy_oracle as O
from oracle.cases import CASES, make_data

from conftest import assert_close, load_golden

LOSS_ID = dict(mse=O.LOSS_MSE, lool=O.LOSS_LOOL, looph=O.LOSS_LOOPH,
               pseudo_huber=O.LOSS_PSEUDO_HUBER, cross_entropy=O.LOSS_CROSS_ENTROPY)


def _ls(case, factor=1.0):
    if case.anisotropic:
        return np.asarray(case.length_scale) * factor
    return case.length_scale * factor


def _2d(x):
    return x[:, None] if x.ndim == 1 else x


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_knn_matches_sklearn(case):
    g = load_golden(case.name)
    data = make_data(case)
    idx, d2 = O.knn_exact(data["train_x"], data["test_x"], case.k)
    assert idx.dtype == np.int64
    np.testing.assert_array_equal(idx, g["test_nn_idx"])
    assert_close(d2, g["test_nn_d2"], rtol=1e-12, what="dist2")
    if case.batch:
        bidx, bd2 = O.knn_batch(data["train_x"], data["batch_idx"], case.k)
        np.testing.assert_array_equal(bidx, g["batch_nn_idx"])
        assert_close(bd2, g["batch_nn_d2"], rtol=1e-12, what="batch dist2")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_stage_tensors(case):
    g = load_golden(case.name)
    data = make_data(case)
    rows = g["stage_Kin"].shape[0]
    nn = g["test_nn_idx"][:rows]
    cd = O.crosswise_tensor(data["test_x"], data["train_x"], np.arange(rows), nn)
    pd = O.pairwise_tensor(data["train_x"], nn)
    if case.anisotropic:
        np.testing.assert_array_equal(cd, g["stage_crosswise"])
        np.testing.assert_array_equal(pd, g["stage_pairwise"])
    else:
        assert_close(O.metric_reduce(case.metric_id, cd), g["stage_crosswise"], 1e-14)
        assert_close(O.metric_reduce(case.metric_id, pd), g["stage_pairwise"], 1e-14)
    Kin, Kcross = O.kernel_tensors(case.kernel_id, case.metric_id, _ls(case),
                                   _2d(data["train_x"]), _2d(data["test_x"]),
                                   np.arange(rows), nn)
    assert_close(Kin, g["stage_Kin"], 1e-13, "Kin")
    assert_close(Kcross, g["stage_Kcross"], 1e-13, "Kcross")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_predict_pipeline(case):
    g = load_golden(case.name)
    data = make_data(case)
    nn = g["test_nn_idx"]
    y = data["train_y"] if case.r > 1 else data["train_y"][:, 0]
    noise = data["hetero_train_noise"][nn] if case.hetero else case.noise
    mean, var = O.predict(case.kernel_id, case.metric_id, _ls(case), noise,
                          float(g["scale_val"]), _2d(data["train_x"]), y,
                          _2d(data["test_x"]), np.arange(case.t), nn)
    assert_close(mean, g["mean"], 1e-12, "mean")
    assert_close(var, g["var"], 1e-12, "var")


@pytest.mark.parametrize("case", [c for c in CASES if c.batch], ids=lambda c: c.name)
def test_loo_objective(case):
    g = load_golden(case.name)
    data = make_data(case)
    y = data["train_y"] if case.r > 1 else data["train_y"][:, 0]
    x = _2d(data["train_x"])
    for lname in case.losses:
        want = g[f"obj_{lname}"]
        got = []
        for f in g["obj_factors"]:
            v, _ = O.loo_objective(LOSS_ID[lname], case.kernel_id, case.metric_id,
                                   _ls(case, f), case.noise, x, y,
                                   data["batch_idx"], g["batch_nn_idx"],
                                   analytic=(case.r == 1), loss_kwargs=case.loss_kwargs)
            got.append(v)
        v, _ = O.loo_objective(LOSS_ID[lname], case.kernel_id, case.metric_id,
                               _ls(case), case.noise * 3.0, x, y,
                               data["batch_idx"], g["batch_nn_idx"],
                               analytic=(case.r == 1), loss_kwargs=case.loss_kwargs,
                               model_noise=case.noise)
        got.append(v)
        assert_close(np.array(got), want, 1e-11, f"obj_{lname}")
    if case.r == 1:
        Kin, _ = O.kernel_tensors(case.kernel_id, case.metric_id, _ls(case), x, x,
                                  data["batch_idx"], g["batch_nn_idx"])
        y_nn = y[g["batch_nn_idx"]]
        assert_close(O.analytic_scale_opt(Kin, y_nn, case.noise),
                     g["analytic_scale"], 1e-12, "scale")
        assert_close(O.analytic_scale_opt(Kin, y_nn, case.noise, 3),
                     g["analytic_scale_it3"], 1e-12, "scale it3")


@pytest.mark.parametrize("case", [c for c in CASES if c.fast], ids=lambda c: c.name)
def test_fast_pipeline(case):
    g = load_golden(case.name)
    data = make_data(case)
    x, tx = _2d(data["train_x"]), _2d(data["test_x"])
    y = data["train_y"] if case.r > 1 else data["train_y"][:, 0]
    tr_nn, _ = O.knn_exact(x, x, case.k)
    np.testing.assert_array_equal(tr_nn[:64], g["fast_train_nn_idx"])
    fast_nn = O.fast_nn_update(tr_nn)
    closest = g["test_nn_idx"][:, 0]
    rows = fast_nn[closest]
    Kin, _ = O.kernel_tensors(case.kernel_id, case.metric_id, _ls(case), x, x,
                              closest, rows)
    coeffs = O.fast_precompute(O.homoscedastic_perturb(Kin, case.noise), y[rows])
    assert_close(coeffs, g["fast_coeffs_closest"], 1e-9, "coeffs")
    cd = O.crosswise_tensor(tx, x, np.arange(case.t), rows)
    Kcross = O.kernel_fn(case.kernel_id, O.apply_length_scale(
        case.metric_id, O.metric_reduce(case.metric_id, cd), case.length_scale))
    assert_close(O.fast_posterior_mean(Kcross, coeffs), g["fast_mean"], 1e-10, "fast mean")


def test_losses():
    g = load_golden("losses")
    for r in (1, 2, 10):
        p, t, v = g[f"pred_r{r}"], g[f"targ_r{r}"], g[f"var_r{r}"]
        assert_close(O.mse(p, t), g[f"mse_r{r}"], 1e-14)
        assert_close(O.pseudo_huber(p, t), g[f"phuber_r{r}"], 1e-13)
        assert_close(O.pseudo_huber(p, t, 2.5), g[f"phuber25_r{r}"], 1e-13)
        if r == 1:
            assert_close(O.lool(p, t, v, 1.3), g["lool_r1"], 1e-13)
            assert_close(O.looph(p, t, v, 1.3), g["looph_r1"], 1e-13)
            assert_close(O.looph(p, t, v, 0.7, 2.0), g["looph2_r1"], 1e-13)
        else:
            oh = g[f"onehot_r{r}"]
            assert_close(O.cross_entropy(p, oh), g[f"ce_r{r}"], 1e-13)
            assert_close(O.cross_entropy(p * 60.0, oh), g[f"ce_big_r{r}"], 1e-13)


def test_chunk_rule():
    # S/_src/mpi_utils.py:36-41: larger chunks go to the LAST ranks
    assert O.chunk_sizes(10, 4) == [2, 2, 3, 3]
    assert O.chunk_sizes(8, 4) == [2, 2, 2, 2]
    assert O.chunk_sizes(3, 8) == [0, 0, 0, 0, 0, 1, 1, 1]
    for count in (0, 1, 17, 1000):
        for size in (1, 2, 3, 8):
            s = O.chunk_sizes(count, size)
            assert sum(s) == count and len(s) == size and max(s) - min(s) <= 1
