#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (numpy backend).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference is imported from /root/reference/src through the import shims in
oracle/ref_shims (package metadata + stubs for the absent bayes_opt/matplotlib;
SURVEY.md section 8c).  Nothing here is used at run time on the GPU box: tests
read only the committed .npz files and regenerate the inputs from the seeds in
oracle/cases.py.
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, ROOT)
os.environ.pop("MUYGPYS_BACKEND", None)

import numpy as np  # noqa: E402

from MuyGPyS.gp import MuyGPS  # noqa: E402
from MuyGPyS.gp.deformation import Anisotropy, Isotropy, F2, l2  # noqa: E402
from MuyGPyS.gp.hyperparameter import (  # noqa: E402
    AnalyticScale,
    FixedScale,
    Parameter,
    VectorParameter,
)
from MuyGPyS.gp.kernels import RBF, Matern  # noqa: E402
from MuyGPyS.gp.noise import HeteroscedasticNoise, HomoscedasticNoise  # noqa: E402
from MuyGPyS.gp.tensors import fast_nn_update  # noqa: E402
from MuyGPyS.neighbors import NN_Wrapper  # noqa: E402
from MuyGPyS.optimize import L_BFGS_B_optimize  # noqa: E402
from MuyGPyS.optimize.loss import (  # noqa: E402
    cross_entropy_fn,
    lool_fn,
    looph_fn,
    mse_fn,
    pseudo_huber_fn,
)
from MuyGPyS.examples.from_indices import (  # noqa: E402
    fast_posterior_mean_from_indices,
    regress_from_indices,
)
from MuyGPyS._src.optimize.loss.numpy import (  # noqa: E402
    _cross_entropy_fn,
    _lool_fn,
    _looph_fn,
    _mse_fn,
    _pseudo_huber_fn,
)

from oracle import numpy_oracle as O  # noqa: E402
from oracle.cases import CASES, make_data  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
LOSS_OBJ = dict(
    mse=mse_fn,
    lool=lool_fn,
    looph=looph_fn,
    pseudo_huber=pseudo_huber_fn,
    cross_entropy=cross_entropy_fn,
)
SMOOTH = {O.KERNEL_MATERN_05: 0.5, O.KERNEL_MATERN_15: 1.5, O.KERNEL_MATERN_25: 2.5,
          O.KERNEL_MATERN_INF: np.inf}


def build_model(case, noise_obj, opt_bounds=False, scale=None):
    metric = l2 if case.metric_id == O.METRIC_L2 else F2
    if case.anisotropic:
        params = [Parameter(v, (v * 0.1, v * 10.0)) if opt_bounds else Parameter(v)
                  for v in case.length_scale]
        deformation = Anisotropy(metric, VectorParameter(*params))
    else:
        v = case.length_scale
        p = Parameter(v, (v * 0.1, v * 10.0)) if opt_bounds else Parameter(v)
        deformation = Isotropy(metric, p)
    if case.kernel_id == O.KERNEL_RBF:
        kernel = RBF(deformation=deformation)
    else:
        kernel = Matern(smoothness=Parameter(SMOOTH[case.kernel_id]), deformation=deformation)
    return MuyGPS(kernel=kernel, noise=noise_obj,
                  scale=scale if scale is not None else FixedScale())


def theta_kwargs(case, factor):
    if case.anisotropic:
        return {f"length_scale{i}": v * factor for i, v in enumerate(case.length_scale)}
    return {"length_scale": case.length_scale * factor}


def run_case(case):
    data = make_data(case)
    train_x, train_y, test_x = data["train_x"], data["train_y"], data["test_x"]
    targets = train_y if case.r > 1 else train_y[:, 0]
    out = {}
    algo = "brute" if case.d > 15 else "ball_tree"
    nbrs = NN_Wrapper(train_x, case.k, nn_method="exact", algorithm=algo)
    nn_idx, nn_d2 = nbrs.get_nns(test_x)
    out["test_nn_idx"] = nn_idx.astype(np.int64)
    out["test_nn_d2"] = nn_d2
    t_idx = np.arange(case.t)

    if case.hetero:
        noise_obj = HeteroscedasticNoise(data["hetero_train_noise"][nn_idx])
    else:
        noise_obj = HomoscedasticNoise(case.noise)
    scale_val = 1.0 + 0.25 * case.seed
    scale = FixedScale()
    scale._set(scale_val)
    muygps = build_model(case, noise_obj, scale=scale)
    out["scale_val"] = np.float64(scale_val)

    # stage-level tensors on the first rows (kept small)
    rows = min(4, case.t)
    cw, pw, nn_targets = muygps.make_predict_tensors(
        t_idx[:rows], nn_idx[:rows], test_x, train_x, targets)
    out["stage_crosswise"] = cw
    out["stage_pairwise"] = pw
    out["stage_Kin"] = muygps.kernel(pw)
    out["stage_Kcross"] = muygps.kernel(cw)
    out["stage_nn_targets"] = nn_targets

    mean, var = regress_from_indices(muygps, t_idx, nn_idx, test_x, train_x, targets)
    out["mean"] = mean
    out["var"] = var

    if case.batch:
        b_idx = data["batch_idx"]
        b_nn, b_d2 = nbrs.get_batch_nns(b_idx)
        out["batch_nn_idx"] = b_nn.astype(np.int64)
        out["batch_nn_d2"] = b_d2
        amodel = build_model(case, HomoscedasticNoise(case.noise), opt_bounds=True,
                             scale=AnalyticScale() if case.r == 1 else FixedScale())
        cwd, pwd, b_t, b_nn_t = amodel.make_train_tensors(b_idx, b_nn, train_x, targets)
        factors = np.array([0.5, 1.0, 1.7])
        out["obj_factors"] = factors
        for lname in case.losses:
            obj_fn = L_BFGS_B_optimize.make_obj_fn(
                amodel, b_t, b_nn_t, cwd, pwd, loss_fn=LOSS_OBJ[lname],
                loss_kwargs=case.loss_kwargs)
            vals = [obj_fn(**theta_kwargs(case, f)) for f in factors]
            # and one evaluation that also overrides the nugget, as the optimiser would
            vals.append(obj_fn(noise=case.noise * 3.0, **theta_kwargs(case, 1.0)))
            out[f"obj_{lname}"] = np.array(vals, dtype=np.float64)
        if case.r == 1:
            Kin = amodel.kernel(pwd)
            out["analytic_scale"] = np.float64(amodel.scale.get_opt_fn(amodel)(Kin, b_nn_t))
            amodel2 = build_model(case, HomoscedasticNoise(case.noise),
                                  scale=AnalyticScale(iteration_count=3))
            out["analytic_scale_it3"] = np.float64(
                amodel2.scale.get_opt_fn(amodel2)(amodel2.kernel(pwd), b_nn_t))
        if "mse" in case.losses and case.name.startswith(("c1", "c2", "c4")):
            opt = L_BFGS_B_optimize(amodel, b_t, b_nn_t, cwd, pwd, loss_fn=mse_fn)
            names, vals, _ = opt.get_opt_params()
            out["opt_mse_names"] = np.array(names)
            out["opt_mse_vals"] = np.array(vals, dtype=np.float64)
            if case.r == 1:
                opt = opt.optimize_scale(pwd, b_nn_t)
                out["opt_mse_scale"] = np.float64(opt.scale())

    if case.fast:
        # tutorial flow (docs/examples/fast_regression_tutorial.ipynb cells 16-18):
        # neighbours of every training point incl. itself -> fast_nn_update
        tr_nn, _ = nbrs.get_nns(train_x)
        tr_nn_fast = fast_nn_update(tr_nn)
        fmodel = build_model(case, HomoscedasticNoise(case.noise))
        pw_fast = fmodel.kernel.deformation.pairwise_tensor(train_x, tr_nn_fast)
        Kin_fast = fmodel.kernel(pw_fast)
        coeffs = fmodel.fast_coefficients(Kin_fast, targets[tr_nn_fast])
        closest = nn_idx[:, 0]
        closest_set = tr_nn_fast[closest]
        fmean = fast_posterior_mean_from_indices(
            fmodel, t_idx, closest_set, test_x, train_x, closest, coeffs)
        out["fast_train_nn_idx"] = tr_nn.astype(np.int64)[:64]
        out["fast_coeffs_head"] = coeffs[:64]
        out["fast_coeffs_closest"] = coeffs[closest]
        out["fast_mean"] = fmean
    return out


def loss_fixture():
    rng = np.random.default_rng(99)
    out = {}
    for r in (1, 2, 10):
        b = 57
        p = rng.normal(size=(b, r)) if r > 1 else rng.normal(size=b)
        t = p + 0.3 * rng.normal(size=p.shape)
        v = rng.uniform(0.05, 2.0, size=b)
        out[f"pred_r{r}"] = p
        out[f"targ_r{r}"] = t
        out[f"var_r{r}"] = v
        out[f"mse_r{r}"] = np.float64(_mse_fn(p, t))
        out[f"phuber_r{r}"] = np.float64(_pseudo_huber_fn(p, t))
        out[f"phuber25_r{r}"] = np.float64(_pseudo_huber_fn(p, t, boundary_scale=2.5))
        if r == 1:
            out["lool_r1"] = np.float64(_lool_fn(p, t, v, 1.3))
            out["looph_r1"] = np.float64(_looph_fn(p, t, v, 1.3))
            out["looph2_r1"] = np.float64(_looph_fn(p, t, v, 0.7, boundary_scale=2.0))
        else:
            lab = rng.integers(0, r, size=b)
            oh = -0.1 * np.ones((b, r))
            oh[np.arange(b), lab] = 0.9
            out[f"onehot_r{r}"] = oh
            out[f"ce_r{r}"] = np.float64(_cross_entropy_fn(p, oh))
            # extreme logits exercise the eps clip inside sklearn.log_loss
            big = p * 60.0
            out[f"ce_big_r{r}"] = np.float64(_cross_entropy_fn(big, oh))
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    for case in CASES:
        res = run_case(case)
        path = os.path.join(GOLD, f"{case.name}.npz")
        np.savez_compressed(path, **res)
        print(f"{case.name}: {len(res)} arrays, {os.path.getsize(path)/1024:.1f} KiB")
    path = os.path.join(GOLD, "losses.npz")
    np.savez_compressed(path, **loss_fixture())
    print(f"losses: {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
