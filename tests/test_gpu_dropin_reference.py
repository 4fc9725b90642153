"""GPU, optional: the UNMODIFIED reference (MuyGPyS 0.9.0) with our CUDA callables
injected through its own `_backend_*` constructor arguments, side by side with the same
reference objects on their numpy defaults -- the pattern of the reference's
tests/backend/torch_correctness.py.  Runs only where MuyGPyS is importable: the build
container installs it under baseline/_ref (git-ignored; it is not part of this repo and
nothing else depends on it)."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for extra in (os.path.join(ROOT, "oracle", "ref_shims"), os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isdir(extra) and extra not in sys.path:
        sys.path.append(extra)

MuyGPyS = pytest.importorskip("MuyGPyS", reason="reference not installed on this machine")

from oracle.cases import by_name, make_data  # noqa: E402

from conftest import assert_close  # noqa: E402

pytestmark = pytest.mark.gpu


def _models(case, R):
    """(reference on numpy defaults, reference with CUDA callables injected)."""
    from MuyGPyS.gp import MuyGPS
    from MuyGPyS.gp.deformation import Anisotropy, Isotropy, l2
    from MuyGPyS.gp.hyperparameter import AnalyticScale, Parameter, VectorParameter
    from MuyGPyS.gp.kernels import Matern
    from MuyGPyS.gp.noise import HomoscedasticNoise

    def ls():
        if case.anisotropic:
            return VectorParameter(*[Parameter(v, (v * 0.1, v * 10)) for v in case.length_scale])
        return Parameter(case.length_scale, (case.length_scale * 0.1, case.length_scale * 10))

    nu = {2: 1.5, 3: 2.5}[case.kernel_id]
    Def, RDef = (Anisotropy, R.Anisotropy) if case.anisotropic else (Isotropy, R.Isotropy)
    plain = MuyGPS(kernel=Matern(smoothness=Parameter(nu), deformation=Def(l2, ls())),
                   noise=HomoscedasticNoise(case.noise), scale=AnalyticScale())
    ours = R.MuyGPS(kernel=R.Matern(smoothness=Parameter(nu), deformation=RDef(R.l2, ls())),
                    noise=R.HomoscedasticNoise(case.noise), scale=R.AnalyticScale())
    return plain, ours


@pytest.mark.parametrize("name", ["c2_m15_2d", "c4_m25_aniso"])
def test_injected_reference_matches_plain_reference(name):
    from MuyGPyS.neighbors import NN_Wrapper
    from MuyGPyS.optimize import L_BFGS_B_optimize
    from MuyGPyS.optimize.loss import lool_fn, mse_fn

    from muygpys_b200.inject import reference_objects
    from muygpys_b200.neighbors import NN_Wrapper as GpuNN

    R = reference_objects()
    case = by_name(name)
    data = make_data(case)
    x, y, q = data["train_x"], data["train_y"][:, 0], data["test_x"]
    plain, ours = _models(case, R)
    nn_ref, d_ref = NN_Wrapper(x, case.k, nn_method="exact", algorithm="ball_tree").get_nns(q)
    nn_gpu, d_gpu = GpuNN(x, case.k).get_nns(q)
    np.testing.assert_array_equal(nn_gpu, nn_ref)
    assert_close(d_gpu, d_ref, 1e-12)
    t_idx = np.arange(case.t)
    outs = []
    for model in (plain, ours):
        cw, pw, nn_t = model.make_predict_tensors(t_idx, nn_ref, q, x, y)
        Kin, Kcross = model.kernel(pw), model.kernel(cw)
        outs.append((cw, pw, Kin, Kcross, model.posterior_mean(Kin, Kcross, nn_t),
                     model.posterior_variance(Kin, Kcross)))
    for a, b, what in zip(outs[0], outs[1], ("crosswise", "pairwise", "Kin", "Kcross", "mean",
                                             "variance")):
        assert isinstance(b, np.ndarray)
        assert_close(b, a, 1e-10, what)
    # the reference's own optimiser chassis driving our kernels through its closures
    bi = data["batch_idx"]
    bnn, _ = NN_Wrapper(x, case.k, nn_method="exact", algorithm="ball_tree").get_batch_nns(bi)
    vals = []
    for model, losses in ((plain, (mse_fn, lool_fn)), (ours, (R.mse_fn, R.lool_fn))):
        cw, pw, b_t, b_nn_t = model.make_train_tensors(bi, bnn, x, y)
        row = []
        for lf in losses:
            obj = L_BFGS_B_optimize.make_obj_fn(model, b_t, b_nn_t, cw, pw, loss_fn=lf)
            kw = ({f"length_scale{i}": v * 1.3 for i, v in enumerate(case.length_scale)}
                  if case.anisotropic else {"length_scale": case.length_scale * 1.3})
            row.append(obj(**kw))
        row.append(model.scale.get_opt_fn(model)(model.kernel(pw), b_nn_t))
        vals.append(row)
    assert_close(np.array(vals[1]), np.array(vals[0]), 1e-10, "objectives and scale")


@pytest.mark.parametrize("name", ["c2_m15_2d", "c4_m25_aniso", "c1_rbf_1d"])
def test_fused_entry_points_take_genuine_reference_objects(name):
    """A reference user changes ONE import (`MuyGPyS.examples.from_indices` ->
    `muygpys_b200.examples.from_indices`) and keeps their own `MuyGPyS.gp.MuyGPS` object, loss
    functor and optimiser: the fused one-launch path must give the plain reference's numbers."""
    from MuyGPyS.examples import from_indices as ref_api
    from MuyGPyS.gp import MuyGPS
    from MuyGPyS.gp.deformation import F2, Anisotropy, Isotropy, l2
    from MuyGPyS.gp.hyperparameter import AnalyticScale, Parameter, VectorParameter
    from MuyGPyS.gp.kernels import RBF, Matern
    from MuyGPyS.gp.noise import HomoscedasticNoise
    from MuyGPyS.neighbors import NN_Wrapper
    from MuyGPyS.optimize import L_BFGS_B_optimize
    from MuyGPyS.optimize.loss import lool_fn, mse_fn

    from muygpys_b200.examples import from_indices as our_api
    from muygpys_b200.optimize.objective import make_fused_loo_crossval_fn

    case = by_name(name)
    data = make_data(case)
    x, y, q = data["train_x"], data["train_y"][:, 0], data["test_x"]

    def ls():
        if case.anisotropic:
            return VectorParameter(*[Parameter(v, (v * 0.1, v * 10)) for v in case.length_scale])
        return Parameter(case.length_scale, (case.length_scale * 0.1, case.length_scale * 10))

    if case.kernel_id == 0:
        kernel = RBF(deformation=Isotropy(F2, ls()))
    else:
        nu = {1: 0.5, 2: 1.5, 3: 2.5}[case.kernel_id]
        Def = Anisotropy if case.anisotropic else Isotropy
        kernel = Matern(smoothness=Parameter(nu), deformation=Def(l2, ls()))
    model = MuyGPS(kernel=kernel, noise=HomoscedasticNoise(case.noise), scale=AnalyticScale())
    nbrs = NN_Wrapper(x, case.k, nn_method="exact", algorithm="ball_tree")
    nn, _ = nbrs.get_nns(q)
    t_idx = np.arange(case.t)
    want_mean, want_var = ref_api.regress_from_indices(model, t_idx, nn, q, x, y)
    got_mean, got_var = our_api.regress_from_indices(model, t_idx, nn, q, x, y)
    assert isinstance(got_mean, np.ndarray)
    assert_close(got_mean, want_mean, 1e-10, "fused mean from a reference object")
    assert_close(got_var, want_var, 1e-10, "fused variance from a reference object")
    assert_close(our_api.posterior_mean_from_indices(model, t_idx, nn, q, x, y), want_mean,
                 1e-10, "posterior_mean_from_indices")
    # objective: the reference's staged closure against the one-launch objective
    bi = data["batch_idx"]
    bnn, _ = nbrs.get_batch_nns(bi)
    cw, pw, b_t, b_nn_t = model.make_train_tensors(bi, bnn, x, y)
    kw = ({f"length_scale{i}": v * 1.3 for i, v in enumerate(case.length_scale)}
          if case.anisotropic else {"length_scale": case.length_scale * 1.3})
    for lf in (mse_fn, lool_fn):
        want = L_BFGS_B_optimize.make_obj_fn(model, b_t, b_nn_t, cw, pw, loss_fn=lf)(**kw)
        got = make_fused_loo_crossval_fn(model, lf, bi, bnn, x, y)(**kw)
        assert_close(np.array([got]), np.array([want]), 1e-10, "fused objective")
    # optimiser: the reference's L_BFGS_B_optimize loop around the fused objective returns a
    # genuine MuyGPS with (nearly) the hyperparameters of the all-reference run
    want_opt = ref_api.optimize_from_indices(model, bi, bnn, x, y, loss_fn=mse_fn,
                                             opt_fn=L_BFGS_B_optimize)
    got_opt = our_api.optimize_from_indices(model, bi, bnn, x, y, loss_fn=mse_fn,
                                            opt_fn=L_BFGS_B_optimize)
    assert type(got_opt) is MuyGPS
    np.testing.assert_allclose(got_opt.get_opt_params()[1], want_opt.get_opt_params()[1],
                               rtol=5e-3)


def test_multivariate_fused_path_matches_reference_mmuygps():
    """MultivariateMuyGPS (S/gp/multivariate_muygps.py:99-340): r models over shared
    neighbourhoods.  Our fused entry points take the GENUINE reference object (and the mirror
    class built from the same dictionaries) and must reproduce the reference's numbers,
    including its quirks (scale applied twice to the variance, nugget applied twice to the fast
    coefficients)."""
    from MuyGPyS.examples import from_indices as ref_api
    from MuyGPyS.gp import MultivariateMuyGPS as RefMM
    from MuyGPyS.gp.deformation import F2, Isotropy, l2
    from MuyGPyS.gp.hyperparameter import FixedScale, Parameter
    from MuyGPyS.gp.kernels import RBF, Matern
    from MuyGPyS.gp.noise import HomoscedasticNoise
    from MuyGPyS.gp.tensors import fast_nn_update, make_fast_predict_tensors
    from MuyGPyS.neighbors import NN_Wrapper

    import muygpys_b200.gp as G
    import muygpys_b200.gp.deformation as GD
    import muygpys_b200.gp.hyperparameter as GH
    import muygpys_b200.gp.kernels as GK
    import muygpys_b200.gp.noise as GN
    from muygpys_b200 import fused
    from muygpys_b200.examples import from_indices as our_api

    rng = np.random.default_rng(21)
    n, t, k = 1500, 200, 20
    x = rng.uniform(size=(n, 3))
    q = rng.uniform(size=(t, 3))
    y = np.stack([np.sin(3 * x[:, 0]) + x[:, 1], np.cos(2 * x[:, 2]) * x[:, 0]], axis=1)
    y += 0.05 * rng.normal(size=y.shape)

    def args(M, D, N, H, kernels, deform, metric_l2, metric_f2):
        s1, s2 = H.FixedScale(), H.FixedScale()
        s1._set(1.7)
        s2._set(0.6)
        return [
            {"kernel": kernels.Matern(smoothness=H.Parameter(1.5),
                                      deformation=deform.Isotropy(metric_l2, H.Parameter(0.4))),
             "noise": N.HomoscedasticNoise(1e-3), "scale": s1},
            {"kernel": kernels.RBF(deformation=deform.Isotropy(metric_f2, H.Parameter(0.7))),
             "noise": N.HomoscedasticNoise(2e-3), "scale": s2},
        ]

    import MuyGPyS.gp.deformation as RD
    import MuyGPyS.gp.hyperparameter as RH
    import MuyGPyS.gp.kernels as RK
    import MuyGPyS.gp.noise as RN

    ref = RefMM(*args(None, None, RN, RH, RK, RD, l2, F2))
    mirror = G.MultivariateMuyGPS(*args(None, None, GN, GH, GK, GD, GD.l2, GD.F2))
    nbrs = NN_Wrapper(x, k, nn_method="exact", algorithm="ball_tree")
    nn, _ = nbrs.get_nns(q)
    t_idx = np.arange(t)
    want_mean, want_var = ref_api.regress_from_indices(ref, t_idx, nn, q, x, y)
    for model in (ref, mirror):
        assert fused.is_multivariate(model)
        got_mean, got_var = our_api.regress_from_indices(model, t_idx, nn, q, x, y)
        assert got_mean.shape == (t, 2) and got_var.shape == (t, 2)
        assert_close(got_mean, want_mean, 1e-10, "multivariate mean")
        assert_close(got_var, want_var, 1e-10, "multivariate variance")
    # fast path: coefficients of every training point, then a k-dot per test point
    train_nn, _ = nbrs.get_batch_nns(np.arange(n))
    nn_fast = fast_nn_update(train_nn)
    pw_fast, y_fast = make_fast_predict_tensors(train_nn, x, y)  # differences (n,k,k,d)
    # The reference's MultivariateMuyGPS.fast_coefficients discards the result of `mm.assign`
    # (multivariate_muygps.py:222-231; numpy's assign copies), so it returns ZEROS.  The values
    # it computes on the way -- each model's coefficients with the nugget applied twice -- are
    # the definition we reproduce.
    assert not np.any(ref.fast_coefficients(l2(pw_fast), y_fast))
    want_coeffs = np.stack(
        [m.fast_coefficients(m.noise.perturb(m.kernel(l2(pw_fast))), y_fast[:, :, i])
         for i, m in enumerate(ref.models)], axis=2)
    got_coeffs = fused.mm_fused_fast_coefficients(ref, nn_fast, x, y)
    assert_close(got_coeffs, want_coeffs, 1e-9, "multivariate fast coefficients")
    closest = nn[:, 0]
    # same defect in MultivariateMuyGPS.fast_posterior_mean (:262-270): Kcross stays zero
    assert not np.any(ref_api.fast_posterior_mean_from_indices(
        ref, t_idx, nn_fast[closest], q, x, closest, want_coeffs))
    cw_fast = ref.models[0].kernel.deformation.crosswise_tensor(q, x, t_idx, nn_fast[closest])
    want_fast = np.stack([np.einsum("bj,bj->b", m.kernel(cw_fast), want_coeffs[closest][:, :, i])
                          for i, m in enumerate(ref.models)], axis=1)
    got_fast = our_api.fast_posterior_mean_from_indices(ref, t_idx, nn_fast[closest], q, x,
                                                        closest, want_coeffs)
    assert_close(got_fast, want_fast, 1e-10, "multivariate fast mean")
    # the mirror object's staged methods (materialised tensors) follow the same definitions
    cw, pw, nn_t = mirror.make_predict_tensors(t_idx, nn, q, x, y)
    assert_close(mirror.posterior_mean(pw, cw, nn_t), want_mean, 1e-10, "staged mv mean")
    assert_close(mirror.posterior_variance(pw, cw), want_var, 1e-10, "staged mv variance")
