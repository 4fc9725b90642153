"""GPU: analytic gradient of the LOO objective from the fused kernel against
central finite differences of (a) the numpy oracle's objective and (b) the fused objective
itself, 1e-6 relative; and L-BFGS-B with jac=True reaching the finite-difference optimum."""

import numpy as np
import pytest
import torch

from oracle import numpy_oracle as O

pytestmark = pytest.mark.gpu


def _setup(seed, n, b, d, k):
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, d))
    y = np.sin(4 * x[:, 0]) + (np.cos(3 * x[:, -1]) if d > 1 else 0.0) + 0.05 * rng.normal(size=n)
    bi = np.sort(rng.choice(n, b, replace=False))
    bnn, _ = O.knn_batch(x, bi, k)
    return x, y, bi, np.ascontiguousarray(bnn)


def _model(kernel, ls, noise, analytic, noise_bounds=None):
    from muygpys_b200.gp import MuyGPS
    from muygpys_b200.gp.deformation import F2, Anisotropy, Isotropy, l2
    from muygpys_b200.gp.hyperparameter import (AnalyticScale, FixedScale, Parameter,
                                                VectorParameter)
    from muygpys_b200.gp.kernels import RBF, Matern
    from muygpys_b200.gp.noise import HomoscedasticNoise

    metric = F2 if kernel == "rbf" else l2
    if isinstance(ls, (list, tuple)):
        deformation = Anisotropy(metric, VectorParameter(*[Parameter(v, (v * 0.2, v * 5)) for v in ls]))
    else:
        deformation = Isotropy(metric, Parameter(ls, (ls * 0.2, ls * 5)))
    kern = RBF(deformation=deformation) if kernel == "rbf" else Matern(
        smoothness=Parameter(kernel), deformation=deformation)
    scale = AnalyticScale() if analytic else FixedScale()
    if not analytic:
        scale._set(1.3)
    nz = HomoscedasticNoise(noise, noise_bounds) if noise_bounds else HomoscedasticNoise(noise)
    return MuyGPS(kernel=kern, noise=nz, scale=scale)


CASES = [
    # kernel, length scale(s), d, k, loss, analytic scale, nugget gradient checked
    (1.5, 0.3, 2, 50, "lool", True, False),
    (1.5, 0.3, 2, 50, "mse", False, True),
    (2.5, [0.3, 0.6], 2, 30, "lool", False, True),
    (0.5, [0.4, 0.3, 0.5], 3, 23, "lool", True, False),
    ("rbf", 0.2, 1, 30, "mse", False, True),
    (np.inf, 0.4, 2, 47, "lool", True, False),   # k % 8 == 7: augmented rows straddle tiles
    (1.5, 0.25, 2, 62, "lool", False, True),
    # k > 62: GRAD instantiations with 9..13 tile rows (C4 is the k = 100 case)
    (2.5, [0.3, 0.6], 2, 100, "lool", True, False),
    (1.5, 0.3, 2, 63, "mse", False, True),
    (0.5, [0.4, 0.3, 0.5], 3, 79, "lool", False, True),
    ("rbf", 0.2, 1, 88, "mse", False, True),
    (np.inf, 0.4, 2, 102, "lool", True, False),
    # looph: Huber-weighted sums in the same epilogue; analytic scale = one extra plain launch
    (1.5, 0.3, 2, 50, "looph", True, False),
    (2.5, [0.3, 0.6], 2, 30, "looph", False, True),
    (0.5, 0.3, 2, 100, "looph", True, False),
    ("rbf", 0.2, 1, 23, "looph", False, True),
    # the nugget under the ANALYTIC scale: sigma^2 is taken at the model's nugget (reference
    # quirk), so it does not move with the optimiser's `noise=`
    (1.5, 0.3, 2, 50, "lool", True, True),
    (2.5, [0.3, 0.6], 2, 38, "looph", True, True),
]


@pytest.mark.parametrize("kernel,ls,d,k,loss,analytic,check_noise", CASES)
def test_analytic_gradient_matches_finite_differences(kernel, ls, d, k, loss, analytic,
                                                      check_noise):
    from muygpys_b200.optimize import loss as losses
    from muygpys_b200.optimize.objective import (make_fused_loo_crossval_fn,
                                                 make_fused_loo_value_and_grad_fn)

    x, y, bi, bnn = _setup(hash((d, k)) % 1000, 3000, 400, d, k)
    noise = 2e-3
    model = _model(kernel, ls, noise, analytic)
    lf = getattr(losses, f"{loss}_fn")
    vg = make_fused_loo_value_and_grad_fn(model, lf, bi, bnn, x, y)
    obj = make_fused_loo_crossval_fn(model, lf, bi, bnn, x, y)
    aniso = isinstance(ls, list)
    theta = ({f"length_scale{i}": v * 1.1 for i, v in enumerate(ls)} if aniso
             else {"length_scale": ls * 1.1})
    val, grads = vg(**theta)
    assert abs(val - obj(**theta)) <= 1e-12 * abs(val)
    names = list(theta) + (["noise"] if check_noise else [])
    for name in names:
        base = theta.get(name, noise)
        h = 1e-5 * base
        up = dict(theta, **{name: base + h})
        dn = dict(theta, **{name: base - h})
        fd = (obj(**up) - obj(**dn)) / (2 * h)
        assert abs(grads[name] - fd) <= 2e-6 * max(abs(fd), 1e-3 * abs(val) / base), (
            name, grads[name], fd)


def test_gradient_against_oracle_objective():
    """Finite differences of the ORACLE's objective (numpy restatement of the reference's
    make_loo_crossval_fn), so the check does not lean on our own kernels."""
    from muygpys_b200.optimize.loss import lool_fn
    from muygpys_b200.optimize.objective import make_fused_loo_value_and_grad_fn

    x, y, bi, bnn = _setup(5, 2000, 150, 2, 50)
    model = _model(1.5, 0.3, 1e-3, True)
    vg = make_fused_loo_value_and_grad_fn(model, lool_fn, bi, bnn, x, y)
    ls = 0.27
    val, grads = vg(length_scale=ls)

    def oracle(l):
        return O.loo_objective(O.LOSS_LOOL, O.KERNEL_MATERN_15, O.METRIC_L2, l, 1e-3, x, y, bi,
                               bnn)[0]

    assert abs(val - oracle(ls)) <= 1e-10 * abs(val)
    h = 1e-5 * ls
    fd = (oracle(ls + h) - oracle(ls - h)) / (2 * h)
    assert abs(grads["length_scale"] - fd) <= 1e-6 * abs(fd), (grads, fd)


@pytest.mark.parametrize("loss,aniso", [("lool", False), ("looph", True), ("lool", True)])
def test_gradient_with_an_optimiser_nugget_that_differs_from_the_models(loss, aniso):
    """Analytic scale while the optimiser's `noise=` differs from the model's nugget: sigma^2 and
    its length-scale derivatives come from a second gradient launch at the MODEL's nugget
    (S/gp/hyperparameter/scale.py:206-208); checked against central differences of the fused
    objective, which reproduces the quirk (golden `obj_*` cases), and of the oracle's."""
    from muygpys_b200.optimize import loss as losses
    from muygpys_b200.optimize.objective import (make_fused_loo_crossval_fn,
                                                 make_fused_loo_value_and_grad_fn)

    x, y, bi, bnn = _setup(21, 3000, 300, 2, 50)
    model_noise = 2e-3
    ls = [0.3, 0.5] if aniso else 0.3
    model = _model(1.5, ls, model_noise, True)
    lf = getattr(losses, f"{loss}_fn")
    vg = make_fused_loo_value_and_grad_fn(model, lf, bi, bnn, x, y)
    obj = make_fused_loo_crossval_fn(model, lf, bi, bnn, x, y)
    theta = ({f"length_scale{i}": v * 1.1 for i, v in enumerate(ls)} if aniso
             else {"length_scale": ls * 1.1})
    theta["noise"] = 5e-3
    val, grads = vg(**theta)
    assert abs(val - obj(**theta)) <= 1e-12 * abs(val)
    lid = O.LOSS_LOOL if loss == "lool" else O.LOSS_LOOPH
    ls_now = np.array([theta[f"length_scale{i}"] for i in range(2)]) if aniso \
        else theta["length_scale"]
    want = O.loo_objective(lid, O.KERNEL_MATERN_15, O.METRIC_L2, ls_now, 5e-3, x, y, bi, bnn,
                           model_noise=model_noise)[0]
    assert abs(val - want) <= 1e-10 * abs(val)
    for name in theta:
        base = theta[name]
        h = 1e-5 * base
        fd = (obj(**dict(theta, **{name: base + h})) - obj(**dict(theta, **{name: base - h}))) \
            / (2 * h)
        assert abs(grads[name] - fd) <= 2e-6 * max(abs(fd), 1e-3 * abs(val) / base), (
            name, grads[name], fd)


def test_looph_gradient_against_oracle_objective():
    """looph (analytic scale, boundary_scale = 2.5 through loss_kwargs) against central
    differences of the oracle's objective."""
    from muygpys_b200.optimize.loss import looph_fn
    from muygpys_b200.optimize.objective import make_fused_loo_value_and_grad_fn

    x, y, bi, bnn = _setup(6, 2000, 150, 2, 50)
    y = y + 0.5 * (np.arange(len(y)) % 97 == 0)  # a few outliers: Huber weights well below 1
    model = _model(1.5, 0.3, 1e-3, True)
    vg = make_fused_loo_value_and_grad_fn(model, looph_fn, bi, bnn, x, y,
                                          loss_kwargs={"boundary_scale": 2.5})
    ls = 0.27
    val, grads = vg(length_scale=ls)

    def oracle(l):
        return O.loo_objective(O.LOSS_LOOPH, O.KERNEL_MATERN_15, O.METRIC_L2, l, 1e-3, x, y, bi,
                               bnn, loss_kwargs={"boundary_scale": 2.5})[0]

    assert abs(val - oracle(ls)) <= 1e-10 * abs(val)
    h = 1e-5 * ls
    fd = (oracle(ls + h) - oracle(ls - h)) / (2 * h)
    assert abs(grads["length_scale"] - fd) <= 1e-6 * abs(fd), (grads, fd)


def test_gradient_against_oracle_objective_c4_shape():
    """The C4 shape (anisotropic Matern 5/2, k = 100, lool with the analytic scale): both length
    scales against central differences of the oracle's objective (isotropic oracle calls are not
    enough here, so the deformation is rebuilt per evaluation)."""
    from muygpys_b200.optimize.loss import lool_fn
    from muygpys_b200.optimize.objective import make_fused_loo_value_and_grad_fn

    x, y, bi, bnn = _setup(11, 2500, 120, 2, 100)
    model = _model(2.5, [0.1, 0.5], 1e-3, True)
    vg = make_fused_loo_value_and_grad_fn(model, lool_fn, bi, bnn, x, y)
    ls = np.array([0.12, 0.45])
    val, grads = vg(length_scale0=ls[0], length_scale1=ls[1])

    def oracle(l):
        return O.loo_objective(O.LOSS_LOOL, O.KERNEL_MATERN_25, O.METRIC_L2, np.asarray(l), 1e-3,
                               x, y, bi, bnn)[0]

    assert abs(val - oracle(ls)) <= 1e-10 * abs(val)
    for f in range(2):
        h = 1e-5 * ls[f]
        e = np.zeros(2)
        e[f] = h
        fd = (oracle(ls + e) - oracle(ls - e)) / (2 * h)
        assert abs(grads[f"length_scale{f}"] - fd) <= 1e-6 * abs(fd), (f, grads, fd)


def test_lbfgsb_with_gradient_reaches_the_finite_difference_optimum():
    from muygpys_b200.examples.from_indices import optimize_from_indices
    from muygpys_b200.optimize.loss import lool_fn

    x, y, bi, bnn = _setup(9, 6000, 1500, 2, 30)
    model = _model(2.5, [0.3, 0.6], 1e-3, True)
    fd_opt = optimize_from_indices(model, bi, bnn, x, y, loss_fn=lool_fn)
    gr_opt = optimize_from_indices(model, bi, bnn, x, y, loss_fn=lool_fn, use_gradient=True)
    a, b = fd_opt.get_opt_params()[1], gr_opt.get_opt_params()[1]
    np.testing.assert_allclose(b, a, rtol=2e-3)


@pytest.mark.parametrize("d,k,kid", [(2, 50, 2), (1, 7, 1), (3, 23, 3), (2, 47, 4), (2, 62, 1),
                                     (3, 38, 2)])
def test_back_substitution_kernels_agree(d, k, kid):
    """The two independent implementations of the back substitution -- the thread-per-tile
    kernel's GRAD instantiation (variant 0, csrc/fused_tp.cuh: factor warp, one slot per tile)
    and the lane-parallel column kernel (variant 4, csrc/fused_col.cuh) -- on the same inputs:
    loss / scale partials, the 20 gradient sums and the fast-mean coefficients agree to
    rounding (the factorisations are bit-identical, the summation orders differ), and repeated
    launches are bit-reproducible."""
    from muygpys_b200 import _lib as L
    from muygpys_b200 import ops

    rng = np.random.default_rng(7000 + 10 * k + d)
    n, b = 6000, 1531  # (not a multiple of any kernel's neighbourhoods per CTA)
    x = dev_t(rng.uniform(size=(n, d)))
    y = dev_t(np.sin(3 * rng.uniform(size=n)) + 0.1 * rng.normal(size=n))
    bi = dev_t(np.sort(rng.choice(n, b, replace=False)))
    nn, _ = ops.knn(x, x[bi], k + 1)
    nn = nn[:, 1:].contiguous()
    mid = 1 if kid == 0 else 0
    ls = [0.2, 0.35, 0.5][:d] if k % 2 else 0.3
    got = {}
    try:
        for variant in (0, 4):
            ops.set_fused_variant(variant)
            loo = ops.FusedLoo(x, y, bi, nn, kernel_id=kid, metric_id=mid, loss_id=L.LOSS_LOOL,
                               want_grad=True)
            rec = loo.record(loo.launch(ls, 1e-3)).copy()
            grad = loo.grad.numpy().copy()
            again = loo.record(loo.launch(ls, 1e-3))
            np.testing.assert_array_equal(again, rec)
            np.testing.assert_array_equal(loo.grad.numpy(), grad)
            co = ops.fused_posterior(x, x, bi, nn, y, kernel_id=kid, metric_id=mid,
                                     length_scale=ls, noise=1e-3, want_coeffs=True)
            got[variant] = (rec, grad, co["coeffs"].cpu().numpy(), co["mean"].cpu().numpy())
    finally:
        ops.set_fused_variant(0)
    for a, b_, name in zip(got[0], got[4], ("partials", "gradient sums", "coefficients", "mean")):
        scale = np.max(np.abs(b_)) + 1e-300
        assert np.max(np.abs(a - b_)) <= 1e-10 * scale, (name, np.max(np.abs(a - b_)) / scale)
    assert np.all(np.isfinite(got[0][2]))


def dev_t(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()
