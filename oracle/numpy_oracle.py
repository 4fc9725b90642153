"""CPU oracle for the MuyGPyS per-neighbourhood hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy *restatement* of the algorithm that the reference's
numpy backend executes on the hot path named in BASELINE.json.  It exists so the
CUDA path can be checked on a machine where `/root/reference` is absent (the GPU
box).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package
`muygpys_b200` never does and has no CPU fallback.

Parity status: **pinned**.  `oracle/make_golden.py` imports the real reference
(numpy backend) in the build container and records its outputs on seeded inputs
into `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
below against those recordings (and, when the reference is importable, against
live calls).  Third-party arithmetic the reference delegates to (LAPACK gesv via
`numpy.linalg.solve`, sklearn NearestNeighbors, sklearn.log_loss + scipy softmax)
is used the same way here (LU solve) or restated (KNN brute force, CE formula)
and pinned through the same fixtures.

All arrays are float64 / int64, C-order, exactly as the reference's `mm.ftype` /
`mm.itype` (S/_src/math/numpy.py:92-96).  `S/` = /root/reference/src/MuyGPyS/.
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# identifiers shared with include/muygpys_b200.h (kept textually in sync by
# tests/test_cabi_header.py)
# --------------------------------------------------------------------------
KERNEL_RBF = 0  # exp(-x/2) on F2/l^2 input             S/_src/gp/kernels/numpy.py:12-13
KERNEL_MATERN_05 = 1  # exp(-x)                               :16-17
KERNEL_MATERN_15 = 2  # (1+sqrt3 x) exp(-sqrt3 x)             :20-22
KERNEL_MATERN_25 = 3  # (1+sqrt5 x+(sqrt5 x)^2/3) exp(-sqrt5 x) :25-27
KERNEL_MATERN_INF = 4  # exp(-x^2/2)                           :30-31

METRIC_L2 = 0  # sqrt(sum diff^2), length scale rule x/l      S/gp/deformation/metric.py:237-250
METRIC_F2 = 1  # sum diff^2,       length scale rule x/l^2    S/gp/deformation/metric.py:252-265

LOSS_NONE = 0
LOSS_MSE = 1
LOSS_LOOL = 2
LOSS_LOOPH = 3
LOSS_PSEUDO_HUBER = 4
LOSS_CROSS_ENTROPY = 5


# --------------------------------------------------------------------------
# a2/a3: difference tensors                       S/_src/gp/tensors/numpy.py:47-69
# --------------------------------------------------------------------------
def crosswise_tensor(data, nn_data, data_indices, nn_indices):
    """(b,k,d) differences `data[data_indices][:,None,:] - nn_data[nn_indices]`.

    Follows S/_src/gp/tensors/numpy.py:47-58, including the 1-D data branch
    which yields a trailing feature axis of length one.
    """
    data = np.asarray(data)
    nn_data = np.asarray(nn_data)
    query_rows = data[data_indices]
    neighbour_rows = nn_data[nn_indices]
    if data.ndim == 1:
        return query_rows[:, None, None] - neighbour_rows[..., None]
    return query_rows[:, None, :] - neighbour_rows


def pairwise_tensor(data, nn_indices):
    """(b,k,k,d) differences with element [b,i,j,:] = P[b,i] - P[b,j].

    S/_src/gp/tensors/numpy.py:61-69: a (b,k,1,d) view minus a (b,1,k,d) view,
    so axis 1 indexes the minuend.  (The sign only matters to consumers of raw
    anisotropic differences; every metric squares it.)  Diagonal is exactly 0.
    """
    data = np.asarray(data)
    rows = data[nn_indices]
    if data.ndim == 1:
        return rows[..., :, None, None] - rows[..., None, :, None]
    return rows[..., None, :] - rows[..., None, :, :]


def F2(diffs):
    """S/_src/gp/tensors/numpy.py:89-90."""
    return np.sum(diffs**2, axis=-1)


def l2(diffs):
    """S/_src/gp/tensors/numpy.py:93-94."""
    return np.sqrt(F2(diffs))


def metric_reduce(metric_id, diffs):
    return l2(diffs) if metric_id == METRIC_L2 else F2(diffs)


def apply_length_scale(metric_id, dists, length_scale):
    """l2: x / l ; F2: x / l**2           S/gp/deformation/metric.py:241,264."""
    if metric_id == METRIC_L2:
        return dists / length_scale
    return dists / length_scale**2


def isotropic_deformation(metric_id, dists, length_scale):
    """S/gp/deformation/isotropy.py:60-89 for a scalar length scale."""
    return apply_length_scale(metric_id, dists, length_scale)


def anisotropic_deformation(metric_id, diffs, length_scales):
    """`metric(diffs / [l_0..l_{d-1}])`       S/gp/deformation/anisotropy.py:43-70."""
    length_scales = np.asarray(length_scales, dtype=np.float64)
    if diffs.shape[-1] != length_scales.shape[0]:
        raise ValueError(
            f"Difference tensor of shape {diffs.shape} must have final "
            f"dimension size of {length_scales.shape[0]}"
        )
    return metric_reduce(metric_id, diffs / length_scales)


# --------------------------------------------------------------------------
# a7: covariance functions of the scaled distance   S/_src/gp/kernels/numpy.py:12-31
# --------------------------------------------------------------------------
def kernel_fn(kernel_id, x):
    x = np.asarray(x, dtype=np.float64)
    if kernel_id == KERNEL_RBF:
        return np.exp(-x / 2.0)
    if kernel_id == KERNEL_MATERN_05:
        return np.exp(-x)
    if kernel_id == KERNEL_MATERN_15:
        s = x * np.sqrt(3)
        return (1.0 + s) * np.exp(-s)
    if kernel_id == KERNEL_MATERN_25:
        s = x * np.sqrt(5)
        return (1.0 + s + s**2 / 3.0) * np.exp(-s)
    if kernel_id == KERNEL_MATERN_INF:
        return np.exp(-(x**2) / 2.0)
    raise ValueError(f"unknown kernel id {kernel_id}")


# --------------------------------------------------------------------------
# a8: nugget                                       S/_src/gp/noise/numpy.py:9-27,56-67
# --------------------------------------------------------------------------
def homoscedastic_perturb(Kin, noise_variance):
    if Kin.ndim != 3:
        raise ValueError(
            "homoscedastic perturbation is not implemented for tensors of "
            f"shape {Kin.shape}"
        )
    k = Kin.shape[1]
    return Kin + noise_variance * np.eye(k)


def heteroscedastic_perturb(Kin, noise_variances):
    out = Kin.copy()
    b, k, _ = Kin.shape
    diag = np.arange(k)
    out[:, diag, diag] += np.asarray(noise_variances).reshape(b, k)
    return out


# --------------------------------------------------------------------------
# a9/a10: posterior mean and diagonal variance      S/_src/gp/muygps/numpy.py:17-67
# --------------------------------------------------------------------------
def posterior_mean(Kin, Kcross, nn_targets):
    """`solve(Kin, Kcross)^T @ nn_targets` (LAPACK gesv LU, like the reference).

    Kin (b,k,k) already perturbed, Kcross (b,k), nn_targets (b,k) or (b,k,r)
    -> (b,) or (b,r).   S/_src/gp/muygps/numpy.py:17-41.
    """
    b, k, _ = Kin.shape
    y = nn_targets.reshape(b, k, -1)
    F = np.linalg.solve(Kin, Kcross.reshape(b, k, 1))  # (b,k,1)
    out = np.swapaxes(F, -2, -1) @ y  # (b,1,r)
    return out.reshape((b,) + nn_targets.shape[2:])


def diagonal_variance(Kin, Kcross, Kout=1.0):
    """`Kout - Kcross^T solve(Kin, Kcross)`      S/_src/gp/muygps/numpy.py:44-67."""
    b, k, _ = Kin.shape
    kc = Kcross.reshape(b, k, 1)
    F = np.linalg.solve(Kin, kc)
    return Kout - (np.swapaxes(F, -2, -1) @ kc).reshape(b)


# --------------------------------------------------------------------------
# a11/a12: fast posterior mean                      S/_src/gp/muygps/numpy.py:70-95
# --------------------------------------------------------------------------
def fast_nn_update(train_nn_indices):
    """[i | nn_0..nn_{k-2}]                  S/_src/gp/tensors/numpy.py:97-108."""
    n = train_nn_indices.shape[0]
    own = np.arange(n, dtype=train_nn_indices.dtype)[:, None]
    return np.concatenate((own, train_nn_indices[:, :-1]), axis=1)


def fast_precompute(Kin, nn_targets_fast):
    """coefficients C = solve(Kin, Y), squeezed.  S/_src/gp/muygps/numpy.py:88-95."""
    y = nn_targets_fast
    if y.ndim == 2:
        y = y[:, :, None]
    return np.squeeze(np.linalg.solve(Kin, y))


def fast_posterior_mean(Kcross, coeffs):
    """einsum('ij,ijk->ik'), squeezed.         S/_src/gp/muygps/numpy.py:70-77."""
    return np.squeeze(np.einsum("ij,ijk->ik", Kcross, np.atleast_3d(coeffs)))


# --------------------------------------------------------------------------
# a15: analytic scale                               S/_src/optimize/scale/numpy.py:9-34
# --------------------------------------------------------------------------
def analytic_scale_unnormalized(Kin, nn_targets):
    y = np.atleast_3d(nn_targets)
    return np.sum(np.einsum("ijk,ijk->ik", y, np.linalg.solve(Kin, y)))


def analytic_scale(Kin, nn_targets):
    b, k, _ = Kin.shape
    return analytic_scale_unnormalized(Kin, nn_targets.reshape(b, k, 1)) / (b * k)


def analytic_scale_opt(Kin_unperturbed, nn_targets, noise, iteration_count=1):
    """AnalyticScale.get_opt_fn                 S/gp/hyperparameter/scale.py:205-217."""
    pK = homoscedastic_perturb(Kin_unperturbed, noise)
    scale = analytic_scale(pK, nn_targets)
    for _ in range(1, iteration_count):
        scale = 0.5 * (scale + analytic_scale(scale * pK, nn_targets))
    return scale


# --------------------------------------------------------------------------
# a14: losses                                       S/_src/optimize/loss/numpy.py:12-112
# --------------------------------------------------------------------------
def mse(predictions, targets):
    return np.sum((predictions - targets) ** 2) / np.prod(predictions.shape)


def lool(predictions, targets, variances, scale):
    v = scale * variances
    return np.sum((predictions - targets) ** 2 / v + np.log(v))


def looph(predictions, targets, variances, scale, boundary_scale=3.0):
    v = scale * variances
    bs2 = boundary_scale**2
    return np.sum(
        2 * bs2 * (np.sqrt(1 + (targets - predictions) ** 2 / (bs2 * v)) - 1)
        + np.log(v)
    )


def pseudo_huber(predictions, targets, boundary_scale=1.5):
    return boundary_scale**2 * np.sum(
        np.sqrt(1 + ((targets - predictions) / boundary_scale) ** 2) - 1
    )


def cross_entropy(predictions, targets):
    """sklearn.log_loss(1[t>0], softmax(p, axis=1), normalize=False) restated.

    The reference (S/_src/optimize/loss/numpy.py:12-19) delegates to
    scikit-learn (>=0.23.2, 1.9.0 in this image) and scipy.special.softmax.
    With row-stochastic probabilities sklearn's multi-label path evaluates
    -sum(y * log(clip(p, eps, 1-eps))) with eps = finfo(float64).eps; pinned
    against the real call in tests/golden/losses.npz.
    """
    p = np.asarray(predictions, dtype=np.float64)
    t = np.asarray(targets, dtype=np.float64)
    z = p - np.max(p, axis=1, keepdims=True)
    e = np.exp(z)
    sm = e / np.sum(e, axis=1, keepdims=True)
    eps = np.finfo(np.float64).eps
    sm = np.clip(sm, eps, 1 - eps)
    onehot = np.where(t > 0.0, 1.0, 0.0)
    return float(-np.sum(onehot * np.log(sm)))


# --------------------------------------------------------------------------
# a1: exact KNN (brute force restatement of sklearn NearestNeighbors, p=2)
#     S/neighbors.py:129-262 -- returns int64 indices sorted by ascending
#     distance and SQUARED l2 distances (:246-250).
# --------------------------------------------------------------------------
def knn_exact(train, queries, k, chunk=2048):
    train = np.asarray(train, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    if train.ndim == 1:
        train = train[:, None]
    if queries.ndim == 1:
        queries = queries[:, None]
    q = queries.shape[0]
    idx = np.empty((q, k), dtype=np.int64)
    d2 = np.empty((q, k), dtype=np.float64)
    for s in range(0, q, chunk):
        block = queries[s : s + chunk]
        # direct differences, never the Gram trick, so near-ties rank exactly
        dist = np.zeros((block.shape[0], train.shape[0]))
        for f in range(train.shape[1]):
            dist += (block[:, f : f + 1] - train[None, :, f]) ** 2
        order = np.argsort(dist, axis=1, kind="stable")[:, :k]
        idx[s : s + chunk] = order
        d2[s : s + chunk] = np.take_along_axis(dist, order, axis=1)
    return idx, d2


def knn_batch(train, batch_indices, k):
    """get_batch_nns: query k+1, drop column 0.    S/neighbors.py:169-211."""
    train2 = train if np.ndim(train) == 2 else np.asarray(train)[:, None]
    idx, d2 = knn_exact(train2, train2[batch_indices], k + 1)
    return idx[:, 1:], d2[:, 1:]


# --------------------------------------------------------------------------
# a17: data-parallel chunk rule                     S/_src/mpi_utils.py:36-41
# --------------------------------------------------------------------------
def chunk_sizes(count, size):
    base = int(count / size)
    extra = count - base * size
    return [base + 1 if i >= size - extra else base for i in range(size)]


# --------------------------------------------------------------------------
# pipelines (a13 + a5/a6 + a7 + a8 + a9 + a10), one neighbourhood per row
# --------------------------------------------------------------------------
def kernel_tensors(
    kernel_id,
    metric_id,
    length_scale,
    train_x,
    query_x,
    query_idx,
    nn_idx,
):
    """Kin (b,k,k) unperturbed and Kcross (b,k) straight from indices.

    `length_scale` scalar -> Isotropy, 1-D array of length d -> Anisotropy.
    Mirrors MuyGPS.make_predict_tensors + kernel(...) (S/gp/muygps.py:405-475,
    S/examples/from_indices.py:22-39).
    """
    cd = crosswise_tensor(query_x, train_x, query_idx, nn_idx)
    pd = pairwise_tensor(train_x, nn_idx)
    if np.ndim(length_scale) == 0:
        xc = isotropic_deformation(metric_id, metric_reduce(metric_id, cd), length_scale)
        xp = isotropic_deformation(metric_id, metric_reduce(metric_id, pd), length_scale)
    else:
        xc = anisotropic_deformation(metric_id, cd, length_scale)
        xp = anisotropic_deformation(metric_id, pd, length_scale)
    return kernel_fn(kernel_id, xp), kernel_fn(kernel_id, xc)


def predict(
    kernel_id,
    metric_id,
    length_scale,
    noise,
    scale,
    train_x,
    train_y,
    query_x,
    query_idx,
    nn_idx,
):
    """regress_from_indices: (mean, scale*variance).  S/examples/from_indices.py:76-90.

    `noise` scalar -> homoscedastic; array (b,k) -> heteroscedastic.
    """
    Kin, Kcross = kernel_tensors(
        kernel_id, metric_id, length_scale, train_x, query_x, query_idx, nn_idx
    )
    if np.ndim(noise) == 0:
        pK = homoscedastic_perturb(Kin, noise)
    else:
        pK = heteroscedastic_perturb(Kin, noise)
    mean = posterior_mean(pK, Kcross, np.asarray(train_y)[nn_idx])
    var = scale * diagonal_variance(pK, Kcross, 1.0)
    return mean, var


def loo_objective(
    loss_id,
    kernel_id,
    metric_id,
    length_scale,
    noise,
    train_x,
    train_y,
    batch_idx,
    batch_nn_idx,
    analytic=True,
    fixed_scale=1.0,
    loss_kwargs=None,
    model_noise=None,
):
    """One `obj_fn(**theta)` evaluation = -loss.     S/optimize/objective.py:20-118,
    S/optimize/loss.py:26-178.  Returns (objective, scale_used).

    Quirk reproduced: the analytic scale inside the objective perturbs with the
    MODEL's stored nugget (`muygps.noise.perturb(Kin)`,
    S/gp/hyperparameter/scale.py:206-208), not with the `noise=` keyword the
    optimiser passes; mean and variance do honour the keyword
    (S/gp/noise/homoscedastic.py:112-113).  `model_noise` defaults to `noise`.
    """
    loss_kwargs = loss_kwargs or {}
    Kin, Kcross = kernel_tensors(
        kernel_id, metric_id, length_scale, train_x, train_x, batch_idx, batch_nn_idx
    )
    pK = homoscedastic_perturb(Kin, noise)
    y_nn = np.asarray(train_y)[batch_nn_idx]
    y_b = np.asarray(train_y)[batch_idx]
    mean = posterior_mean(pK, Kcross, y_nn)
    if loss_id == LOSS_MSE:
        return -mse(mean, y_b), None
    if loss_id == LOSS_PSEUDO_HUBER:
        return -pseudo_huber(mean, y_b, **loss_kwargs), None
    if loss_id == LOSS_CROSS_ENTROPY:
        return -cross_entropy(mean, y_b), None
    if analytic:
        sK = pK if model_noise is None else homoscedastic_perturb(Kin, model_noise)
        scale = analytic_scale(sK, y_nn)
    else:
        scale = fixed_scale
    var = diagonal_variance(pK, Kcross, 1.0)
    if loss_id == LOSS_LOOL:
        return -lool(mean, y_b, var, scale), scale
    if loss_id == LOSS_LOOPH:
        return -looph(mean, y_b, var, scale, **loss_kwargs), scale
    raise ValueError(f"unknown loss id {loss_id}")


# ---- batch filters (S/optimize/batch.py, S/examples/classify.py) -------------------------
def nonconstant_mask(labels, nn_indices):
    """Rows whose neighbours do not all carry the same label.  S/optimize/batch.py:104-110
    (1-D labels) and S/examples/classify.py:577-583 (column 0 of one-hot labels)."""
    lab = np.asarray(labels)
    col = lab if lab.ndim == 1 else lab[:, 0]
    nn_labels = col[nn_indices]
    return np.max(nn_labels, axis=1) != np.min(nn_labels, axis=1)


def full_filtered_batch(train, labels, k):
    """(batch_indices, batch_nn_indices).  S/optimize/batch.py:67-113."""
    indices = np.arange(len(labels))
    nn_indices, _ = knn_batch(train, indices, k)
    mask = nonconstant_mask(labels, nn_indices)
    return indices[mask], nn_indices[mask, :]


def classify_any(kernel_id, metric_id, length_scale, noise, train_x, train_labels, test_x, k):
    """Surrogate class scores.  S/examples/classify.py:536-608: rows whose neighbours agree
    take the nearest neighbour's one-hot row, the others the posterior mean."""
    nn, _ = knn_exact(train_x, test_x, k)
    mask = nonconstant_mask(train_labels, nn)
    pred = train_labels[nn[:, 0]].copy()
    rows = np.where(mask)[0]
    if len(rows):
        Kin, Kcross = kernel_tensors(kernel_id, metric_id, length_scale, train_x, test_x, rows,
                                     nn[rows])
        pred[rows] = posterior_mean(homoscedastic_perturb(Kin, noise), Kcross,
                                    train_labels[nn[rows]])
    return pred
