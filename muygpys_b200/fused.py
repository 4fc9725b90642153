"""The fused (indices in, posterior out) calls over ANY MuyGPS object.

Each function takes either a genuine `MuyGPyS.gp.MuyGPS` or this package's mirror object
(`adapt.ModelSpec.of`) and runs the whole neighbourhood pipeline -- gather, distances,
covariances, nugget, factorisation, posterior -- in one K1 launch (`mgp_fused_posterior` /
`mgp_fused_posterior_host`).  Nothing of size (b,k,k[,d]) exists at any point.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._arrays import fdev, idev, is_host, like_input
from .adapt import ModelSpec


def _noise_arg(spec: ModelSpec, theta: dict, rows: int):
    noise = spec.noise(theta.get("noise"))
    if spec.heteroscedastic:
        noise = fdev(noise)
        if noise.dim() == 3 and noise.shape[-1] == 1:
            noise = noise[:, :, 0]
        if noise.shape[0] != rows:
            raise ValueError(
                f"heteroscedastic noise has {noise.shape[0]} rows, the batch has {rows}")
    return noise


def fused_call(muygps, indices, nn_indices, test_features, train_features, train_targets, *,
               theta: Optional[dict] = None, scale: Optional[float] = None, **want):
    """One K1 launch; returns the dict of device tensors of `ops.fused_posterior`."""
    spec = ModelSpec.of(muygps)
    theta = theta or {}
    nn = idev(nn_indices)
    if test_features is None:
        test_features = train_features
    return ops.fused_posterior(
        fdev(train_features), fdev(test_features),
        None if indices is None else idev(indices), nn,
        None if train_targets is None else fdev(train_targets),
        kernel_id=spec.kernel_id, metric_id=spec.metric_id,
        length_scale=spec.length_scale_arg(**theta), noise=_noise_arg(spec, theta, nn.shape[0]),
        scale=spec.scale() if scale is None else scale, **want)


def _fused_pipelined(spec: ModelSpec, indices, nn_indices, test_features, train_features,
                     train_targets, want_mean, want_var):
    """Host-resident index batches: `mgp_fused_posterior_host` uploads chunk i+1 on a side
    stream while chunk i is in the fused kernel (csrc/pipeline.cu)."""
    x, y = fdev(train_features), fdev(train_targets)
    q_src = test_features if test_features is not None else train_features
    q_h = torch.as_tensor(q_src)
    q_dev = q_h if q_h.is_cuda else q_h.to(x.device, non_blocking=True)
    out = ops.fused_posterior_host(
        x, q_dev, indices, nn_indices, y, kernel_id=spec.kernel_id, metric_id=spec.metric_id,
        length_scale=spec.length_scale_arg(), noise=spec.noise(None), scale=spec.scale(),
        want_mean=want_mean, want_var=want_var)
    # No host synchronisation here: the pipeline is joined back into the current stream, so
    # whatever the caller enqueues next (a device-to-host copy of the results, the next batch)
    # is ordered behind it, and the host is free to prepare the next call meanwhile.  The staged
    # HOST tensors must outlive the asynchronous uploads: fused_regress ties them to the result
    # tensors it returns (whoever reads the results synchronises first).
    return {"mean": out.get("mean"), "var": out.get("var"),
            "_keepalive": out.get("_host_buffers")}


def _squeeze_response(mean: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    return mean[:, 0] if targets.dim() == 1 else mean


def fused_regress(muygps, indices, nn_indices, test_features, train_features, train_targets,
                  want_mean=True, want_var=True):
    """Posterior mean and scaled variance straight from indices
    (S/examples/from_indices.py:22-90 in one launch, or a copy/compute pipeline when the index
    batch still lives in host memory)."""
    spec = ModelSpec.of(muygps)
    if (is_host(nn_indices) and not is_host(train_features) and not is_host(train_targets)
            and not spec.heteroscedastic and len(nn_indices) >= 16384
            and (indices is None or (test_features is not None and is_host(indices)))):
        out = _fused_pipelined(spec, indices, nn_indices, test_features, train_features,
                               train_targets, want_mean, want_var)
    else:
        out = fused_call(spec, indices, nn_indices, test_features, train_features,
                         train_targets, want_mean=want_mean, want_var=want_var)
    host = (indices, nn_indices, test_features, train_features, train_targets)
    res = []
    if want_mean:
        res.append(like_input(_squeeze_response(out["mean"], fdev(train_targets)), *host))
    if want_var:
        res.append(like_input(out["var"], *host))
    keep = out.get("_keepalive")
    if keep is not None:
        for t in res:
            if isinstance(t, torch.Tensor):  # (numpy results were copied out: already synced)
                t._mgp_keepalive = keep
    return tuple(res) if len(res) > 1 else res[0]


def fused_fast_coefficients(muygps, nn_indices_fast, train_features, train_targets):
    """(K+eps)^-1 Y for every row of `nn_indices_fast` (K3), never building Kin
    (S/gp/muygps.py fast_coefficients + S/_src/gp/muygps/numpy.py:88-95)."""
    out = fused_call(muygps, None, nn_indices_fast, train_features, train_features,
                     train_targets, want_mean=False, want_var=False, want_coeffs=True)["coeffs"]
    y = fdev(train_targets)
    out = out[:, :, 0] if y.dim() == 1 else out
    return like_input(out, nn_indices_fast, train_features, train_targets)


def fused_optimize_scale(muygps, batch_indices, batch_nn_indices, train_features,
                         train_targets):
    """`MuyGPS.optimize_scale` (S/gp/muygps.py:373-403) without the pairwise tensor: the fused
    kernel's y^T K^-1 y output is all the analytic scale needs.  Sets the scale in place."""
    spec = ModelSpec.of(muygps)
    if spec.analytic:
        out = fused_call(spec, batch_indices, batch_nn_indices, train_features, train_features,
                         train_targets, want_mean=False, want_var=False, want_yky=True)
        b, k = idev(batch_nn_indices).shape
        spec.set_scale(spec.sigma_from_mean_quadratic_form(float(out["yky"].sum()) / (b * k)))
    return muygps


# ---- MultivariateMuyGPS (S/gp/multivariate_muygps.py:99-340): one model per response --------
def is_multivariate(muygps) -> bool:
    return hasattr(muygps, "models") and not hasattr(muygps, "kernel")


def _column(targets, i):
    t = fdev(targets)
    return t[:, i].contiguous()


def _model_specs(mmuygps):
    """One spec per response, each seen through the tensors of models[0]'s deformation (the
    reference builds ONE pairwise / crosswise tensor with models[0] and hands it to every
    model's kernel, multivariate_muygps.py:136-139; from_indices.py:111-113)."""
    specs = [ModelSpec.of(m) for m in mmuygps.models]
    if len({s.anisotropic for s in specs}) != 1:
        raise NotImplementedError("MultivariateMuyGPS mixing Isotropy and Anisotropy models")
    return [s.seen_through(specs[0].metric_id) for s in specs]


def mm_fused_regress(mmuygps, indices, nn_indices, test_features, train_features, train_targets,
                     want_mean=True, want_var=True):
    """(b,r) posterior means and (b,r) variances of r independent models over the SAME
    neighbourhoods: one K1 launch per response on the shared index arrays.  Reproduces the
    reference's variance quirk: `posterior_variance` already carries the model's scale and
    MultivariateMuyGPS multiplies by it again (multivariate_muygps.py:183-192), i.e. scale^2."""
    nn = idev(nn_indices)
    idx = None if indices is None else idev(indices)
    x = fdev(train_features)
    q = x if test_features is None else fdev(test_features)
    means, variances = [], []
    for i, spec in enumerate(_model_specs(mmuygps)):
        out = fused_call(spec, idx, nn, q, x, _column(train_targets, i), scale=spec.scale() ** 2,
                         want_mean=want_mean, want_var=want_var)
        if want_mean:
            means.append(out["mean"][:, 0])
        if want_var:
            variances.append(out["var"])
    host = (indices, nn_indices, test_features, train_features, train_targets)
    res = []
    if want_mean:
        res.append(like_input(torch.stack(means, dim=1), *host))
    if want_var:
        res.append(like_input(torch.stack(variances, dim=1), *host))
    return tuple(res) if len(res) > 1 else res[0]


def mm_fused_fast_coefficients(mmuygps, nn_indices_fast, train_features, train_targets):
    """(n,k,r) fast-mean coefficients, one K3 launch per response.  The reference perturbs Kin
    before handing it to `fast_coefficients`, which perturbs again
    (multivariate_muygps.py:224-231): the nugget enters twice, and so it does here.  (The
    reference then DISCARDS each column -- `mm.assign` returns a copy that is never bound -- and
    returns zeros; we return the columns it computed.)"""
    cols = []
    for i, spec in enumerate(_model_specs(mmuygps)):
        noise = spec.noise(None)
        theta = {} if spec.heteroscedastic else {"noise": 2.0 * noise}
        if spec.heteroscedastic:
            raise NotImplementedError("multivariate fast coefficients with heteroscedastic noise")
        out = fused_call(spec, None, nn_indices_fast, train_features, train_features,
                         _column(train_targets, i), theta=theta, want_mean=False,
                         want_var=False, want_coeffs=True)["coeffs"]
        cols.append(out[:, :, 0])
    return like_input(torch.stack(cols, dim=2), nn_indices_fast, train_features, train_targets)


def mm_fast_posterior_mean(mmuygps, indices, nn_indices, test_features, train_features,
                           closest_index, coeffs_tensor):
    """_mmuygps_fast_posterior_mean (S/_src/gp/muygps/numpy.py:80-85): per response, the
    crosswise covariances of ITS kernel dotted with its coefficient column.  (As with the
    coefficients, the reference's wrapper never binds the `mm.assign` results and feeds zeros
    to this einsum, multivariate_muygps.py:262-270; we compute what it defines.)"""
    coeffs = fdev(coeffs_tensor)
    cols = []
    for i, spec in enumerate(_model_specs(mmuygps)):
        out = ops.fast_mean(fdev(train_features), fdev(test_features), idev(indices),
                            idev(nn_indices), idev(closest_index),
                            coeffs[:, :, i].contiguous(), kernel_id=spec.kernel_id,
                            metric_id=spec.metric_id, length_scale=spec.length_scale_arg())
        cols.append(out[:, 0])
    return like_input(torch.stack(cols, dim=1), indices, nn_indices, test_features,
                      train_features, coeffs_tensor)
