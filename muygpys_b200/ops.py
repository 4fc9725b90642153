"""Thin torch-tensor wrappers over the C ABI: device memory and streams only.

Every function takes/returns CUDA float64 / int64 tensors, launches on torch's
current stream and never synchronises.  There is no CPU path: tensors that are
not on a CUDA device raise.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import torch

from . import _lib as L

f64 = torch.float64
i64 = torch.int64


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """Handle of torch's current CUDA stream (the raw getter is ~10x cheaper than building a
    torch.cuda.Stream object, which matters for 10 k-row objective evaluations)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first CUDA tensor argument current, so that the
    launch, the stream handle and every allocation target the tensors' device even when the
    caller's current device is another GPU (one process may drive several)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)

    return wrapped


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise L.MgpError(
            f"{name} lives on {t.device}: muygpys_b200 kernels only run on CUDA tensors "
            "(no CPU fallback)"
        )


def fdev(t, name="tensor") -> torch.Tensor:
    """float64, contiguous, CUDA."""
    _need_cuda(t, name)
    if t.dtype != f64:
        t = t.to(f64)
    return t.contiguous()


def idev(t, name="indices") -> torch.Tensor:
    """int64, contiguous, CUDA."""
    _need_cuda(t, name)
    if t.dtype != i64:
        t = t.to(i64)
    return t.contiguous()


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _host_doubles(vals: Sequence[float]):
    arr = (C.c_double * len(vals))(*[float(v) for v in vals])
    return arr


def _ls_list(length_scale) -> list:
    if isinstance(length_scale, torch.Tensor):
        length_scale = length_scale.detach().cpu().reshape(-1).tolist()
    try:
        return [float(v) for v in length_scale]
    except TypeError:
        return [float(length_scale)]


def check_index_range(idx: Optional[torch.Tensor], upper: int, name: str) -> None:
    """ValueError for an index outside [0, upper) -- what numpy's fancy indexing raises in the
    reference (it also wraps negatives; here they are rejected).  One min/max reduction and a
    host read, so the hot entry points only do it when MGP_CHECK_INDICES=1; the one-time
    set-up of an objective always does."""
    if idx is None or idx.numel() == 0:
        return
    lo, hi = torch.aminmax(idx)
    lo, hi = int(lo), int(hi)
    if lo < 0 or hi >= upper:
        raise ValueError(f"{name} holds values in [{lo}, {hi}], valid rows are 0..{upper - 1}")


_CHECK_INDICES = __import__("os").environ.get("MGP_CHECK_INDICES") == "1"


def as_2d(x: torch.Tensor) -> torch.Tensor:
    """The reference treats 1-D feature arrays as (n,1) (S/neighbors.py:84-85)."""
    return x[:, None] if x.dim() == 1 else x


# --------------------------------------------------------------------------
# fused path
# --------------------------------------------------------------------------
def set_fused_variant(variant: int) -> None:
    """0 = auto, 1 = generic shared-memory kernel, 2 = register-tile DMMA kernel, 3 = the
    column-direct kernels (thread-per-tile where built), 4 = the column-direct kernel with
    lane-parallel column steps."""
    L.check(L.lib().mgp_set_fused_variant(int(variant)))


@_on_tensor_device
def fused_posterior(
    train_x: torch.Tensor,
    query_x: torch.Tensor,
    query_idx: Optional[torch.Tensor],
    nn_idx: torch.Tensor,
    train_y: Optional[torch.Tensor],
    *,
    kernel_id: int,
    metric_id: int,
    length_scale,
    noise: Union[float, torch.Tensor] = 0.0,
    scale: float = 1.0,
    want_mean: bool = True,
    want_var: bool = True,
    want_yky: bool = False,
    want_coeffs: bool = False,
    want_status: bool = False,
    out_mean: Optional[torch.Tensor] = None,
    out_var: Optional[torch.Tensor] = None,
    _host: Optional[tuple] = None,
):
    """One launch of K1 over a batch of neighbourhoods.  Returns a dict of tensors.

    `noise` is a python float (homoscedastic) or a (b,k) tensor (heteroscedastic).
    mean is (b,r), var (b,), yky (b,), coeffs (b,k,r), status (b,) int32.
    """
    lib = L.lib()
    train_x = as_2d(fdev(train_x, "train_features"))
    query_x = as_2d(fdev(query_x, "test_features"))
    nn_idx = idev(nn_idx, "nn_indices")
    if nn_idx.dim() != 2:
        raise ValueError(f"nn_indices must be (batch_count, nn_count), not {tuple(nn_idx.shape)}")
    b, k = nn_idx.shape
    n, d = train_x.shape
    t = query_x.shape[0]
    if query_x.shape[1] != d:
        raise ValueError(f"feature counts differ: {query_x.shape[1]} vs {d}")
    if query_idx is not None:
        query_idx = idev(query_idx, "indices")
        if query_idx.shape[0] != b:
            raise ValueError("indices and nn_indices disagree on batch_count")
    elif b > t:
        raise ValueError("more neighbourhood rows than query points")
    r = 1
    y2 = None
    if train_y is not None:
        y2 = fdev(train_y, "train_targets")
        y2 = y2[:, None] if y2.dim() == 1 else y2
        if y2.shape[0] != n:
            raise ValueError("train_targets and train_features disagree on train_count")
        y2 = y2.contiguous()
        r = y2.shape[1]
    dev = train_x.device
    if _CHECK_INDICES and _host is None:
        check_index_range(nn_idx, n, "nn_indices")
        check_index_range(query_idx, t, "indices")
    ls = _ls_list(length_scale)
    ls_host = _host_doubles(ls)
    noise_bk = None
    noise_val = 0.0
    if isinstance(noise, torch.Tensor) and noise.dim() > 0:
        noise_bk = fdev(noise, "noise")
        if tuple(noise_bk.shape) != (b, k):
            raise ValueError(f"heteroscedastic noise must be (batch_count, nn_count)={b, k}")
    else:
        noise_val = float(noise)
    out = {}
    if want_mean:
        if out_mean is not None:  # caller-provided (b,r) slice, e.g. of a pipelined batch
            assert out_mean.is_contiguous() and tuple(out_mean.shape) == (b, r)
        out["mean"] = out_mean if out_mean is not None else torch.empty((b, r), dtype=f64,
                                                                        device=dev)
    if want_var:
        if out_var is not None:
            assert out_var.is_contiguous() and tuple(out_var.shape) == (b,)
        out["var"] = out_var if out_var is not None else torch.empty((b,), dtype=f64, device=dev)
    if want_yky:
        out["yky"] = torch.empty((b,), dtype=f64, device=dev)
    if want_coeffs:
        out["coeffs"] = torch.empty((b, k, r), dtype=f64, device=dev)
    if want_status:
        out["status"] = torch.empty((b,), dtype=torch.int32, device=dev)
    p = L.MgpProblem(
        train_x=_p(train_x), query_x=_p(query_x), query_idx=_p(query_idx), nn_idx=_p(nn_idx),
        train_y=_p(y2), n=n, t=t, b=b, k=k, d=d, r=r, kernel_id=int(kernel_id),
        metric_id=int(metric_id), length_scale_count=len(ls),
        length_scale=C.cast(ls_host, C.POINTER(C.c_double)), noise=noise_val,
        noise_bk=_p(noise_bk), scale=float(scale), mean=_p(out.get("mean")),
        var=_p(out.get("var")), yky=_p(out.get("yky")), coeffs=_p(out.get("coeffs")),
        status=_p(out.get("status")),
    )
    ws = None
    ws_bytes = lib.mgp_fused_workspace_bytes(C.byref(p))
    if ws_bytes:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    if _host is not None:  # (nn_idx, query_idx, mean, var) CPU tensors; see fused_posterior_host
        hp = [None if h is None else C.c_void_p(h.data_ptr()) for h in _host[:4]]
        if _host[0].dtype == torch.int32:  # 32-bit indices: int32 device staging buffer last
            L.check(lib.mgp_fused_posterior_host32(C.byref(p), hp[0], _p(_host[4]), hp[1], hp[2],
                                                   hp[3], _p(ws), ws_bytes, _stream()))
        else:
            L.check(lib.mgp_fused_posterior_host(C.byref(p), hp[0], hp[1], hp[2], hp[3], _p(ws),
                                                 ws_bytes, _stream()))
        out["_host_buffers"] = _host  # keep them alive until the stream has been synchronised
        return out
    L.check(lib.mgp_fused_posterior(C.byref(p), _p(ws), ws_bytes, _stream()))
    return out


@_on_tensor_device
def fused_posterior_host(train_x, query_x, query_idx_host, nn_idx_host, train_y, *,
                         mean_host: Optional[torch.Tensor] = None,
                         var_host: Optional[torch.Tensor] = None, **kw):
    """K1 with the neighbour (and batch) indices in HOST memory: `mgp_fused_posterior_host`
    (`mgp_fused_posterior_host32` when the neighbour indices are int32: they cross the host link
    as they are and are widened on the device)
    uploads them chunk by chunk on two internal streams under the kernels of the previous
    chunks and, when `mean_host` / `var_host` (CPU float64 tensors, ideally pinned) are given,
    streams the results back the same way.  Returns the device tensors like fused_posterior;
    synchronise the current stream before reading the host outputs."""
    train_x = as_2d(fdev(train_x, "train_features"))
    dev = train_x.device

    def host_i64(a, name, keep32=False):
        t = torch.as_tensor(a)
        if t.is_cuda:
            raise TypeError(f"{name} must be a host array")
        if keep32 and t.dtype == torch.int32:  # travels as is: half the bytes on the host link
            return t.contiguous()
        return t.to(i64).contiguous()

    nn_h = host_i64(nn_idx_host, "nn_indices", keep32=True)
    if nn_h.dim() != 2:
        raise ValueError(f"nn_indices must be (batch_count, nn_count), not {tuple(nn_h.shape)}")
    idx_h = None if query_idx_host is None else host_i64(query_idx_host, "indices")
    for name, h in (("mean_host", mean_host), ("var_host", var_host)):
        if h is not None and (h.is_cuda or h.dtype != f64 or not h.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous host float64 tensor")
    nn_stage = torch.empty(nn_h.shape, dtype=i64, device=dev)
    idx_stage = None if idx_h is None else torch.empty(idx_h.shape, dtype=i64, device=dev)
    nn_stage32 = (torch.empty(nn_h.shape, dtype=torch.int32, device=dev)
                  if nn_h.dtype == torch.int32 else None)
    return fused_posterior(train_x, query_x, idx_stage, nn_stage, train_y,
                           _host=(nn_h, idx_h, mean_host, var_host, nn_stage32), **kw)


class FusedLoo:
    """One-launch leave-one-out objective evaluations over a fixed training batch.

    Wraps `mgp_fused_loo`: everything that does not change between evaluations (device tensors,
    the problem record, the zeroed workspace, the partials record and a page-locked copy of it)
    is set up once; `launch(length_scale, noise)` then costs one kernel launch, and `record()`
    one 64-byte device-to-host copy plus a stream synchronisation."""

    @_on_tensor_device
    def __init__(self, train_x, train_y, batch_idx, nn_idx, *, kernel_id, metric_id, loss_id,
                 boundary_scale=1.0, partials: Optional[torch.Tensor] = None,
                 want_grad: bool = False):
        lib = L.lib()
        self.x = as_2d(fdev(train_x, "train_features"))
        y = fdev(train_y, "train_targets")
        self.y = (y[:, None] if y.dim() == 1 else y).contiguous()
        self.bi = idev(batch_idx, "batch_indices")
        self.nn = idev(nn_idx, "batch_nn_indices")
        n, d = self.x.shape
        b, k = self.nn.shape
        if self.y.shape[1] != 1:
            raise NotImplementedError("mgp_fused_loo handles one response (r == 1)")
        check_index_range(self.nn, n, "batch_nn_indices")
        check_index_range(self.bi, n, "batch_indices")
        dev = self.x.device
        self.ls_host = (C.c_double * max(d, 1))()
        # The record is written by the kernel's last block.  By default it lives in page-locked
        # HOST memory (device-accessible under unified addressing): the 64 bytes cross the link
        # as the kernel's own stores, no copy is enqueued and the host reads them as soon as
        # the stream is synchronised.
        self.pinned = torch.zeros((L.MGP_PARTIALS,), dtype=f64).pin_memory()
        self.partials = partials if partials is not None else self.pinned
        self.p = L.MgpProblem(
            train_x=_p(self.x), query_x=_p(self.x), query_idx=_p(self.bi), nn_idx=_p(self.nn),
            train_y=_p(self.y), n=n, t=n, b=b, k=k, d=d, r=1, kernel_id=int(kernel_id),
            metric_id=int(metric_id), length_scale_count=1,
            length_scale=C.cast(self.ls_host, C.POINTER(C.c_double)), noise=0.0, noise_bk=None,
            scale=1.0, mean=None, var=None, yky=None, coeffs=None, status=None)
        self.ws_bytes = lib.mgp_fused_loo_workspace_bytes(C.byref(self.p))
        self.ws = torch.zeros((self.ws_bytes,), dtype=torch.uint8, device=dev)  # counter = 0
        self.loss_id = int(loss_id)
        self.boundary_scale = float(boundary_scale)
        self._lib = lib
        self._fn = lib.mgp_fused_loo_grad
        # gradient sums (MGP_GRAD_DOUBLES), written by the same kernel into pinned host memory
        self.grad = torch.zeros((L.MGP_GRAD_DOUBLES,), dtype=f64).pin_memory() if want_grad \
            else None
        self._grad_ptr = _p(self.grad)
        self._pref = C.byref(self.p)
        self._out = _p(self.partials)
        self._ws = _p(self.ws)
        self._np = self.pinned.numpy()
        self._stream_obj = None

    def launch(self, length_scale, noise: float, peers=None,
               scale: Optional[float] = None) -> torch.Tensor:
        """Enqueue one evaluation on the current stream; returns the device partials record.
        `peers` (a distributed.PeerChannel): the record is summed across the GPUs of the NVLink
        domain inside the same kernel (`mgp_fused_loo_peers`).  `scale`: the variance scale
        sigma^2 the looph epilogue needs (MGP_LOSS_LOOPH reads it from p->scale)."""
        if self.x.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.x.device):
                return self.launch(length_scale, noise, peers, scale)
        if scale is not None:
            self.p.scale = float(scale)
        if isinstance(length_scale, float):
            self.ls_host[0] = length_scale
            self.p.length_scale_count = 1
        else:
            ls = _ls_list(length_scale)
            for i, v in enumerate(ls):
                self.ls_host[i] = v
            self.p.length_scale_count = len(ls)
        self.p.noise = float(noise)
        g = None if peers is None else C.byref(peers.group_struct())
        rc = self._fn(self._pref, self.loss_id, self.boundary_scale, self._out, self._grad_ptr,
                      self._ws, self.ws_bytes, g, _stream())
        if rc != 0:
            L.check(rc)
        return self.partials

    def record(self, device_record: Optional[torch.Tensor] = None):
        """Host copy (numpy, 8 doubles) of the partials record after a stream synchronise."""
        src = self.partials if device_record is None else device_record
        if src is not self.pinned:
            self.pinned.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._np.copy()


def fused_loo_supported(d: int, k: int, r: int, kernel_id: int, metric_id: int,
                        heteroscedastic: bool, grad: bool = False) -> bool:
    """Shapes `mgp_fused_loo` / `mgp_fused_loo_grad` take (mirrors col_shape_ok in
    csrc/fused_col.cu): k = 7..102 (thread-per-tile kernel, up to 13 tile rows), with or without
    the analytic gradient (GRAD instantiations: back substitution on the stored factor)."""
    if r != 1 or d > 3 or heteroscedastic or not 7 <= k <= 102:
        return False
    if metric_id == L.METRIC_L2:
        return kernel_id in (L.KERNEL_MATERN_05, L.KERNEL_MATERN_15, L.KERNEL_MATERN_25,
                             L.KERNEL_MATERN_INF)
    return kernel_id == L.KERNEL_RBF


@_on_tensor_device
def fast_mean(train_x, query_x, query_idx, nn_idx, coeff_row, coeffs, *, kernel_id, metric_id,
              length_scale) -> torch.Tensor:
    """K4: mean (b,r) = sum_j kernel(|q - x_nn_j|) * coeffs[coeff_row, j, :]."""
    lib = L.lib()
    train_x = as_2d(fdev(train_x))
    query_x = as_2d(fdev(query_x))
    nn_idx = idev(nn_idx)
    b, k = nn_idx.shape
    n, d = train_x.shape
    coeffs = fdev(coeffs, "coeffs_tensor")
    if coeffs.dim() == 2:
        coeffs = coeffs[:, :, None].contiguous()
    if coeffs.shape[1] != k:
        raise ValueError("coeffs_tensor and nn_indices disagree on nn_count")
    r = coeffs.shape[2]
    query_idx = None if query_idx is None else idev(query_idx)
    coeff_row = None if coeff_row is None else idev(coeff_row)
    ls = _ls_list(length_scale)
    ls_host = _host_doubles(ls)
    mean = torch.empty((b, r), dtype=f64, device=train_x.device)
    p = L.MgpProblem(
        train_x=_p(train_x), query_x=_p(query_x), query_idx=_p(query_idx), nn_idx=_p(nn_idx),
        train_y=None, n=n, t=query_x.shape[0], b=b, k=k, d=d, r=r, kernel_id=int(kernel_id),
        metric_id=int(metric_id), length_scale_count=len(ls),
        length_scale=C.cast(ls_host, C.POINTER(C.c_double)), noise=0.0, noise_bk=None, scale=1.0,
        mean=_p(mean), var=None, yky=None, coeffs=None, status=None,
    )
    L.check(lib.mgp_fast_mean(C.byref(p), _p(coeff_row), _p(coeffs), _stream()))
    return mean


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
@_on_tensor_device
def loss_partials(loss_id, pred, targets, var=None, yky=None, scale_dev=None,
                  boundary_scale=1.0, partials=None) -> torch.Tensor:
    """Accumulate an MGP_PARTIALS record (device tensor of 8 doubles)."""
    lib = L.lib()
    pred = fdev(pred, "predictions")
    b = pred.shape[0]
    r = 1 if pred.dim() == 1 else pred.shape[1]
    dev = pred.device
    if targets is not None:
        targets = fdev(targets, "targets")
        if targets.numel() != pred.numel():
            raise ValueError("predictions and targets differ in size")
    var = None if var is None else fdev(var, "variances")
    yky = None if yky is None else fdev(yky)
    scale_dev = None if scale_dev is None else fdev(scale_dev)
    if partials is None:
        partials = torch.zeros((L.MGP_PARTIALS,), dtype=f64, device=dev)
    ws_bytes = lib.mgp_loss_workspace_bytes(b, r)
    ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
    L.check(lib.mgp_loss_partials(int(loss_id), _p(pred), _p(targets), _p(var), _p(yky),
                                  _p(scale_dev), float(boundary_scale), b, r, _p(partials),
                                  _p(ws), ws_bytes, _stream()))
    return partials


# --------------------------------------------------------------------------
# KNN
# --------------------------------------------------------------------------
@_on_tensor_device
def knn(train, queries, k, self_idx=None):
    lib = L.lib()
    train = as_2d(fdev(train, "train"))
    queries = as_2d(fdev(queries, "queries"))
    n, d = train.shape
    q = queries.shape[0]
    if queries.shape[1] != d:
        raise ValueError(f"query feature count {queries.shape[1]} != train feature count {d}")
    self_idx = None if self_idx is None else idev(self_idx)
    out_idx = torch.empty((q, k), dtype=i64, device=train.device)
    out_d2 = torch.empty((q, k), dtype=f64, device=train.device)
    ws_bytes = lib.mgp_knn_workspace_bytes(n, q, d, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=train.device) if ws_bytes else None
    L.check(lib.mgp_knn(_p(train), n, _p(queries), q, d, int(k), 0 if self_idx is None else 1,
                        _p(self_idx), _p(out_idx), _p(out_d2), _p(ws), ws_bytes, _stream()))
    return out_idx, out_d2


class KnnGrid:
    """Uniform cell grid over a low-dimensional training set (d <= 3) for exact KNN.

    Built once per training set: cell ids from the CUDA kernel, a device radix sort
    (torch.sort, plumbing) to bucket the points, and the per-cell offsets."""

    POINTS_PER_CELL = 6.0  # measured: 4-8 is best for k = 10..100 (tools/sweep_grid_density.py)
    MAX_CELLS = 1 << 27

    @_on_tensor_device
    def __init__(self, train: torch.Tensor):
        lib = L.lib()
        train = as_2d(fdev(train, "train"))
        n, d = train.shape
        if not 1 <= d <= 3:
            raise NotImplementedError("the grid index supports 1 <= d <= 3")
        self.n, self.d = n, d
        lo = train.min(dim=0).values
        hi = train.max(dim=0).values
        extent = (hi - lo).clamp_min(0.0).cpu().tolist()
        cells = min(max(n / self.POINTS_PER_CELL, 1.0), float(self.MAX_CELLS))
        # cell edge from the volume of the axes that are wider than a cell; an axis that is
        # (nearly) degenerate -- a constant feature, or one with a tiny extent next to the
        # others -- collapses to a single layer instead of blowing the other axes' cell counts up
        span = sorted((e for e in extent if e > 0.0), reverse=True)
        h = 1.0
        while span:
            volume = 1.0
            for e in span:
                volume *= e
            h = (volume / cells) ** (1.0 / len(span))
            if span[-1] >= h:
                break
            span.pop()  # thinner than a cell: treat as collapsed and recompute
        self.h = float(h)
        self.dims = [max(1, int(e / self.h) + 1) for e in extent]
        ncells = 1
        for v in self.dims:
            ncells *= v
        while ncells > self.MAX_CELLS:  # rounding up per axis can overshoot the cap
            self.h *= 1.1
            self.dims = [max(1, int(e / self.h) + 1) for e in extent]
            ncells = 1
            for v in self.dims:
                ncells *= v
        if max(self.dims) >= 2 ** 31 or ncells >= 2 ** 31:
            raise ValueError(f"grid of {self.dims} cells does not fit 32-bit cell ids")
        self.origin = [float(v) for v in lo.cpu().tolist()]
        self._dims_c = (C.c_int32 * d)(*self.dims)
        self._origin_c = (C.c_double * d)(*self.origin)
        ncells = 1
        for v in self.dims:
            ncells *= v
        cell = torch.empty((n,), dtype=torch.int32, device=train.device)
        L.check(lib.mgp_knn_grid_cells(_p(train), n, d, self._dims_c, self._origin_c, self.h,
                                       _p(cell), _stream()))
        sorted_cell, perm = torch.sort(cell)
        self.sorted_points = train[perm].contiguous()
        self.sorted_ids = perm.to(torch.int32).contiguous()
        edges = torch.arange(ncells + 1, dtype=torch.int32, device=train.device)
        self.cell_start = torch.searchsorted(sorted_cell, edges).to(torch.int32).contiguous()

    @_on_tensor_device
    def query(self, queries: torch.Tensor, k: int, self_idx=None):
        lib = L.lib()
        queries = as_2d(fdev(queries, "queries"))
        q, d = queries.shape
        if d != self.d:
            raise ValueError(f"query feature count {d} != train feature count {self.d}")
        dev = queries.device
        out_idx = torch.empty((q, k), dtype=i64, device=dev)
        out_d2 = torch.empty((q, k), dtype=f64, device=dev)
        order = None
        if q > 1:
            qcell = torch.empty((q,), dtype=torch.int32, device=dev)
            L.check(lib.mgp_knn_grid_cells(_p(queries), q, d, self._dims_c, self._origin_c,
                                           self.h, _p(qcell), _stream()))
            order = torch.argsort(qcell).to(torch.int32).contiguous()
        self_idx = None if self_idx is None else idev(self_idx)
        L.check(lib.mgp_knn_grid_query(_p(self.sorted_points), _p(self.sorted_ids),
                                       _p(self.cell_start), self.n, d, self._dims_c,
                                       self._origin_c, self.h, _p(queries), _p(order), q, int(k),
                                       _p(self_idx), _p(out_idx), _p(out_d2), _stream()))
        return out_idx, out_d2


# --------------------------------------------------------------------------
# staged ops
# --------------------------------------------------------------------------
@_on_tensor_device
def crosswise_diffs(data, nn_data, data_idx, nn_idx) -> torch.Tensor:
    lib = L.lib()
    data, nn_data = as_2d(fdev(data)), as_2d(fdev(nn_data))
    nn_idx = idev(nn_idx)
    data_idx = None if data_idx is None else idev(data_idx)
    b, k = nn_idx.shape
    d = data.shape[1]
    out = torch.empty((b, k, d), dtype=f64, device=data.device)
    L.check(lib.mgp_crosswise_diffs(_p(data), _p(nn_data), _p(data_idx), _p(nn_idx), b, k, d,
                                    _p(out), _stream()))
    return out


@_on_tensor_device
def pairwise_diffs(data, nn_idx) -> torch.Tensor:
    lib = L.lib()
    data = as_2d(fdev(data))
    nn_idx = idev(nn_idx)
    b, k = nn_idx.shape
    d = data.shape[1]
    out = torch.empty((b, k, k, d), dtype=f64, device=data.device)
    L.check(lib.mgp_pairwise_diffs(_p(data), _p(nn_idx), b, k, d, _p(out), _stream()))
    return out


@_on_tensor_device
def metric_reduce(metric_id, diffs, length_scale=None) -> torch.Tensor:
    lib = L.lib()
    diffs = fdev(diffs, "diffs")
    d = diffs.shape[-1]
    rows = diffs.numel() // d if d else 0
    out = torch.empty(diffs.shape[:-1], dtype=f64, device=diffs.device)
    ls_host = None
    if length_scale is not None:
        ls = _ls_list(length_scale)
        if len(ls) != d:
            raise ValueError(
                f"Difference tensor of shape {tuple(diffs.shape)} must have final dimension "
                f"size of {len(ls)}"
            )
        ls_host = C.cast(_host_doubles(ls), C.POINTER(C.c_double))
    L.check(lib.mgp_metric_reduce(int(metric_id), _p(diffs), rows, d, ls_host, _p(out),
                                  _stream()))
    return out


@_on_tensor_device
def crosswise_dists(metric_id, data, nn_data, data_idx, nn_idx) -> torch.Tensor:
    lib = L.lib()
    data, nn_data = as_2d(fdev(data)), as_2d(fdev(nn_data))
    nn_idx = idev(nn_idx)
    data_idx = None if data_idx is None else idev(data_idx)
    b, k = nn_idx.shape
    out = torch.empty((b, k), dtype=f64, device=data.device)
    L.check(lib.mgp_crosswise_dists(int(metric_id), _p(data), _p(nn_data), _p(data_idx),
                                    _p(nn_idx), b, k, data.shape[1], _p(out), _stream()))
    return out


@_on_tensor_device
def pairwise_dists(metric_id, data, nn_idx) -> torch.Tensor:
    lib = L.lib()
    data = as_2d(fdev(data))
    nn_idx = idev(nn_idx)
    b, k = nn_idx.shape
    out = torch.empty((b, k, k), dtype=f64, device=data.device)
    L.check(lib.mgp_pairwise_dists(int(metric_id), _p(data), _p(nn_idx), b, k, data.shape[1],
                                   _p(out), _stream()))
    return out


@_on_tensor_device
def kernel_apply(kernel_id, x, pre_scale=1.0) -> torch.Tensor:
    lib = L.lib()
    x = fdev(x, "dists")
    out = torch.empty_like(x)
    L.check(lib.mgp_kernel_apply(int(kernel_id), _p(x), float(pre_scale), x.numel(), _p(out),
                                 _stream()))
    return out


@_on_tensor_device
def perturb(Kin, noise) -> torch.Tensor:
    lib = L.lib()
    Kin = fdev(Kin, "Kin")
    if Kin.dim() != 3 or Kin.shape[1] != Kin.shape[2]:
        raise ValueError(
            f"homoscedastic perturbation is not implemented for tensors of shape "
            f"{tuple(Kin.shape)}"
        )
    b, k, _ = Kin.shape
    out = torch.empty_like(Kin)
    noise_bk = None
    noise_val = 0.0
    if isinstance(noise, torch.Tensor) and noise.dim() > 0:
        noise_bk = fdev(noise).reshape(b, k).contiguous()
    else:
        noise_val = float(noise)
    L.check(lib.mgp_perturb(_p(Kin), b, k, noise_val, _p(noise_bk), _p(out), _stream()))
    return out


@_on_tensor_device
def solve(Kin, Kcross=None, Y=None, kout=1.0, *, want_mean=False, want_var=False,
          want_yky=False, want_coeffs=False):
    """Batched SPD solve on materialised tensors (K5).  Y is (b,k) or (b,k,r)."""
    lib = L.lib()
    Kin = fdev(Kin, "Kin")
    if Kin.dim() != 3 or Kin.shape[1] != Kin.shape[2]:
        raise ValueError(f"Kin must be (batch_count, nn_count, nn_count), not {tuple(Kin.shape)}")
    b, k, _ = Kin.shape
    dev = Kin.device
    r = 0
    if Kcross is not None:
        Kcross = fdev(Kcross, "Kcross").reshape(b, k).contiguous()
    if Y is not None:
        Y = fdev(Y, "nn_targets")
        Y = Y.reshape(b, k, -1).contiguous()
        r = Y.shape[2]
    out = {}
    if want_mean:
        out["mean"] = torch.empty((b, r), dtype=f64, device=dev)
    if want_var:
        out["var"] = torch.empty((b,), dtype=f64, device=dev)
    if want_yky:
        out["yky"] = torch.empty((b,), dtype=f64, device=dev)
    if want_coeffs:
        out["coeffs"] = torch.empty((b, k, r), dtype=f64, device=dev)
    ws_bytes = lib.mgp_solve_workspace_bytes(b, k, r)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
    L.check(lib.mgp_solve(_p(Kin), _p(Kcross), _p(Y), b, k, r, float(kout), _p(out.get("mean")),
                          _p(out.get("var")), _p(out.get("yky")), _p(out.get("coeffs")), None,
                          _p(ws), ws_bytes, _stream()))
    return out


@_on_tensor_device
def nn_label_mask(labels, nn_idx) -> torch.Tensor:
    """bool (b,): the labels of row i's neighbours are not all equal.  `labels` is (n,) or a
    one-hot style (n, class_count) matrix whose column 0 is inspected (classify.py:577-583)."""
    lib = L.lib()
    labels = fdev(labels, "labels")
    nn_idx = idev(nn_idx, "nn_indices")
    b, k = nn_idx.shape
    stride = 1 if labels.dim() == 1 else labels.shape[1]
    mask = torch.empty((b,), dtype=torch.uint8, device=labels.device)
    L.check(lib.mgp_nn_label_mask(_p(labels), stride, _p(nn_idx), b, k, _p(mask), _stream()))
    return mask.bool()


@_on_tensor_device
def rowdot(Kcross, coeffs) -> torch.Tensor:
    lib = L.lib()
    Kcross = fdev(Kcross, "Kcross")
    coeffs = fdev(coeffs, "coeffs_tensor")
    b, k = Kcross.shape
    coeffs = coeffs.reshape(b, k, -1).contiguous()
    r = coeffs.shape[2]
    out = torch.empty((b, r), dtype=f64, device=Kcross.device)
    L.check(lib.mgp_rowdot(_p(Kcross), _p(coeffs), b, k, r, _p(out), _stream()))
    return out


def fp64_probe(mode: int, blocks: int, threads: int, iters: int) -> torch.Tensor:
    lib = L.lib()
    sink = torch.empty((blocks * threads,), dtype=f64, device="cuda")
    L.check(lib.mgp_fp64_probe(mode, blocks, threads, iters, _p(sink), _stream()))
    return sink
