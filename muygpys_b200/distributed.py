"""Data parallelism over neighbourhoods: one process per GPU, NCCL over NVLink.

The reference's MPI backend computes whole tensors on rank 0, scatters row
chunks with pickled point-to-point messages and all-reduces python scalars
(S/_src/mpi_utils.py:36-159, S/_src/optimize/loss/mpi.py, scale/mpi.py).  Here
the training set is replicated in every GPU's HBM, each rank slices its own row
range with the reference's chunk rule (no scatter at all), and the only
collective on the path is ONE SUM all-reduce of an 8-double partials record per
objective evaluation (two for looph, whose loss is nonlinear in sigma^2).
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def get_chunk_sizes(count: int, size: int) -> List[int]:
    """floor(count/size) rows per rank, the remainder going to the LAST ranks
    (S/_src/mpi_utils.py:36-41)."""
    base = int(count / size)
    extra = count - base * size
    return [base + 1 if i >= size - extra else base for i in range(size)]


def rank_and_size(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def local_range(count: int, group=None, rank: Optional[int] = None,
                size: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of this rank's rows among `count` batch/test rows."""
    if rank is None or size is None:
        rank, size = rank_and_size(group)
    sizes = get_chunk_sizes(count, size)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


def shard_rows(*tensors, group=None):
    """Slice the leading axis of every tensor to this rank's chunk (replaces the
    rank-0 scatter of `@mpi_chunk`, S/_src/mpi_utils.py:99-115)."""
    lo, hi = local_range(tensors[0].shape[0], group)
    out = tuple(t[lo:hi] for t in tensors)
    return out if len(out) > 1 else out[0]


def allreduce_partials(record: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of a partials record in place (no-op for a single process).

    Every slot is a plain sum (include/muygpys_b200.h, MGP_P_*), so this single
    call replaces the 1-2 scalar allreduces each reference loss performs."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(record, op=dist.ReduceOp.SUM, group=group)
    return record


class PartialsReducer:
    """Cross-rank SUM of 8-double partials records, delivered to the host.

    One objective evaluation ends in exactly one of these reductions (two for looph and for
    the analytic-scale nugget quirk).  `slot()` hands out the device record a kernel should
    write into; `sum_to_host(record)` returns the summed record as numpy on every rank.  The
    reduction is a plain `all_reduce(SUM)` on the process group (NCCL over NVLink on the GPU
    box, gloo in the CPU tests); the copy to the host goes through one page-locked buffer."""

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self._pin = None
        if self.device.type == "cuda":
            self._pin = torch.empty((8,), dtype=torch.float64).pin_memory()

    def slot(self) -> torch.Tensor:
        return torch.zeros((8,), dtype=torch.float64, device=self.device)

    def sum_to_host(self, record: torch.Tensor):
        allreduce_partials(record, self.group)
        if self._pin is None:
            return record.detach().cpu().numpy().copy()
        self._pin.copy_(record, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pin.numpy().copy()


def allgather_rows(local: torch.Tensor, count: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row chunks (uneven sizes allowed) in rank order --
    the analogue of `_consistent_unchunk_tensor` (S/_src/mpi_utils.py:118-143)."""
    rank, size = rank_and_size(group)
    if size == 1:
        return local
    sizes = get_chunk_sizes(count, size)
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(size)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
