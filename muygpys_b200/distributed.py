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


class PeerChannel:
    """One slot of the peer-mapped exchange buffer: who we are, where every rank's copy of the
    slot is mapped, and the epoch counter of `mgp_peer_group` (starts at 1, +1 per call)."""

    def __init__(self, rank: int, world: int, ptrs: List[int]):
        self.rank, self.world, self.ptrs = rank, world, ptrs
        self.epoch = 0

    def group_struct(self):
        from . import _lib as L

        self.epoch += 1
        g = L.MgpPeerGroup(rank=self.rank, world=self.world, epoch=self.epoch)
        for i, p in enumerate(self.ptrs):
            g.peer_buf[i] = p
        return g


class _PeerBuffers:
    """Symmetric (peer-mapped) memory shared by all objectives of a process group: `CHANNELS`
    exchange slots of mgp_peer_buffer_bytes() each, handed out round-robin in program order
    (every rank creates its objectives in the same order)."""

    CHANNELS = 64
    _cache = {}

    def __init__(self, device, group):
        import torch.distributed._symmetric_memory as symm

        from . import _lib as L

        self.rank, self.world = rank_and_size(group)
        self.stride = int(L.lib().mgp_peer_buffer_bytes())
        assert self.stride % 8 == 0
        self.buf = symm.empty((self.CHANNELS * self.stride // 8,), dtype=torch.float64,
                              device=device)
        self.buf.zero_()
        hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        self._hdl = hdl
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's buffer is zeroed before anyone pushes into it
        self.channels = [None] * self.CHANNELS
        self.next = 0

    @classmethod
    def get(cls, device, group):
        key = (torch.device(device).index, id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(device, group)
        return cls._cache[key]

    def channel(self) -> PeerChannel:
        i = self.next % self.CHANNELS
        self.next += 1
        if self.channels[i] is None:
            self.channels[i] = PeerChannel(self.rank, self.world,
                                           [p + i * self.stride for p in self.ptrs])
        return self.channels[i]


class PartialsReducer:
    """Cross-rank SUM of 8-double partials records, delivered to the host.

    One objective evaluation ends in exactly one of these reductions (two for looph and for
    the analytic-scale nugget quirk).  On the GPUs of one NVLink domain the sum is a one-shot
    exchange through peer-mapped memory (`mgp_peer_sum8`, or fused into the epilogue of the
    objective kernel itself: `channel()` + `mgp_fused_loo_peers`) -- no collective launch, no
    host synchronisation besides the final 64-byte read.  Everywhere else (gloo in the CPU
    tests, MGP_NO_PEER=1, symmetric memory unavailable) it is a plain `all_reduce(SUM)`."""

    def __init__(self, device, group=None):
        import os

        self.device = torch.device(device)
        self.group = group
        self._pin = None
        self.peers = None
        self._own_channel = None
        if self.device.type == "cuda":
            self._pin = torch.empty((8,), dtype=torch.float64).pin_memory()
            _, size = rank_and_size(group)
            if size > 1 and size <= 8 and os.environ.get("MGP_NO_PEER") != "1":
                try:
                    self.peers = _PeerBuffers.get(self.device, group)
                except Exception as exc:  # noqa: BLE001  (no P2P / symmetric memory here)
                    import warnings

                    warnings.warn(f"peer-memory reduction unavailable ({exc}); using all_reduce")
                    self.peers = None

    def channel(self) -> Optional[PeerChannel]:
        """A fresh exchange slot for a kernel that sums across ranks itself (None: it cannot)."""
        return None if self.peers is None else self.peers.channel()

    def slot(self) -> torch.Tensor:
        return torch.zeros((8,), dtype=torch.float64, device=self.device)

    def to_host(self, record: torch.Tensor):
        """Host copy of an already summed device record."""
        if self._pin is None:
            return record.detach().cpu().numpy().copy()
        self._pin.copy_(record, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pin.numpy().copy()

    def sum_to_host(self, record: torch.Tensor):
        """Sum a per-rank device record across ranks (in place) and return it on the host."""
        if self.peers is not None:
            from . import _lib as L
            from .ops import _stream

            if self._own_channel is None:
                self._own_channel = self.peers.channel()
            g = self._own_channel.group_struct()
            with torch.cuda.device(self.device):
                import ctypes as C

                L.check(L.lib().mgp_peer_sum8(record.data_ptr(), C.byref(g), _stream()))
        else:
            allreduce_partials(record, self.group)
        return self.to_host(record)


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
