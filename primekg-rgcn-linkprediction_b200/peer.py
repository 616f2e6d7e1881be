"""Peer-mapped device buffers for the GPUs of one node (plumbing of the fused partitioned path, ``dist_fused.py``).

``PeerBuffer(nbytes)`` allocates the same-sized buffer on every rank and maps all of them into every process, so
that the kernels of ``csrc/p2p.cu`` and the transform epilogue can load / store other GPUs' memory over NVLink.
The mapping comes from ``torch.distributed._symmetric_memory`` (CUDA VMM handles exchanged through the process
group's store); its signal pads also give a device-side barrier that needs no collective library.  The barrier is
stream-ordered (no host synchronisation) and sits between kernel boundaries, which is what makes peer stores visible
to the kernels that follow it on the other GPUs.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class PeerBuffer:
    def __init__(self, nbytes: int, device: torch.device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerBuffer needs an initialised process group")
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device
        self.nfloat = (int(nbytes) + 15) // 16 * 4
        self.kind = "symm"
        try:
            self.local = symm.empty(self.nfloat, dtype=torch.float32, device=device)
            self._hdl = symm.rendezvous(self.local, self.group)
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("peer-mapped memory is not available on this node (torch symmetric memory failed); "
                               "use the collective-library form dist.PartitionedRGCN") from e
        self.ptrs: List[int] = [int(p) for p in self._hdl.buffer_ptrs]
        if len(self.ptrs) != self.world or self.ptrs[self.rank] != self.local.data_ptr():
            raise RuntimeError("symmetric memory handle does not map the local buffer at its own address")
        self._chan = 0

    def view(self, rows: int, cols: int, offset_floats: int = 0) -> torch.Tensor:
        """[rows, cols] fp32 view of the LOCAL buffer."""
        return self.local[offset_floats: offset_floats + rows * cols].view(rows, cols)

    def peer_ptrs(self, offset_floats: int = 0) -> List[int]:
        """Device address of float ``offset_floats`` of every rank's buffer, in rank order."""
        return [p + 4 * offset_floats for p in self.ptrs]

    def barrier(self) -> None:
        """Cross-GPU barrier on the current stream: everything the ranks enqueued before it (peer stores included)
        is complete and visible before anything enqueued after it starts."""
        self._hdl.barrier(channel=self._chan)
        self._chan ^= 1


class PeerAllReduce:
    """All-reduce (average) of one flat fp32 buffer over the ranks by OUR kernels (``rgcn_p2p_allreduce``: two-shot over
    peer-mapped memory, rank-ordered sums, device-side epoch flags) — no collective call, CUDA-graph capturable.

    ``inp`` (write the local contribution here) and ``out`` (read the average here) are views of one ``PeerBuffer``; the
    data-parallel ``GraphedTrainStep(flat_grads="arena", allreduce="peer")`` lets the backward kernels write the parameter
    gradients straight into ``inp`` and runs the exchange as the last kernels of the captured step."""

    def __init__(self, numel: int, device: torch.device, group=None):
        from . import _lib
        self.n = (int(numel) + 63) // 64 * 64
        flag_floats = 64                                           # >= rgcn_p2p_allreduce_flag_bytes() / 4
        assert int(_lib.load().rgcn_p2p_allreduce_flag_bytes()) <= 4 * flag_floats
        self.buf = PeerBuffer((2 * self.n + flag_floats) * 4, device, group)
        self.buf.local.zero_()
        self.inp = self.buf.local[: self.n]
        self.out = self.buf.local[self.n: 2 * self.n]
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.rank, self.world = self.buf.rank, self.buf.world
        torch.cuda.synchronize(device)
        dist.barrier(self.buf.group)                               # every rank's flags are zero before anybody signals
        self._in = self.buf.peer_ptrs(0)
        self._out = self.buf.peer_ptrs(self.n)
        self._flags = self.buf.peer_ptrs(2 * self.n)

    def __call__(self, n_floats: int = None, scale: float = None) -> torch.Tensor:
        """Exchange the first ``n_floats`` (default: all) entries of ``inp``; returns ``out`` (valid in stream order)."""
        import ctypes as C
        from . import _lib
        from .graph import _ptr, _stream
        lib = _lib.load()
        n = self.n if n_floats is None else (int(n_floats) + 3) // 4 * 4
        if n > self.n:
            raise ValueError("n_floats exceeds the buffer")
        arr = lambda ptrs: (C.c_void_p * len(ptrs))(*[int(a) for a in ptrs])     # noqa: E731
        _lib.check(lib.rgcn_p2p_allreduce(arr(self._in), arr(self._out), arr(self._flags), self.world, self.rank, n,
                                          float(1.0 / self.world if scale is None else scale), _ptr(self.epoch),
                                          _ptr(self.status), _stream(self.inp.device)), "rgcn_p2p_allreduce")
        return self.out

    def check(self) -> None:
        """Raise if a wait gave up (a peer never arrived).  Synchronises."""
        if int(self.status.item()) & 2:
            raise RuntimeError("peer all-reduce: a rank did not arrive within the wait bound")
