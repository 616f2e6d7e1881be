"""Peer-mapped device buffers for the GPUs of one node (plumbing of the fused partitioned path, ``dist_fused.py``).

``PeerBuffer(nbytes)`` allocates the same-sized buffer on every rank and maps all of them into every process, so
that the kernels of ``csrc/p2p.cu`` and the transform epilogue can load / store other GPUs' memory over NVLink.
Two mechanisms, tried in this order (``PRIMEKG_RGCN_PEER=symm|ipc`` forces one):

* ``symm``: ``torch.distributed._symmetric_memory`` (CUDA VMM handles exchanged through the process group's store);
  its signal pads also give a device-side barrier that needs no collective library;
* ``ipc`` : plain allocations shared with ``cudaIpc*`` handles through ``torch.multiprocessing.reductions`` and an
  ``all_gather_object``; the barrier is then a one-element NCCL all-reduce on the current stream.

Either way the barrier is stream-ordered (no host synchronisation) and comes after kernel boundaries, which is what
makes peer writes visible to the kernels that follow it on the other GPUs.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist

from . import _lib


class PeerBuffer:
    def __init__(self, nbytes: int, device: torch.device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerBuffer needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device
        self.nfloat = (int(nbytes) + 15) // 16 * 4
        want = os.environ.get("PRIMEKG_RGCN_PEER", "").lower()
        self.kind = None
        err = None
        if want in ("", "symm"):
            try:
                self._init_symm()
                self.kind = "symm"
            except Exception as e:  # noqa: BLE001 — fall through to the IPC mechanism, report if that fails too
                err = e
                if want == "symm":
                    raise
        if self.kind is None:
            try:
                self._init_ipc()
                self.kind = "ipc"
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(f"no peer-memory mechanism works on this node (symmetric memory: {err!r}; "
                                   f"CUDA IPC: {e!r})") from e
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self._chan = 0

    # -- torch symmetric memory ---------------------------------------------------------------------
    def _init_symm(self) -> None:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(self.nfloat, dtype=torch.float32, device=self.device)
        hdl = symm.rendezvous(t, self.group)
        self.local, self._hdl = t, hdl
        self.ptrs: List[int] = [int(p) for p in hdl.buffer_ptrs]
        if self.ptrs[self.rank] != t.data_ptr():
            raise RuntimeError("symmetric memory handle does not map the local buffer at its own address")

    # -- CUDA IPC -----------------------------------------------------------------------------------
    def _init_ipc(self) -> None:
        from torch.multiprocessing.reductions import reduce_tensor
        lib = _lib.load()
        self.local = torch.empty(self.nfloat, dtype=torch.float32, device=self.device)
        self._hdl = None
        fn, args = reduce_tensor(self.local)
        infos = [None] * self.world
        dist.all_gather_object(infos, (fn, args, self.device.index), group=self.group)
        self._peers = []
        self.ptrs = []
        for q, (f, a, dev_q) in enumerate(infos):
            if q == self.rank:
                self._peers.append(self.local)
            else:
                with torch.cuda.device(self.device):
                    _lib.check(lib.rgcn_enable_peer_access(int(dev_q)), "rgcn_enable_peer_access")
                self._peers.append(f(*a))              # maps the peer allocation into this process
            self.ptrs.append(int(self._peers[-1].data_ptr()))
        dist.barrier(group=self.group)

    # -----------------------------------------------------------------------------------------------
    def view(self, rows: int, cols: int, offset_floats: int = 0) -> torch.Tensor:
        """[rows, cols] fp32 view of the LOCAL buffer."""
        return self.local[offset_floats: offset_floats + rows * cols].view(rows, cols)

    def peer_ptrs(self, offset_floats: int = 0) -> List[int]:
        """Device address of float ``offset_floats`` of every rank's buffer, in rank order."""
        return [p + 4 * offset_floats for p in self.ptrs]

    def barrier(self) -> None:
        """Cross-GPU barrier on the current stream: everything the ranks enqueued before it (peer stores included)
        is complete and visible before anything enqueued after it starts."""
        if self.kind == "symm":
            self._hdl.barrier(channel=self._chan)
            self._chan ^= 1
        else:
            dist.all_reduce(self._flag, group=self.group)
