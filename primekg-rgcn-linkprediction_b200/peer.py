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
