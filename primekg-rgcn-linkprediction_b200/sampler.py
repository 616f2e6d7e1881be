"""Device-side negative sampler with the interface of the reference's ``NegativeSampler`` (src/train.py:43-97).

``sample(pos_head, pos_tail, pos_rel)`` returns the corrupted triples like the reference; ``batch(...)`` returns the whole
mini-batch the training loop assembles from them (src/train.py:281-288: positives, then negatives, labels 1 / 0) from one
kernel.  Random numbers come from a counter-based generator (seed drawn from torch's generator at construction, device
counter advanced by every call), so a captured CUDA graph draws fresh negatives on every replay.  Same distribution as
the reference (head or tail with probability 1/2, replacement uniform over all nodes, ``repeat_interleave`` order), not
the same stream of numbers as ``torch.rand`` / ``torch.randint``.
"""
from __future__ import annotations

import torch

from . import ops


class NegativeSampler:
    def __init__(self, num_nodes: int, num_neg_samples: int = 1):
        self.num_nodes = int(num_nodes)
        self.num_neg_samples = int(num_neg_samples)
        self._seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        self._ctr = None

    def _counter(self, device):
        if self._ctr is None or self._ctr.device != device:
            self._ctr = ops.rng_counter(device)
        return self._ctr

    def batch(self, pos_head: torch.Tensor, pos_tail: torch.Tensor, pos_rel: torch.Tensor, out=None):
        """(all_heads, all_tails, all_rels, labels) of src/train.py:281-288."""
        if not pos_head.is_cuda:
            raise RuntimeError("the device-side sampler needs CUDA tensors")
        return ops.link_batch(pos_head, pos_tail, pos_rel, self.num_nodes, self.num_neg_samples, self._seed,
                              self._counter(pos_head.device), out=out)

    def sample(self, pos_head: torch.Tensor, pos_tail: torch.Tensor, pos_rel: torch.Tensor):
        n = pos_head.numel()
        h, t, r, _ = self.batch(pos_head, pos_tail, pos_rel)
        return h[n:], t[n:], r[n:]
