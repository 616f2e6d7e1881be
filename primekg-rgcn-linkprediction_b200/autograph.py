"""The unmodified training call ``model(edge_index, edge_type, heads, tails, rels)`` ... ``loss.backward()`` (reference
src/train.py:291-306) as two CUDA graphs behind the module API.

Driven eagerly from Python, one cfg2 step is ~30 kernels of 3-80 us and ~0.7 ms of host work (autograd bookkeeping, ctypes
argument marshalling, allocator calls) against 0.41 ms of GPU time: the reference's loop is bound by the host.  After
``WARM`` identical calls (same graph tensors, batch size, parameters, modes) the model captures its forward and its
backward with ``torch.cuda.make_graphed_callables`` and from then on replays them: the call keeps its signature and its
autograd semantics (scores with a ``grad_fn``; ``backward()`` accumulates into ``p.grad``), the caller's loss, optimiser
and clipping code stay as they are.  Fresh dropout masks on every replay (the fused dropout hashes a device-side step
counter that a kernel of the captured step advances).

Caveat of CUDA graphs, as for ``GraphedTrainStep``: the returned scores live in a static buffer that the next training
call overwrites.  ``PRIMEKG_RGCN_AUTOGRAPH=0`` switches the capture off; graphs above ``MAX_EDGES`` edges are left eager
(their steps are not host-bound and the private pool of a captured step would double the activation memory).
"""
from __future__ import annotations

import gc
import os
import threading
import weakref
from collections import OrderedDict

import torch

WARM = 3                 # identical eager calls before the capture
MAX_EDGES = 4_000_000
MAX_CACHED = 4           # captured (graph, batch size) pairs kept per model


def enabled() -> bool:
    return os.environ.get("PRIMEKG_RGCN_AUTOGRAPH", "1") != "0"


_TLS = threading.local()     # .bypass: the call comes from the capture itself and must run the eager body


def _capture(model, edge_index, edge_type, heads, tails, rels):
    """Forward and backward graphs of the call, differentiable in the model's parameters.

    The capture runs on ALIASES of the parameters (fresh leaves over the same storage).  The caller's previous step is
    usually still alive when the next ``model(...)`` arrives (``scores`` / ``loss`` of a ``for`` loop), and with it the
    parameters' AccumulateGrad nodes, which are bound to the stream they were created on (the default stream); a capture
    whose backward feeds those nodes would make the legacy stream wait on the capturing one and die.  The aliases get
    their own accumulators on the capture's streams; at replay time the REAL parameters are the inputs of the graphed
    function (same storage: nothing is copied) and receive the gradients through ordinary autograd.

    Nothing captured here references the model strongly (the graphs die with it, by reference count — a cycle would leave
    their destruction to the garbage collector, which may run in the middle of somebody's capture and invalidate it), and
    the collector is off while the streams are capturing."""
    mref = weakref.ref(model)
    names = [n for n, _ in model.named_parameters()]
    params = [p for _, p in model.named_parameters()]

    def fn(h, t, r, *ps):
        _TLS.bypass = True
        try:
            return torch.func.functional_call(mref(), dict(zip(names, ps)), (edge_index, edge_type, h, t, r))
        finally:
            _TLS.bypass = False

    aliases = tuple(p.detach().requires_grad_(p.requires_grad) for p in params)
    sample = (heads.clone(), tails.clone(), rels.clone()) + aliases
    gc.collect()
    was_on = gc.isenabled()
    gc.disable()
    try:
        graphed = torch.cuda.make_graphed_callables(fn, sample, num_warmup_iters=3, allow_unused_input=True)
    finally:
        if was_on:
            gc.enable()

    def run(h, t, r):
        return graphed(h, t, r, *params)

    return run


def _key(model, edge_index, edge_type, heads, tails, rels):
    modes = tuple(getattr(m, "mode", None) for m in model.encoder._layers())
    return (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), edge_type.data_ptr(), edge_type._version,
            heads.numel(), heads.dtype, tails.dtype, rels.dtype, heads.device.index, modes,
            tuple(p.data_ptr() for p in model.parameters()))


def lookup(model, edge_index, edge_type, heads, tails, rels):
    """The graphed callable for this call signature, or None (run eagerly)."""
    if getattr(_TLS, "bypass", False) or not (model.training and torch.is_grad_enabled() and enabled()):
        return None
    if not (edge_index.is_cuda and heads.is_cuda and tails.is_cuda and rels.is_cuda):
        return None
    if edge_type.numel() > MAX_EDGES or heads.dim() != 1 or heads.shape != tails.shape or heads.shape != rels.shape:
        return None
    if torch.cuda.is_current_stream_capturing():          # an outer capture (GraphedTrainStep) owns the step
        return None
    state = model.__dict__.setdefault("_autograph", OrderedDict())
    key = _key(model, edge_index, edge_type, heads, tails, rels)
    entry = state.get(key)
    if entry is None:
        while len(state) >= MAX_CACHED:
            state.popitem(last=False)
        entry = state[key] = [0, None]
    else:
        state.move_to_end(key)
    if entry[1] is not None:
        return entry[1]
    entry[0] += 1
    if entry[0] <= WARM:
        return None
    if entry[1] is None and entry[0] == WARM + 1:
        try:
            entry[1] = _capture(model, edge_index, edge_type, heads, tails, rels)
        except Exception as e:                             # leave the call eager, say why once
            entry[1] = None
            entry[0] = 1 << 30
            import warnings
            if os.environ.get("PRIMEKG_RGCN_AUTOGRAPH_DEBUG"):
                import traceback
                traceback.print_exc()
            warnings.warn(f"primekg_rgcn_linkprediction_b200: CUDA-graph capture of the training call failed, staying eager: {e}")
    return entry[1]


def invalidate(model) -> None:
    model.__dict__.pop("_autograph", None)
