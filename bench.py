#!/usr/bin/env python
"""Benchmark of the RGCN message-passing hot path (BASELINE.json metric: full-batch RGCN fwd+bwd
edges/sec; aggregation GB/s against the measured HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode fp32|bf16]

One *step* = train-mode ``DrugDiseaseModel.forward`` on the full graph + BCEWithLogits + ``backward()``
(reference src/train.py:291-306; optimiser, clipping and sampling excluded).
Workload at every N: cfg2 of BASELINE.json — the synthetic PrimeKG-shaped graph (30,926 nodes /
849,456 directed edges / 3 relations), 2-layer RGCN 64 -> 256 -> 256, batch 1,024 positives + 1,024
negatives, dropout 0.5 / decoder dropout 0.1, seed 42.  N > 1: data-parallel replicas — every rank holds
the graph and the model, processes its own mini-batch and the parameter gradients are all-reduced over
NCCL inside the timed step (weak scaling: N * E edges per step).

Prints ONE JSON line (rank 0).  See the module docstring of the contract in DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(num_nodes=30_926, num_edges=849_456, num_relations=3, embedding_dim=64, hidden_dim=256,
           batch_pos=1024, dropout=0.5, decoder_dropout=0.1, seed=42)
WORKLOAD = ("cfg2: synthetic PrimeKG-shaped KG 30,926 nodes / 849,456 edges / 3 relations, "
            "2-layer RGCN 64->256->256, batch 1024+1024, fwd+loss+bwd")
METRIC = "rgcn_fwd_bwd_edges_per_sec"
UNIT = "edges/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def make_workload(rank: int):
    from primekg_rgcn_linkprediction_b200 import synth
    kg = synth.primekg_subgraph(CFG["num_edges"], seed=CFG["seed"])
    batch = synth.link_batch(kg, CFG["batch_pos"], seed=CFG["seed"] + 1000 * rank)   # a different mini-batch per rank
    return kg, batch


def l2_cap(achieved_gbs, clocks):
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 6300.0 * mhz * 1e6 / 1e9
    return {"peak": round(peak, 1), "unit": "GB/s", "frac": round(achieved_gbs / peak, 4),
            "source": "6300 B/clk (B300_MICROARCH.md, LTS throughput cap) x %.0f MHz" % mhz}


def algorithmic_bytes(E, N, R, d_in):
    """SURVEY.md §8d per-layer figures for the aggregation kernels (bytes per launch)."""
    fwd = E * (d_in * 4 + 4) + (N * R + 1) * 4
    bwd = E * (d_in * 4 + 8) + (N * R + 1) * 4
    return fwd, bwd


def time_dominant_kernel(pkg, graph, d, iters, flush, comp=None):
    """CUDA-event duration of the dominant op (forward / backward aggregation = hub chunks +
    aggregate_rows_kernel, gather width d) on the launching stream, L2 flushed between launches."""
    from primekg_rgcn_linkprediction_b200 import ops
    x = torch.randn(graph.n_src, d, device="cuda")
    gA = torch.randn(graph.n_dst, (graph.R + 1) * d, device="cuda")
    out = {}
    for name, fn in (("aggregate_fwd", lambda: ops.aggregate_fwd(graph, x, comp=comp)),
                     ("aggregate_bwd", lambda: ops.aggregate_bwd(graph, gA, d, init=gA[:, graph.R * d:]))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        out[name] = statistics.mean(ts)
    return out


def run_ours(args, rank, world, local_rank):
    import primekg_rgcn_linkprediction_b200 as pkg
    from primekg_rgcn_linkprediction_b200 import _lib
    import torch.distributed as dist

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    kg, (heads, tails, rels, labels) = make_workload(rank)
    torch.manual_seed(CFG["seed"])
    model = pkg.DrugDiseaseModel(kg.num_nodes, kg.num_relations, CFG["embedding_dim"], CFG["hidden_dim"],
                                 dropout=CFG["dropout"], decoder_dropout=CFG["decoder_dropout"]).to(dev)
    for c in (model.encoder.conv1, model.encoder.conv2):
        c.mode = args.mode
    model.train()
    ei, et = kg.edge_index.to(dev), kg.edge_type.to(dev)
    params = [p for p in model.parameters()]
    flush_buf = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def flush():
        flush_buf.fill_(1.0)            # 512 MiB write > 126 MB L2

    d_batch = [t.to(dev) for t in (heads, tails, rels, labels)]

    def eager_step(b):
        for p in params:
            p.grad = None
        scores = model(ei, et, b[0], b[1], b[2])
        loss = F.binary_cross_entropy_with_logits(scores, b[3])
        loss.backward()
        return loss

    def allreduce_grads():
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eager_step(d_batch)
        allreduce_grads()
    barrier()
    l0 = lib.rgcn_launch_count()
    eager_step(d_batch)
    torch.cuda.synchronize()
    launches_per_step = int(lib.rgcn_launch_count() - l0)

    def timed(run_one, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps."""
        evs = []
        for _ in range(steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run_one(); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    # ---- eager module API (what an unmodified src/train.py drives): device-resident, then from pinned host ----
    barrier()
    eager_ms = timed(lambda: (eager_step(d_batch), allreduce_grads()), args.steps)
    pinned = [t.pin_memory() for t in (heads, tails, rels, labels)]
    h2d = sum(t.numel() * t.element_size() for t in pinned)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_eager():
        b = [t.to(dev, non_blocking=True) for t in pinned]
        loss_host.copy_(eager_step(b).detach().reshape(1), non_blocking=True)
        allreduce_grads()

    for _ in range(2):
        e2e_eager()
    barrier()
    e2e_eager_value = world * kg.num_edges / (timed(e2e_eager, args.steps) * 1e-3)
    for p in params:
        p.grad = None

    # ---- the same step captured once into a CUDA graph (GraphedTrainStep) ----
    # N > 1: the backward kernels write every parameter gradient into ONE flat buffer (ops.GradArena), so the replicas
    # exchange a single tensor (one NCCL all-reduce: 63-70 us for the 9.2 MB against 107-124 us as a coalesced group)
    gstep = pkg.GraphedTrainStep(model, ei, et, batch_size=d_batch[0].numel(), flat_grads="arena" if world > 1 else False)
    gstep.load_batch(*d_batch)
    flat_holder = [gstep.flat_grad]

    def allreduce_graphed():
        if world > 1:
            if flat_holder[0] is not None:
                dist.all_reduce(flat_holder[0], op=dist.ReduceOp.AVG)   # NCCL averages in the reduction: no scaling pass
                return
            grads = [p.grad for p in params]
            with dist._coalescing_manager(device=dev):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG)

    allreduce_grads = allreduce_graphed     # noqa: F811
    for _ in range(max(args.warmup, 3)):
        gstep(); allreduce_grads()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    ms_per_step = timed(lambda: (gstep(), allreduce_grads()), args.steps)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    value = world * kg.num_edges / (ms_per_step * 1e-3)

    # ---- end to end: batch in pinned host memory, H2D of the step's inputs and D2H of the loss inside ----
    # the step with its host I/O captured: H2D of the batch (one pinned int64 [4, B] block) as the graph's first node,
    # D2H of the loss (+ correct count) as its last — one graph launch per step is all the host does
    packed = pkg.GraphedTrainStep.pack_batch(heads, tails, rels, labels)
    h2d_packed = packed.numel() * packed.element_size()
    for p in params:
        p.grad = None
    ghost = pkg.GraphedTrainStep(model, ei, et, batch_size=d_batch[0].numel(), host_io=True,
                                 flat_grads="arena" if world > 1 else False)
    ghost.host_batch.copy_(packed)
    flat_value = flat_holder[0]
    flat_holder[0] = ghost.flat_grad

    def e2e_graphed():
        ghost.replay_host()
        allreduce_grads()

    for _ in range(2):
        e2e_graphed()
    barrier()
    e2e_value = world * kg.num_edges / (timed(e2e_graphed, args.steps) * 1e-3)
    e2e_loss = float(ghost.host_loss)                 # read after the timed region's synchronisation
    flat_holder[0] = flat_value
    del ghost

    # ---- the same graphed step with the row-sparse hand-over off: the last layer's backward over all N rows, i.e.
    #      the dense formulation SURVEY.md §8d's byte counts describe (same gradients, see tests) ----
    sparse_env = os.environ.get("PRIMEKG_RGCN_SPARSE_BWD")
    dense_ms = None
    if sparse_env != "0":
        os.environ["PRIMEKG_RGCN_SPARSE_BWD"] = "0"
        try:
            del gstep
            for p in params:
                p.grad = None
            gdense = pkg.GraphedTrainStep(model, ei, et, batch_size=d_batch[0].numel(),
                                          flat_grads="arena" if world > 1 else False)
            gdense.load_batch(*d_batch)
            flat_holder[0] = gdense.flat_grad
            for _ in range(3):
                gdense(); allreduce_grads()
            barrier()
            dense_ms = timed(lambda: (gdense(), allreduce_grads()), args.steps)
        finally:
            if sparse_env is None:
                del os.environ["PRIMEKG_RGCN_SPARSE_BWD"]
            else:
                os.environ["PRIMEKG_RGCN_SPARSE_BWD"] = sparse_env

    if rank != 0:
        return None
    # ---- roofline of the dominant kernel ----
    graph = pkg.get_graph(ei, et, kg.num_nodes, kg.num_relations)
    d2 = CFG["hidden_dim"]
    kt = time_dominant_kernel(pkg, graph, d2, max(5, min(args.steps, 20)), flush)
    fwd_b, bwd_b = algorithmic_bytes(kg.num_edges, kg.num_nodes, kg.num_relations, d2)
    peak, peak_src = measured_peaks()
    achieved = fwd_b / (kt["aggregate_fwd"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("aggregate_rows_fwd_d256_bytes")
    roofline = {"bound": "hbm", "kernel": "hub_partial_kernel + aggregate_rows_kernel (layer-2 forward gather, d=256)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": fwd_b,
                "avg_launch_ms": round(kt["aggregate_fwd"], 5),
                "bwd": {"achieved": round(bwd_b / (kt["aggregate_bwd"] * 1e-3) / 1e9, 1),
                        "avg_launch_ms": round(kt["aggregate_bwd"], 5), "algorithmic_bytes_per_launch": bwd_b},
                "note": "cfg2 working set is L2-resident (features 31.7 MB < 126 MB L2): algorithmic GB/s may exceed the HBM peak",
                # the bound that binds on this working set: the L2 slices' throughput cap, ~6300 B/clk full chip
                # (B300_MICROARCH.md, 'LTS throughput cap'; same LTS count on B200) at the SM clock sampled in this run
                "l2_cap": l2_cap(achieved, clocks)}
    # the same kernel where the gather working set exceeds L2 (cfg3-sized graph: 129,375 x 256 fp32 = 132 MB):
    # there the HBM roofline is the binding one
    hbm_case = None
    try:
        from primekg_rgcn_linkprediction_b200 import synth
        big = synth.primekg_full()
        gb = pkg.RelGraph.from_edges(big.edge_index.to(dev), big.edge_type.to(dev), big.num_nodes, big.num_relations)
        comp = torch.randn(big.num_relations, 8, device=dev)            # cfg3 uses basis decomposition, B = 8
        kb_ = time_dominant_kernel(pkg, gb, d2, 5, flush, comp=comp)
        fb, bb = algorithmic_bytes(big.num_edges, big.num_nodes, big.num_relations, d2)
        out_bytes = big.num_nodes * 8 * d2 * 4                           # the basis-mixed output write, not in SURVEY's figure
        hbm_case = {"workload": "cfg3-shaped KG 129,375 nodes / 8,100,498 edges / 30 relations, gather width 256",
                    "fwd": {"achieved": round(fb / (kb_["aggregate_fwd"] * 1e-3) / 1e9, 1),
                            "avg_launch_ms": round(kb_["aggregate_fwd"], 4), "algorithmic_bytes_per_launch": fb,
                            "output_bytes_not_counted": out_bytes},
                    "bwd": {"achieved": round(bb / (kb_["aggregate_bwd"] * 1e-3) / 1e9, 1),
                            "avg_launch_ms": round(kb_["aggregate_bwd"], 4), "algorithmic_bytes_per_launch": bb},
                    "peak": peak, "unit": "GB/s"}
        hbm_case["fwd"]["frac"] = round(hbm_case["fwd"]["achieved"] / peak, 4)
        hbm_case["bwd"]["frac"] = round(hbm_case["bwd"]["achieved"] / peak, 4)
        del gb, big
    except Exception as ex:  # pragma: no cover
        hbm_case = {"error": repr(ex)[:200]}
    roofline["hbm_bound_case"] = hbm_case
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16-transform/f32-accumulate",
           "data": "synthetic",
           "config": {"workload": WORKLOAD, "mode": args.mode, "l2": "flushed between steps (512 MiB write)",
                      "parallelism": "single GPU" if world == 1 else f"dp{world} replicas, parameter gradients written into one flat buffer and all-reduced in one NCCL call",
                      "timing": "CUDA events per step on the launching stream, max over ranks",
                      "step": "one CUDA-graph replay of forward + BCE loss + backward (GraphedTrainStep)",
                      "last_layer_backward": ("dense over all N rows (PRIMEKG_RGCN_SPARSE_BWD=0)" if sparse_env == "0" else
                                              "row-sparse: on the 2*batch rows of the encoder output the loss reads "
                                              "(reference src/models/rgcn.py:325-326); identical gradients, "
                                              "tests/test_gpu_parity.py::test_layer_bwd_rows_equals_dense")},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": 8,
                   "api": "GraphedTrainStep(model, edge_index, edge_type, host_io=True): host_batch <- pack_batch(heads, tails, rels, labels); replay_host(); host_loss",
                   "loss_read_back": e2e_loss,
                   "eager_module_api_value": e2e_eager_value,
                   "note": "batch (heads, tails, rels, labels; one int64 [4, B] block) copied from pinned host memory and the loss copied back to pinned host memory inside every step (memcpy nodes of the captured graph); "
                           "the graph and the model stay device-resident as in reference src/train.py:122-135; "
                           "eager_module_api_value = the unmodified reference call model(...); loss; backward()"},
           "eager_ms_per_step": eager_ms,
           "dense_last_layer_bwd": (None if dense_ms is None else
                                    {"ms_per_step": dense_ms, "value": world * kg.num_edges / (dense_ms * 1e-3), "unit": UNIT,
                                     "note": "same graphed step with the last layer's backward over all N rows "
                                             "(PRIMEKG_RGCN_SPARSE_BWD=0)"}),
           "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
           "wall_s_timed_region": t_wall, "clocks": clocks, "roofline": roofline}
    return out


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, budget_s, threads=None):
    """Times the oracle (the restated reference path, oracle/rgcn_ref.py) on the host cores.
    Returns (edges_per_s, cores, sample description, ms_per_step)."""
    from oracle import rgcn_ref
    from primekg_rgcn_linkprediction_b200 import synth
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    kg = synth.primekg_subgraph(CFG["num_edges"], seed=CFG["seed"])
    heads, tails, rels, labels = synth.link_batch(kg, CFG["batch_pos"], seed=CFG["seed"])
    torch.manual_seed(CFG["seed"])
    model = rgcn_ref.ModelRef(kg.num_nodes, kg.num_relations, CFG["embedding_dim"], CFG["hidden_dim"],
                              CFG["dropout"], CFG["decoder_dropout"])
    model.train()
    ei, et = kg.edge_index, kg.edge_type

    def step(ei, et):
        model.zero_grad(set_to_none=True)
        rgcn_ref.train_step_ref(model, ei, et, heads, tails, rels, labels)

    t0 = time.perf_counter(); step(ei, et); t_first = time.perf_counter() - t0
    frac = min(1.0, budget_s / max(1e-9, (steps + warmup) * t_first))
    sample = f"full step: all {kg.num_edges:,} edges"
    if frac < 1.0:
        keep = max(2, int(kg.num_edges * frac) // 2 * 2)
        ei, et = ei[:, :keep].contiguous(), et[:keep].contiguous()      # whole (a->b),(b->a) pairs
        sample = f"first {keep:,} of {kg.num_edges:,} edges per step (bounded sample, {cores} threads)"
    for _ in range(max(0, warmup - 1)):
        step(ei, et)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(ei, et); ts.append(time.perf_counter() - t0)
    ms = statistics.mean(ts) * 1e3
    return et.numel() / (ms * 1e-3), cores, sample, ms


def run_reference(args):
    value, cores, sample, ms = cpu_reference_steps(args.steps, args.warmup, budget_s=150.0)
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": "fp32", "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "PyG is not installable here; this is the restated reference path (oracle/rgcn_ref.py) in "
                    "PyTorch on the host cores"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("PRIMEKG_RGCN_MODE", "fp32"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, ms = cpu_reference_steps(3, 1, budget_s=25.0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                   "ms_per_step": ms}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
