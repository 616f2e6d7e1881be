"""Partitioned (destination-range) RGCN on N GPUs of one node: correctness against the single-GPU model on the same
graph, then timing of the forward + loss + backward step.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/run_partitioned.py --nodes 2000000 --edges 40000000 --relations 30 --hidden 128 --layers 3
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=200_000)
    ap.add_argument("--edges", type=int, default=4_000_000)
    ap.add_argument("--relations", type=int, default=30)
    ap.add_argument("--embedding", type=int, default=64)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", action="store_true", help="compare with the single-GPU model on rank 0")
    ap.add_argument("--mode", default="fp32")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "fused"],
                    help="nccl: all-gather / reduce-scatter calls (dist.py); fused: our kernels over peer memory (dist_fused.py)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import primekg_rgcn_linkprediction_b200 as pkg
    from primekg_rgcn_linkprediction_b200 import dist as D
    from primekg_rgcn_linkprediction_b200 import synth

    os.environ["PRIMEKG_RGCN_MODE"] = args.mode
    kg = synth.scaled_kg(args.nodes, args.edges, args.relations, seed=42, device=dev)     # same graph on every rank
    ei, et = kg.edge_index, kg.edge_type
    plan = D.plan_partition(ei[1], kg.num_nodes, world)
    from primekg_rgcn_linkprediction_b200 import dist_fused as DF
    cls = DF.FusedPartitionedModel if args.exchange == "fused" else D.PartitionedModel
    model = cls(plan, rank, args.relations, args.embedding, args.hidden, dropout=0.0,
                decoder_dropout=0.0, num_layers=args.layers, seed=42).to(dev)
    model.encoder.build_graph(ei, et)
    model.train()
    B = 2048
    g = torch.Generator(device=dev).manual_seed(7)
    heads = torch.randint(0, args.nodes, (B,), generator=g, device=dev)
    tails = torch.randint(0, args.nodes, (B,), generator=g, device=dev)
    rels = torch.randint(0, args.relations, (B,), generator=g, device=dev)
    labels = (torch.rand(B, generator=g, device=dev) < 0.5).float()
    sl = slice(rank * B // world, (rank + 1) * B // world)

    def step():
        for p in model.parameters():
            p.grad = None
        s = model(heads[sl], tails[sl], rels[sl])
        loss = F.binary_cross_entropy_with_logits(s, labels[sl], reduction="sum") / B
        loss.backward()
        model.allreduce_decoder_grads()
        return loss.detach(), s.detach()

    loss, scores = step()
    total = loss.clone()
    dist.all_reduce(total)
    report = {"world": world, "nodes": args.nodes, "edges": int(et.numel()), "relations": args.relations,
              "layers": args.layers, "hidden": args.hidden, "loss": float(total),
              "shard_rows": [plan.size(p) for p in range(world)], "max_n": plan.max_n}

    if args.check:
        # single-GPU model with the same parameters on rank 0
        shards = [torch.zeros_like(model.encoder.node_embeddings.data) for _ in range(world)]
        dist.all_gather(shards, model.encoder.node_embeddings.data)
        gshards = [torch.zeros_like(model.encoder.node_embeddings.grad) for _ in range(world)]
        dist.all_gather(gshards, model.encoder.node_embeddings.grad)
        if rank == 0:
            ref = pkg.DrugDiseaseModel(args.nodes, args.relations, args.embedding, args.hidden, dropout=0.0,
                                       decoder_dropout=0.0, num_layers=args.layers).to(dev)
            with torch.no_grad():
                ref.encoder.node_embeddings.weight.copy_(torch.cat([shards[p][: plan.size(p)] for p in range(world)]))
                for mine, theirs in zip(model.encoder.convs, ref.encoder._layers()):
                    theirs.weight.copy_(mine.weight); theirs.root.copy_(mine.root); theirs.bias.copy_(mine.bias)
                ref.decoder.relation_embeddings.weight.copy_(model.decoder.relation_embeddings.weight)
            ref.train()
            rs = ref(ei, et, heads, tails, rels)
            rl = F.binary_cross_entropy_with_logits(rs, labels)
            rl.backward()
            ggrad = torch.cat([gshards[p][: plan.size(p)] for p in range(world)])
            want = ref.encoder.node_embeddings.weight.grad
            report["check"] = {
                "loss_ref": float(rl), "loss_abs_err": abs(float(rl) - float(total)),
                "score_max_abs_err": float((rs[sl] - scores).abs().max()),
                "emb_grad_rel_fro": float((ggrad - want).norm() / (want.norm() + 1e-30)),
                "w_grad_rel_fro": [float((m.weight.grad - t.weight.grad).norm() / (t.weight.grad.norm() + 1e-30))
                                   for m, t in zip(model.encoder.convs, ref.encoder._layers())]}
            del ref
            torch.cuda.empty_cache()
    dist.barrier()
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize(); dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        report.update({"ms_per_step": float(t), "edges_per_sec": int(et.numel()) / (float(t) * 1e-3), "mode": args.mode,
                       "exchange": args.exchange,
                       "peer_memory": getattr(getattr(model.encoder, "_ex", None), "x", None) and model.encoder._ex.x.kind})
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
