"""All-reduce of the step's 9.2 MB of gradients: one flat tensor vs the 8 parameter tensors coalesced (NCCL, AVG)."""
import os, sys, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
shapes = [(30926, 64), (3, 64, 256), (64, 256), (256,), (3, 256, 256), (256, 256), (256,), (3, 256)]
tens = [torch.randn(*s, device=dev) for s in shapes]
flat = torch.randn(sum(t.numel() for t in tens), device=dev)
def coalesced():
    with dist._coalescing_manager(device=dev):
        for g in tens:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
def one(): dist.all_reduce(flat, op=dist.ReduceOp.AVG)
def two():
    dist.all_reduce(tens[0], op=dist.ReduceOp.AVG)
    with dist._coalescing_manager(device=dev):
        for g in tens[1:]:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
for name, fn in (("coalesced 8 tensors", coalesced), ("one flat tensor", one), ("big one + coalesced rest", two)):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); dist.barrier()
    t = torch.tensor([sorted(ts)[len(ts)//2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"world {world}: {name}: {float(t)*1e3:.1f} us (median, max over ranks)", flush=True)
dist.destroy_process_group()
