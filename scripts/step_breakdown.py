"""Per-kernel breakdown of ONE graph-replayed training step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv`): the launches between two consecutive link_loss_fwd_kernel launches are one
period of the step (rotated: loss, backward, next forward).  usage: step_breakdown.py launches.csv"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
for r in rd:
    rows.append((r[ki], r[gi], float(r[vi].replace(",", "")) / 1e3))
marks = [i for i, r in enumerate(rows) if "link_loss_fwd_kernel" in r[0]]
if len(marks) < 2:
    sys.exit("need two link_loss_fwd_kernel launches in the list")
step = rows[marks[-2]:marks[-1]]


def cat(name):
    if "aggregate_rows" in name or "hub_partial" in name:
        return "agg"
    if "gemm_" in name:
        return "gemm"
    if "split_planes" in name or "split_weights" in name or "rows_" in name:
        return "split/convert/compact"
    if "reduce" in name:
        return "reduce"
    return "other"


print("# one training step of bench.py (cfg2, fp32, CUDA-graph replay) under ncu --metrics gpu__time_duration.sum")
print("# (cold caches, kernels serialised: shares, not absolute times, compare with the live step)")
tot, by = 0.0, {}
for name, grid, us in step:
    short = re.sub(r"\(.*", "", name.replace("void ", "").replace("rgcn::", "").replace("at::native::", ""))[:64]
    print(f"{short:66s} {grid:>16s} {us:8.2f} us")
    tot += us
    by[cat(name)] = by.get(cat(name), 0.0) + us
print(f"\ntotal {tot:.1f} us, {len(step)} launches")
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f"{k:24s} {v:8.1f} us  {100 * v / tot:5.1f} %")
