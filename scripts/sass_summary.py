"""Instruction evidence for the tensor-core / TMA path: per kernel of csrc/librgcn_b200.so, the counts of the SASS
mnemonics B200_PROFILING.md names (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, FADD2 / FFMA2 = packed fp32, LDG.E.128 = 128-bit global loads, ATOMG / RED =
global atomics).  usage: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "primekg-rgcn-linkprediction_b200", "csrc", "librgcn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "FADD2", "FFMA2", "LDG.E.128", "LDG.E.64", "STG.E.128",
        "ATOMG", "RED.", "ATOMS", "F2FP", "HMMA"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("rgcn::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k) or (k.endswith(".") is False and k in op and k.startswith("LDG") is False and op.startswith(k)):
                counts[cur][k] += 1
        for k in ("LDG.E.128", "LDG.E.64", "STG.E.128"):
            if op.startswith(k.split(".")[0]) and k.split(".", 1)[1] in op and not op.startswith(k):
                counts[cur][k] += 1
print("# cuobjdump -sass csrc/librgcn_b200.so: instruction counts per kernel (sm_100a)")
print("# %-62s %7s  %s" % ("kernel", "instrs", "mnemonics"))
for k, c in counts.items():
    items = " ".join(f"{n}={c[n]}" for n in KEYS if c[n])
    print("%-64s %7d  %s" % (k[:64], c["_total"], items))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# total: " + " ".join(f"{n}={tot[n]}" for n in KEYS if tot[n]))
