"""Developer tool: opcode mix and hottest SASS windows of one kernel from `ncu --page source --csv --print-source sass`.

    ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name K > k.csv; python scripts/ncu_hotspots.py k.csv [block]
"""
import collections
import csv
import sys

import numpy as np

rows = list(csv.reader(open(sys.argv[1])))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = b["hdr"], b["data"]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
ex = np.array([int(r[ia]) for r in data])
sm = np.array([int(r[isamp]) for r in data])
tot, ts = ex.sum(), sm.sum()
print("%s: %d SASS instructions, %.3f G warp-instructions executed, %d samples" % (b["name"], len(data), tot / 1e9, ts))
op, ops = collections.Counter(), collections.Counter()
for r in data:
    s = r[isrc].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    o = s.split()[0].split(".")[0]
    op[o] += int(r[ia])
    ops[o] += int(r[isamp])
for o, c in op.most_common(22):
    print("  %-10s %6.2f%% of executed  %6.2f%% of samples" % (o, 100 * c / tot, 100 * ops[o] / ts))
W = 40
win = sorted(((ex[i:i + W].sum(), i) for i in range(0, len(ex), W)), reverse=True)
print("hottest %d-instruction windows:" % W)
for c, i in win[:20]:
    mix = collections.Counter(r[isrc].strip().split()[0 if not r[isrc].strip().startswith("@") else 1].split(".")[0] for r in data[i:i + W])
    print("  @%5d  %5.2f%% executed %5.2f%% samples   %s" % (i, 100 * c / tot, 100 * sm[i:i + W].sum() / ts,
                                                             " ".join("%s:%d" % kv for kv in mix.most_common(5))))
