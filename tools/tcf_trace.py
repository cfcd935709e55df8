"""one time step's event trace of the tensor-core kernel -> per-stage timeline (development).
On the GPU box:  VQWN_PROFILE=1 VQWN_TRACE_FILE=gpurun_out/tcf_trace.txt python tools/tc_time.py 64 2048 tc
Here:            python tools/tcf_trace.py gpurun_out/tcf_trace.txt [first_layer last_layer]
events: 1 step start | issuing warps 4-7: 2 first-chain weights ready, 3 gather complete, 4 gate chain issued, 5 residual +
skip chain issued, 6 taps landed, 7 tap chain issued | epilogue warps: 10 gate accumulator ready, 11 gate epilogue done,
12 gate slice published (warp 3), 13 residual + skip accumulator ready, 14 x staged (warp 0), 15 x published (warp 11) | loader lanes 8-9: 20 slot
free for chunk n, 21 chunk n requested | tap loader: 30 pair buffer free, 31 pair block requested"""
import sys
from collections import defaultdict
ev = defaultdict(dict)
rows = [tuple(int(x) for x in line.split()) for line in open(sys.argv[1])]
t0 = min(r[3] for r in rows)
l0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
l1 = int(sys.argv[3]) if len(sys.argv) > 3 else 13
names = {2: "weights ready", 3: "gather complete", 4: "gate chain issued", 5: "res+skip chain issued", 6: "taps landed", 7: "tap chain issued",
         10: "accA ready", 11: "gate epilogue done", 12: "gate published", 13: "accB ready", 14: "x staged", 15: "x published"}
by_layer = defaultdict(list)
for w, e, l, c in rows:
    if e in names:
        by_layer[l].append((c - t0, w, e))
ref = None
for l in range(l0, l1 + 1):
    evs = sorted(by_layer[l])
    g = [c for c, w, e in evs if e == 3]
    base = min(g) if g else evs[0][0]
    print("layer %d (gather complete at +%d since step start%s)" % (l, base, "" if ref is None else ", stage length %d" % (base - ref)))
    ref = base
    for c, w, e in evs:
        print("   %+7d  warp %2d  %s" % (c - base, w, names[e]))
ld = sorted((c - t0, w, e, l) for w, e, l, c in rows if e in (20, 21))
print("loader lanes: chunk n -> (slot free, requested) relative to step start")
last = {}
for c, w, e, n in ld:
    last.setdefault((w, n), {})[e] = c
for (w, n), d in sorted(last.items(), key=lambda kv: kv[1].get(20, 0))[:60]:
    print("   warp %d chunk %3d: free %7d requested %7d" % (w, n, d.get(20, -1), d.get(21, -1)))
