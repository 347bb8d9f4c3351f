"""Per-lane view of a NOBS_WHISPER_TRACE file (multi-launch decode path): which lane ran what when, how long each kernel class
took, how long a lane's layer takes, and how much of the time two lanes' projection kernels overlapped.
usage: trace_lanes.py file [t_lo_us t_hi_us]"""
import sys
from collections import defaultdict

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
kid, tag, t0, t1 = a[:, 0].astype(int), a[:, 1], a[:, 2].astype(np.int64), a[:, 3].astype(np.int64)
ok = (t0 > 0) & (kid < 100) & (t1 > 0)
kid, tag, t0, t1 = kid[ok], tag[ok], t0[ok], t1[ok]
base = t0.min()
t0 = (t0 - base) / 1e3
t1 = (t1 - base) / 1e3
lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.0, 1e12)
names = {1: "gemm", 2: "reduce", 3: "self", 4: "cross", 7: "k6", 8: "chain", 9: "proj"}
# lane of a kernel: skinny GEMM / reduce carry the lane's split-K workspace, the attention kernels its attention buffer
lane_of = {}
for group in ((1, 2), (3, 4)):
    tags = sorted(set(int(t) for k, t in zip(kid, tag) if k in group))
    # the logits GEMM (kid 1) writes to the lane's logits buffer: keep only the most frequent tags (one per lane)
    cnt = defaultdict(int)
    for k, t in zip(kid, tag):
        if k in group:
            cnt[int(t)] += 1
    top = sorted(cnt, key=lambda t: -cnt[t])
    n_l = max(1, sum(1 for t in top if cnt[t] > 0.2 * cnt[top[0]]))
    for i, t in enumerate(sorted(top[:n_l])):
        lane_of[(group, t)] = i
ev = []
for k, t, b, e in zip(kid, tag, t0, t1):
    g = (1, 2) if k in (1, 2) else (3, 4) if k in (3, 4) else None
    ln = lane_of.get((g, int(t)), -1) if g else -1
    ev.append((b, e, k, ln))
ev.sort()
sel = [x for x in ev if lo <= x[0] <= hi]
n_lanes = 1 + max([x[3] for x in sel] + [0])
print(f"# {len(sel)} kernels in window, {n_lanes} lanes")
for ln in range(n_lanes):
    mine = [x for x in sel if x[3] == ln]
    cross = [x for x in mine if x[2] == 4]
    if len(cross) > 2:
        per = np.diff([c[0] for c in cross])
        per = per[per < 2000]
        print(f"lane {ln}: layer period (cross begin -> next cross begin) mean {per.mean():.1f} p50 {np.median(per):.1f} us over {len(per)} layers")
    for k in (1, 2, 3, 4, 7):
        d = np.array([e - b for b, e, kk, _ in mine if kk == k])
        if len(d):
            print(f"   {names[k]:7s} n={len(d):5d} block-0 residency mean {d.mean():7.2f} p50 {np.median(d):7.2f} p90 {np.percentile(d, 90):7.2f}")
# overlap of cross-attention between lanes and idle time of the cross-attention "stream"
cr = sorted((b, e) for b, e, k, _ in sel if k == 4)
if cr:
    span = cr[-1][1] - cr[0][0]
    busy, cur_b, cur_e = 0.0, cr[0][0], cr[0][1]
    for b, e in cr[1:]:
        if b > cur_e:
            busy += cur_e - cur_b
            cur_b, cur_e = b, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_b
    print(f"# some cross-attention block 0 resident {100 * busy / span:.1f} % of {span:.0f} us; sum of residencies {sum(e - b for b, e in cr) / span:.2f} x span")
for k in (9,):   # kernels whose tag does not identify the lane
    d = np.array([e - b for b, e, kk, _ in sel if kk == k])
    if len(d):
        print(f"all lanes {names[k]:7s} n={len(d):5d} block-0 residency mean {d.mean():7.2f} p50 {np.median(d):7.2f} p90 {np.percentile(d, 90):7.2f}")
if "--dump" in sys.argv:
    for b, e, k, ln in sel:
        print(f"{b:10.2f} {e - b:8.2f}  lane {ln:2d}  {names.get(k, k)}")
