"""Dump a NOBS_WHISPER_TRACE file as a flat timeline (block 0 of every instrumented decode kernel).
Kernel ids: 1 skinny GEMM, 2 split-K epilogue, 3 self-attention, 4 cross-attention, 5 tiled GEMM, 7 K6, 8 fused projection chain,
9 cluster projection (150 / 151 / 152: block 0 parked its tile / the cluster did / block 0 wrote its share);
100 + id: the moment the kernel's PDL wait returned; chain marks: 110 + s = all partial sums of step s are in (first barrier passed),
120 + s = step s reduced (second barrier passed).
usage: trace_dump.py file [t_lo_us t_hi_us]"""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
kid, tag, t0, t1 = a[:, 0].astype(int), a[:, 1], a[:, 2].astype(np.int64), a[:, 3].astype(np.int64)
ok = t0 > 0
kid, tag, t0, t1 = kid[ok], tag[ok], t0[ok], t1[ok]
base = t0.min()
names = {1: "skinny_gemm", 2: "reduce", 3: "self_attn", 4: "cross_attn", 5: "gemm", 7: "k6", 8: "chain", 9: "proj"}
lo, hi = (float(sys.argv[2]) * 1e3, float(sys.argv[3]) * 1e3) if len(sys.argv) > 3 else (0, 400e3)
order = np.argsort(t0)
lanes = {}
for i in order:
    t = t0[i] - base
    if t < lo or t > hi:
        continue
    k = kid[i]
    if k in names:
        nm, dur = names[k], (t1[i] - t0[i]) / 1e3 if t1[i] > 0 else float("nan")
        print(f"{t/1e3:10.2f}  {nm:12s} begin          dur {dur:8.2f}  tag {int(tag[i]) & 0xffffffff:08x}")
    elif 100 < k < 110:
        print(f"{t/1e3:10.2f}  {names.get(k-100, k):12s} wait-returned            tag {int(tag[i]) & 0xffffffff:08x}")
    elif 110 <= k < 120:
        print(f"{t/1e3:10.2f}  chain        step {k-110} partials in      tag {int(tag[i]) & 0xffffffff:08x}")
    elif 120 <= k < 130:
        print(f"{t/1e3:10.2f}  chain        step {k-120} reduced          tag {int(tag[i]) & 0xffffffff:08x}")
    elif 130 <= k < 140:
        print(f"{t/1e3:10.2f}  chain        step {k-130} block 0 GEMM done tag {int(tag[i]) & 0xffffffff:08x}")
    elif 140 <= k < 150:
        print(f"{t/1e3:10.2f}  chain        step {k-140} block 0 rows done tag {int(tag[i]) & 0xffffffff:08x}")
    elif 150 <= k < 153:
        what = {150: "tile parked", 151: "cluster parked", 152: "reduced + written"}[k]
        print(f"{t/1e3:10.2f}  proj         {what:24s} tag {int(tag[i]) & 0xffffffff:08x}")
for k in sorted(set(kid)):
    if k in names:
        m = (kid == k) & (t1 > 0)
        d = (t1[m] - t0[m]) / 1e3
        if len(d):
            print(f"# {names[k]:12s} n={len(d):7d} dur us: mean {d.mean():7.2f} p50 {np.median(d):7.2f} p90 {np.percentile(d, 90):7.2f}")
