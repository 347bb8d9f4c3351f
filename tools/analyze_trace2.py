"""Pair each instrumented kernel's launch entry with its 'dependency satisfied' entry and print the timeline:
launch time, time the PDL wait returned, end time (all of block 0), per kernel, in order of wait-return."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
kid, tag, t0, t1 = a[:, 0].astype(int), a[:, 1], a[:, 2].astype(np.int64), a[:, 3].astype(np.int64)
base = t0.min(); t0 = t0 - base; t1 = t1 - base
names = {1: "gemm_skinny", 2: "reduce", 3: "self_attn", 4: "cross_attn", 5: "gemm", 7: "process_logits"}
rows = []
for k in (1, 2, 3, 4, 7):
    for tg in np.unique(tag[kid == k]):
        b = np.where((kid == k) & (tag == tg))[0]; w = np.where((kid == 100 + k) & (tag == tg))[0]
        b = b[np.argsort(t0[b])]; w = w[np.argsort(t0[w])]
        # FIFO: kernels of one type on one stream launch and start in order
        ev = sorted([(t0[i], 0, i) for i in b] + [(t0[i], 1, i) for i in w])
        q = []
        for t, typ, i in ev:
            if typ == 0: q.append(i)
            elif q:
                bi = q.pop(0)
                rows.append((t0[i], k, int(tg) & 0xffffffff, t0[bi], t1[bi]))
rows.sort()
lo, hi = float(sys.argv[2]) * 1e3, float(sys.argv[3]) * 1e3
prev_end = {}
for tw, k, tg, tb, te in rows:
    if tw < lo or tw > hi: continue
    print(f"start {tw/1e3:10.2f}  {names[k]:12s} lane {tg:08x} launched {(tw-tb)/1e3:7.2f} us earlier, active {(te-tw)/1e3:7.2f} us, end {te/1e3:10.2f}")
# summary of active durations
import collections
act = collections.defaultdict(list)
for tw, k, tg, tb, te in rows: act[k].append((te - tw) / 1e3)
for k, v in act.items():
    v = np.array(v); print(names[k], "active us: mean %.2f p50 %.2f p90 %.2f" % (v.mean(), np.median(v), np.percentile(v, 90)), "n", len(v))
