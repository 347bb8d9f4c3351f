"""Read a NOBS_WHISPER_TRACE dump: entries {kernel id, tag, start ns, end ns} written by block 0 of the
instrumented decode kernels (ids: 1 skinny GEMM, 2 split-K epilogue, 3 self-attention, 4 cross-attention,
5 tiled GEMM; +100 = the moment the kernel's programmatic-dependent-launch wait returned)."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
kid, tag, t0, t1 = a[:, 0].astype(int), a[:, 1], a[:, 2].astype(np.int64), a[:, 3].astype(np.int64)
base = t0.min()
t0 -= base; t1 -= base
names = {1: "skinny_gemm", 2: "reduce", 3: "self_attn", 4: "cross_attn", 5: "gemm"}
print("entries", len(a), "span ms", (t1.max()) / 1e6)
for k in sorted(set(kid)):
    m = kid == k
    if k < 100:
        d = (t1[m] - t0[m]) / 1e3
        print(f"{names.get(k,k):12s} n={m.sum():7d} dur us: mean {d.mean():7.2f} p50 {np.median(d):7.2f} p90 {np.percentile(d,90):7.2f} max {d.max():8.2f}")
# time from kernel start to pdl wait return: join 100+k entries with k entries by tag & nearest start
order = np.argsort(t0)
kid, tag, t0, t1 = kid[order], tag[order], t0[order], t1[order]
if len(sys.argv) > 2:
    lo = float(sys.argv[2]) * 1e3; hi = float(sys.argv[3]) * 1e3
    tags = {}
    for i in range(len(kid)):
        if t0[i] < lo or t0[i] > hi: continue
        # lane id by cross/self-attn out pointer or partial pointer: group tags by appearance
        print(f"{t0[i]/1e3:10.2f} {('+wait ' if kid[i]>=100 else '') + names.get(kid[i]%100, str(kid[i])):18s} dur {(t1[i]-t0[i])/1e3:8.2f} tag {int(tag[i])&0xffffffff:08x}")
