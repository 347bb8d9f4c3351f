"""Micro-benchmark of the device-wide barrier (whisper_b200_debug_grid_sync): us per barrier by variant, CTA count and store burst."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from nobs_whisper_b200 import _lib
L = _lib.lib()
us = C.c_float(0)
for ctas in (148, 100, 32):
    for store in (0, 16, 55, 110):
        row = []
        for v in (0, 1, 2):
            rc = L.whisper_b200_debug_grid_sync(ctas, 2000, v, store, C.byref(us))
            row.append("%.2f" % us.value if rc == 0 else "rc%d" % rc)
        print(f"ctas {ctas:3d} store {store:3d} floats/thread: variant0 {row[0]} us, variant1 {row[1]} us, variant2 {row[2]} us", flush=True)
