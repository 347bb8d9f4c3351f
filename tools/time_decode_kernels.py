"""Device time per launch of the decoder-step kernels (micro-benchmark through the debug hook)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nobs_whisper_b200 import _lib  # noqa: E402

NAMES = ["skinny QKV (N=3d,K=d)", "skinny out (N=d,K=d)", "skinny FC1 (N=4d,K=d)", "skinny FC2 (N=d,K=4d)", "reduce plain N=3d",
         "reduce resid+LN N=d", "reduce gelu N=4d", "self-attn (100 keys)", "cross-attn (1500 keys)", "generic GEMM N=d,K=d", "layernorm", "cross-attn bulk stream SIMT", "cross-attn stream tcgen05"]
L = _lib.lib()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 120
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
out = np.zeros(16, np.float32)
rc = L.whisper_b200_debug_time_decode_kernels(R, d, 200, out.ctypes.data_as(C.POINTER(C.c_float)))
assert rc == 0, (rc, L.whisper_b200_last_error())
for n, v in zip(NAMES, out):
    print(f"{n:28s} {v:8.2f} us")
layer = out[0] + out[4] + out[7] + out[1] + out[5] + out[1] + out[5] * 0 + out[4] * 0 + out[8] + out[1] + out[5] + out[2] + out[6] + out[3] + out[5]
print(f"approx per-layer sum {layer:.1f} us")
