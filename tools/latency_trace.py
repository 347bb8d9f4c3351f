"""Stage split and kernel timeline of one 5-s utterance on large-v3-turbo (BASELINE configs[4]).
usage: NOBS_WHISPER_TRACE=out.bin python tools/latency_trace.py   (the trace holds the LAST utterance's kernels when TRACE_SKIP is set)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nobs_whisper_b200 as nw  # noqa: E402
from nobs_whisper_b200 import ggml_synth, synth_audio  # noqa: E402

path = ggml_synth.ensure_model(os.environ.get("NOBS_TEST_MODEL_DIR", "/tmp/nobs_whisper_models"), "large-v3-turbo", seed=0, ftype=1, init="survey")
eng = nw.WhisperEngine()
eng.load_model(path)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ms = []
for i in range(n):
    c = synth_audio.synth_clip(5000 + i, 5.0)
    t0 = time.perf_counter()
    eng.transcribe(c, "en", None, None)
    ms.append(1e3 * (time.perf_counter() - t0))
    s = eng.last_stats()
    print(f"clip {i}: {ms[-1]:7.2f} ms wall; windows {s.n_windows} rounds {s.n_decode_rounds} rows {s.n_decode_rows} fallbacks {s.n_fallbacks} launches {s.n_kernel_launches}; "
          f"gpu ms mel {s.gpu_ms_mel:.2f} encode {s.gpu_ms_encode:.2f} decode {s.gpu_ms_decode:.2f}")
print("p50", float(np.percentile(ms[2:], 50)))
eng.close()
