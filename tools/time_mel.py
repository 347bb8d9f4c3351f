"""Device time of the log-mel kernel for a batch of 30-s windows (library stats: CUDA events around the launch)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nobs_whisper_b200 as nw
from nobs_whisper_b200 import ggml_synth, synth_audio
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
ctx = nw.WhisperContext.new_with_params(ggml_synth.ensure_model("/tmp/nobs_whisper_models", "micro", init="fanin"), nw.WhisperContextParameters.default(), precision="bf16")
pcm = np.concatenate([synth_audio.synth_clip(i, 30.0) for i in range(8)])
import time
for rep in range(3):
    st = ctx.create_state()
    big = np.tile(pcm, n // 8)
    t0 = time.perf_counter(); st.pcm_to_mel(big); dt = time.perf_counter() - t0
    print(f"pcm_to_mel of {len(big)/16000:.0f} s of audio: {dt*1e3:.2f} ms wall (incl. H2D of {big.nbytes/1e6:.0f} MB)")
    st.close()
