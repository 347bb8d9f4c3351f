"""Python mirror of the reference's silence chunker API (src-tauri/src/audio.rs:400-507) over the
library's C ABI.  The RMS scan runs on the GPU (csrc/audio_chunker.cu); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

WHISPER_SAMPLE_RATE = 16000   # audio.rs:7
CHUNK_OVERLAP_MS = 200        # audio.rs:15


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {_lib.lib().whisper_b200_last_error().decode(errors='replace')}")


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def window_rms(audio, window: int) -> np.ndarray:
    """calculate_rms (audio.rs:364-370) of every full `window`-sample block, bit-identical to the reference."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    n = C.c_size_t(0)
    out = np.zeros(max(1, a.size // max(window, 1)), np.float32)
    _check(_lib.lib().whisper_b200_window_rms(_fp(a), a.size, window, _fp(out), out.size, C.byref(n)), "window_rms")
    return out[: n.value]


def find_silence_boundaries(audio, sample_rate: int) -> list[int]:
    """audio.rs:400-463."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    cap = a.size // max(sample_rate, 1) + 2        # boundaries are >= 1 s apart
    out = (C.c_size_t * cap)()
    n = C.c_size_t(0)
    _check(_lib.lib().nobs_find_silence_boundaries(_fp(a), a.size, sample_rate, out, cap, C.byref(n)), "find_silence_boundaries")
    return [int(out[i]) for i in range(min(n.value, cap))]


def split_at_silences_with_overlap(audio, boundaries, sample_rate: int) -> list[np.ndarray]:
    """audio.rs:473-507."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    nb = len(boundaries)
    b = (C.c_size_t * max(nb, 1))(*[int(x) for x in boundaries])
    ranges = (C.c_size_t * (2 * (nb + 1)))()
    n = C.c_size_t(0)
    _check(_lib.lib().nobs_split_at_silences_with_overlap(a.size, b, nb, sample_rate, ranges, C.byref(n)), "split_at_silences")
    return [a[int(ranges[2 * k]): int(ranges[2 * k + 1])].copy() for k in range(n.value)]


def split_at_silences(audio, boundaries) -> list[np.ndarray]:
    """audio.rs:467-469."""
    return split_at_silences_with_overlap(audio, boundaries, WHISPER_SAMPLE_RATE)
