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


def resample_audio(audio, from_rate: int, to_rate: int) -> np.ndarray:
    """audio.rs:509-563 (rubato FftFixedIn(from, to, 1024, 2, 1) restated as one GEMM per recording on the GPU)."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    n = C.c_size_t(0)
    L = _lib.lib()
    _check(L.nobs_resample_audio(_fp(a), a.size, from_rate, to_rate, None, 0, C.byref(n)), "resample_audio")
    out = np.zeros(max(n.value, 1), np.float32)
    if n.value:
        _check(L.nobs_resample_audio(_fp(a), a.size, from_rate, to_rate, _fp(out), out.size, C.byref(n)), "resample_audio")
    return out[: n.value]


def resample_chunk(audio, input_sample_rate: int) -> np.ndarray:
    """audio.rs:329-334."""
    if input_sample_rate == WHISPER_SAMPLE_RATE:
        return np.ascontiguousarray(audio, dtype=np.float32).copy()
    return resample_audio(audio, input_sample_rate, WHISPER_SAMPLE_RATE)


def mix_to_mono(interleaved, channels: int) -> np.ndarray:
    """state.rs:590-594."""
    a = np.ascontiguousarray(interleaved, dtype=np.float32)
    frames = a.size // channels
    out = np.zeros(max(frames, 1), np.float32)
    _check(_lib.lib().nobs_mix_to_mono(_fp(a), frames, channels, _fp(out)), "mix_to_mono")
    return out[:frames]


def calculate_rms(samples) -> float:
    """audio.rs:364-370 (host)."""
    a = np.ascontiguousarray(samples, dtype=np.float32)
    return float(_lib.lib().nobs_calculate_rms(_fp(a), a.size))


class AudioBuffer:
    """audio.rs:29-244: the streaming capture buffer (host logic of the library, csrc/host/audio_buffer.cpp)."""

    def __init__(self, sample_rate: int = 48000):
        self._L = _lib.lib()
        self._h = self._L.nobs_audio_buffer_new(sample_rate)

    @classmethod
    def with_sample_rate(cls, sample_rate: int) -> "AudioBuffer":
        return cls(sample_rate)

    def push_samples(self, samples) -> None:
        a = np.ascontiguousarray(samples, dtype=np.float32)
        self._L.nobs_audio_buffer_push_samples(self._h, _fp(a), a.size)

    def has_silence_boundary(self) -> bool:
        return bool(self._L.nobs_audio_buffer_has_silence_boundary(self._h))

    def _take(self, fn):
        n = C.c_size_t(0)
        p = fn(self._h, C.byref(n))
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)

    def take_chunk_at_silence(self):
        return self._take(self._L.nobs_audio_buffer_take_chunk_at_silence)

    def take_forced_chunk(self):
        return self._take(self._L.nobs_audio_buffer_take_forced_chunk)

    def take(self):
        return self._take(self._L.nobs_audio_buffer_take)

    def __len__(self) -> int:
        return int(self._L.nobs_audio_buffer_len(self._h))

    @property
    def overlap_len(self) -> int:
        return int(self._L.nobs_audio_buffer_overlap_len(self._h))

    def get_noise_floor(self) -> float:
        return float(self._L.nobs_audio_buffer_noise_floor(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.nobs_audio_buffer_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
