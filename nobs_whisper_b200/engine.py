"""Python mirror of the reference's WhisperEngine (src-tauri/src/whisper.rs:16-260) over the
C++ host wrapper exported by the library (csrc/host/whisper_engine.cpp, nobs_engine_* shims)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class WhisperEngineError(Exception):
    pass


class LoadError(WhisperEngineError):  # whisper.rs:8-9
    pass


class TranscriptionError(WhisperEngineError):  # whisper.rs:10-11
    pass


class NoModel(WhisperEngineError):  # whisper.rs:12-13
    pass


_ERR = {1: LoadError, 2: TranscriptionError, 3: NoModel}


def filter_hallucinations(text: str) -> str:  # whisper.rs:233-260
    return _lib.lib().nobs_filter_hallucinations(text.encode()).decode()


class WhisperEngine:
    def __init__(self):  # whisper.rs:22-27
        self._L = _lib.lib()
        self._h = self._L.nobs_engine_new()

    @classmethod
    def from_file(cls, model_path: str) -> "WhisperEngine":  # whisper.rs:30-34
        e = cls()
        e.load_model(model_path)
        return e

    def _raise(self, kind: int):
        raise _ERR[kind](self._L.nobs_engine_last_error(self._h).decode(errors="replace"))

    def load_model(self, model_path: str) -> None:  # whisper.rs:36-52
        k = self._L.nobs_engine_load_model(self._h, str(model_path).encode())
        if k:
            self._raise(k)

    def unload_model(self) -> None:  # whisper.rs:55-59
        self._L.nobs_engine_unload_model(self._h)

    def is_loaded(self) -> bool:  # whisper.rs:62-64
        return bool(self._L.nobs_engine_is_loaded(self._h))

    @staticmethod
    def _s(v):
        return None if v is None else v.encode()

    def transcribe(self, audio, language=None, vocabulary=None, context=None) -> str:  # whisper.rs:66-148
        a = np.ascontiguousarray(audio, dtype=np.float32)
        out = C.c_char_p()
        k = self._L.nobs_engine_transcribe(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), int(a.size), self._s(language),
                                           self._s(vocabulary), self._s(context), C.byref(out))
        if k:
            self._raise(k)
        return (out.value or b"").decode("utf-8", errors="replace")

    def transcribe_chunked(self, chunks, language=None, vocabulary=None) -> str:  # whisper.rs:152-197
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in chunks]
        n = len(arrs)
        ptrs = (C.POINTER(C.c_float) * max(n, 1))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        ns = (C.c_int * max(n, 1))(*[int(a.size) for a in arrs])
        out = C.c_char_p()
        k = self._L.nobs_engine_transcribe_chunked(self._h, ptrs, ns, n, self._s(language), self._s(vocabulary), C.byref(out))
        if k:
            self._raise(k)
        return (out.value or b"").decode("utf-8", errors="replace")

    def transcribe_chunked_parallel(self, chunks, language=None, vocabulary=None, abort_on_error: bool = True):
        """whisper.rs:152-197 decoded data-parallel: the same text as transcribe_chunked — every chunk's prompt carries
        the previous non-empty chunk's text — reached by speculative batches.  Returns (text, n_decodes, n_rounds)."""
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in chunks]
        n = len(arrs)
        ptrs = (C.POINTER(C.c_float) * max(n, 1))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        ns = (C.c_int * max(n, 1))(*[int(a.size) for a in arrs])
        out = C.c_char_p()
        nd, nr = C.c_int(0), C.c_int(0)
        k = self._L.nobs_engine_transcribe_chunked_parallel(self._h, ptrs, ns, n, self._s(language), self._s(vocabulary), int(abort_on_error),
                                                            C.byref(out), C.byref(nd), C.byref(nr))
        if k:
            self._raise(k)
        return (out.value or b"").decode("utf-8", errors="replace"), nd.value, nr.value

    def transcribe_recording(self, audio, language=None, vocabulary=None, parallel=False) -> str:  # state.rs:757-792
        """The audio left when a recording stops: longer than 30 s it is cut at silences (audio.rs) and the pieces
        are transcribed in order with the previous text as context (parallel=True / 1: together, no chaining; parallel=2 or "chained":
        together WITH the chaining, same text as the sequential loop)."""
        parallel = 2 if parallel == "chained" else int(parallel)
        a = np.ascontiguousarray(audio, dtype=np.float32)
        out = C.c_char_p()
        k = self._L.nobs_engine_transcribe_recording(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), int(a.size), self._s(language),
                                                     self._s(vocabulary), int(parallel), C.byref(out))
        if k:
            self._raise(k)
        return (out.value or b"").decode("utf-8", errors="replace")

    def transcribe_batch(self, audios, language=None, vocabulary=None, beam_size: int = 0) -> list[str]:
        """Independent windows decoded together on the GPU; no context chaining."""
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in audios]
        n = len(arrs)
        if n == 0:
            return []
        ptrs = (C.POINTER(C.c_float) * n)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        ns = (C.c_int * n)(*[int(a.size) for a in arrs])
        texts = (C.c_char_p * n)()
        k = self._L.nobs_engine_transcribe_batch(self._h, ptrs, ns, n, self._s(language), self._s(vocabulary), int(beam_size), texts)
        if k:
            self._raise(k)
        return [(t or b"").decode("utf-8", errors="replace") for t in texts]

    def last_stats(self) -> "_lib.B200Stats":
        s = _lib.B200Stats()
        self._L.nobs_engine_last_stats(self._h, C.byref(s))
        return s

    def n_lanes(self) -> int:
        """Decode lanes of the loaded context (0 without a model)."""
        ctx = self._L.nobs_engine_context(self._h)
        return int(self._L.whisper_b200_decode_lanes(ctx)) if ctx else 0

    def set_profiling(self, on: bool) -> None:
        ctx = self._L.nobs_engine_context(self._h)
        if ctx:
            self._L.whisper_b200_set_profiling(ctx, int(on))

    def event_record(self, slot: int) -> None:
        """CUDA event on the library's stream (device-side timing of whole calls)."""
        self._L.whisper_b200_event_record(self._L.nobs_engine_context(self._h), slot)

    def event_elapsed_ms(self, a: int, b: int) -> float:
        return self._L.whisper_b200_event_elapsed_ms(self._L.nobs_engine_context(self._h), a, b)

    def transcribe_batch_ptrs(self, ptrs, lengths, language=None, vocabulary=None, beam_size: int = 0) -> list[str]:
        """transcribe_batch over raw addresses (pinned host or device memory), e.g. torch tensors' data_ptr()."""
        n = len(ptrs)
        p = (C.POINTER(C.c_float) * n)(*[C.cast(C.c_void_p(int(a)), C.POINTER(C.c_float)) for a in ptrs])
        ns = (C.c_int * n)(*[int(x) for x in lengths])
        texts = (C.c_char_p * n)()
        k = self._L.nobs_engine_transcribe_batch(self._h, p, ns, n, self._s(language), self._s(vocabulary), int(beam_size), texts)
        if k:
            self._raise(k)
        return [(t or b"").decode("utf-8", errors="replace") for t in texts]

    def close(self):
        if getattr(self, "_h", None):
            self._L.nobs_engine_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
