"""whisper-rs 0.15 API surface used by the reference (src-tauri/src/whisper.rs:3), as ctypes
wrappers over the C ABI: WhisperContextParameters, WhisperContext, WhisperState, FullParams,
SamplingStrategy, WhisperSegment.  Names and argument meaning follow the Rust crate so that
the parity tests read like the reference's call sequence (whisper.rs:39-45, 83-141)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class WhisperError(Exception):
    """whisper-rs `WhisperError` (only ever stringified by the reference: whisper.rs:45,85,129)."""


PRECISION = {"default": 0, "fp32": 1, "bf16": 2}


class WhisperContextParameters:
    def __init__(self):
        self._p = _lib.lib().whisper_context_default_params()

    @classmethod
    def default(cls) -> "WhisperContextParameters":
        return cls()

    def use_gpu(self, v: bool) -> "WhisperContextParameters":  # whisper.rs:40
        self._p.use_gpu = bool(v)
        return self

    def gpu_device(self, d: int) -> "WhisperContextParameters":
        self._p.gpu_device = int(d)
        return self


@dataclass
class SamplingStrategy:
    kind: str
    best_of: int = 1
    beam_size: int = 5
    patience: float = -1.0

    @classmethod
    def Greedy(cls, best_of: int = 1) -> "SamplingStrategy":  # whisper.rs:88
        return cls("greedy", best_of=best_of)

    @classmethod
    def BeamSearch(cls, beam_size: int = 5, patience: float = -1.0) -> "SamplingStrategy":
        return cls("beam", beam_size=beam_size, patience=patience)


class FullParams:
    """whisper-rs `FullParams`: whisper_full_default_params(strategy) plus setters."""

    def __init__(self, strategy: SamplingStrategy):
        L = _lib.lib()
        if strategy.kind == "greedy":
            self._p = L.whisper_full_default_params(0)
            self._p.greedy.best_of = strategy.best_of
        else:
            self._p = L.whisper_full_default_params(1)
            self._p.beam_search.beam_size = strategy.beam_size
            self._p.beam_search.patience = strategy.patience
        self._keep = {}

    @classmethod
    def new(cls, strategy: SamplingStrategy) -> "FullParams":
        return cls(strategy)

    def _set_str(self, field: str, v):
        b = None if v is None else (v.encode() if isinstance(v, str) else bytes(v))
        self._keep[field] = b  # the params borrow the string for the duration of `full`
        setattr(self._p, field, b)

    def set_language(self, lang):  # whisper.rs:91-95; None => auto-detect
        self._set_str("language", lang)

    def set_initial_prompt(self, prompt: str):  # whisper.rs:106-107
        self._set_str("initial_prompt", prompt)

    def set_print_special(self, v): self._p.print_special = bool(v)          # whisper.rs:112
    def set_print_progress(self, v): self._p.print_progress = bool(v)        # whisper.rs:113
    def set_print_realtime(self, v): self._p.print_realtime = bool(v)        # whisper.rs:114
    def set_print_timestamps(self, v): self._p.print_timestamps = bool(v)    # whisper.rs:115
    def set_translate(self, v): self._p.translate = bool(v)                  # whisper.rs:116
    def set_no_context(self, v): self._p.no_context = bool(v)                # whisper.rs:117
    def set_single_segment(self, v): self._p.single_segment = bool(v)        # whisper.rs:118
    def set_suppress_blank(self, v): self._p.suppress_blank = bool(v)        # whisper.rs:121
    def set_no_speech_thold(self, v): self._p.no_speech_thold = float(v)     # whisper.rs:122
    def set_entropy_thold(self, v): self._p.entropy_thold = float(v)         # whisper.rs:123
    def set_logprob_thold(self, v): self._p.logprob_thold = float(v)         # whisper.rs:124
    def set_temperature(self, v): self._p.temperature = float(v)
    def set_temperature_inc(self, v): self._p.temperature_inc = float(v)
    def set_no_timestamps(self, v): self._p.no_timestamps = bool(v)
    def set_max_initial_ts(self, v): self._p.max_initial_ts = float(v)
    def set_length_penalty(self, v): self._p.length_penalty = float(v)
    def set_n_max_text_ctx(self, v): self._p.n_max_text_ctx = int(v)
    def set_max_tokens(self, v): self._p.max_tokens = int(v)
    def set_detect_language(self, v): self._p.detect_language = bool(v)
    def set_offset_ms(self, v): self._p.offset_ms = int(v)
    def set_duration_ms(self, v): self._p.duration_ms = int(v)


class WhisperSegment:
    def __init__(self, state: "WhisperState", i: int):
        self._s, self._i = state, i

    def to_bytes(self) -> bytes:
        L = _lib.lib()
        p = L.whisper_full_get_segment_text_from_state(self._s._h, self._i)
        return C.string_at(p) if p else b""

    def to_str_lossy(self) -> str:  # whisper.rs:137
        return self.to_bytes().decode("utf-8", errors="replace")

    def start_timestamp(self) -> int:
        return _lib.lib().whisper_full_get_segment_t0_from_state(self._s._h, self._i)

    def end_timestamp(self) -> int:
        return _lib.lib().whisper_full_get_segment_t1_from_state(self._s._h, self._i)

    def no_speech_probability(self) -> float:
        return _lib.lib().whisper_full_get_segment_no_speech_prob_from_state(self._s._h, self._i)

    def n_tokens(self) -> int:
        return _lib.lib().whisper_full_n_tokens_from_state(self._s._h, self._i)

    def token_ids(self) -> list[int]:
        L = _lib.lib()
        return [L.whisper_full_get_token_id_from_state(self._s._h, self._i, t) for t in range(self.n_tokens())]

    def token_data(self, t: int):
        return _lib.lib().whisper_full_get_token_data_from_state(self._s._h, self._i, t)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class WhisperContext:
    """whisper-rs `WhisperContext` (Send + Sync in Rust; here the C side serialises GPU work)."""

    def __init__(self, path: str, params: WhisperContextParameters | None = None, precision: str = "default", host_only: bool = False):
        L = _lib.lib()
        params = params or WhisperContextParameters()
        if host_only:  # vocabulary / tokenizer queries only; compute entry points fail on this handle
            self._h = L.whisper_b200_init_host_only(path.encode())
        else:
            self._h = L.whisper_b200_init_from_file(path.encode(), params._p, PRECISION[precision])
        if not self._h:
            raise WhisperError("InitError: " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))

    @classmethod
    def new_with_params(cls, path: str, params: WhisperContextParameters, precision: str = "default") -> "WhisperContext":
        return cls(path, params, precision)  # whisper.rs:41-45

    def create_state(self) -> "WhisperState":  # whisper.rs:83-85
        return WhisperState(self)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().whisper_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # vocabulary / model queries
    def n_vocab(self): return _lib.lib().whisper_n_vocab(self._h)
    def n_audio_ctx(self): return _lib.lib().whisper_n_audio_ctx(self._h)
    def n_audio_state(self): return _lib.lib().whisper_model_n_audio_state(self._h)
    def n_text_state(self): return _lib.lib().whisper_model_n_text_state(self._h)
    def n_text_layer(self): return _lib.lib().whisper_model_n_text_layer(self._h)
    def n_mels(self): return _lib.lib().whisper_model_n_mels(self._h)
    def precision(self): return {1: "fp32", 2: "bf16"}.get(_lib.lib().whisper_b200_precision(self._h))
    def decode_lanes(self) -> int: return int(_lib.lib().whisper_b200_decode_lanes(self._h))
    def token_eot(self): return _lib.lib().whisper_token_eot(self._h)
    def token_sot(self): return _lib.lib().whisper_token_sot(self._h)
    def token_beg(self): return _lib.lib().whisper_token_beg(self._h)
    def token_transcribe(self): return _lib.lib().whisper_token_transcribe(self._h)
    def token_nosp(self): return _lib.lib().whisper_token_nosp(self._h)
    def token_lang(self, i): return _lib.lib().whisper_token_lang(self._h, i)

    def token_to_bytes(self, tid: int) -> bytes:
        p = _lib.lib().whisper_token_to_str(self._h, tid)
        return C.string_at(p) if p else b""

    def tokenize(self, text) -> list[int]:
        L = _lib.lib()
        b = text.encode() if isinstance(text, str) else bytes(text)
        n = L.whisper_token_count(self._h, b)
        buf = (C.c_int32 * max(n, 1))()
        m = L.whisper_tokenize(self._h, b, buf, n)
        if m < 0:
            raise WhisperError("tokenize failed")
        return list(buf[:m])

    def set_logits_hook(self, fn):
        """Scripted-logits test hook (whisper_b200_set_logits_hook): fn(seek, i_temp, step, decoder, n_prompt, logits: np.ndarray view)
        returns truthy if it overwrote the logits of that sample row; None removes the hook."""
        L = _lib.lib()
        if fn is None:
            L.whisper_b200_set_logits_hook(self._h, C.cast(None, _lib.LOGITS_HOOK), None)
            self._hook = None
            return

        def tramp(_user, seek, it, step, dec, n_prompt, n_vocab, ptr):
            return 1 if fn(seek, it, step, dec, n_prompt, np.ctypeslib.as_array(ptr, shape=(n_vocab,))) else 0

        self._hook = _lib.LOGITS_HOOK(tramp)     # keep the trampoline alive
        L.whisper_b200_set_logits_hook(self._h, self._hook, None)

    def process_logits(self, params: FullParams, logits, hist=(), has_ts=False, seek_delta=3000, temperature=0.0, mode=0, u=0.0, k=0):
        """Stage hook: K6 (filter + log-softmax + sampling) on explicit logits."""
        L = _lib.lib()
        lg = _f32(logits)
        n = len(lg)
        lp = np.zeros(n, np.float32)
        pr = np.zeros(n, np.float32)
        h = np.ascontiguousarray(hist, dtype=np.int32)
        res = _lib.B200SampleResult()
        rc = L.whisper_b200_process_logits(self._h, params._p, _fptr(lg), h.ctypes.data_as(C.POINTER(C.c_int32)), len(h), int(has_ts),
                                           int(seek_delta), float(temperature), int(mode), float(u), int(k), _fptr(lp), _fptr(pr), C.byref(res))
        if rc != 0:
            raise WhisperError("process_logits failed: " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))
        return lp, pr, res


class WhisperState:
    def __init__(self, ctx: WhisperContext):
        L = _lib.lib()
        self._ctx = ctx
        self._h = L.whisper_init_state(ctx._h)
        if not self._h:
            raise WhisperError("InitError: failed to create state")

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            _lib.lib().whisper_free_state(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def full(self, params: FullParams, audio) -> int:  # whisper.rs:127-129
        L = _lib.lib()
        a = _f32(audio)
        if a.size == 0:
            raise WhisperError("NoSamples: Input sample buffer was empty.")
        rc = L.whisper_full_with_state(self._ctx._h, self._h, params._p, _fptr(a), int(a.size))
        if rc != 0:
            name = {-1: "UnableToCalculateSpectrogram", 7: "FailedToEncode", 8: "FailedToDecode"}.get(rc, f"GenericError({rc})")
            raise WhisperError(name + ": " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))
        return rc

    def full_n_segments(self) -> int:  # whisper.rs:132
        return _lib.lib().whisper_full_n_segments_from_state(self._h)

    def get_segment(self, i: int):  # whisper.rs:136
        return WhisperSegment(self, i) if 0 <= i < self.full_n_segments() else None

    def full_lang_id(self) -> int:
        return _lib.lib().whisper_full_lang_id_from_state(self._h)

    def stats(self) -> "_lib.B200Stats":
        s = _lib.B200Stats()
        _lib.lib().whisper_b200_get_stats(self._h, C.byref(s))
        return s

    # ---- stage-level API (upstream whisper.h names) used by the parity tests
    def pcm_to_mel(self, audio) -> None:
        a = _f32(audio)
        if _lib.lib().whisper_pcm_to_mel_with_state(self._ctx._h, self._h, _fptr(a), int(a.size), 1) != 0:
            raise WhisperError("UnableToCalculateSpectrogram")

    def n_len(self) -> int:
        return _lib.lib().whisper_n_len_from_state(self._h)

    def get_mel(self) -> np.ndarray:
        L = _lib.lib()
        need = -L.whisper_b200_get_mel(self._h, None, 0)
        out = np.zeros(need, np.float32)
        n_len = L.whisper_b200_get_mel(self._h, _fptr(out), need)
        if n_len <= 0:
            raise WhisperError("no mel")
        return out.reshape(-1, n_len)

    def encode(self, offset: int = 0) -> None:
        if _lib.lib().whisper_encode_with_state(self._ctx._h, self._h, int(offset), 1) != 0:
            raise WhisperError("FailedToEncode: " + (_lib.lib().whisper_b200_last_error() or b"").decode(errors="replace"))

    def encoder_output(self) -> np.ndarray:
        c = self._ctx
        out = np.zeros((c.n_audio_ctx(), c.n_audio_state()), np.float32)
        if _lib.lib().whisper_b200_get_encoder_output(c._h, self._h, _fptr(out), out.size) != 0:
            raise WhisperError("no encoder output")
        return out

    def cross_kv(self, layer: int):
        c = self._ctx
        k = np.zeros((c.n_audio_ctx(), c.n_text_state()), np.float32)
        v = np.zeros_like(k)
        if _lib.lib().whisper_b200_get_cross_kv(c._h, self._h, layer, _fptr(k), _fptr(v), k.size) != 0:
            raise WhisperError("no cross kv")
        return k, v

    def decode(self, tokens, n_past: int = 0) -> np.ndarray:
        L = _lib.lib()
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        if L.whisper_decode_with_state(self._ctx._h, self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), len(t), int(n_past), 1) != 0:
            raise WhisperError("FailedToDecode: " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))
        p = L.whisper_get_logits_from_state(self._h)
        return np.ctypeslib.as_array(p, shape=(self._ctx.n_vocab(),)).copy()

    def lang_auto_detect(self):
        probs = np.zeros(100, np.float32)
        lid = _lib.lib().whisper_lang_auto_detect_with_state(self._ctx._h, self._h, 0, 1, _fptr(probs))
        if lid < 0:
            raise WhisperError(f"lang detect failed ({lid})")
        return lid, probs

    def segments(self) -> list[dict]:
        out = []
        for i in range(self.full_n_segments()):
            s = self.get_segment(i)
            out.append({"t0": s.start_timestamp(), "t1": s.end_timestamp(), "text": s.to_bytes(), "tokens": s.token_ids()})
        return out


def decode_batch(ctx: WhisperContext, states: list[WhisperState], tokens: list, n_past: list[int], lane: int = 0) -> np.ndarray:
    """B200 extension: one decoder round over several encoded states (whisper_b200_decode_batch); tokens[i] are the rows
    state i contributes at positions n_past[i]..; returns the last-row logits [len(states)][n_vocab]."""
    L = _lib.lib()
    n = len(states)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.int32).ravel() for t in tokens]), dtype=np.int32)
    nt = np.ascontiguousarray([len(t) for t in tokens], dtype=np.int32)
    npast = np.ascontiguousarray(n_past, dtype=np.int32)
    handles = (C.c_void_p * n)(*[s._h for s in states])
    out = np.zeros((n, ctx.n_vocab()), np.float32)
    rc = L.whisper_b200_decode_batch(ctx._h, handles, n, flat.ctypes.data_as(C.POINTER(C.c_int32)), nt.ctypes.data_as(C.POINTER(C.c_int32)),
                                     npast.ctypes.data_as(C.POINTER(C.c_int32)), int(lane), _fptr(out))
    if rc != 0:
        raise WhisperError("FailedToDecode: " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))
    return out


def full_batch(ctx: WhisperContext, states: list[WhisperState], params: FullParams, audios: list) -> list[int]:
    """B200 extension: `full` over independent audios in lock step (whisper_b200_full_batch)."""
    L = _lib.lib()
    n = len(states)
    arrs = [_f32(a) for a in audios]
    st = (C.c_void_p * n)(*[s._h for s in states])
    ptrs = (C.POINTER(C.c_float) * n)(*[_fptr(a) for a in arrs])
    ns = (C.c_int * n)(*[int(a.size) for a in arrs])
    rc = (C.c_int * n)()
    r = L.whisper_b200_full_batch(ctx._h, st, n, params._p, ptrs, ns, rc)
    if r != 0:
        raise WhisperError(f"full_batch failed ({r}): " + (L.whisper_b200_last_error() or b"").decode(errors="replace"))
    return list(rc)
