"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY — see whisper_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "whisper_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class WoParams(C.Structure):
    _fields_ = [
        ("strategy", C.c_int), ("best_of", C.c_int), ("beam_size", C.c_int),
        ("language", C.c_char_p), ("initial_prompt", C.c_char_p),
        ("translate", C.c_int), ("no_context", C.c_int), ("single_segment", C.c_int),
        ("no_timestamps", C.c_int), ("suppress_blank", C.c_int),
        ("temperature", C.c_float), ("temperature_inc", C.c_float), ("max_initial_ts", C.c_float),
        ("length_penalty", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float),
        ("no_speech_thold", C.c_float), ("n_max_text_ctx", C.c_int), ("max_tokens", C.c_int),
        ("beam_sampled", C.c_int),
    ]


def reference_params(language: str | None = "en", initial_prompt: str | None = None, beam_size: int = 0,
                     temperature_inc: float = 0.2) -> WoParams:
    """The parameter block the reference builds (src-tauri/src/whisper.rs:88-124) on top of
    whisper_full_default_params: Greedy{best_of:1}; beam_size>0 selects BeamSearch."""
    p = WoParams()
    p.strategy = 1 if beam_size > 0 else 0
    p.best_of = -1 if beam_size > 0 else 1
    p.beam_size = beam_size if beam_size > 0 else -1
    p.language = language.encode() if language else None
    p.initial_prompt = initial_prompt.encode() if initial_prompt else None
    p.translate = 0
    p.no_context = 0
    p.single_segment = 0
    p.no_timestamps = 0
    p.suppress_blank = 1
    p.temperature = 0.0
    p.temperature_inc = temperature_inc
    p.max_initial_ts = 1.0
    p.length_penalty = -1.0
    p.entropy_thold = 2.4
    p.logprob_thold = -1.0
    p.no_speech_thold = 0.6
    p.n_max_text_ctx = 16384
    p.max_tokens = 0
    p.beam_sampled = 0
    return p


# int hook(user, seek, i_temp, step, decoder, n_prompt, n_vocab, logits*): scripted-logits hook of the control-flow tests
LOGITS_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float))

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.wo_load.restype = vp; L.wo_load.argtypes = [C.c_char_p]
        L.wo_free.argtypes = [vp]
        L.wo_set_gelu_erf.argtypes = [vp, C.c_int]
        L.wo_set_ggml_faithful.argtypes = [vp, C.c_int]
        L.wo_gelu.restype = C.c_float; L.wo_gelu.argtypes = [vp, C.c_float]
        L.wo_set_threads.argtypes = [C.c_int]
        L.wo_max_threads.restype = C.c_int
        L.wo_hparams.argtypes = [vp, ip]
        L.wo_special_tokens.argtypes = [vp, ip]
        L.wo_tokenize.restype = C.c_int; L.wo_tokenize.argtypes = [vp, C.c_char_p, ip, C.c_int]
        L.wo_token_str.restype = C.c_void_p; L.wo_token_str.argtypes = [vp, C.c_int]
        L.wo_token_len.restype = C.c_int; L.wo_token_len.argtypes = [vp, C.c_int]
        L.wo_pcm_to_mel.restype = C.c_int; L.wo_pcm_to_mel.argtypes = [vp, fp, C.c_int]
        L.wo_mel_n_len.restype = C.c_int; L.wo_mel_n_len.argtypes = [vp]
        L.wo_mel_n_len_org.restype = C.c_int; L.wo_mel_n_len_org.argtypes = [vp]
        L.wo_mel_data.restype = fp; L.wo_mel_data.argtypes = [vp]
        L.wo_encode.restype = C.c_int; L.wo_encode.argtypes = [vp, C.c_int]
        L.wo_encoder_out.restype = fp; L.wo_encoder_out.argtypes = [vp]
        L.wo_cross_k.restype = fp; L.wo_cross_k.argtypes = [vp, C.c_int]
        L.wo_cross_v.restype = fp; L.wo_cross_v.argtypes = [vp, C.c_int]
        L.wo_decode.restype = C.c_int; L.wo_decode.argtypes = [vp, ip, C.c_int, C.c_int, C.c_int]
        L.wo_logits.restype = fp; L.wo_logits.argtypes = [vp]
        L.wo_lang_detect.restype = C.c_int; L.wo_lang_detect.argtypes = [vp, fp]
        L.wo_full.restype = C.c_int; L.wo_full.argtypes = [vp, C.POINTER(WoParams), fp, C.c_int]
        L.wo_n_segments.restype = C.c_int; L.wo_n_segments.argtypes = [vp]
        L.wo_segment_text.restype = C.c_void_p; L.wo_segment_text.argtypes = [vp, C.c_int]
        L.wo_segment_text_len.restype = C.c_int; L.wo_segment_text_len.argtypes = [vp, C.c_int]
        L.wo_segment_t0.restype = C.c_long; L.wo_segment_t0.argtypes = [vp, C.c_int]
        L.wo_segment_t1.restype = C.c_long; L.wo_segment_t1.argtypes = [vp, C.c_int]
        L.wo_segment_n_tokens.restype = C.c_int; L.wo_segment_n_tokens.argtypes = [vp, C.c_int]
        L.wo_segment_token_id.restype = C.c_int; L.wo_segment_token_id.argtypes = [vp, C.c_int, C.c_int]
        L.wo_segment_token_plog.restype = C.c_float; L.wo_segment_token_plog.argtypes = [vp, C.c_int, C.c_int]
        L.wo_no_speech_prob.restype = C.c_float; L.wo_no_speech_prob.argtypes = [vp]
        L.wo_lang_id.restype = C.c_int; L.wo_lang_id.argtypes = [vp]
        L.wo_stats.argtypes = [vp, C.POINTER(C.c_long)]
        L.wo_process_logits.restype = C.c_int
        L.wo_process_logits.argtypes = [vp, C.POINTER(WoParams), ip, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp]
        L.wo_set_logits.argtypes = [vp, fp]
        L.wo_set_logits_hook.argtypes = [vp, LOGITS_HOOK, vp]
        L.wo_canonical_stream.argtypes = [C.POINTER(C.c_double), C.c_int]
        L.wo_sample_stream.argtypes = [fp, C.c_int, ip, C.c_int]
        _lib = L
    return _lib


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """One loaded model + one state (the reference creates a fresh state per call,
    src-tauri/src/whisper.rs:83-85; `full` resets everything a fresh state would)."""

    def __init__(self, model_path: str):
        self.L = lib()
        self.h = self.L.wo_load(model_path.encode())
        if not self.h:
            raise RuntimeError(f"oracle: failed to load {model_path}")
        hp = np.zeros(11, np.int32)
        self.L.wo_hparams(self.h, _iptr(hp))
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer,
         self.n_text_ctx, self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels, self.ftype) = map(int, hp)
        st = np.zeros(9, np.int32)
        self.L.wo_special_tokens(self.h, _iptr(st))
        (self.token_eot, self.token_sot, self.token_translate, self.token_transcribe, self.token_solm,
         self.token_prev, self.token_nosp, self.token_not, self.token_beg) = map(int, st)

    def close(self):
        if self.h:
            self.L.wo_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_logits_hook(self, fn):
        """fn(seek, i_temp, step, decoder, n_prompt, logits: np.ndarray view) -> truthy if it overwrote the logits; None removes the hook."""
        if fn is None:
            self._hook = None
            self.L.wo_set_logits_hook(self.h, C.cast(None, LOGITS_HOOK), None)
            return

        def tramp(_user, seek, it, step, dec, n_prompt, n_vocab, ptr):
            return 1 if fn(seek, it, step, dec, n_prompt, np.ctypeslib.as_array(ptr, shape=(n_vocab,))) else 0

        self._hook = LOGITS_HOOK(tramp)     # keep the trampoline alive
        self.L.wo_set_logits_hook(self.h, self._hook, None)

    def set_ggml_faithful(self, on: bool):
        """ggml's CPU roundings on top of the ideal-fp32 graph: f16 mul_mat activations (f16 models), f16 im2col, f16 GELU table,
        f16 KV caches, f16 attention operands."""
        self.L.wo_set_ggml_faithful(self.h, int(on))

    def gelu(self, x: float) -> float:
        return float(self.L.wo_gelu(self.h, float(x)))

    def set_gelu_erf(self, on: bool):
        self.L.wo_set_gelu_erf(self.h, int(on))

    def tokenize(self, text: str | bytes) -> list[int]:
        b = text.encode() if isinstance(text, str) else text
        out = np.zeros(max(16, 2 * len(b) + 16), np.int32)
        n = self.L.wo_tokenize(self.h, b, _iptr(out), len(out))
        assert n >= 0
        return out[:n].tolist()

    def token_bytes(self, tid: int) -> bytes:
        n = self.L.wo_token_len(self.h, tid)
        return C.string_at(self.L.wo_token_str(self.h, tid), n)

    def mel(self, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, np.float32)
        n_len = self.L.wo_pcm_to_mel(self.h, _fptr(pcm), len(pcm))
        data = np.ctypeslib.as_array(self.L.wo_mel_data(self.h), shape=(self.n_mels, n_len)).copy()
        return data, self.L.wo_mel_n_len_org(self.h)

    def encode(self, mel_offset: int = 0) -> np.ndarray:
        rc = self.L.wo_encode(self.h, mel_offset)
        assert rc == 0, rc
        return np.ctypeslib.as_array(self.L.wo_encoder_out(self.h), shape=(self.n_audio_ctx, self.n_audio_state)).copy()

    def cross_kv(self, layer: int):
        shp = (self.n_audio_ctx, self.n_text_state)
        return (np.ctypeslib.as_array(self.L.wo_cross_k(self.h, layer), shape=shp).copy(),
                np.ctypeslib.as_array(self.L.wo_cross_v(self.h, layer), shape=shp).copy())

    def decode(self, tokens, n_past: int = 0, seq: int = 0) -> np.ndarray:
        t = np.ascontiguousarray(tokens, np.int32)
        rc = self.L.wo_decode(self.h, _iptr(t), len(t), n_past, seq)
        assert rc == 0, rc
        return np.ctypeslib.as_array(self.L.wo_logits(self.h), shape=(self.n_vocab,)).copy()

    def lang_detect(self):
        probs = np.zeros(100, np.float32)
        lid = self.L.wo_lang_detect(self.h, _fptr(probs))
        return lid, probs

    def process_logits(self, params: WoParams, logits: np.ndarray, hist, has_ts: bool, seek_delta: int, temperature: float):
        lg = np.ascontiguousarray(logits, np.float32)
        self.L.wo_set_logits(self.h, _fptr(lg))
        h = np.ascontiguousarray(hist, np.int32)
        lp = np.zeros(self.n_vocab, np.float32)
        pr = np.zeros(self.n_vocab, np.float32)
        self.L.wo_process_logits(self.h, C.byref(params), _iptr(h), len(h), int(has_ts), seek_delta, temperature, _fptr(lp), _fptr(pr))
        return lp, pr

    def full(self, params: WoParams, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, np.float32)
        rc = self.L.wo_full(self.h, C.byref(params), _fptr(pcm), len(pcm))
        if rc != 0:
            raise RuntimeError(f"oracle full failed: {rc}")
        segs = []
        for i in range(self.L.wo_n_segments(self.h)):
            n = self.L.wo_segment_text_len(self.h, i)
            text = C.string_at(self.L.wo_segment_text(self.h, i), n)
            toks = [self.L.wo_segment_token_id(self.h, i, j) for j in range(self.L.wo_segment_n_tokens(self.h, i))]
            plogs = [self.L.wo_segment_token_plog(self.h, i, j) for j in range(len(toks))]
            segs.append({"t0": self.L.wo_segment_t0(self.h, i), "t1": self.L.wo_segment_t1(self.h, i),
                         "text": text, "tokens": toks, "plog": plogs})
        return segs

    def stats(self):
        s = (C.c_long * 5)()
        self.L.wo_stats(self.h, s)
        return {"n_encode": s[0], "n_decode_calls": s[1], "n_decode_tokens": s[2], "n_fail_p": s[3], "n_fail_h": s[4]}

    def no_speech_prob(self) -> float:
        return float(self.L.wo_no_speech_prob(self.h))


def canonical_stream(n: int) -> np.ndarray:
    out = np.zeros(n, np.float64)
    lib().wo_canonical_stream(out.ctypes.data_as(C.POINTER(C.c_double)), n)
    return out


def sample_stream(probs: np.ndarray, n: int) -> np.ndarray:
    pr = np.ascontiguousarray(probs, np.float32)
    out = np.zeros(n, np.int32)
    lib().wo_sample_stream(_fptr(pr), len(pr), _iptr(out), n)
    return out
