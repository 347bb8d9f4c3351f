"""Synthetic ggml-format Whisper model files (random-init, named architectures).

The reference loads `ggml-<id>.bin` files (reference: src-tauri/src/model.rs:50-188,
src-tauri/src/lib.rs:29, src-tauri/src/config.rs:144) through
`WhisperContext::new_with_params` (src-tauri/src/whisper.rs:36-52).  There is no network
here, so tests and the bench write files of the same on-disk layout with seeded random
weights of the named architecture (SURVEY.md §8a row a1 describes the layout):

    u32 magic 0x67676d6c
    11 x i32 hparams (n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer,
                      n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype)
    i32 n_mel, i32 n_fft, f32[n_mel*n_fft] mel filterbank
    i32 n_tokens, then (u32 len, bytes) per token
    tensors until EOF: i32 n_dims, i32 name_len, i32 ttype, i32 ne[n_dims] (innermost
                       first), name bytes, raw data (f32 or f16)

This module is a file *writer* used by tests, `bench.py` and `smoke()`; it performs no
inference.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

GGML_MAGIC = 0x67676D6C


@dataclass(frozen=True)
class Arch:
    name: str
    n_vocab: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int
    n_mels: int


def _a(name, nv, d, h, le, ld, mels):
    return Arch(name, nv, 1500, d, h, le, 448, d, h, ld, mels)


# reference catalogue ids: src-tauri/src/model.rs:54,65,76,87,98,109.  "micro" is a
# test-only architecture (same structure, d_head 64) small enough for second-scale CPU runs.
ARCHS = {
    "micro": _a("micro", 51865, 128, 2, 2, 3, 80),  # 3 decoder layers: 2 would trip the distil rule
    "micro128": _a("micro128", 51866, 128, 2, 2, 3, 128),  # test-only: v3 vocabulary layout + 128 mel bins
    "tiny": _a("tiny", 51865, 384, 6, 4, 4, 80),
    "base": _a("base", 51865, 512, 8, 6, 6, 80),
    "small": _a("small", 51865, 768, 12, 12, 12, 80),
    "medium": _a("medium", 51865, 1024, 16, 24, 24, 80),
    "large-v3": _a("large-v3", 51866, 1280, 20, 32, 32, 128),
    "large-v3-turbo": _a("large-v3-turbo", 51866, 1280, 20, 32, 4, 128),
}

N_BASE_TOKENS = 50257  # tokens stored in the file; specials are derived by the loader


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f)


def slaney_filterbank(n_mels: int, n_fft: int = 400, sr: int = 16000) -> np.ndarray:
    """Slaney-normalised triangular mel filterbank [n_mels, n_fft//2+1] (f32)."""
    n_freq = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    fb = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2:] - mel_pts[:-2])
    fb *= enorm[:, None]
    return fb.astype(np.float32)


_SYL = ["ka", "to", "mi", "re", "su", "no", "la", "vi", "de", "po", "an", "el", "or", "ti", "us", "en"]


def synthetic_vocab(n: int = N_BASE_TOKENS) -> list[bytes]:
    """Deterministic byte-string vocabulary: 256 single bytes, then unique pseudo-words.

    Contains b" " (needed by suppress_blank) and ASCII letters/digits/punctuation as single
    bytes, so any text is tokenisable by greedy longest-match.
    """
    toks = [bytes([i]) for i in range(256)]
    seen = set(toks)
    i = 0
    while len(toks) < n:
        k = i
        parts = []
        for _ in range(1 + (i % 3)):
            parts.append(_SYL[k % 16])
            k //= 16
        w = "".join(parts)
        if k:
            w += str(k)
        variants = (" " + w, w, " " + w.capitalize(), w.upper())
        cand = variants[(i // 7) % 4].encode()
        i += 1
        if cand in seen:
            continue
        seen.add(cand)
        toks.append(cand)
    return toks


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def tensor_specs(a: Arch):
    """Yield (name, torch-order shape, kind) for every tensor in file order."""
    d, dt = a.n_audio_state, a.n_text_state
    yield "decoder.positional_embedding", (a.n_text_ctx, dt), "pos_dec"
    yield "encoder.positional_embedding", (a.n_audio_ctx, d), "pos_enc"
    yield "decoder.token_embedding.weight", (a.n_vocab, dt), "mat"
    yield "encoder.conv1.weight", (d, a.n_mels, 3), "conv"
    yield "encoder.conv1.bias", (d, 1), "convb"
    yield "encoder.conv2.weight", (d, d, 3), "conv"
    yield "encoder.conv2.bias", (d, 1), "convb"
    for i in range(a.n_audio_layer):
        p = f"encoder.blocks.{i}."
        yield p + "attn_ln.weight", (d,), "gamma"
        yield p + "attn_ln.bias", (d,), "vec"
        yield p + "attn.query.weight", (d, d), "mat"
        yield p + "attn.query.bias", (d,), "vec"
        yield p + "attn.key.weight", (d, d), "mat"
        yield p + "attn.value.weight", (d, d), "mat"
        yield p + "attn.value.bias", (d,), "vec"
        yield p + "attn.out.weight", (d, d), "mat"
        yield p + "attn.out.bias", (d,), "vec"
        yield p + "mlp_ln.weight", (d,), "gamma"
        yield p + "mlp_ln.bias", (d,), "vec"
        yield p + "mlp.0.weight", (4 * d, d), "mat"
        yield p + "mlp.0.bias", (4 * d,), "vec"
        yield p + "mlp.2.weight", (d, 4 * d), "mat"
        yield p + "mlp.2.bias", (d,), "vec"
    yield "encoder.ln_post.weight", (d,), "gamma"
    yield "encoder.ln_post.bias", (d,), "vec"
    for i in range(a.n_text_layer):
        p = f"decoder.blocks.{i}."
        for att in ("attn", "cross_attn"):
            yield p + att + "_ln.weight", (dt,), "gamma"
            yield p + att + "_ln.bias", (dt,), "vec"
            yield p + att + ".query.weight", (dt, dt), "mat"
            yield p + att + ".query.bias", (dt,), "vec"
            yield p + att + ".key.weight", (dt, dt), "mat"
            yield p + att + ".value.weight", (dt, dt), "mat"
            yield p + att + ".value.bias", (dt,), "vec"
            yield p + att + ".out.weight", (dt, dt), "mat"
            yield p + att + ".out.bias", (dt,), "vec"
        yield p + "mlp_ln.weight", (dt,), "gamma"
        yield p + "mlp_ln.bias", (dt,), "vec"
        yield p + "mlp.0.weight", (4 * dt, dt), "mat"
        yield p + "mlp.0.bias", (4 * dt,), "vec"
        yield p + "mlp.2.weight", (dt, 4 * dt), "mat"
        yield p + "mlp.2.bias", (dt,), "vec"
    yield "decoder.ln.weight", (dt,), "gamma"
    yield "decoder.ln.bias", (dt,), "vec"


def gen_tensor(rng: np.random.Generator, a: Arch, shape, kind, init: str = "survey") -> np.ndarray:
    """`init`:
    "survey" — SURVEY.md §8d: N(0, 0.02) matrices and biases, decoder positions N(0, 0.01),
               sinusoidal encoder positions (LN gamma is 1 + N(0, 0.02), beta N(0, 0.02) so that
               both LN parameters are exercised).  Logits come out nearly flat, so the timestamp
               rules dominate decoding.
    "fanin"  — matrices N(0, 1/fan_in): activations and logits are O(1), so text tokens win
               and the decode loop runs long (the other regime the parity tests cover).
    """
    f32 = np.float32
    if kind == "pos_enc":
        return sinusoids(a.n_audio_ctx, a.n_audio_state)
    if kind == "pos_dec":
        return rng.standard_normal(shape, dtype=f32) * f32(0.01)
    if kind == "gamma":
        return f32(1.0) + rng.standard_normal(shape, dtype=f32) * f32(0.02)
    if kind in ("conv", "mat") and init == "fanin":
        fan_in = int(np.prod(shape[1:]))
        return rng.standard_normal(shape, dtype=f32) * f32(1.0 / np.sqrt(fan_in))
    return rng.standard_normal(shape, dtype=f32) * f32(0.02)


def generate_weights(arch: str | Arch, seed: int = 0, init: str = "survey"):
    """Yield (name, f32 ndarray in torch order) for every tensor, deterministically."""
    a = ARCHS[arch] if isinstance(arch, str) else arch
    rng = np.random.default_rng(seed)
    for name, shape, kind in tensor_specs(a):
        yield name, gen_tensor(rng, a, shape, kind, init)


_F32_NAMES = (
    "encoder.conv1.bias",
    "encoder.conv2.bias",
    "encoder.positional_embedding",
    "decoder.positional_embedding",
)


# whisper.cpp `quantize` file types -> ggml tensor type of the quantised matrices (reference catalogue: model.rs:155-186)
QUANT_TTYPE = {2: 2, 3: 3, 7: 8, 8: 6, 9: 7}   # ftype: q4_0, q4_1, q8_0, q5_0, q5_1


def quantize_blocks(w: np.ndarray, ttype: int) -> bytes:
    """ggml reference quantisers (quantize_row_q{4_0,4_1,5_0,5_1,8_0}_ref): 32 weights per block."""
    x = np.ascontiguousarray(w, dtype=np.float32).reshape(-1, 32)
    nb = x.shape[0]
    if ttype == 8:
        amax = np.abs(x).max(axis=1)
        d = (amax / np.float32(127.0)).astype(np.float32)
        inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), np.float32(0)).astype(np.float32)
        q = np.rint(x * inv[:, None]).astype(np.int8)
        out = np.zeros((nb, 34), np.uint8)
        out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = q.view(np.uint8)
        return out.tobytes()
    sym = ttype in (2, 6)
    levels = 16 if ttype in (2, 3) else 32
    if sym:
        idx = np.abs(x).argmax(axis=1)
        mx = x[np.arange(nb), idx]                      # the value with the largest magnitude, sign kept
        d = (mx / np.float32(-(levels // 2))).astype(np.float32)
        inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), np.float32(0)).astype(np.float32)
        q = np.minimum(levels - 1, (x * inv[:, None] + np.float32(levels // 2 + 0.5)).astype(np.int32)).astype(np.uint8)
        hdr = d.astype(np.float16).view(np.uint8).reshape(nb, 2)
    else:
        mn, mx = x.min(axis=1), x.max(axis=1)
        d = ((mx - mn) / np.float32(levels - 1)).astype(np.float32)
        inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), np.float32(0)).astype(np.float32)
        q = np.minimum(levels - 1, ((x - mn[:, None]) * inv[:, None] + np.float32(0.5)).astype(np.int32)).astype(np.uint8)
        hdr = np.concatenate([d.astype(np.float16).view(np.uint8).reshape(nb, 2), mn.astype(np.float16).view(np.uint8).reshape(nb, 2)], axis=1)
    qs = (q[:, :16] & 15) | ((q[:, 16:] & 15) << 4)
    parts = [hdr]
    if levels == 32:
        bits = ((q >> 4) & 1).astype(np.uint32)
        qh = (bits << np.arange(32, dtype=np.uint32)[None, :]).sum(axis=1).astype(np.uint32)
        parts.append(qh.view(np.uint8).reshape(nb, 4))
    parts.append(qs.astype(np.uint8))
    return np.concatenate(parts, axis=1).tobytes()


def dequantize_blocks(raw: bytes, ttype: int, count: int) -> np.ndarray:
    """ggml dequantize_row_q*: the float32 values a loader must produce from `raw` (used by the tests)."""
    bs = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34}[ttype]
    b = np.frombuffer(raw, np.uint8).reshape(-1, bs)
    nb = b.shape[0]
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(nb)
    off = 2
    m = np.zeros(nb, np.float32)
    if ttype in (3, 7):
        m = b[:, 2:4].copy().view(np.float16).astype(np.float32).reshape(nb)
        off = 4
    if ttype == 8:
        return (b[:, 2:].copy().view(np.int8).astype(np.float32) * d[:, None]).reshape(-1)[:count]
    hi_bits = np.zeros((nb, 32), np.int32)
    if ttype in (6, 7):
        qh = b[:, off:off + 4].copy().view(np.uint32).reshape(nb)
        hi_bits = ((qh[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.int32) << 4
        off += 4
    qs = b[:, off:off + 16].astype(np.int32)
    q = np.concatenate([qs & 15, qs >> 4], axis=1) | hi_bits
    if ttype == 2:
        q = q - 8
    if ttype == 6:
        q = q - 16
    y = q.astype(np.float32) * d[:, None]
    if ttype in (3, 7):
        y = y + m[:, None]
    return y.reshape(-1)[:count]


def write_model(path: str, arch: str | Arch, seed: int = 0, ftype: int = 0, vocab: list[bytes] | None = None,
                init: str = "survey", dequantized: bool = False) -> str:
    """Write `ggml-<arch>.bin`-layout file.  ftype 0 = all f32, 1 = f16 matrices, 2/3/7/8/9 = q4_0/q4_1/q8_0/q5_0/q5_1
    2-D `*.weight` matrices (what whisper.cpp's `quantize` tool produces) with the remaining tensors as for ftype 1.
    dequantized=True writes the SAME model with every quantised matrix replaced by its float32 dequantisation
    (the file a loader must treat identically: the parity tests compare the two)."""
    a = ARCHS[arch] if isinstance(arch, str) else arch
    vocab = vocab if vocab is not None else synthetic_vocab()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<I", GGML_MAGIC))
        f.write(
            struct.pack(
                "<11i", a.n_vocab, a.n_audio_ctx, a.n_audio_state, a.n_audio_head, a.n_audio_layer,
                a.n_text_ctx, a.n_text_state, a.n_text_head, a.n_text_layer, a.n_mels, ftype,
            )
        )
        fb = slaney_filterbank(a.n_mels)
        f.write(struct.pack("<2i", fb.shape[0], fb.shape[1]))
        f.write(fb.tobytes())
        f.write(struct.pack("<i", len(vocab)))
        for t in vocab:
            f.write(struct.pack("<I", len(t)))
            f.write(t)
        for name, w in generate_weights(a, seed, init):
            use_f16 = ftype >= 1 and w.ndim >= 2 and name not in _F32_NAMES
            qt = QUANT_TTYPE.get(ftype, 0) if (w.ndim == 2 and name.endswith("weight") and w.shape[-1] % 32 == 0) else 0
            nb = name.encode()
            if qt and dequantized:
                w = dequantize_blocks(quantize_blocks(w, qt), qt, w.size).reshape(w.shape)
                qt, use_f16 = 0, False
            f.write(struct.pack("<3i", w.ndim, len(nb), qt if qt else (1 if use_f16 else 0)))
            f.write(struct.pack(f"<{w.ndim}i", *reversed(w.shape)))
            f.write(nb)
            f.write(quantize_blocks(w, qt) if qt else (w.astype(np.float16) if use_f16 else w).tobytes())
    os.replace(tmp, path)
    return path


def model_path(dirname: str, arch: str, seed: int = 0, ftype: int = 0, init: str = "survey", dequantized: bool = False) -> str:
    """Path following the reference naming `ggml-<id>.bin` (lib.rs:29), tagged by seed/ftype/init."""
    tag = "" if (seed == 0 and ftype == 0 and init == "survey") else f"-s{seed}-f{ftype}-{init}"
    return os.path.join(dirname, f"ggml-{arch}{tag}{'-deq' if dequantized else ''}.bin")


def ensure_model(dirname: str, arch: str, seed: int = 0, ftype: int = 0, init: str = "survey", dequantized: bool = False) -> str:
    os.makedirs(dirname, exist_ok=True)
    p = model_path(dirname, arch, seed, ftype, init, dequantized)
    if not os.path.exists(p):
        write_model(p, arch, seed, ftype, init=init, dequantized=dequantized)
    return p
