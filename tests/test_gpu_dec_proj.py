"""Fused decoder-step projection (csrc/decode_proj_sm100.cu) against float64 numpy on the same bf16-rounded operands:
cluster split-K reduced through distributed shared memory, fused bias / GELU / residual epilogues, and the LayerNorm tail
(per-tile statistics + "last cluster normalises") that replaces the split-K GEMM + epilogue-kernel pair of round 1 on the
decoder's latency chain."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_round(a):
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def gelu(x):
    return 0.5 * x * (1.0 + np.tanh(0.79788456080286535587989211986876 * x * (1.0 + 0.044715 * x * x)))


def tile_stats(x):
    """per 128-feature tile and row: (mean, M2)"""
    R, N = x.shape
    t = x.reshape(R, N // 128, 128).astype(np.float64)
    mean = t.mean(axis=2)
    m2 = ((t - mean[:, :, None]) ** 2).sum(axis=2)
    return mean.T, m2.T          # [tiles][R]


def run(lib, R, N, K, *, resid, ln=False, act=0, bias=True, seed=0, iters=0):
    rng = np.random.default_rng(seed)
    W = bf16_round(rng.standard_normal((N, K), dtype=np.float32) * 0.05)
    x = bf16_round(rng.standard_normal((R, K), dtype=np.float32))
    b = rng.standard_normal(N).astype(np.float32) if bias else None
    res = (rng.standard_normal((R, N)).astype(np.float32) * 2.0 + rng.standard_normal((R, 1)).astype(np.float32)) if resid else None
    g = rng.standard_normal(N).astype(np.float32) if ln else None
    be = (rng.standard_normal(N).astype(np.float32) * 0.1) if ln else None
    out = np.full((R, N), np.nan, np.float32)
    y = np.full((R, N), np.nan, np.float32) if ln else None
    st = np.zeros((N // 128, 128, 2), np.float32) if ln else None
    us = C.c_float(0)
    rc = lib.whisper_b200_debug_dec_proj(R, N, K, fp(x), fp(W), fp(b), act, fp(res), fp(g), fp(be), fp(out), fp(y), fp(st), iters, C.byref(us))
    assert rc == 0, (rc, lib.whisper_b200_last_error())
    want = x.astype(np.float64) @ W.astype(np.float64).T
    if bias:
        want = want + b
    if act:
        want = gelu(want)
    if resid:
        want = want + res
    want_y = None
    if ln:
        want_y = (want - want.mean(1, keepdims=True)) / np.sqrt(want.var(1, keepdims=True) + 1e-5) * g + be
    return out, want, y, want_y, st, us.value


SHAPES = [(1280, 1280), (3840, 1280), (5120, 1280), (1280, 5120), (384, 384), (1152, 384), (1536, 384), (384, 1536), (1024, 1024), (1024, 4096)]


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("R", [1, 7, 33, 64, 101, 128])
def test_residual_epilogue(N, K, R):
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    out, want, *_ = run(lib, R, N, K, resid=True, seed=R + N)
    err = np.abs(out - want).max() / max(1.0, np.abs(want).max())
    assert err < 2e-5, err          # fp32 accumulation of bf16 products: only the summation order differs


@pytest.mark.parametrize("N,K", [(1280, 1280), (1280, 5120), (384, 384), (384, 1536), (1024, 1024), (768, 3072)])
@pytest.mark.parametrize("R", [1, 20, 60, 64, 100, 128])
def test_layernorm_tail(N, K, R):
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    out, want, y, want_y, st, _ = run(lib, R, N, K, resid=True, ln=True, seed=3 * R + K)
    assert np.abs(out - want).max() / max(1.0, np.abs(want).max()) < 2e-5
    mean, m2 = tile_stats(want.astype(np.float32))
    assert np.abs(st[:, :R, 0] - mean).max() < 1e-4
    assert np.abs(st[:, :R, 1] - m2).max() < 1e-3 * max(1.0, m2.max())
    # y is stored as bf16: half an ulp (2^-9 relative) of the largest normalised value, plus the fp32 statistics
    assert np.abs(y - want_y).max() < 6e-3 * max(1.0, np.abs(want_y).max())
    assert np.abs(y - want_y).mean() < 2e-3


def test_bf16_out_gelu_no_bias():
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    out, want, *_ = run(lib, 48, 5120, 1280, resid=False, act=1, bias=False, seed=5)
    assert np.abs(out - want).max() / np.abs(want).max() < 8e-3      # bf16 store + tanh.approx


def test_bf16_out_qkv_shape():
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    out, want, *_ = run(lib, 61, 3840, 1280, resid=False, seed=6)
    assert np.abs(out - want).max() / np.abs(want).max() < 6e-3


def test_deterministic_whichever_cluster_is_last():
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    ref = run(lib, 57, 1280, 5120, resid=True, ln=True, seed=9)
    for _ in range(5):
        again = run(lib, 57, 1280, 5120, resid=True, ln=True, seed=9)
        assert np.array_equal(ref[0], again[0]) and np.array_equal(ref[2], again[2])


def test_timing_report():
    """Not an assertion on speed: prints the per-launch time of the projection shapes of large-v3 at 60 rows."""
    from nobs_whisper_b200 import _lib
    lib = _lib.lib()
    for name, N, K, kw in [("out / cross-out + LN", 1280, 1280, dict(resid=True, ln=True)), ("cross-q", 1280, 1280, dict(resid=False)),
                           ("QKV", 3840, 1280, dict(resid=False)), ("FC1 + GELU", 5120, 1280, dict(resid=False, act=1)),
                           ("FC2 + LN", 1280, 5120, dict(resid=True, ln=True)), ("FC2 (no LN)", 1280, 5120, dict(resid=True))]:
        us = run(lib, 60, N, K, iters=200, **kw)[5]
        print(f"dec_proj {name:22s} N={N} K={K}: {us:.2f} us per launch (back to back, PDL)")
