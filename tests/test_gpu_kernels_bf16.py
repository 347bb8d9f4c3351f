"""Kernel-level GPU tests of the bf16 tcgen05 GEMM against a numpy fp32 reference of the same op
(inputs rounded to bf16 exactly as the kernel sees them).  Tolerance: fp32 accumulation over K
plus one bf16 rounding of the output (2^-8 relative) when C is bf16."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def gelu_tanh(x):
    return 0.5 * x * (1.0 + np.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def run_gemm(M, N, K, lda=None, bias=False, act=0, res=False, res_mod=0, win_rows=0, valid_rows=0, out_f32=True, seed=0):
    from nobs_whisper_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(seed)
    lda = lda or K
    a_elems = (M - 1) * lda + K
    A_flat = rng.standard_normal(a_elems).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if bias else None
    rrows = res_mod if res_mod else M
    r = rng.standard_normal((rrows, N)).astype(np.float32) if res else None
    out = np.zeros((M, N), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None
    rc = L.whisper_b200_debug_gemm_bf16(M, N, K, lda, fp(A_flat), a_elems, fp(W), fp(b), act, fp(r), res_mod, win_rows, valid_rows, int(out_f32), fp(out))
    assert rc == 0, (rc, L.whisper_b200_last_error())
    Ab = to_bf16(A_flat).astype(np.float64)
    A = np.lib.stride_tricks.as_strided(Ab, shape=(M, K), strides=(lda * 8, 8))
    ref = A @ to_bf16(W).astype(np.float64).T
    if bias:
        ref = ref + b
    if act:
        ref = gelu_tanh(ref)
    if res:
        ref = ref + (r[np.arange(M) % res_mod] if res_mod else r)
    if win_rows:
        ref[(np.arange(M) % win_rows) >= valid_rows] = 0.0
    return out, ref


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 1280), (1536, 1280, 1280), (3072, 384, 240), (4096, 5120, 1280),
                                   (1, 1280, 1280), (7, 3840, 1280), (120, 5120, 1280), (33, 51865, 384), (300, 200, 448)])
def test_gemm_plain(M, N, K):
    out, ref = run_gemm(M, N, K, out_f32=True, seed=M + N)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


def test_gemm_epilogues():
    out, ref = run_gemm(1536, 512, 512, bias=True, act=1, out_f32=False, seed=1)
    assert np.abs(out - ref).max() < 1.5e-2 * max(1.0, np.abs(ref).max())          # bf16 output + tanh.approx
    out, ref = run_gemm(1536, 512, 2048, bias=True, res=True, out_f32=True, seed=2)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
    out, ref = run_gemm(3072, 384, 1152, bias=True, act=1, res=True, res_mod=1536, out_f32=True, seed=3)
    assert np.abs(out - ref).max() < 5e-3 * max(1.0, np.abs(ref).max())
    out, ref = run_gemm(6144, 384, 240, bias=True, act=1, win_rows=3072, valid_rows=3000, out_f32=False, seed=4)
    assert np.abs(out - ref).max() < 1.5e-2 * max(1.0, np.abs(ref).max())
    assert np.all(out[3000:3072] == 0.0)


def test_gemm_overlapping_rows_like_the_stem_convolutions():
    # conv1: row stride n_mels, K = 3*n_mels;  conv2: row stride 2d, K = 3d
    out, ref = run_gemm(3072, 384, 240, lda=80, bias=True, act=1, out_f32=False, seed=5)
    assert np.abs(out - ref).max() < 1.5e-2 * max(1.0, np.abs(ref).max())
    out, ref = run_gemm(1536, 384, 1152, lda=768, bias=True, out_f32=True, seed=6)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
    out, ref = run_gemm(3072, 1280, 384, lda=128, out_f32=True, seed=7)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n_win,n_head,use_simt", [(1, 1, 0), (2, 3, 0), (1, 2, 1)])
def test_encoder_attention_kernel(n_win, n_head, use_simt):
    """tcgen05 flash attention (and the CUDA-core variant) against a float64 softmax(QK^T/8)V."""
    from nobs_whisper_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(n_win * 10 + n_head)
    d = 64 * n_head
    rows = n_win * 1536
    qkv = rng.standard_normal((rows, 3 * d)).astype(np.float32)
    qkv[:, :d] *= 1.5          # a softmax that is neither flat nor one-hot
    out = np.zeros((rows, d), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = L.whisper_b200_debug_enc_attention(n_win, n_head, fp(qkv), fp(out), use_simt)
    assert rc == 0, (rc, L.whisper_b200_last_error())
    qb = to_bf16(qkv).astype(np.float64)
    for w in range(n_win):
        blk = qb[w * 1536:(w + 1) * 1536]
        for h in range(n_head):
            q = blk[:, h * 64:(h + 1) * 64]
            k = blk[:1500, d + h * 64: d + (h + 1) * 64]
            v = blk[:1500, 2 * d + h * 64: 2 * d + (h + 1) * 64]
            s = q @ k.T * 0.125
            p = np.exp(s - s.max(axis=1, keepdims=True))
            ref = (p / p.sum(axis=1, keepdims=True)) @ v
            got = out[w * 1536:(w + 1) * 1536, h * 64:(h + 1) * 64]
            assert np.abs(got[:1500] - ref[:1500]).max() < 2e-2 * max(1.0, np.abs(ref).max())
            assert np.isfinite(got).all()


@pytest.mark.parametrize("R,n_head,n_slots,n_keys,streaming", [(3, 2, 2, 1500, 2), (5, 20, 3, 1500, 2), (130, 6, 4, 1500, 2), (2, 1, 1, 77, 2),
                                                                (4, 3, 2, 1280, 2), (3, 2, 2, 1500, 1), (130, 6, 4, 1500, 1), (2, 1, 1, 77, 1),
                                                                (3, 2, 2, 1500, 0),
                                                                (6, 3, 3, 1500, 22), (121, 20, 31, 1500, 22), (7, 2, 2, 1500, 23), (10, 6, 2, 1500, 25),
                                                                (9, 2, 1, 77, 24), (130, 6, 5, 1280, 24)])
def test_decoder_cross_attention_kernel(R, n_head, n_slots, n_keys, streaming):
    """Streaming cross-attention kernels (2: tcgen05, 1: SIMT over a cp.async.bulk ring, 0: block-per-head SIMT)
    against a float64 softmax(q K^T / 8) V over head-major panels; more items than SMs, ragged last chunk,
    rows sharing a slot.  Pad rows past n_keys hold large finite garbage that must not reach the result.
    streaming 20 + g: the tcgen05 kernel with ROW GROUPS — runs of g consecutive rows per audio slot (a pass and its speculative
    successor: 2; beams: 5) are one work item of up to 4 rows whose panels are streamed once; ragged last run, runs longer than 4."""
    run = streaming - 20 if streaming > 20 else 1
    from nobs_whisper_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(R * 100 + n_head)
    d = 64 * n_head
    q = (1.5 * rng.standard_normal((R, d))).astype(np.float32)
    k = rng.standard_normal((n_slots, n_head, 1536, 64)).astype(np.float32)
    v = rng.standard_normal((n_slots, n_head, 1536, 64)).astype(np.float32)
    k[:, :, n_keys:] = 1e4      # rows past n_keys must never reach the result
    v[:, :, n_keys:] = -3e3
    out = np.zeros((R, d), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = L.whisper_b200_debug_dec_cross_attention(R, n_head, n_slots, n_keys, fp(q), fp(k), fp(v), fp(out), streaming)
    assert rc == 0, (rc, L.whisper_b200_last_error())
    assert np.isfinite(out).all()
    qb, kb, vb = to_bf16(q).astype(np.float64), to_bf16(k[:, :, :n_keys]).astype(np.float64), to_bf16(v[:, :, :n_keys]).astype(np.float64)
    for r in range(R):
        s_ = (r // run) % n_slots
        for h in range(n_head):
            sc = kb[s_, h] @ qb[r, h * 64:(h + 1) * 64] * 0.125
            p = np.exp(sc - sc.max())
            ref = (p / p.sum()) @ vb[s_, h]
            assert np.abs(out[r, h * 64:(h + 1) * 64] - ref).max() < (2e-2 if streaming >= 2 else 1e-2) * max(1.0, np.abs(ref).max())
