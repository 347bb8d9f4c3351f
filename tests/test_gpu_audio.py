"""Silence chunker: the CUDA path through the C ABI against the CPU oracle
(oracle/audio_oracle.py, itself pinned by the reference's unit tests in tests/test_audio_oracle.py).
Integer / threshold work: the bar is bit-exact."""
import numpy as np
import pytest

from oracle import audio_oracle as ao

pytestmark = pytest.mark.gpu
SR = 16000


def tone(seconds, amp, w=0.01):
    i = np.arange(int(seconds * SR), dtype=np.float32)
    return (np.sin(i * np.float32(w)) * np.float32(amp)).astype(np.float32)


def recording(seed, seconds):
    """speech-like bursts separated by gaps of 0.2-1.5 s over a noisy floor"""
    rng = np.random.default_rng(seed)
    parts = [tone(0.5, 0.002, 0.1)]
    t = 0.5
    while t < seconds:
        d = float(rng.uniform(1.0, 6.0))
        parts.append(tone(d, float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.005, 0.05))) + rng.normal(0, 0.003, int(d * SR)).astype(np.float32))
        g = float(rng.uniform(0.2, 1.5))
        parts.append(rng.normal(0, float(rng.uniform(0.0005, 0.004)), int(g * SR)).astype(np.float32))
        t += d + g
    return np.concatenate(parts).astype(np.float32)


@pytest.mark.parametrize("n,window", [(320 * 50, 320), (320 * 7 + 13, 320), (441 * 20 + 5, 441), (960 * 33, 960), (100, 320), (0, 320)])
def test_window_rms_bit_exact(n, window):
    from nobs_whisper_b200 import audio
    x = np.random.default_rng(n + window).standard_normal(n).astype(np.float32) * np.float32(0.2)
    got = audio.window_rms(x, window)
    want = ao.window_rms(x, window)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_reference_unit_test_cases_on_the_gpu():
    """audio.rs:620-803: the reference's own known-answer cases through the CUDA path."""
    from nobs_whisper_b200 import audio
    q, sp, z = lambda s, a=0.002: tone(s, a, 0.1), lambda s: tone(s, 0.3), lambda s: np.zeros(int(s * SR), np.float32)
    assert len(audio.find_silence_boundaries(np.concatenate([q(0.5), sp(2), z(1), sp(2), z(1), sp(2)]), SR)) == 2
    a = np.concatenate([q(0.5), sp(10)])
    chunks = audio.split_at_silences(a, audio.find_silence_boundaries(a, SR))
    assert len(chunks) == 1 and len(chunks[0]) == len(a)
    a = np.concatenate([q(0.5), sp(2), z(1), sp(2)])
    b = audio.find_silence_boundaries(a, SR)
    assert len(b) == 1 and len(audio.split_at_silences(a, b)) == 2
    assert audio.find_silence_boundaries(np.concatenate([q(0.5), sp(2), z(0.5), sp(2)]), SR) == []
    assert len(audio.find_silence_boundaries(np.concatenate([q(0.5, 0.005), sp(2), q(1.0, 0.005), sp(2)]), SR)) == 1
    six = (np.sin(np.arange(SR * 6, dtype=np.float32) * np.float32(0.001)) * np.float32(0.1)).astype(np.float32)
    assert [len(c) for c in audio.split_at_silences_with_overlap(six, [SR * 2, SR * 4], SR)] == [SR * 2, SR * 2 + 3200, SR * 2 + 3200]


@pytest.mark.parametrize("seed,seconds,sr", [(1, 60, 16000), (2, 300, 16000), (3, 45, 48000), (4, 0.3, 16000)])
def test_boundaries_and_chunks_match_the_oracle(seed, seconds, sr):
    from nobs_whisper_b200 import audio
    x = recording(seed, seconds)
    got = audio.find_silence_boundaries(x, sr)
    want = ao.find_silence_boundaries(x, sr)
    assert got == want
    gc, wc = audio.split_at_silences_with_overlap(x, got, sr), ao.split_at_silences_with_overlap(x, want, sr)
    assert len(gc) == len(wc) and all(np.array_equal(a, b) for a, b in zip(gc, wc))


def test_one_hour_recording_properties():
    """BASELINE size (1 h = 57.6 M samples): oracle equality plus size-independent properties of the cut."""
    from nobs_whisper_b200 import audio
    x = np.tile(recording(7, 360), 10)[: 3600 * SR]
    b = audio.find_silence_boundaries(x, SR)
    assert b == ao.find_silence_boundaries(x, SR)
    assert len(b) > 100 and all(b[i + 1] - b[i] >= SR for i in range(len(b) - 1)) and b[0] >= SR
    chunks = audio.split_at_silences(x, b)
    assert len(chunks) == len(b) + 1
    # dropping each chunk's 200 ms overlap gives back the recording exactly
    rebuilt = np.concatenate([chunks[0]] + [c[3200:] for c in chunks[1:]])
    assert np.array_equal(rebuilt, x)


def test_chunker_feeds_the_engine(model_dir):
    """state.rs:757-780: a recording longer than 30 s is cut at silences and the pieces go through transcribe()."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import audio, ggml_synth
    eng = nw.WhisperEngine()
    eng.load_model(ggml_synth.ensure_model(model_dir, "micro", init="fanin"))
    x = recording(11, 70)
    chunks = audio.split_at_silences(x, audio.find_silence_boundaries(x, SR))
    assert len(chunks) > 1
    chained = eng.transcribe_chunked(chunks, "en", None)
    batch = eng.transcribe_batch(chunks, "en", None)
    assert isinstance(chained, str) and len(batch) == len(chunks)
    eng.close()


def test_transcribe_recording_follows_the_reference_flow(model_dir):
    """state.rs:757-792 through the engine: > 30 s is cut at silences and chained; == the same steps done by hand."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import audio, ggml_synth
    eng = nw.WhisperEngine()
    eng.load_model(ggml_synth.ensure_model(model_dir, "micro", init="fanin"))
    x = recording(21, 50)
    assert len(x) > 30 * SR
    chunks = audio.split_at_silences(x, audio.find_silence_boundaries(x, SR))
    results = []
    for c in chunks:                       # the loop of state.rs:764-777
        t = eng.transcribe(c, "en", None, results[-1] if results else None)
        if t:
            results.append(t)
    assert eng.transcribe_recording(x, "en", None) == " ".join(results).strip()
    # data-parallel variant == transcribe_batch over the same pieces
    assert eng.transcribe_recording(x, "en", None, parallel=True) == " ".join(t for t in eng.transcribe_batch(chunks, "en", None) if t).strip()
    short = recording(22, 8)[: 8 * SR]
    assert eng.transcribe_recording(short, "en", None) == eng.transcribe(short, "en", None, None).strip()
    eng.close()


@pytest.mark.parametrize("fs_in,seconds", [(48000, 1.0), (48000, 12.3), (44100, 7.7), (22050, 5.0), (8000, 3.0), (32000, 0.01)])
def test_resampler_matches_the_oracle(fs_in, seconds):
    """the GEMM-form block resampler against the FFT-form oracle (float32 tolerance 1e-4
    relative to the signal peak; lengths exact, including the reference's own case audio.rs:570-583)."""
    from nobs_whisper_b200 import audio
    rng = np.random.default_rng(fs_in)
    n = int(fs_in * seconds)
    t = np.arange(n) / fs_in
    x = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t + 1.0) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got, want = audio.resample_audio(x, fs_in, 16000), ao.resample_audio(x, fs_in, 16000)
    assert got.shape == want.shape
    if len(want):
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())
    if fs_in == 48000 and seconds == 1.0:
        assert len(got) == 15872


def test_resample_chunk_and_mono_mix():
    from nobs_whisper_b200 import audio
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    assert np.array_equal(audio.resample_chunk(x, 16000), x)                      # audio.rs:330-332
    assert np.allclose(audio.resample_chunk(x, 48000), ao.resample_chunk(x, 48000), atol=1e-4)
    st = np.random.default_rng(1).standard_normal(2000).astype(np.float32)
    assert np.array_equal(audio.mix_to_mono(st, 2), ao.mix_to_mono(st, 2))        # state.rs:590-594
