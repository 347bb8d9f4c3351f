"""Streaming capture buffer (reference src-tauri/src/audio.rs:29-244): the library's host-side `AudioBuffer`
(csrc/host/audio_buffer.cpp through the C ABI) against the CPU oracle, which tests/test_audio_oracle.py pins with
the reference's own unit test (audio.rs:806-830).  Host logic: runs without a GPU.  Bar: bit-exact."""
import numpy as np
import pytest

from oracle import audio_oracle as ao

SR = 16000


def tone(seconds, amp, w=0.01, sr=SR):
    i = np.arange(int(seconds * sr), dtype=np.float32)
    return (np.sin(i * np.float32(w)) * np.float32(amp)).astype(np.float32)


def test_reference_unit_test_audio_buffer_overlap():   # audio.rs:806-830
    from nobs_whisper_b200 import audio
    buf = audio.AudioBuffer.with_sample_rate(SR)
    buf.push_samples(tone(3, 0.3))
    buf.push_samples(np.zeros(int(1.5 * SR), np.float32))
    assert buf.has_silence_boundary()
    chunk = buf.take_chunk_at_silence()
    assert chunk is not None
    assert buf.overlap_len == SR * 200 // 1000
    buf.close()


def test_calculate_rms_matches_the_reference_cases():   # audio.rs:586-594
    from nobs_whisper_b200 import audio
    assert audio.calculate_rms(np.zeros(100, np.float32)) < 0.001
    assert audio.calculate_rms(np.full(100, 0.5, np.float32)) > 0.4
    assert audio.calculate_rms(np.zeros(0, np.float32)) == 0.0
    x = np.random.default_rng(0).standard_normal(777).astype(np.float32)
    assert np.float32(audio.calculate_rms(x)) == ao.calculate_rms(x)


@pytest.mark.parametrize("seed,sr,push", [(0, 16000, 160), (1, 16000, 512), (2, 48000, 480), (3, 16000, 333), (4, 44100, 1024)])
def test_streaming_session_matches_the_oracle(seed, sr, push):
    """A capture session in callback-sized pushes (state.rs:586-605): every chunk either side hands out, the noise
    floor and the buffer length agree bit for bit after every push."""
    from nobs_whisper_b200 import audio
    rng = np.random.default_rng(seed)
    parts = []
    for _ in range(6):
        d = float(rng.uniform(0.6, 7.0))
        parts.append(tone(d, float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.004, 0.05)), sr) + rng.normal(0, 0.002, int(d * sr)).astype(np.float32))
        parts.append(rng.normal(0, float(rng.uniform(0.0003, 0.003)), int(rng.uniform(0.3, 1.6) * sr)).astype(np.float32))
    parts.append(tone(27.0, 0.25, 0.02, sr))     # long continuous speech: forces a split (MAX_BUFFER_DURATION_S = 25)
    x = np.concatenate(parts).astype(np.float32)
    got, want = audio.AudioBuffer(sr), ao.AudioBuffer(sr)
    n_chunks = n_forced = 0
    for off in range(0, len(x), push):
        blk = x[off:off + push]
        got.push_samples(blk)
        want.push_samples(blk)
        assert np.float32(got.get_noise_floor()) == want.noise_floor
        assert got.has_silence_boundary() == want.has_silence_boundary()
        a, b = got.take_chunk_at_silence(), want.take_chunk_at_silence()
        if a is None and b is None:                 # state.rs:600-603: forced split only if no silence chunk
            a, b = got.take_forced_chunk(), want.take_forced_chunk()
            n_forced += a is not None
        assert (a is None) == (b is None)
        if a is not None:
            n_chunks += 1
            assert np.array_equal(a, b)
        assert len(got) == len(want) and got.overlap_len == len(want.overlap_buffer)
    assert n_chunks >= 4 and n_forced >= 1
    assert np.array_equal(got.take(), want.take())
    assert len(got) == 0 and got.overlap_len == 0
    got.close()


def test_split_at_silences_host_function_matches_the_oracle():
    """audio.rs:467-507 through the C ABI (index arithmetic only: no GPU involved)."""
    from nobs_whisper_b200 import audio
    rng = np.random.default_rng(5)
    x = rng.standard_normal(SR * 9).astype(np.float32)
    cases = [[], [SR * 2, SR * 4], [100, SR * 3, SR * 3, SR * 8], [0, SR * 9, SR * 20], [SR * 5, SR * 2]]
    for b in cases:
        for sr in (16000, 48000):
            got = audio.split_at_silences_with_overlap(x, b, sr)
            want = ao.split_at_silences_with_overlap(x, b, sr)
            assert len(got) == len(want) and all(np.array_equal(g, w) for g, w in zip(got, want)), (b, sr)
    # the reference's own case (audio.rs:663-683)
    six = (np.sin(np.arange(SR * 6, dtype=np.float32) * np.float32(0.001)) * np.float32(0.1)).astype(np.float32)
    assert [len(c) for c in audio.split_at_silences(six, [SR * 2, SR * 4])] == [SR * 2, SR * 2 + 3200, SR * 2 + 3200]


def test_mix_to_mono_host_function():
    """state.rs:590-594."""
    from nobs_whisper_b200 import audio
    st = np.random.default_rng(2).standard_normal(3001).astype(np.float32)
    for ch in (1, 2, 3):
        assert np.array_equal(audio.mix_to_mono(st, ch), ao.mix_to_mono(st, ch))
