"""The reference's own unit tests for the silence chunker (src-tauri/src/audio.rs:586-831), restated
against the CPU oracle (oracle/audio_oracle.py).  These are the known-answer tests that pin the oracle
for SURVEY.md §8f row N3; tests/test_gpu_audio.py then holds the CUDA path to the oracle bit for bit."""
import numpy as np

from oracle import audio_oracle as ao

SR = 16000


def quiet(seconds, amp=0.002):   # audio.rs:603-605
    i = np.arange(int(seconds * SR), dtype=np.float32)
    return (np.sin(i * np.float32(0.1)) * np.float32(amp)).astype(np.float32)


def speech(seconds):             # audio.rs:608-610
    i = np.arange(int(seconds * SR), dtype=np.float32)
    return (np.sin(i * np.float32(0.01)) * np.float32(0.3)).astype(np.float32)


def silence(seconds):
    return np.zeros(int(seconds * SR), np.float32)


def test_calculate_rms():        # audio.rs:586-594
    assert ao.calculate_rms(np.zeros(100, np.float32)) < 0.001
    assert ao.calculate_rms(np.full(100, 0.5, np.float32)) > 0.4
    assert ao.calculate_rms(np.zeros(0, np.float32)) == 0.0


def test_estimate_noise_floor():  # audio.rs:597-617
    audio = np.concatenate([quiet(0.5), speech(2)])
    assert ao.estimate_noise_floor(audio, SR) < 0.01


def test_find_silence_in_audio():  # audio.rs:620-660
    audio = np.concatenate([quiet(0.5), speech(2), silence(1), speech(2), silence(1), speech(2)])
    assert len(ao.find_silence_boundaries(audio, SR)) == 2


def test_split_at_silences_with_overlap():  # audio.rs:663-683
    audio = (np.sin(np.arange(SR * 6, dtype=np.float32) * np.float32(0.001)) * np.float32(0.1)).astype(np.float32)
    chunks = ao.split_at_silences_with_overlap(audio, [SR * 2, SR * 4], SR)
    overlap = SR * ao.CHUNK_OVERLAP_MS // 1000
    assert [len(c) for c in chunks] == [SR * 2, SR * 2 + overlap, SR * 2 + overlap]


def test_no_silence_returns_single_chunk():  # audio.rs:686-708
    audio = np.concatenate([quiet(0.5), speech(10)])
    chunks = ao.split_at_silences(audio, ao.find_silence_boundaries(audio, SR))
    assert len(chunks) == 1 and len(chunks[0]) == len(audio)


def test_audio_with_silence_is_chunked():  # audio.rs:711-741
    audio = np.concatenate([quiet(0.5), speech(2), silence(1), speech(2)])
    b = ao.find_silence_boundaries(audio, SR)
    assert len(b) == 1
    assert len(ao.split_at_silences(audio, b)) == 2


def test_short_silence_not_split():  # audio.rs:744-771
    audio = np.concatenate([quiet(0.5), speech(2), silence(0.5), speech(2)])
    assert ao.find_silence_boundaries(audio, SR) == []


def test_adaptive_threshold_with_noisy_background():  # audio.rs:774-803
    audio = np.concatenate([quiet(0.5, 0.005), speech(2), quiet(1.0, 0.005), speech(2)])
    assert len(ao.find_silence_boundaries(audio, SR)) == 1


def test_audio_buffer_overlap():  # audio.rs:806-830
    buf = ao.AudioBuffer(SR)
    buf.push_samples(speech(3))
    buf.push_samples(silence(1.5))
    assert buf.has_silence_boundary()
    chunk = buf.take_chunk_at_silence()
    assert chunk is not None
    assert len(buf.overlap_buffer) == SR * ao.CHUNK_OVERLAP_MS // 1000
    # the split is in the middle of the silence: 3 s of speech + half of the 1.5 s gap
    assert len(chunk) == 3 * SR + (SR * 3 // 2) // 2


def test_forced_chunk_splits_at_the_quietest_window():  # audio.rs:163-227 (no reference test: property check)
    buf = ao.AudioBuffer(SR)
    audio = speech(26).copy()
    audio[int(23.5 * SR): int(23.5 * SR) + 320] = 0.0     # one silent 20 ms window inside the last 5 s
    buf.push_samples(audio)
    assert buf.take_chunk_at_silence() is None
    chunk = buf.take_forced_chunk()
    assert chunk is not None and len(chunk) == int(23.5 * SR) + 160
    assert len(buf) == len(audio) - len(chunk)


def test_window_rms_is_the_sequential_float32_sum():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(320 * 7 + 11).astype(np.float32)
    r = ao.window_rms(x, 320)
    assert r.shape == (7,)
    for i in range(7):
        s = np.float32(0.0)
        for v in x[i * 320:(i + 1) * 320]:
            s = np.float32(s + np.float32(v * v))
        assert r[i] == np.sqrt(np.float32(s / np.float32(320)))


def test_resample_ratio():   # audio.rs:570-583, the reference's only test of the resampler
    x = np.sin(np.arange(48000, dtype=np.float32) * np.float32(0.001)).astype(np.float32)
    y = ao.resample_audio(x, 48000, 16000)
    expected = len(x) // 3
    assert abs(len(y) - expected) < expected // 10
    assert len(y) == 15872        # 47 chunks of 1024 -> 31 complete 1536-frame blocks -> 31 * 512 frames


def test_resampler_oracle_is_a_low_pass_decimator():
    """Signal-level properties of the restated design (it cannot be pinned against rubato itself offline)."""
    t = np.arange(96000) / 48000.0
    for f, keep in ((440.0, True), (3000.0, True), (7000.0, True), (9000.0, False), (15000.0, False)):
        y = ao.resample_audio(np.sin(2 * np.pi * f * t).astype(np.float32), 48000, 16000)
        body = y[2000:-200]
        if keep:   # same tone at 16 kHz, delayed by half the 1536-tap filter (256 output frames)
            ref = np.sin(2 * np.pi * f * (np.arange(len(y)) - 256) / 16000.0)[2000:-200]
            assert np.abs(body - ref).max() < 1e-4
        else:      # above the new Nyquist: rejected, not aliased
            assert np.abs(body).max() < 1e-5
    assert np.array_equal(ao.resample_chunk(np.ones(5, np.float32), 16000), np.ones(5, np.float32))
    assert np.array_equal(ao.mix_to_mono(np.array([1, 3, 2, 4], np.float32), 2), np.array([2, 3], np.float32))
