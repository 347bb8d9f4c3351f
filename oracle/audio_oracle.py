"""CPU restatement of the reference's silence chunker (src-tauri/src/audio.rs) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module; the product
path (nobs_whisper_b200/csrc/audio_chunker.cu + nobs_whisper_b200/audio.py) never does.

Every function cites the reference lines it follows.  Arithmetic is float32 in the reference's own
order: `calculate_rms` adds the squares sequentially (Rust `iter().map(|s| s * s).sum()`), which a
float32 `cumsum` reproduces bit for bit.  Pinned by the reference's own unit tests
(audio.rs:570-831), restated in tests/test_audio_oracle.py.
"""
from __future__ import annotations

import numpy as np

F = np.float32

# audio.rs:7-15, 337-360
WHISPER_SAMPLE_RATE = 16000
MAX_BUFFER_DURATION_S = 25
CHUNK_OVERLAP_MS = 200
SILENCE_THRESHOLD = F(0.01)
MIN_SILENCE_DURATION_MS = 700
MIN_CHUNK_DURATION_MS = 1000
NOISE_FLOOR_UPDATE_MAX_FRAMES = 100
ADAPTIVE_THRESHOLD_NOISE_FACTOR = F(3.0)
MIN_THRESHOLD_FACTOR = F(0.5)
NOISE_FLOOR_EMA_DECAY = F(0.95)
NOISE_FLOOR_UPDATE_THRESHOLD_FACTOR = F(0.5)
NOISE_FLOOR_ESTIMATION_WINDOWS = 25
NOISE_FLOOR_PERCENTILE = F(0.1)
MIN_NOISE_FLOOR_FACTOR = F(0.3)


def calculate_rms(samples) -> np.float32:
    """audio.rs:364-370."""
    x = np.asarray(samples, dtype=F)
    if x.size == 0:
        return F(0.0)
    s = np.cumsum(x * x, dtype=F)[-1]          # sequential float32 sum
    return np.sqrt(F(s / F(x.size)), dtype=F)


def window_rms(audio, window_size: int) -> np.ndarray:
    """RMS of every full window [i*w, (i+1)*w) — the values the loops at audio.rs:376-381 and :432-433 see."""
    x = np.asarray(audio, dtype=F)
    n_win = x.size // window_size
    if n_win == 0:
        return np.zeros(0, F)
    w = x[: n_win * window_size].reshape(n_win, window_size)
    s = np.cumsum(w * w, axis=1, dtype=F)[:, -1]
    return np.sqrt((s / F(window_size)).astype(F), dtype=F)


def estimate_noise_floor(audio, sample_rate: int, rms: np.ndarray | None = None) -> np.float32:
    """audio.rs:373-395."""
    window_size = sample_rate // 50
    if rms is None:
        rms = window_rms(np.asarray(audio, dtype=F)[: NOISE_FLOOR_ESTIMATION_WINDOWS * window_size], window_size)
    vals = sorted(float(v) for v in rms[:NOISE_FLOOR_ESTIMATION_WINDOWS])
    if not vals:
        return SILENCE_THRESHOLD
    idx = int(F(len(vals)) * NOISE_FLOOR_PERCENTILE)
    nf = F(vals[idx]) if idx < len(vals) else SILENCE_THRESHOLD
    return max(nf, F(SILENCE_THRESHOLD * MIN_NOISE_FLOOR_FACTOR))


def find_silence_boundaries(audio, sample_rate: int) -> list[int]:
    """audio.rs:400-463."""
    x = np.asarray(audio, dtype=F)
    min_silence = sample_rate * MIN_SILENCE_DURATION_MS // 1000
    min_chunk = sample_rate * MIN_CHUNK_DURATION_MS // 1000
    w = sample_rate // 50
    rms = window_rms(x, w)
    noise_floor = estimate_noise_floor(x, sample_rate, rms)
    thr = max(F(noise_floor * ADAPTIVE_THRESHOLD_NOISE_FACTOR), F(SILENCE_THRESHOLD * MIN_THRESHOLD_FACTOR))
    boundaries: list[int] = []
    last_boundary = 0
    silence_start = None

    def try_add(start: int, end: int):
        nonlocal last_boundary
        dur = end - start
        if dur >= min_silence:
            split = start + dur // 2
            if split - last_boundary >= min_chunk:
                boundaries.append(split)
                last_boundary = split

    pos = 0
    for r in rms:
        if r < thr:
            if silence_start is None:
                silence_start = pos
        else:
            if silence_start is not None:
                try_add(silence_start, pos)
            silence_start = None
        pos += w
    if silence_start is not None:
        try_add(silence_start, x.size)
    return boundaries


def split_at_silences_with_overlap(audio, boundaries, sample_rate: int) -> list[np.ndarray]:
    """audio.rs:473-507."""
    x = np.asarray(audio, dtype=F)
    if len(boundaries) == 0:
        return [x.copy()]
    overlap = sample_rate * CHUNK_OVERLAP_MS // 1000
    chunks = []
    start = 0
    for b in boundaries:
        if start < b < x.size:
            chunks.append(x[max(0, start - overlap): b].copy())
            start = b
    if start < x.size:
        chunks.append(x[max(0, start - overlap):].copy())
    return chunks


def split_at_silences(audio, boundaries) -> list[np.ndarray]:
    """audio.rs:467-469."""
    return split_at_silences_with_overlap(audio, boundaries, WHISPER_SAMPLE_RATE)


class AudioBuffer:
    """audio.rs:29-244 (the streaming side: adaptive noise floor, silence boundary, forced split, overlap)."""

    def __init__(self, sample_rate: int = 48000):
        self.samples = np.zeros(0, F)
        self.last_speech_pos = 0
        self.sample_rate = sample_rate
        self.noise_floor = SILENCE_THRESHOLD
        self.noise_floor_frames = 0
        self.overlap_buffer = np.zeros(0, F)

    def push_samples(self, samples):  # audio.rs:59-86
        s = np.asarray(samples, dtype=F)
        start_pos = self.samples.size
        self.samples = np.concatenate([self.samples, s])
        w = self.sample_rate // 50
        for i in range(0, (s.size + w - 1) // w):
            rms = calculate_rms(s[i * w:(i + 1) * w])          # the last chunk may be short (slice::chunks)
            if rms < F(self.noise_floor * NOISE_FLOOR_UPDATE_THRESHOLD_FACTOR) and self.noise_floor_frames < NOISE_FLOOR_UPDATE_MAX_FRAMES:
                self.noise_floor = F(F(self.noise_floor * NOISE_FLOOR_EMA_DECAY) + F(rms * F(F(1.0) - NOISE_FLOOR_EMA_DECAY)))
                self.noise_floor_frames += 1
            thr = max(F(self.noise_floor * ADAPTIVE_THRESHOLD_NOISE_FACTOR), F(SILENCE_THRESHOLD * MIN_THRESHOLD_FACTOR))
            if rms >= thr:
                self.last_speech_pos = start_pos + (i + 1) * w

    def take(self):  # audio.rs:88-92
        self.last_speech_pos = 0
        self.overlap_buffer = np.zeros(0, F)
        out, self.samples = self.samples, np.zeros(0, F)
        return out

    def has_silence_boundary(self) -> bool:  # audio.rs:96-105
        if self.samples.size == 0 or self.last_speech_pos == 0:
            return False
        silence = max(0, self.samples.size - self.last_speech_pos)
        return silence >= self.sample_rate * MIN_SILENCE_DURATION_MS // 1000

    def _cut(self, split_point: int):
        overlap = self.sample_rate * CHUNK_OVERLAP_MS // 1000
        chunk = np.concatenate([self.overlap_buffer, self.samples[:split_point]])
        self.overlap_buffer = self.samples[max(0, split_point - overlap): split_point].copy()
        self.samples = self.samples[split_point:].copy()
        return chunk

    def take_chunk_at_silence(self):  # audio.rs:110-158
        if not self.has_silence_boundary():
            return None
        if self.last_speech_pos < self.sample_rate // 2:
            return None
        silence_start = self.last_speech_pos
        split_point = silence_start + (self.samples.size - silence_start) // 2
        chunk = self._cut(split_point)
        self.last_speech_pos = 0
        return chunk

    def take_forced_chunk(self):  # audio.rs:163-227
        max_samples = self.sample_rate * MAX_BUFFER_DURATION_S
        if self.samples.size <= max_samples:
            return None
        w = self.sample_rate // 50
        search_start = max(0, self.samples.size - self.sample_rate * 5)
        quietest_pos, quietest = search_start, F(np.finfo(F).max)
        pos = search_start
        while pos + w <= self.samples.size:
            r = calculate_rms(self.samples[pos:pos + w])
            if r < quietest:
                quietest, quietest_pos = r, pos
            pos += w
        split_point = min(quietest_pos + w // 2, self.samples.size)
        if split_point < self.sample_rate // 2:
            return None
        chunk = self._cut(split_point)
        self.last_speech_pos = self.last_speech_pos - split_point if self.last_speech_pos > split_point else 0
        return chunk

    def __len__(self):
        return int(self.samples.size)


# ------------------------------------------------------------------------------------------------------
# Resampler: reference audio.rs:509-563 `resample_audio` = rubato 0.15.0 (src-tauri/Cargo.lock:3896-3898)
# `FftFixedIn::<f32>::new(from, to, 1024, 2, 1)` fed 1024-frame chunks (last one zero-padded), output truncated to
# floor(n * to / from).  rubato is NOT vendored in /root/reference (Cargo.lock only), so this is a restatement of
# its published algorithm from the crate's documented design — PARITY UNPINNED beyond the reference's own test
# (audio.rs:570-583: output length within 10 % of n/3 for 48 kHz -> 16 kHz):
#   FftFixedIn: fft_chunks = ceil((chunk_size_in / sub_chunks) / (to / gcd)); fft_size_in/out = fft_chunks * from|to / gcd;
#     input is buffered and every complete fft_size_in block yields fft_size_out frames (a call may yield none).
#   FftResampler: sinc low-pass of fft_size_in taps (cutoff 0.4^(16/fft_size_in) [* out/in when decimating],
#     Blackman-Harris^2 window, unit DC gain) scaled by 1/(2 fft_size_in); per block: zero-pad to 2 fft_size_in, real
#     FFT, multiply the first min(fft_size_in + 1, fft_size_out) bins, inverse real FFT of 2 fft_size_out points,
#     overlap-add the second half into the next block.
def _rubato_sizes(fs_in: int, fs_out: int, chunk_size_in: int = 1024, sub_chunks: int = 2):
    from math import gcd
    g = gcd(fs_in, fs_out)
    fft_chunks = int(np.ceil(F(chunk_size_in // sub_chunks) / F(fs_out // g)))
    return fft_chunks * fs_in // g, fft_chunks * fs_out // g


def _rubato_filter(n_in: int, n_out: int) -> np.ndarray:
    """make_sincs(n_in, 1, cutoff, BlackmanHarris2)[0] in float64 (the crate computes it in f32)."""
    cutoff = float(F(0.4) ** F(16.0 / n_in)) * (n_out / n_in if n_in > n_out else 1.0)
    x = np.arange(n_in, dtype=np.float64)
    xf = x / n_in
    bh = 0.35875 - 0.48829 * np.cos(2 * np.pi * xf) + 0.14128 * np.cos(4 * np.pi * xf) - 0.01168 * np.cos(6 * np.pi * xf)
    h = bh * bh * np.sinc((x - n_in // 2) * cutoff)
    return h / h.sum()


def resample_audio(audio, from_rate: int, to_rate: int) -> np.ndarray:
    """audio.rs:509-563."""
    x = np.asarray(audio, dtype=np.float64)
    n = x.size
    n_in, n_out = _rubato_sizes(from_rate, to_rate)
    h = _rubato_filter(n_in, n_out)
    filt = np.fft.rfft(np.concatenate([h / (2 * n_in), np.zeros(n_in)]))
    new_len = n_in + 1 if n_in < n_out else n_out
    padded = np.concatenate([x, np.zeros((-n) % 1024)])               # chunks of 1024, the last one zero-padded
    n_blocks = padded.size // n_in                                    # frames left over are never flushed
    out = np.zeros(n_blocks * n_out)
    overlap = np.zeros(n_out)
    for b in range(n_blocks):
        spec = np.fft.rfft(np.concatenate([padded[b * n_in:(b + 1) * n_in], np.zeros(n_in)]))
        o = np.zeros(n_out + 1, np.complex128)
        o[:new_len] = spec[:new_len] * filt[:new_len]
        o[0] = o[0].real                                              # the inverse real FFT ignores Im(DC)
        y = np.fft.irfft(o, 2 * n_out) * (2 * n_out)                  # realfft's inverse is unnormalised
        out[b * n_out:(b + 1) * n_out] = y[:n_out] + overlap
        overlap = y[n_out:]
    expected = int(n * (to_rate / from_rate))
    return out[:expected].astype(F)


def resample_chunk(audio, input_sample_rate: int) -> np.ndarray:
    """audio.rs:329-334."""
    if input_sample_rate == WHISPER_SAMPLE_RATE:
        return np.asarray(audio, dtype=F).copy()
    return resample_audio(audio, input_sample_rate, WHISPER_SAMPLE_RATE)


def mix_to_mono(interleaved, channels: int) -> np.ndarray:
    """state.rs:590-594: mono = sum of the frame's channels / channels (float32, left to right)."""
    x = np.asarray(interleaved, dtype=F)
    fr = x[: x.size // channels * channels].reshape(-1, channels)
    s = np.cumsum(fr, axis=1, dtype=F)[:, -1]
    return (s / F(channels)).astype(F)
