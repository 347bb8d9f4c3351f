"""Seeded synthetic 16 kHz mono PCM in [-1, 1].

Follows the reference's own synthetic-signal idiom (src-tauri/src/audio.rs:627-654):
sinusoid "speech" at ~0.3 peak, low-level background, and zero-valued silences.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000


def synth_clip(clip_index: int, seconds: float = 30.0, base_seed: int = 1234) -> np.ndarray:
    n = int(round(seconds * SAMPLE_RATE))
    rng = np.random.default_rng(base_seed + clip_index)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    k = int(rng.integers(3, 6))
    freqs = rng.uniform(100.0, 4000.0, size=k)
    phases = rng.uniform(0.0, 2 * np.pi, size=k)
    amps = rng.uniform(0.3, 1.0, size=k)
    sig = np.zeros(n, dtype=np.float64)
    for f, p, a in zip(freqs, phases, amps):
        sig += a * np.sin(2 * np.pi * f * t + p)
    sig /= amps.sum()
    am = 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(0.5, 3.0) * t + rng.uniform(0, 2 * np.pi))
    sig = 0.3 * sig * am + rng.standard_normal(n) * 0.01
    # zero gaps of 0.3-1.0 s every few seconds
    pos = int(rng.uniform(1.0, 3.0) * SAMPLE_RATE)
    while pos < n:
        gap = int(rng.uniform(0.3, 1.0) * SAMPLE_RATE)
        sig[pos:pos + gap] = 0.0
        pos += gap + int(rng.uniform(2.0, 5.0) * SAMPLE_RATE)
    return np.clip(sig, -1.0, 1.0).astype(np.float32)


def synth_batch(n_clips: int, seconds: float = 30.0, base_seed: int = 1234, first: int = 0) -> list[np.ndarray]:
    return [synth_clip(first + i, seconds, base_seed) for i in range(n_clips)]
