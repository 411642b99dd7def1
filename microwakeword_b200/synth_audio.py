"""Deterministic synthetic 16 kHz int16 audio used by tests/, bench.py and smoke()
(Gaussian noise with log-uniform level + tone bursts, plus adversarial edge cases)."""

import numpy as np


def synth_audio(n_samples: int, seed: int) -> np.ndarray:
    """Gaussian noise with log-uniform level plus two tone bursts."""
    rng = np.random.default_rng(seed)
    sigma = np.exp(rng.uniform(np.log(50.0), np.log(8000.0)))
    x = rng.normal(0.0, sigma, n_samples)
    t = np.arange(n_samples)
    for _ in range(2):
        f = rng.uniform(200.0, 4000.0)
        a = rng.uniform(500.0, 12000.0)
        length = int(rng.integers(min(1600, n_samples // 2), max(min(8000, n_samples), 1601)))
        start = int(rng.integers(0, max(n_samples - length, 1)))
        x[start:start + length] += a * np.sin(2 * np.pi * f * t[start:start + length] / 16000.0)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def edge_case_audio(n_samples: int) -> np.ndarray:
    """Adversarial streams: silence, full-scale square waves, DC, impulse, int16 extremes, ramp."""
    t = np.arange(n_samples)
    rows = [
        np.zeros(n_samples),
        np.where((t // 8) % 2 == 0, 32767, -32768),
        np.where((t // 2) % 2 == 0, 32767, -32768),
        np.where(t % 2 == 0, 32767, -32768),
        np.full(n_samples, 12345),
        np.full(n_samples, -32768),
        np.full(n_samples, 32767),
        (t == n_samples // 3) * 32767,
        ((t * 37) % 65536) - 32768,
        np.where((t // 1000) % 2 == 0, 1, -1),
        32767 * np.sin(2 * np.pi * 4000.0 * t / 16000.0),
        32767 * np.sin(2 * np.pi * 1000.0 * t / 16000.0 + 0.3),
        np.where(t < n_samples // 2, 0, 20000 * np.sin(2 * np.pi * 440.0 * t / 16000.0)),
        # full-scale complex-exponential pattern that stresses the int16 wrap inside the FFT butterflies
        32767 * np.sign(np.sin(2 * np.pi * (t // 2) / 4.0 + np.pi / 4 + (t % 2) * np.pi / 2) + 1e-9),
    ]
    return np.clip(np.round(np.stack(rows)), -32768, 32767).astype(np.int16)
