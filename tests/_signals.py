"""Synthetic IQ generators shared by the tests and bench.py.

All return u8 offset-binary interleaved I,Q as the RTL-SDR delivers it, i.e. the
input of IqDataProcessor::acceptIqData, with the wanted signal at -Fs/4 (the
radio tunes Fs/4 high, Radio.cc:617-618).
"""
import numpy as np

FS = 256000.0
MODE_AM, MODE_FM, MODE_WBFM, MODE_LSB, MODE_USB = 1, 2, 3, 4, 5


def noise(n_channels, nbytes, seed=0xB200):
    """iid uniform bytes: exercises every wrap / clamp quirk."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n_channels, nbytes), dtype=np.uint8)


def tone(mode, n_samples, seed=0, start=0):
    """One channel of a modulated carrier for `mode`, samples [start, start+n)."""
    rng = np.random.default_rng([seed, start])
    t = (np.arange(n_samples, dtype=np.float64) + start) / FS
    if mode == MODE_AM:
        env, ph = 1.0 + 0.5 * np.sin(2 * np.pi * 1000 * t), 0.0 * t
    elif mode == MODE_FM:
        env, ph = 1.0, (5000.0 / 1000.0) * np.sin(2 * np.pi * 1000 * t)
    elif mode == MODE_WBFM:
        env, ph = 1.0, (75000.0 / 1000.0) * np.sin(2 * np.pi * 1000 * t)
    else:  # SSB: a 1.5 kHz tone in the wanted sideband
        sign = -1.0 if mode == MODE_LSB else 1.0
        env, ph = 1.0, sign * 2 * np.pi * 1500 * t
    z = (100.0 / 1.5) * env * np.exp(1j * (ph - 2 * np.pi * (FS / 4) * t))
    i = np.clip(np.round(128 + z.real + rng.normal(0, 2, n_samples)), 0, 255)
    q = np.clip(np.round(128 + z.imag + rng.normal(0, 2, n_samples)), 0, 255)
    return np.stack([i, q], axis=1).reshape(-1).astype(np.uint8)


def tone_bank(modes, nbytes, seed=0, start=0):
    return np.stack([tone(int(m), nbytes // 2, seed=seed + ch, start=start) for ch, m in enumerate(modes)])
