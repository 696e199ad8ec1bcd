#!/usr/bin/env python
"""Generates tests/golden/golden_multirate_v1.npz with the UNMODIFIED reference classes
Decimator, Interpolator, Decimator_int16 and Interpolator_int16 compiled in place
(oracle/_ref, built by oracle/Makefile from /root/reference).

Run in the build container, where /root/reference is mounted:
    make -C oracle && python tests/golden/make_golden_multirate.py

All taps and inputs are synthetic (windowed-sinc prototypes designed here, a chirp plus noise);
no reference data file or tap table is copied. Per case c (see CASES): taps_c, x_c, y_c, and
kind/factor in `meta`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402


def lowpass(n_taps, cutoff, gain=1.0):
    """Hamming-windowed sinc, float32 taps (cutoff as a fraction of the sample rate)."""
    k = np.arange(n_taps) - (n_taps - 1) / 2.0
    h = 2 * cutoff * np.sinc(2 * cutoff * k) * np.hamming(n_taps)
    return (gain * h / h.sum()).astype(np.float32)


def audio(n, seed, full_scale=False):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 9000 * np.sin(2 * np.pi * (0.01 + 0.2 * t / n) * t / 2) + rng.normal(0, 800, n)
    if full_scale:  # drives the Q15 accumulator into its per-tap clamp
        x = 32767 * np.sign(np.sin(2 * np.pi * t / 37.0)) + rng.integers(-3, 3, n)
    return np.clip(np.round(x), -32768, 32767)


# (name, kind, taps, factor, input)
CASES = [
    ("dec_f32_80x4", O.MR_DECIMATOR_F32, lowpass(80, 0.11), 4, audio(4099, 1).astype(np.float32)),
    ("dec_f32_7x1", O.MR_DECIMATOR_F32, lowpass(7, 0.2), 1, audio(600, 2).astype(np.float32)),
    ("int_f32_64x2", O.MR_INTERPOLATOR_F32, lowpass(64, 0.22, 2.0), 2, audio(1501, 3).astype(np.float32)),
    ("int_f32_48x8", O.MR_INTERPOLATOR_F32, lowpass(48, 0.05, 8.0), 8, audio(333, 4).astype(np.float32)),
    ("dec_i16_80x4", O.MR_DECIMATOR_I16, lowpass(80, 0.11), 4, audio(4099, 5).astype(np.int16)),
    ("dec_i16_clamp", O.MR_DECIMATOR_I16, lowpass(24, 0.2, 1.9), 3, audio(2000, 6, True).astype(np.int16)),
    ("dec_i16_unity", O.MR_DECIMATOR_I16, np.array([0, 0, 1.0], dtype=np.float32), 1, audio(100, 7).astype(np.int16)),
    ("int_i16_64x2", O.MR_INTERPOLATOR_I16, lowpass(64, 0.22, 1.99), 2, audio(1501, 8).astype(np.int16)),
    ("int_i16_clamp", O.MR_INTERPOLATOR_I16, lowpass(36, 0.1, 5.5), 3, audio(900, 9, True).astype(np.int16)),
]


def main():
    if O.ref("radiodiags") is None:
        sys.exit("oracle/_ref is not built: run `make -C oracle` where /root/reference is mounted")
    out, meta = {}, []
    for name, kind, taps, factor, x in CASES:
        y = O.Multirate(kind, taps, factor, impl="ref").run(x)
        out["taps_" + name], out["x_" + name], out["y_" + name] = taps, x, y
        meta.append((name, kind, factor))
        print("%-14s kind %d  %d taps  factor %d  %d -> %d samples" % (name, kind, taps.size, factor, x.size, y.size))
    out["meta"] = np.array(["%s,%d,%d" % m for m in meta])
    np.savez_compressed(os.path.join(HERE, "golden_multirate_v1.npz"), **out)


if __name__ == "__main__":
    main()
