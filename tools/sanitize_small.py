"""Tiny run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtlsdrdiags_b200 as R
n = 23
modes = np.array([1 + c % 5 for c in range(n)], dtype=np.uint8)
e = R.Engine(n, 0, 32768)
e.set_modes(modes)
rng = np.random.default_rng(1)
for nbytes in (32768, 4096 + 64, 64):
    iq = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    pcm, counts = e.demodulate(iq)
print("ok", int(pcm.astype(np.int64).sum()))
