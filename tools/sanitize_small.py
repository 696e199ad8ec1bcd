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

# the WBFM kernels of every generation on a bank that fills CTAs (generation 4: every M-block, the hole at the
# recurrence warp's slot, the MMA warp as warp 15; both geometries; the legacy mma.sync pre-filter of 2 and 3),
# input without clipping bytes so that the tensor-core paths run, plus one channel that clips
n = 61
for gen, per_cta in ((5, 28), (6, 14), (3 | 16, 28), (2 | 16, 14), (3, 28), (2, 14)):
    e = R.Engine(n, 0, 8192)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    e.set_launch_shape(R.KIND_WBFM, per_cta, 0)
    e.debug_set_wbfm_kernel(gen)
    iq = np.clip(np.round(128 + 45 * rng.standard_normal((n, 8192 + 1024 + 64))), 1, 255).astype(np.uint8)
    iq[7, 3000:3100] = 0
    pcm, counts = e.demodulate(iq)
    print("ok wbfm generation", gen, int(pcm.astype(np.int64).sum()))
    e.close()
