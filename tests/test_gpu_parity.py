"""GPU parity: the CUDA engine, called through the C ABI, against the oracle on
the same seeded inputs. Bit-exact for every mode (the float stages are written
to round exactly like the reference, so the +-1 LSB allowance is not used).
"""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu

KINDS = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}


def _engine(n, max_bytes=32768):
    import rtlsdrdiags_b200 as R
    return R, R.Engine(n, 0, max_bytes)


def _oracle_rows(modes, iq, gains=None):
    rows = []
    for ch, m in enumerate(modes):
        c = O.OracleChain()
        c.set_mode(int(m))
        if gains is not None and int(m):
            c.set_gain(O.MODE_TO_KIND[int(m)], float(gains[ch]))
        rows.append(c.accept_u8(iq[ch]))
    return rows


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("signal", ["noise", "tone"])
def test_single_mode_bank_matches_oracle(mode, signal):
    n, nbytes = 37, 32768 * 3
    R, e = _engine(n)
    e.set_modes(np.full(n, mode, dtype=np.uint8))
    iq = S.noise(n, nbytes, seed=mode) if signal == "noise" else S.tone_bank([mode] * n, nbytes, seed=mode)
    iq[0, : nbytes // 2] = 0
    iq[1, : nbytes // 2] = 255
    pcm, counts = e.demodulate(iq)
    assert (counts == 512).all()
    exp = _oracle_rows([mode] * n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d differs (max |d| = %d)" % (
            ch, np.abs(pcm[ch].astype(int) - exp[ch].astype(int)).max())
    # AM / SSB: a FIR kernel and a recurrence kernel per call; FM / WBFM: one kernel
    assert e.launch_count == 3 * (2 if mode in (1, 4, 5) else 1)


def test_mixed_mode_bank_and_none():
    n, nbytes = 64, 32768 * 2
    R, e = _engine(n)
    modes = np.array([ch % 6 for ch in range(n)], dtype=np.uint8)
    e.set_modes(modes)
    iq = S.noise(n, nbytes, seed=99)
    pcm, counts = e.demodulate(iq)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        if modes[ch] == 0:
            assert counts[ch] == 0 and exp[ch].size == 0
        else:
            assert np.array_equal(pcm[ch], exp[ch]), "channel %d mode %d" % (ch, modes[ch])


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(1, 32), (5, 96), (32, 1024), (7, 1024)])
def test_launch_shapes(mode, shape):
    n, nbytes = 45, 32768
    R, e = _engine(n)
    e.set_modes(np.full(n, mode, dtype=np.uint8))
    e.set_launch_shape(R.MODE_TO_KIND[mode], *shape)
    iq = S.noise(n, nbytes * 2, seed=7)
    pcm, _ = e.demodulate(iq)
    exp = _oracle_rows([mode] * n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch])


@pytest.mark.parametrize("mode", [1, 2, 3, 5])
def test_gains_large_and_small(mode):
    n, nbytes = 12, 32768 * 2
    R, e = _engine(n)
    e.set_modes(np.full(n, mode, dtype=np.uint8))
    base = {1: 300.0, 2: 10185.916, 3: 40743.664, 5: 300.0}[mode]
    gains = [base * g for g in (0.0, 0.01, 0.5, 1.0, 2.0, 3.7, 10.0, 100.0, 1e4, 1e7, 1e12, 1e30)]
    for ch, g in enumerate(gains):
        e.set_gain(ch, R.MODE_TO_KIND[mode], g)
    iq = S.noise(n, nbytes, seed=3)
    pcm, _ = e.demodulate(iq)
    exp = _oracle_rows([mode] * n, iq, gains=[np.float32(g) for g in gains])
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "gain %g" % gains[ch]


def test_block_partition_invariance_and_ragged_blocks():
    """cutting the stream differently must not change the PCM."""
    n = 10
    R, e1 = _engine(n)
    _, e2 = _engine(n)
    modes = np.array([1 + ch % 5 for ch in range(n)], dtype=np.uint8)
    e1.set_modes(modes)
    e2.set_modes(modes)
    # rotation phase restarts per call: pieces must be multiples of 8 bytes; the
    # engine needs multiples of 64
    sizes = [64, 128, 32768, 4096 + 64, 192, 8192]
    iq = S.noise(n, sum(sizes), seed=5)
    a = []
    off = 0
    for s in sizes:
        p, _ = e1.demodulate(iq[:, off:off + s])
        a.append(p)
        off += s
    a = np.concatenate(a, axis=1)
    b, _ = e2.demodulate(iq)
    assert np.array_equal(a, b)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(b[ch], exp[ch])


def test_mode_switch_preserves_idle_state_and_reset():
    """IqDataProcessor.cc:793-835: idle demodulators keep their state; reset quirks."""
    n, nb = 6, 32768
    R, e = _engine(n)
    chains = [O.OracleChain() for _ in range(n)]
    rng = np.random.default_rng(17)
    script = [("mode", 3), ("run",), ("mode", 2), ("run",), ("mode", 3), ("run",), ("reset", 3), ("run",),
              ("mode", 4), ("run",), ("mode", 1), ("run",), ("mode", 5), ("run",), ("reset", 4), ("run",),
              ("mode", 0), ("run",), ("mode", 2), ("reset", 2), ("run",)]
    for step in script:
        if step[0] == "mode":
            for ch in range(n):
                e.set_mode(ch, step[1])
                chains[ch].set_mode(step[1])
        elif step[0] == "reset":
            kind = {1: 1, 2: 2, 3: 3, 4: 4}[step[1]]
            for ch in range(0, n, 2):
                e.reset(ch, kind)
                chains[ch].reset(kind)
        else:
            iq = rng.integers(0, 256, size=(n, nb), dtype=np.uint8)
            pcm, counts = e.demodulate(iq)
            for ch in range(n):
                exp = chains[ch].accept_u8(iq[ch])
                assert counts[ch] == exp.size
                if exp.size:
                    assert np.array_equal(pcm[ch], exp)


def test_signed_rotated_entry_and_research_scaling():
    """The demodulator classes' own entry (signed, rotated IQ) and the research tree's scaling."""
    n, nb = 8, 32768 * 2
    R, e = _engine(n)
    e.set_scaling(R.SCALING_RESEARCH)
    modes = np.array([1, 2, 3, 4, 5, 2, 3, 1], dtype=np.uint8)
    e.set_modes(modes)
    iq = np.random.default_rng(23).integers(-128, 128, size=(n, nb), dtype=np.int8)
    pcm, _ = e.demodulate(iq, fmt=R.IQ_S8_ROTATED)
    for ch in range(n):
        c = O.OracleChain(O.VARIANT_RESEARCH)
        assert np.array_equal(pcm[ch], c.accept_s8(int(modes[ch]), iq[ch]))


def test_device_resident_input_and_argument_errors():
    import torch
    n, nb = 16, 32768
    R, e = _engine(n)
    e.set_modes(np.full(n, 2, dtype=np.uint8))
    iq = S.noise(n, nb, seed=2)
    d = torch.from_numpy(iq).cuda()
    e.accept_iq_device(d)
    pcm, _ = e.get_pcm()
    exp = _oracle_rows([2] * n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch])
    with pytest.raises(R.SdrError):
        e.accept_iq_ptr(d.data_ptr(), 100, nb, R.IQ_DEVICE)          # not a multiple of 64
    with pytest.raises(R.SdrError):
        e.accept_iq_ptr(d.data_ptr(), nb * 2, nb * 2, R.IQ_DEVICE)   # longer than max_bytes
    with pytest.raises(R.SdrError):
        e.set_mode(n, 1)
    with pytest.raises(R.SdrError):
        e.set_mode(0, 6)


def test_channel_permutation_and_duplicates():
    """Identical channels give identical rows; permuting channels permutes rows."""
    n, nb = 40, 32768
    R, e = _engine(n)
    modes = np.array([1 + ch % 5 for ch in range(n)], dtype=np.uint8)
    e.set_modes(modes)
    base = S.noise(5, nb, seed=77)
    iq = base[np.arange(n) % 5]
    pcm, _ = e.demodulate(iq)
    for ch in range(5, n):
        assert np.array_equal(pcm[ch], pcm[ch % 5])


@pytest.mark.parametrize("fmt", ["u8", "s8"])
def test_fm_tensor_core_tuner_and_clipping_fallback(fmt):
    """NBFM's tuner decimators run as an int8 Toeplitz GEMM on the tensor cores unless a raw
    byte 0 sits where the Fs/4 rotation negates (-(-128) = -128 is not linear); then the tile
    takes the SIMT path. Clean tone input (GEMM path), the same with clipping bytes sprinkled
    into some tiles (mixed paths within one stream), block sizes that leave partial tiles, the
    signed/rotated input format, and state carried across calls: all bit-exact."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 6, 3 * 32768
    rng = np.random.default_rng(99)
    iq = np.stack([S.tone(2, nbytes // 2, seed=40 + ch) for ch in range(n)])
    assert iq.min() > 0                      # no clipping anywhere: GEMM path only
    # channel 1: zeros at negated positions (group bytes 3..6) of a few tiles
    for pos in (3, 2048 * 5 + 4, 2048 * 5 + 13, 2048 * 17 + 2046, 2048 * 30 + 6, nbytes - 3):
        iq[1, pos] = 0
    # channel 2: zeros only at positions the rotation does not negate, and 255s
    for pos in (0, 1, 2, 7, 2048 * 9 + 8, 2048 * 9 + 15):
        iq[2, pos] = 0
    iq[2, 5000:5032] = 255
    # channel 3: a burst of clipping across a tile boundary (history of the next tile)
    iq[3, 2048 * 11 - 40:2048 * 11 + 8] = 0
    # channel 4: noise in the middle of a tone
    iq[4, 30000:50000] = rng.integers(0, 256, 20000, dtype=np.uint8)
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 2, dtype=np.uint8))
    chains = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(2)
        chains.append(c)
    cuts = [0, 32768, 32768 + 2048 * 3 + 64, 32768 * 2 + 640, nbytes]
    for a, b in zip(cuts[:-1], cuts[1:]):
        piece = np.ascontiguousarray(iq[:, a:b])
        if fmt == "u8":
            e.accept_iq_host(piece)
        else:
            e.accept_iq_host(np.stack([O.front_end(piece[ch]) for ch in range(n)]), R.IQ_S8_ROTATED)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            want = chains[ch].accept_u8(piece[ch])
            assert counts[ch] == want.size
            assert np.array_equal(pcm[ch][:counts[ch]], want), (fmt, a, ch)


def test_first_generation_wbfm_kernel_still_matches_the_oracle():
    """wbfm_tile_kernel (atan2 gathered from global memory) is the fallback when the host libm's
    atan2 is not odd in q; SDR_WB_KERNEL=1 selects it. The choice is read once per process, so the
    check runs in a child process."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import _oracle as O, _signals as S
import rtlsdrdiags_b200 as R
n, nbytes = 21, 32768 * 2 + 4096
e = R.Engine(n, 0, 32768)
e.set_modes(np.full(n, 3, dtype=np.uint8))
iq = np.concatenate([S.noise(7, nbytes, seed=3), S.tone_bank([3] * 14, nbytes, seed=4)])
pcm, counts = e.demodulate(iq)
for ch in range(n):
    c = O.OracleChain(); c.set_mode(3)
    assert np.array_equal(pcm[ch], c.accept_u8(iq[ch])), ch
print("ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, SDR_WB_KERNEL="1"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
