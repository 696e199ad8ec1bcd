"""The AM / SSB DC-removal recurrence (IirFilter.cc:161-176, AmDemodulator.cc:460-471,
SsbDemodulator.cc:586-598) runs segment-parallel on the GPU: a segment starts early from y = 0,
and its state on entering its own rows is compared bit for bit with its predecessor's final state;
a segment that does not verify is redone from the true state (dc_block_kernel). These tests
compare with the oracle at the default segmentation, force the redo path, and feed the input
that can never verify (y stuck on a denormal under all-zero numerators)."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu
BLOCK = 32768


def _oracle_rows(modes, iq):
    rows = []
    for ch, m in enumerate(modes):
        c = O.OracleChain()
        c.set_mode(int(m))
        rows.append(c.accept_u8(iq[ch]))
    return rows


@pytest.mark.parametrize("mode", [1, 4, 5])
@pytest.mark.parametrize("signal", ["noise", "tone"])
def test_long_calls_default_segmentation(mode, signal):
    """Sixteen reference blocks per call (bench.py's AM shape): sixteen segments per channel."""
    import rtlsdrdiags_b200 as R
    n, blocks, calls = 9, 16, 3
    e = R.Engine(n, 0, blocks * BLOCK)
    e.set_modes(np.full(n, mode, dtype=np.uint8))
    nbytes = blocks * BLOCK * calls
    iq = S.noise(n, nbytes, seed=40 + mode) if signal == "noise" else S.tone_bank([mode] * n, nbytes, seed=mode)
    pcm, counts = e.demodulate(iq)
    exp = _oracle_rows([mode] * n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    # trajectories merge before the 896-step warm-up ends (tools/iir_merge.py: latest 717-807 steps in
    # 200,000 trials per input class): no segment needed the serial redo
    assert e.debug_dc_redo_count() == 0
    e.close()


@pytest.mark.parametrize("seg_count,warm_rows", [(8, 1), (32, 0), (4, 2), (16, 3)])
def test_short_warm_up_forces_the_serial_redo(seg_count, warm_rows):
    """A warm-up of 0-96 steps cannot merge: segments fail the boundary check and are redone."""
    import rtlsdrdiags_b200 as R
    n, blocks = 11, 16
    modes = np.array([(1, 4, 5)[ch % 3] for ch in range(n)], dtype=np.uint8)
    e = R.Engine(n, 0, blocks * BLOCK)
    e.set_modes(modes)
    e.debug_set_dc_shape(seg_count, warm_rows)
    iq = S.noise(n, 2 * blocks * BLOCK, seed=77)
    pcm, _ = e.demodulate(iq)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d mode %d" % (ch, modes[ch])
    assert e.debug_dc_redo_count() >= n * (seg_count - 1), e.debug_dc_redo_count()
    e.close()


@pytest.mark.parametrize("seg_count", [0, 4, 32])
def test_ragged_calls_with_segments(seg_count):
    """Calls whose length is not a whole number of rows (32 PCM samples) or of segments; the
    partial last row ends the recurrence at the right sample."""
    import rtlsdrdiags_b200 as R
    n = 7
    modes = np.array([(1, 4, 5)[ch % 3] for ch in range(n)], dtype=np.uint8)
    e = R.Engine(n, 0, 16 * BLOCK)
    e.set_modes(modes)
    e.debug_set_dc_shape(seg_count, 1 if seg_count else 32)
    sizes = [64 * (32 * 5 + 7), 64, 64 * 33, 16 * BLOCK, 64 * (32 * 70 + 31), 64 * 31, 4 * BLOCK + 64]
    iq = S.noise(n, sum(sizes), seed=3)
    out, off = [], 0
    for sz in sizes:
        e.accept_iq_host(np.ascontiguousarray(iq[:, off:off + sz]))
        out.append(e.get_pcm()[0])
        off += sz
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d mode %d" % (ch, modes[ch])
    e.close()


def test_silence_after_signal_never_verifies_and_is_still_exact():
    """After a signal, constant input makes every numerator 0 and y decays onto a denormal it
    cannot leave (0.95 * y rounds back to y), while a segment started from 0 stays at 0: the
    boundary check fails in every segment, every call, and the kernel falls back to the serial
    order. PCM and the carried state must still be the reference's."""
    import rtlsdrdiags_b200 as R
    n, blocks = 6, 16
    modes = np.array([1, 4, 5, 1, 4, 5], dtype=np.uint8)
    e = R.Engine(n, 0, blocks * BLOCK)
    e.set_modes(modes)
    rng = np.random.default_rng(9)
    loud = rng.integers(0, 256, size=(n, blocks * BLOCK), dtype=np.uint8)
    quiet = np.full((n, 2 * blocks * BLOCK), 128, dtype=np.uint8)
    quiet[3:] = 200                      # a constant carrier instead of silence
    again = rng.integers(0, 256, size=(n, blocks * BLOCK), dtype=np.uint8)
    iq = np.concatenate([loud, quiet, again], axis=1)
    pcm, _ = e.demodulate(iq)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    assert e.debug_dc_redo_count() > 0
    e.close()


def test_squelched_channels_keep_their_recurrence_state():
    import rtlsdrdiags_b200 as R
    n, blocks = 8, 16
    modes = np.array([1, 4] * 4, dtype=np.uint8)
    e = R.Engine(n, 0, blocks * BLOCK)
    e.set_modes(modes)
    chains = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(int(modes[ch]))
        thr = -10 if ch % 4 < 2 else -200
        c.set_threshold(thr)
        e.set_squelch_threshold(ch, thr)
        chains.append(c)
    rng = np.random.default_rng(21)
    for amp in (50.0, 1.0, 1.0, 50.0):
        x = 128 + amp * rng.standard_normal((n, blocks * BLOCK))
        iq = np.clip(np.round(x), 0, 255).astype(np.uint8)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            # the gate is per call on the GPU: give the oracle the same one-block-per-call view
            c = chains[ch]
            exp = c.accept_u8(iq[ch])
            assert counts[ch] == exp.size, "channel %d" % ch
            if exp.size:
                assert np.array_equal(pcm[ch], exp), "channel %d" % ch
    e.close()
