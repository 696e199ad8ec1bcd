"""The three generations of the WBFM kernel (WbFmDemodulator.cc:383-562) against the oracle:
1 = atan2 table gathered from global memory, 2 = table in shared memory and one channel per
worker warp, 3 (default) = two channels per worker warp, 512 samples per channel and round.
They share the carry blob, so a stream may change kernel between calls."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def _oracle_rows(n, iq, gains=None):
    rows = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(3)
        if gains is not None:
            c.set_gain(O.KIND_WBFM, float(gains[ch]))
        rows.append(c.accept_u8(iq[ch]))
    return rows


@pytest.mark.parametrize("gen", [2, 3])
@pytest.mark.parametrize("n", [1, 2, 29, 57, 300])
def test_generations_match_oracle(gen, n):
    import rtlsdrdiags_b200 as R
    nbytes = 3 * 32768
    e = R.Engine(n, 0, 32768)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    e.debug_set_wbfm_kernel(gen)
    iq = S.noise(n, nbytes, seed=n)
    iq[0] = S.tone(3, nbytes // 2, seed=4)
    pcm, counts = e.demodulate(iq)
    assert (counts == 512).all()
    exp = _oracle_rows(n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()


def test_ragged_calls_and_kernel_switches():
    """Call lengths that leave partial tiles and partial half-tiles (any multiple of 64 bytes),
    the kernel generation changed between calls, a large gain on one channel (the wrapping
    (int16_t) conversion and the clamp path of the audio decimator)."""
    import rtlsdrdiags_b200 as R
    n = 11
    e = R.Engine(n, 0, 4 * 32768)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    gains = [40743.664] * n
    gains[3] = 40743.664 * 40
    gains[4] = 1e9
    for ch in range(n):
        e.set_gain(ch, R.KIND_WBFM, gains[ch])
    sizes = [64, 64 * 15, 64 * 16, 64 * 17, 32768, 64 * 33, 4 * 32768, 64 * 31, 1024 + 64, 2048, 64 * 47, 32768 + 64 * 5]
    gens = [3, 3, 2, 3, 1, 3, 3, 2, 3, 3, 1, 3]
    iq = S.noise(n, sum(sizes), seed=12)
    iq[1] = S.tone(3, sum(sizes) // 2, seed=2)
    out, off = [], 0
    for sz, g in zip(sizes, gens):
        e.debug_set_wbfm_kernel(g)
        e.accept_iq_host(np.ascontiguousarray(iq[:, off:off + sz]))
        out.append(e.get_pcm()[0])
        off += sz
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(n, iq, gains=[np.float32(g) for g in gains])
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()


def test_squelched_half_and_reset():
    """An odd channel count leaves the last worker warp's second half without a channel; a
    squelched channel shares a warp with an open one; resetDemodulator keeps the de-emphasis state
    (WbFmDemodulator.cc:304-320)."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 5, 32768
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    chains = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(3)
        thr = -10 if ch in (1, 2) else -200
        c.set_threshold(thr)
        e.set_squelch_threshold(ch, thr)
        chains.append(c)
    rng = np.random.default_rng(5)
    for step, amp in enumerate((60.0, 1.0, 1.0, 60.0, 60.0)):
        if step == 3:
            e.reset(0, R.KIND_WBFM)
            chains[0].reset(O.KIND_WBFM)
        x = 128 + amp * rng.standard_normal((n, nbytes))
        iq = np.clip(np.round(x), 0, 255).astype(np.uint8)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            exp = chains[ch].accept_u8(iq[ch])
            assert counts[ch] == exp.size, "step %d channel %d" % (step, ch)
            if exp.size:
                assert np.array_equal(pcm[ch], exp), "step %d channel %d" % (step, ch)
    e.close()
