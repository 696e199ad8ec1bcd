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


@pytest.mark.parametrize("gen", [0, 3, 5, 6])
def test_squelched_half_and_reset(gen):
    """An odd channel count leaves the last worker warp's second half without a channel; a
    squelched channel shares a warp with an open one; resetDemodulator keeps the de-emphasis state
    (WbFmDemodulator.cc:304-320). Generations 3 and 4 (both geometries; the quiet steps take the tensor
    cores, the loud ones clip)."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 5, 32768
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    e.debug_set_wbfm_kernel(gen)
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


def _quiet(n, nbytes, seed, amp=45.0):
    """Band-limited-ish noise that never reaches the rails: no raw byte 0, the tensor-core path's case."""
    rng = np.random.default_rng(seed)
    x = 128 + amp * rng.standard_normal((n, nbytes))
    return np.clip(np.round(x), 1, 255).astype(np.uint8)


@pytest.mark.parametrize("gen", [2 | 16, 3 | 16, 5, 6])
@pytest.mark.parametrize("n", [1, 2, 29, 61, 330])
def test_tensor_core_prefilter(gen, n):
    """The pre-filter as an int8 GEMM on the raw bytes: generation 4 (tcgen05.mma, accumulators in tensor
    memory; 5 / 6 = with two / one channel(s) per worker warp) and the legacy mma.sync option of generations
    2 and 3 (+ 16). Streams without a clipping byte take the tensor cores for every full (half-)tile; a raw
    byte 0 where the Fs/4 rotation negates sends exactly the tiles that see it -- the one that holds it and,
    through the 15 samples of history, the next -- down the CUDA-core path. Bit-exact against the oracle
    either way, and against the CUDA-core kernel."""
    import rtlsdrdiags_b200 as R
    nbytes = 2 * 32768
    iq = _quiet(n, nbytes, seed=100 + n)
    iq[0, 2048 * 3 + 64 * 5 + 3] = 0          # Q1: negated -> the tile falls back
    iq[0, 2048 * 7 + 2047] = 0                # the tile's last byte (Q3, not negated): no fallback
    iq[0, 2048 * 9 + 2044] = 0                # I2 of the tile's last group: this tile and the next fall back
    iq[n - 1, 1024 * 21 + 7] = 255
    exp = _oracle_rows(n, iq)
    two = gen in (3 | 16, 5)                  # two channels per worker warp: half-tiles, one count per warp
    outs = {}
    for g in (gen, 3):
        e = R.Engine(n, 0, nbytes)
        e.set_modes(np.full(n, 3, dtype=np.uint8))
        e.debug_set_wbfm_kernel(g)
        e.debug_wb_prefilter_counts()
        pcm, counts = e.demodulate(iq)
        assert (counts == nbytes // 64).all()
        outs[g] = pcm.copy()
        mma, simt = e.debug_wb_prefilter_counts()
        units = ((n + 1) // 2) * (nbytes // 1024) if two else n * (nbytes // 2048)
        if g == 3:
            assert (mma, simt) == (0, 0)     # the CUDA-core kernel does not count
        else:
            assert mma + simt == units
            assert simt == 3, (mma, simt)      # the tile with the Q1 byte; the one with the I2 byte and its successor
        e.close()
    for ch in range(n):
        assert np.array_equal(outs[gen][ch], exp[ch]), "channel %d" % ch
    assert np.array_equal(outs[gen], outs[3])


def test_tensor_core_prefilter_ragged_calls_and_switches():
    """Call lengths that leave partial (half-)tiles, every kernel generation and pre-filter build
    alternating between calls on the same streams: the raw history in shared memory and the planes in the
    carry blob are two views of the same 16 samples."""
    import rtlsdrdiags_b200 as R
    n = 7
    e = R.Engine(n, 0, 4 * 32768)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    sizes = [2048, 64, 64 * 15, 1024, 64 * 17, 32768, 64 * 33, 2 * 32768, 64 * 31, 1024 + 64, 4096, 64 * 47, 32768 + 64 * 5]
    gens = [5, 2 | 16, 3, 6, 2, 5, 3 | 16, 6, 5, 3, 4, 1, 5]
    iq = _quiet(n, sum(sizes), seed=31)
    iq[2] = S.noise(1, sum(sizes), seed=3)[0]          # full-scale bytes: clipping bytes all over
    iq[3] = S.tone(3, sum(sizes) // 2, seed=2)
    iq[4, ::4099] = 0
    out, off = [], 0
    e.debug_wb_prefilter_counts()
    for sz, g in zip(sizes, gens):
        e.debug_set_wbfm_kernel(g)
        e.accept_iq_host(np.ascontiguousarray(iq[:, off:off + sz]))
        out.append(e.get_pcm()[0])
        off += sz
    mma, simt = e.debug_wb_prefilter_counts()
    assert mma > 0 and simt > 0
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()


@pytest.mark.parametrize("gen,per_cta", [(5, 28), (6, 14), (5, 10), (6, 5)])
def test_generation4_full_ctas(gen, per_cta):
    """Generation 4 with as many channels per CTA as a large bank gets (28 = 14 worker warps x 2, or 14 x 1):
    every M-block, the hole at the recurrence warp's slot, the MMA warp as warp 15."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 61, 32768
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    e.set_launch_shape(R.KIND_WBFM, per_cta, 0)
    e.debug_set_wbfm_kernel(gen)
    iq = _quiet(n, 2 * nbytes, seed=77)
    iq[7, 5000] = 0
    iq[33] = S.noise(1, 2 * nbytes, seed=5)[0]
    e.debug_wb_prefilter_counts()
    pcm, counts = e.demodulate(iq)
    mma, simt = e.debug_wb_prefilter_counts()
    assert mma > simt > 0
    exp = _oracle_rows(n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()


def test_clipping_bank_moves_to_generation_3():
    """The default kernel of a large bank is generation 4; when more than a quarter of a launch's (half-)tiles
    held a clipping byte and fell back to the CUDA cores, the engine takes generation 3 for the next calls.
    The streams do not notice."""
    import rtlsdrdiags_b200 as R
    n, nbytes, calls = 2200, 4096, 3            # more than 14 x 148 channels: a 'large' bank
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    iq = _quiet(n, calls * nbytes, seed=9)
    iq[:, nbytes:] = S.noise(n, (calls - 1) * nbytes, seed=10)      # from the second call on: full-scale bytes
    e.debug_wb_prefilter_counts()
    seen, out = [], []
    for c in range(calls):
        e.accept_iq_host(np.ascontiguousarray(iq[:, c * nbytes:(c + 1) * nbytes]))
        out.append(e.get_pcm()[0])
        seen.append(e.debug_wb_prefilter_counts())
    units = (n // 2) * (nbytes // 1024)
    assert seen[0] == (units, 0), (seen, units)                     # clean input: the tensor cores
    # clipping input: generation 4 fell back (but for the odd half-tile pair without a byte 0 in a negated position) ...
    assert seen[1][0] + seen[1][1] == 2 * units and seen[1][1] > 0.9 * units, (seen, units)
    assert seen[2] == seen[1], (seen, units)                        # ... and the third call ran generation 3
    pcm = np.concatenate(out, axis=1)
    some = [0, 1, 2, 777, 1023, 1024, 2198, 2199]
    exp = _oracle_rows(len(some), iq[some])
    for i, ch in enumerate(some):
        assert np.array_equal(pcm[ch], exp[i]), "channel %d" % ch
    e.close()


@pytest.mark.parametrize("gen", [5, 6])
def test_generation4_signed_input_and_format_switches(gen):
    """Input that is already signed and rotated (IQ_S8_ROTATED) never takes the tensor cores: generation 4 runs
    its CUDA-core path for those calls (the engine's default would pick generation 3 for them), and a stream may
    change format between calls -- the history record is raw bytes of whichever format the call had."""
    import rtlsdrdiags_b200 as R
    n, nbytes, calls = 9, 8192, 4
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 3, dtype=np.uint8))
    e.debug_set_wbfm_kernel(gen)
    iq = _quiet(n, calls * nbytes, seed=44)
    out = []
    for c in range(calls):
        piece = np.ascontiguousarray(iq[:, c * nbytes:(c + 1) * nbytes])
        if c % 2 == 0:
            e.accept_iq_host(piece)
        else:
            e.accept_iq_host(np.stack([O.front_end(piece[ch]) for ch in range(n)]), R.IQ_S8_ROTATED)
        out.append(e.get_pcm()[0])
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(n, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()
