"""BASELINE.json's full bank sizes through a size-independent property: a bank built by repeating a
few dozen distinct radios (input and mode) must give every copy the PCM of its original, bit for
bit, and the originals must equal the oracle -- so every one of the 1024 / 8192 / 16384 / 65536
channels is checked against the reference's arithmetic at the cost of a few dozen oracle runs.
Also at full size: cutting the stream into calls differently does not change a byte."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu

BLOCK = 32768


def _bank(modes_base, n_channels, n_blocks, seed):
    """(engine, device IQ [n_channels][bytes], base IQ, base modes): channel c is a copy of base c % B."""
    import torch
    import rtlsdrdiags_b200 as R
    B, nbytes = len(modes_base), n_blocks * BLOCK
    half = B // 2
    base = np.concatenate([S.noise(half, nbytes, seed=seed), S.tone_bank(modes_base[half:], nbytes, seed=seed + 1)])
    base[0, : nbytes // 4] = 0      # the quirk runs the reference's tests cannot reach: -(-128), full-scale wrap
    base[1, : nbytes // 4] = 255
    idx = torch.arange(n_channels, device="cuda") % B
    iq = torch.from_numpy(base).cuda()[idx].contiguous()
    e = R.Engine(n_channels, 0, nbytes)
    e.set_modes(np.asarray(modes_base, dtype=np.uint8)[np.arange(n_channels) % B])
    return e, iq, base


def _oracle(modes_base, base):
    rows = []
    for m, x in zip(modes_base, base):
        c = O.OracleChain()
        c.set_mode(int(m))
        rows.append(c.accept_u8(x))
    return np.stack(rows)


def _check(e, iq, modes_base, base):
    B, n = len(modes_base), e.n
    e.accept_iq_device(iq)
    pcm, counts = e.get_pcm()
    assert (counts == pcm.shape[1]).all()
    assert n % B == 0
    copies = pcm.reshape(n // B, B, pcm.shape[1])
    assert (copies == copies[0][None]).all(), "a copy differs from its original"
    assert np.array_equal(copies[0], _oracle(modes_base, base)), "the originals differ from the oracle"
    return pcm


def test_am_1024_channels_16_blocks():           # BASELINE configs[1], the bench shape
    modes = [1] * 64
    e, iq, base = _bank(modes, 1024, 16, seed=11)
    _check(e, iq, modes, base)
    e.close()


def test_wbfm_8192_channels():                   # BASELINE configs[2]
    modes = [3] * 64
    e, iq, base = _bank(modes, 8192, 2, seed=12)
    pcm = _check(e, iq, modes, base)
    # the same two blocks as two calls on a fresh engine: same bytes
    import rtlsdrdiags_b200 as R
    e2 = R.Engine(8192, 0, BLOCK)
    e2.set_modes(np.full(8192, 3, dtype=np.uint8))
    halves = []
    for b in range(2):
        e2.accept_iq_device(iq[:, b * BLOCK:(b + 1) * BLOCK])
        halves.append(e2.get_pcm()[0])
    assert np.array_equal(np.concatenate(halves, axis=1), pcm)
    e.close()
    e2.close()


def test_ssb_16384_channels():                   # BASELINE configs[3], both sidebands, on one GPU
    modes = [4, 5] * 32
    e, iq, base = _bank(modes, 16384, 1, seed=13)
    _check(e, iq, modes, base)
    e.close()


def test_mixed_65536_channels():                 # BASELINE configs[4]: the whole 8-GPU bank on one GPU
    modes = [1 + c % 5 for c in range(64)]      # 64 originals over the five modes; copies keep the original's mode
    e, iq, base = _bank(modes, 65536, 1, seed=14)
    _check(e, iq, modes, base)
    e.close()
