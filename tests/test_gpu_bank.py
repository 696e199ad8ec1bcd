"""The multi-device bank (sdr_bank_*): contiguous channel shards, one pinned tick array
and one pinned PCM array shared by all devices. With one GPU the bank has one shard; with more
(the scaling box) the same test spreads the channels over all of them."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return list(range(torch.cuda.device_count()))


def _oracle_rows(modes, iq):
    rows = []
    for ch, m in enumerate(modes):
        c = O.OracleChain()
        c.set_mode(int(m))
        rows.append(c.accept_u8(iq[ch]))
    return rows


@pytest.mark.parametrize("n", [7, 64, 203])
def test_bank_matches_oracle_over_all_devices(n):
    import rtlsdrdiags_b200 as R
    devs = _devices()[: max(1, min(len(_devices()), n))]
    block = 32768
    b = R.Bank(n, devs, block, n_slots=3)
    shards = b.shards()
    assert [s[0] for s in shards] == devs
    assert shards[0][1] == 0 and sum(s[2] for s in shards) == n
    for i in range(1, len(shards)):
        assert shards[i][1] == shards[i - 1][1] + shards[i - 1][2]
    modes = np.array([ch % 6 for ch in range(n)], dtype=np.uint8)
    b.set_modes(modes)
    b.set_gain(n - 1, R.MODE_TO_KIND.get(int(modes[n - 1]), 1), 123.0)
    ticks = 5
    iq = S.noise(n, ticks * block, seed=n)
    chains = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(int(modes[ch]))
        if modes[n - 1] and ch == n - 1:
            c.set_gain(O.MODE_TO_KIND[int(modes[ch])], 123.0)
        chains.append(c)
    got = []
    for k in range(ticks):
        b.acquire()[:] = iq[:, k * block:(k + 1) * block]
        b.commit(1000 + k)
        if k >= 2:
            got.append(b.retire())
    got.append(b.retire())
    got.append(b.retire())
    with pytest.raises(R.SdrError):
        b.retire()
    for k, (ts, pcm, counts) in enumerate(got):
        assert ts == 1000 + k
        for ch in range(n):
            exp = chains[ch].accept_u8(iq[ch, k * block:(k + 1) * block])
            assert counts[ch] == exp.size
            if exp.size:
                assert np.array_equal(pcm[ch], exp), "tick %d channel %d" % (k, ch)
    b.close()


def test_bank_short_tick_and_errors():
    import rtlsdrdiags_b200 as R
    devs = _devices()
    n, block = 12, 8192
    b = R.Bank(n, devs[: min(len(devs), n)], block, n_slots=2)
    modes = np.full(n, 1, dtype=np.uint8)
    b.set_modes(modes)
    iq = S.noise(n, block, seed=1)
    b.acquire()[:] = iq
    b.commit(7, bytes_per_channel=4096)          # a short block: rows of 64 samples
    ts, pcm, counts = b.retire()
    assert ts == 7 and pcm.shape == (n, 64) and (counts == 64).all()
    exp = _oracle_rows(modes, iq[:, :4096])
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch])
    with pytest.raises(R.SdrError):
        b.set_mode(n, 1)
    with pytest.raises(R.SdrError):
        R.Bank(4, [0, 0], block)                  # a device listed twice
    with pytest.raises(R.SdrError):
        R.Bank(4, [99], block)
    b.close()
