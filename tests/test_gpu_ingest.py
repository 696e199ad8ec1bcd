"""Ingest ring: the bank's DataConsumer. Ticks pipelined through pinned slots
give the PCM the oracle gives block by block; timestamps, short-block accounting and clipping
follow DataConsumer::acceptData (DataConsumer.cc:232-246)."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def _chains(modes):
    out = []
    for m in modes:
        c = O.OracleChain()
        c.set_mode(int(m))
        out.append(c)
    return out


@pytest.mark.parametrize("n_slots,lag", [(2, 1), (3, 2), (4, 0)])
def test_ring_matches_oracle_tick_by_tick(n_slots, lag):
    import rtlsdrdiags_b200 as R
    n, block = 12, 8192
    modes = np.array([ch % 6 for ch in range(n)], dtype=np.uint8)
    e = R.Engine(n, 0, block)
    e.set_modes(modes)
    q = R.Ingest(e, n_slots, block)
    chains = _chains(modes)
    ticks = [S.noise(n, block, seed=100 + t) for t in range(9)]
    sizes = [block, block, 4096, block, 64, block, block, block - 64, block]
    exp = []
    for t, iq in enumerate(ticks):
        exp.append([chains[ch].accept_u8(iq[ch, :sizes[t]]) for ch in range(n)])
    retired = 0

    def check():
        nonlocal retired
        ts, pcm, counts = q.retire()
        assert ts == 1000 + retired
        for ch in range(n):
            want = exp[retired][ch]
            assert counts[ch] == want.size, (retired, ch)
            assert np.array_equal(pcm[ch][:counts[ch]], want), (retired, ch)
        retired += 1

    for t, iq in enumerate(ticks):
        if t % 2:
            q.accept(1000 + t, iq[:, :sizes[t]])
        else:                         # zero-copy: fill the pinned slot in place
            slot = q.acquire()
            slot[:, :sizes[t]] = iq[:, :sizes[t]]
            q.commit(1000 + t, sizes[t])
        if t >= lag:
            check()
    while retired < len(ticks):
        check()
    st = q.stats()
    assert st == {"last_timestamp": 1000 + len(ticks) - 1, "short_blocks": 3, "ticks": 9, "in_flight": 0}
    with pytest.raises(R.SdrError):
        q.retire()


def test_ring_full_clip_and_squelch():
    import rtlsdrdiags_b200 as R
    n, block = 4, 4096
    e = R.Engine(n, 0, 32768)
    e.set_modes(np.array([2, 2, 1, 3], dtype=np.uint8))
    e.set_squelch_threshold(1, 0)      # never opens on noise of this level
    q = R.Ingest(e, 2, block)
    chains = _chains([2, 2, 1, 3])
    chains[1].set_threshold(0)
    big = S.noise(n, block + 1024, seed=1)   # longer than the slot: clipped like the reference
    q.accept(7, big)
    q.accept(8, big)
    with pytest.raises(R.SdrError):          # both slots in flight
        q.accept(9, big)
    for want_ts in (7, 8):
        ts, pcm, counts = q.retire()
        assert ts == want_ts
        for ch in range(n):
            want = chains[ch].accept_u8(big[ch, :block])
            assert counts[ch] == want.size
            assert np.array_equal(pcm[ch][:counts[ch]], want)
        assert counts[1] == 0
    assert q.stats()["short_blocks"] == 0
    # signed, rotated ticks (.iq format) through the same ring
    s8 = np.stack([O.front_end(big[ch, :block]) for ch in range(n)])
    q.accept(10, s8, R.IQ_S8_ROTATED)
    ts, pcm, counts = q.retire()
    for ch in (0, 2, 3):
        assert np.array_equal(pcm[ch][:counts[ch]], chains[ch].accept_u8(big[ch, :block]))
