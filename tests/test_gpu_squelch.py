"""Squelch gate (IqDataProcessor -> Squelch -> SignalDetector / SignalTracker) on the GPU
against the oracle: thresholds, tuner gain, the one-block tail, frozen demodulator state
while squelched, signal reports without a closing threshold."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu


def _block(rng, n, amps, nbytes):
    """One block per channel: gaussian noise of the given per-channel amplitude around 128."""
    x = 128 + np.asarray(amps, dtype=np.float64)[:, None] * rng.standard_normal((n, nbytes))
    return np.clip(np.round(x), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("nbytes", [32768, 4096])
def test_squelch_matches_oracle(nbytes):
    import rtlsdrdiags_b200 as R
    n = 24
    rng = np.random.default_rng(11)
    modes = np.array([ch % 6 for ch in range(n)], dtype=np.uint8)
    thresholds = [-200, -30, -20, -15, -12, -9, -6, -3][:8] * 3
    gains = [0, 0, 5, 0, 3, 0, 0, 10] * 3
    e = R.Engine(n, 0, nbytes)
    e.set_modes(modes)
    chains = []
    for ch in range(n):
        c = O.OracleChain()
        c.set_mode(int(modes[ch]))
        c.set_threshold(thresholds[ch])
        c.set_rx_gain(gains[ch])
        chains.append(c)
        e.set_squelch_threshold(ch, thresholds[ch])
        e.set_receive_gain_db(ch, gains[ch])
    # amplitude schedule: quiet, loud, quiet (tail), quiet, medium ...
    schedule = [1.0, 60.0, 2.0, 1.0, 12.0, 30.0, 0.5, 0.5, 90.0, 4.0]
    for step, base in enumerate(schedule):
        amps = base * (0.5 + np.arange(n) / n)
        iq = _block(rng, n, amps, nbytes)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        allowed, mag = e.get_signal()
        for ch in range(n):
            exp = chains[ch].accept_u8(iq[ch])
            ea, em = chains[ch].signal()
            assert bool(allowed[ch]) == ea and int(mag[ch]) == em, "step %d channel %d" % (step, ch)
            assert counts[ch] == exp.size, "step %d channel %d: count %d vs %d" % (step, ch, counts[ch], exp.size)
            if exp.size:
                assert np.array_equal(pcm[ch], exp), "step %d channel %d" % (step, ch)


def test_arming_after_open_blocks_and_reports_only():
    """Blocks processed before any threshold is set leave the trackers in `Tracking`;
    signal reports work while nothing can close."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 6, 8192
    rng = np.random.default_rng(5)
    e = R.Engine(n, 0, nbytes)
    e.set_modes(np.full(n, 1, dtype=np.uint8))
    chains = [O.OracleChain() for _ in range(n)]
    for c in chains:
        c.set_mode(1)
    for step in range(3):
        iq = _block(rng, n, [20.0] * n, nbytes)
        if step == 1:
            e.enable_signal_reports(True)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            assert np.array_equal(pcm[ch], chains[ch].accept_u8(iq[ch]))
        if step >= 1:
            allowed, mag = e.get_signal()
            for ch in range(n):
                assert (bool(allowed[ch]), int(mag[ch])) == chains[ch].signal()
    # now close the squelch on quiet input: first quiet block is the tail, then silence
    for ch in range(n):
        e.set_squelch_threshold(ch, -10)
        chains[ch].set_threshold(-10)
    for step in range(4):
        iq = _block(rng, n, [1.0] * n, nbytes)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            exp = chains[ch].accept_u8(iq[ch])
            assert counts[ch] == exp.size
            if exp.size:
                assert np.array_equal(pcm[ch], exp)
        assert (counts == 0).all() == (step >= 1)


def test_squelch_matches_golden_vectors():
    """All seven golden configurations as seven channels of one engine fed the same blocks:
    decisions, magnitudes, counts and PCM as the compiled reference produced them."""
    import os
    import rtlsdrdiags_b200 as R
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_squelch_v1.npz"))
    n = len(g["cfg"])
    e = R.Engine(n, 0, 4096)
    e.set_modes(np.full(n, 2, dtype=np.uint8))
    for i, (thr, gain) in enumerate(g["cfg"]):
        e.set_squelch_threshold(i, int(thr))
        e.set_receive_gain_db(i, int(gain))
    out = [[] for _ in range(n)]
    for b, blk in enumerate(g["blocks"]):
        e.accept_iq_host(np.ascontiguousarray(np.broadcast_to(blk, (n, blk.size))))
        pcm, counts = e.get_pcm()
        allowed, mag = e.get_signal()
        assert np.array_equal(allowed.astype(np.uint8), g["allowed"][:, b])
        assert np.array_equal(mag, g["magnitude"][:, b])
        assert np.array_equal(counts, g["counts"][:, b])
        for i in range(n):
            out[i].append(pcm[i][:counts[i]])
    for i in range(n):
        assert np.array_equal(np.concatenate(out[i]), g["pcm_%d" % i])


def test_rearming_after_blocks_accepted_while_disarmed():
    """SignalTracker runs on every block in the reference (SignalTracker.cc:104-145), also while no
    threshold can close. Arm and squelch a channel (tracker in NoSignal), disarm, accept a block
    (the reference's tracker moves to Tracking), re-arm and feed noise: the first quiet block is the
    one-block tail and must pass."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 8, 4096
    rng = np.random.default_rng(17)
    e = R.Engine(n, 0, nbytes)
    modes = np.array([1 + ch % 5 for ch in range(n)], dtype=np.uint8)
    e.set_modes(modes)
    chains = [O.OracleChain() for _ in range(n)]
    for ch, c in enumerate(chains):
        c.set_mode(int(modes[ch]))

    def set_threshold(t):
        for ch in range(n):
            e.set_squelch_threshold(ch, t)
            chains[ch].set_threshold(t)

    def step(amp, what):
        iq = _block(rng, n, [amp] * n, nbytes)
        e.accept_iq_host(iq)
        pcm, counts = e.get_pcm()
        for ch in range(n):
            exp = chains[ch].accept_u8(iq[ch])
            assert counts[ch] == exp.size, "%s: channel %d count %d vs %d" % (what, ch, counts[ch], exp.size)
            if exp.size:
                assert np.array_equal(pcm[ch], exp), "%s: channel %d" % (what, ch)
        return counts

    set_threshold(-10)
    step(40.0, "armed, loud")
    step(1.0, "armed, tail")
    assert (step(1.0, "armed, squelched") == 0).all()
    set_threshold(-200)                       # disarmed: no squelch kernel runs
    assert (step(1.0, "disarmed") == 512 * nbytes // 32768).all()
    set_threshold(-10)                        # re-armed: quiet block = end-of-signal tail, allowed
    assert (step(1.0, "re-armed, tail") == 512 * nbytes // 32768).all()
    assert (step(1.0, "re-armed, squelched") == 0).all()
