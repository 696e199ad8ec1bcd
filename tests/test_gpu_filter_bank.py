"""GPU parity of the batched filter banks through the C ABI: bit-identical to
the oracle's Decimator / Interpolator / Decimator_int16 / Interpolator_int16 and to the golden
vectors the compiled reference produced, for every kind, however the stream is cut."""
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_multirate_v1.npz"))
CASES = [m.split(",") for m in GOLD["meta"]]
f32 = np.float32


def _bank(kind, rows, taps, factor):
    import rtlsdrdiags_b200 as R
    return R.FilterBank(kind, rows, taps, factor)


@pytest.mark.parametrize("name,kind,factor", CASES)
def test_golden_vectors(name, kind, factor):
    x, y = GOLD["x_" + name], GOLD["y_" + name]
    b = _bank(int(kind), 3, GOLD["taps_" + name], int(factor))
    got = b.run(np.stack([x, x[::-1], x]))
    assert got[0].tobytes() == y.tobytes() and got[2].tobytes() == y.tobytes()
    exp1 = O.Multirate(int(kind), GOLD["taps_" + name], int(factor)).run(x[::-1])
    assert got[1].tobytes() == exp1.tobytes()
    assert b.launch_count == 1


@pytest.mark.parametrize("name,kind,factor", CASES)
def test_stream_cut_into_ragged_calls(name, kind, factor):
    x, y = GOLD["x_" + name], GOLD["y_" + name]
    b = _bank(int(kind), 2, GOLD["taps_" + name], int(factor))
    cuts = [0, 1, 2, 3, 10, 11, 77, 300, 301, x.size]
    parts = [b.run(np.stack([x[a:c], x[a:c]])) for a, c in zip(cuts[:-1], cuts[1:])]
    got = np.concatenate(parts, axis=1)
    assert got[0].tobytes() == y.tobytes() and got[1].tobytes() == y.tobytes()
    # resetFilterState: the same input again gives the same output
    b.reset()
    assert b.run(np.stack([x, x]))[1].tobytes() == y.tobytes()


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_random_filters_many_rows(kind):
    rng = np.random.default_rng(40 + kind)
    for trial in range(6):
        factor = int(rng.integers(1, 9))
        q = int(rng.integers(1, 40))
        n_taps = q * factor if kind in (2, 4) else int(rng.integers(1, 200))
        taps = (rng.normal(0, 0.4, n_taps) * rng.choice([0.1, 1.0, 2.5])).astype(f32)
        rows, n = int(rng.integers(1, 40)), int(rng.integers(1, 9000))
        x = rng.integers(-32768, 32768, (rows, n)).astype(np.int16 if kind >= 3 else f32)
        b = _bank(kind, rows, taps, factor)
        half = n // 2
        got = np.concatenate([b.run(x[:, :half]), b.run(x[:, half:])], axis=1) if half else b.run(x)
        for r in range(rows):
            exp = O.Multirate(kind, taps, factor).run(x[r])
            assert got[r].tobytes() == exp.tobytes(), (kind, trial, r)


def test_q15_tap_quantisation_matches_the_oracle():
    taps = np.array([1.0, -1.0, 0.99998, 0.5, -0.5, 1.5e-5, 4.6e-5, -4.6e-5, 0.25000763, 0.123456], dtype=f32)
    b = _bank(3, 1, taps, 1)
    # an impulse of 2 through the Q15 FIR returns (q*2 + 16384) >> 15 ... use the oracle's taps directly
    x = np.zeros(40, dtype=np.int16)
    x[0] = 16384
    exp = O.Multirate(3, taps, 1).run(x)
    assert b.run(x[None, :])[0].tobytes() == exp.tobytes()
    q = b.taps_q15()
    assert q[0] == -32768 and q[1] == -32768 and q[3] == 16384 and q[5] == 0 and q[6] == 2 and q[7] == -2


def test_device_resident_large_bank_is_cut_invariant():
    """Full-size property: 4096 rows x 64 Ki samples through the 80-tap 4:1 float decimator in one
    call and in three ragged calls give the same bytes; a sample of rows equals the oracle."""
    import torch
    import rtlsdrdiags_b200 as R
    name = "dec_f32_80x4"
    taps = GOLD["taps_" + name]
    rows, n = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.rand((rows, n), device="cuda", generator=g) * 20000 - 10000).round().contiguous()
    b = R.FilterBank(R.FILTER_DECIMATOR_F32, rows, taps, 4)
    y = torch.zeros((rows, n // 4), device="cuda")
    assert b.run_device(x.data_ptr(), n, n, y.data_ptr(), n // 4) == n // 4
    b.sync()
    b2 = R.FilterBank(R.FILTER_DECIMATOR_F32, rows, taps, 4)
    y2 = torch.zeros((rows, n // 4 + 8), device="cuda")
    done = 0
    for a, c in ((0, 1001), (1001, 30002), (30002, n)):
        done += b2.run_device(x.data_ptr() + 4 * a, n, c - a, y2.data_ptr() + 4 * done, n // 4 + 8)
    b2.sync()
    assert done == n // 4
    assert torch.equal(y, y2[:, : n // 4])
    xs, ys = x[::512].cpu().numpy(), y[::512].cpu().numpy()
    for r in range(xs.shape[0]):
        assert ys[r].tobytes() == O.Multirate(1, taps, 4).run(xs[r]).tobytes()


def test_argument_checks():
    import ctypes as C
    import rtlsdrdiags_b200 as R
    L = R.load_library()
    h = C.c_void_p()
    t = (C.c_float * 8)(*([0.1] * 8))
    assert L.sdr_filter_bank_create(0, 9, 1, t, 8, 2, C.byref(h)) == -1
    assert L.sdr_filter_bank_create(0, 2, 1, t, 7, 2, C.byref(h)) == -1   # interpolator: N % L != 0
    assert L.sdr_filter_bank_create(0, 1, 0, t, 8, 2, C.byref(h)) == -1
    assert L.sdr_filter_bank_run(None, None, 0, 0, None, 0, None, 0) == -1
    b = R.FilterBank(R.FILTER_DECIMATOR_F32, 2, [0.5, 0.5], 4)
    assert b.out_count(3) == 0 and b.run(np.ones((2, 3), dtype=f32)).shape == (2, 0)
    assert b.out_count(1) == 1
    got = b.run(np.ones((2, 1), dtype=f32))
    assert got.shape == (2, 1) and got[0, 0] == 1.0


@pytest.mark.parametrize("factor", [2, 4])
def test_float_decimator_block_path_edges(factor):
    """Float decimators with M = 2, 4 take the path with 16-byte window loads: tap counts around the
    block size (4 M), with and without taps above the highest full block, short and ragged streams."""
    rng = np.random.default_rng(900 + factor)
    for n_taps in (1, 3, 4, 5, 4 * factor - 1, 4 * factor, 4 * factor + 1, 15, 16, 17, 37, 80, 81, 83, 127):
        taps = rng.normal(0, 0.3, n_taps).astype(f32)
        rows, n = 5, int(rng.integers(1, 6000))
        x = (rng.normal(0, 3000, (rows, n))).astype(f32)
        b = _bank(1, rows, taps, factor)
        cut = int(rng.integers(0, n))
        got = np.concatenate([b.run(x[:, :cut]), b.run(x[:, cut:])], axis=1) if cut else b.run(x)
        for r in range(rows):
            exp = O.Multirate(1, taps, factor).run(x[r])
            assert got[r].tobytes() == exp.tobytes(), (factor, n_taps, r)


@pytest.mark.parametrize("factor", [2, 4])
def test_q15_decimator_block_path_edges(factor):
    """Q15 decimators with M = 2, 4 take the blocked path in tiles whose samples cannot reach the clamp
    (moderate amplitude here), the per-tap clamp path otherwise (one full-scale row per case)."""
    rng = np.random.default_rng(950 + factor)
    for n_taps in (1, 3, 4, 5, 4 * factor - 1, 4 * factor, 4 * factor + 1, 16, 17, 37, 80, 83):
        taps = (rng.normal(0, 0.3, n_taps) / max(1.0, np.sqrt(n_taps) / 3)).astype(f32)
        rows, n = 5, int(rng.integers(1, 7000))
        x = rng.integers(-3000, 3000, (rows, n)).astype(np.int16)
        x[4] = rng.integers(-32768, 32768, n).astype(np.int16)   # this row's tiles keep the clamp
        b = _bank(3, rows, taps, factor)
        cut = int(rng.integers(0, n))
        got = np.concatenate([b.run(x[:, :cut]), b.run(x[:, cut:])], axis=1) if cut else b.run(x)
        for r in range(rows):
            exp = O.Multirate(3, taps, factor).run(x[r])
            assert got[r].tobytes() == exp.tobytes(), (factor, n_taps, r)


def test_create_refuses_what_no_tile_can_serve():
    """A bank whose taps and factor need more shared memory than any tile leaves is refused when it
    is created, not at its first run (create and run share one launch plan)."""
    import rtlsdrdiags_b200 as R
    for kind, n_taps, factor in [(R.FILTER_DECIMATOR_F32, 8000, 1), (R.FILTER_DECIMATOR_F32, 300, 300),
                                 (R.FILTER_DECIMATOR_I16, 10000, 2)]:
        with pytest.raises(R.SdrError):
            R.FilterBank(kind, 4, np.ones(n_taps, dtype=np.float32) / n_taps, factor)
    # the largest filters the reference itself ships still fit
    b = R.FilterBank(R.FILTER_DECIMATOR_F32, 4, np.ones(80, dtype=np.float32) / 80, 4)
    assert b.run(np.zeros((4, 64), dtype=np.float32)).shape == (4, 16)
    b.close()
