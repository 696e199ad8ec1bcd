"""The CUDA pipelines, compiled for the host (tests/emu, -DSDR_EMU: phases run serially,
a phase boundary stands in for __syncthreads), against the oracle. This checks the
kernels' indexing, history moves, state save/restore and clamp-path selection on the
CPU-only build box; the -m gpu tests repeat it on the real device through the C ABI."""
import numpy as np
import pytest

import _emu as E
import _oracle as O
import _signals as S

KIND_MODES = {E.KIND_AM: [1], E.KIND_FM: [2], E.KIND_WBFM: [3], E.KIND_SSB: [4, 5]}
DEF_GAIN = {E.KIND_AM: 300.0, E.KIND_FM: np.float32(64000 / (2 * np.pi)),
            E.KIND_WBFM: np.float32(256000 / (2 * np.pi)), E.KIND_SSB: 300.0}


def run_case(kind, n_ch, G, NT, gmul, sizes, seed, fmt=E.FMT_U8):
    rng = np.random.default_rng(seed)
    bank = E.EmuBank(kind, n_ch, G=G, NT=NT)
    modes = [KIND_MODES[kind][c % len(KIND_MODES[kind])] for c in range(n_ch)]
    chains = []
    for c in range(n_ch):
        ch = O.OracleChain()
        ch.set_mode(modes[c])
        g = np.float32(DEF_GAIN[kind] * gmul)
        ch.set_gain(O.MODE_TO_KIND[modes[c]], float(g))
        chains.append(ch)
        bank.scale[c] = E.scale_for(kind, g)
        bank.lsb[c] = 1 if modes[c] == 4 else 0
    start = 0
    for nbytes in sizes:
        if fmt == E.FMT_U8:
            iq = rng.integers(0, 256, size=(n_ch, nbytes), dtype=np.uint8)
            iq[0] = 0
            if n_ch > 1:
                iq[1] = 255
            if n_ch > 2:
                iq[2] = S.tone(modes[2], nbytes // 2, seed=seed, start=start)
            got = bank.run(iq, fmt)
            exp = [chains[c].accept_u8(iq[c]) for c in range(n_ch)]
        else:
            iq = rng.integers(-128, 128, size=(n_ch, nbytes), dtype=np.int8)
            got = bank.run(iq, fmt)
            exp = [chains[c].accept_s8(modes[c], iq[c]) for c in range(n_ch)]
        for c in range(n_ch):
            assert np.array_equal(got[c], exp[c]), "kind %d channel %d block of %d bytes" % (kind, c, nbytes)
        start += nbytes // 2


@pytest.mark.parametrize("kind", [E.KIND_AM, E.KIND_SSB, E.KIND_FM, E.KIND_WBFM])
@pytest.mark.parametrize("shape", [(32, 1024, 33), (1, 32, 2), (7, 96, 9)])
@pytest.mark.parametrize("gmul", [1.0, 3.7, 1e7])
def test_emulated_kernel_matches_oracle(kind, shape, gmul):
    G, NT, n_ch = shape
    run_case(kind, n_ch, G, NT, gmul, [2048 * 2 * 2 + 64, 8192, 64, 32768], seed=11 * kind + G)


@pytest.mark.parametrize("kind", [E.KIND_AM, E.KIND_SSB, E.KIND_FM, E.KIND_WBFM])
def test_emulated_kernel_signed_rotated_entry(kind):
    run_case(kind, 3, 2, 64, 1.0, [32768 + 128, 4096], seed=5, fmt=E.FMT_S8)


@pytest.mark.parametrize("kind", [E.KIND_FM, E.KIND_WBFM])
def test_clamp_path_switches_mid_stream(kind):
    """Quiet input (fast dot-product path) followed by full-scale noise (per-tap clamp
    path) and back: the sticky flag and the history scan must keep parity."""
    n_ch = 3
    bank = E.EmuBank(kind, n_ch, G=3, NT=64)
    mode = KIND_MODES[kind][0]
    chains = [O.OracleChain() for _ in range(n_ch)]
    for c in range(n_ch):
        chains[c].set_mode(mode)
        g = np.float32(DEF_GAIN[kind] * 8)
        chains[c].set_gain(O.MODE_TO_KIND[mode], float(g))
        bank.scale[c] = E.scale_for(kind, g)
    rng = np.random.default_rng(2)
    for quiet in [True, False, True, True, False]:
        if quiet:
            iq = np.full((n_ch, 32768), 128, dtype=np.uint8) + rng.integers(0, 2, size=(n_ch, 32768), dtype=np.uint8)
        else:
            iq = rng.integers(0, 256, size=(n_ch, 32768), dtype=np.uint8)
        got = bank.run(iq, E.FMT_U8)
        for c in range(n_ch):
            assert np.array_equal(got[c], chains[c].accept_u8(iq[c]))


def test_table_wrap_without_fp64_equals_the_double_wrap_for_every_table_pair():
    """wbfm_tile2_kernel wraps theta differences with two float subtractions (wrap_pi_table)
    instead of the reference's double arithmetic (WbFmDemodulator.cc:472-480). Exhaustive over
    all pairs of the atan2 table's distinct values: 39,920^2 = 1.59e9 differences."""
    import ctypes as C
    L = E.lib()
    L.emu_wrap_table_check.restype = C.c_uint64
    L.emu_wrap_table_check.argtypes = [C.c_void_p, C.c_uint32]
    # ... and of the NBFM kernel's 280 x 280 table (fm_tile_kernel's discriminator): 47,808^2 = 2.29e9
    for kind, distinct in ((E.KIND_WBFM, 39920), (E.KIND_FM, 47808)):
        vals = np.unique(E.lut(kind).ravel())
        assert vals.size == distinct
        assert L.emu_wrap_table_check(vals.ctypes.data_as(C.c_void_p), vals.size) == 0


def test_wbfm_table_is_odd_in_q():
    """The second WBFM kernel keeps the half plane q >= 0 only: theta(-q, i) == -theta(q, i) bit for
    bit for every entry of the reference's table (WbFmDemodulator.cc:159-170)."""
    t = E.lut(E.KIND_WBFM)            # [q + 128][i + 128]
    for q in range(1, 128):
        assert np.array_equal(t[128 - q], -t[128 + q])
        assert np.all(np.signbit(t[128 - q]) != np.signbit(t[128 + q]))
    assert np.array_equal(t[0], -np.array([O.oracle().sdro_atan2f(128, i) for i in range(-128, 128)], dtype=np.float32))


def test_grouped_front_end_is_the_reference_front_end():
    """front_end_ab (AM / SSB stage 1) groups a rotation period's eight bytes by what has to be done to them
    -- the four the Fs/4 rotation negates in one word -- instead of by arm. Against the oracle's front end
    (IqDataProcessor.cc:735-738, 567-611) for every byte value in every position, -(-128) = -128 included,
    and raw_from_ab as its inverse; for input that is already signed and rotated it is a plain regrouping."""
    import ctypes as C
    L = E.lib()
    L.emu_front_end_ab.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11)
    out = np.zeros(8, dtype=np.int8)
    cases = [np.full(8, v, dtype=np.uint8) for v in range(256)]
    for pos in range(8):
        for v in (0, 1, 127, 128, 129, 255):
            c = rng.integers(0, 256, size=8, dtype=np.uint8)
            c[pos] = v
            cases.append(c)
    cases += [rng.integers(0, 256, size=8, dtype=np.uint8) for _ in range(2000)]
    for raw in cases:
        raw = np.ascontiguousarray(raw)
        assert L.emu_front_end_ab(E.FMT_U8, raw.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        want = O.front_end(raw)                      # I'0 Q'0 I'1 Q'1 ... as the reference leaves them
        assert np.array_equal(out[:4], want[0::2]) and np.array_equal(out[4:], want[1::2]), raw
        s8 = raw.view(np.int8)
        assert L.emu_front_end_ab(E.FMT_S8, raw.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out[:4], s8[0::2]) and np.array_equal(out[4:], s8[1::2]), raw
