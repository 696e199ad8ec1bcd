"""The AM / SSB FIR kernel fetches full tiles with TMA (cp.async.bulk.tensor through a tensor map
of the caller's IQ array, 128-byte hardware swizzle) and a partial last tile with cp.async, and
with the TMA slot layout can run stage 1 (the 8-tap 4:1 decimators, AmDemodulator.cc:349-374) on the
tensor cores. Loader codes (sdr_debug_set_tile_loader): 0 = cp.async, 2..4 = TMA with that many
slot buffers, +8 = stage 1 on the tensor cores, +16 = the 80-register build. All against the
oracle: strided and offset inputs, ragged lengths, device-resident input, both input formats, and
clipping bytes (raw 0x00 where the Fs/4 rotation negates: -(-128) = -128, which the GEMM cannot
represent) placed where the fallback logic has its edges."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def _oracle_rows(modes, iq):
    rows = []
    for ch, m in enumerate(modes):
        c = O.OracleChain()
        c.set_mode(int(m))
        rows.append(c.accept_u8(iq[ch]))
    return rows


@pytest.mark.parametrize("tma", [0, 2, 3, 4, 10, 12, 16, 18])
@pytest.mark.parametrize("nbytes", [2048, 64 * 33, 4096 + 64, 32768, 32768 * 3 + 64 * 5, 64 * 31])
def test_loaders_match_oracle(tma, nbytes):
    import rtlsdrdiags_b200 as R
    n = 13
    modes = np.array([(1, 4, 5)[ch % 3] for ch in range(n)], dtype=np.uint8)
    e = R.Engine(n, 0, 4 * 32768)
    e.set_modes(modes)
    e.debug_set_tile_loader(tma)
    iq = S.noise(n, 2 * nbytes, seed=nbytes)
    out = []
    for k in range(2):
        e.accept_iq_host(np.ascontiguousarray(iq[:, k * nbytes:(k + 1) * nbytes]))
        out.append(e.get_pcm()[0])
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d mode %d" % (ch, modes[ch])
    e.close()


@pytest.mark.parametrize("tma", [0, 2, 4, 12])
def test_device_input_with_stride_and_offset(tma):
    """The tensor map is built from the caller's pointer and channel stride: a view into a larger
    device array (stride > bytes, base offset by 16 bytes), changed between calls."""
    import torch
    import rtlsdrdiags_b200 as R
    n, nbytes = 9, 32768 + 2048 + 64 * 3
    modes = np.array([(1, 5, 4)[ch % 3] for ch in range(n)], dtype=np.uint8)
    e = R.Engine(n, 0, 65536)
    e.set_modes(modes)
    e.debug_set_tile_loader(tma)
    iq = S.noise(n, 3 * nbytes, seed=5)
    big = torch.zeros((n, 3 * nbytes + 4096), dtype=torch.uint8, device="cuda")
    out = []
    for k, off in enumerate((16, 48, 2048)):
        view = big[:, off:off + nbytes]
        view.copy_(torch.from_numpy(iq[:, k * nbytes:(k + 1) * nbytes]).cuda())
        torch.cuda.synchronize()
        e.accept_iq_device(view)
        out.append(e.get_pcm()[0])
    pcm = np.concatenate(out, axis=1)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d" % ch
    e.close()


def _carrier(n, nbytes, seed):
    """Noisy carriers that never reach the rails: every tile takes the tensor-core path."""
    rng = np.random.default_rng(seed)
    t = np.arange(nbytes // 2)
    iq = np.empty((n, nbytes), dtype=np.uint8)
    for ch in range(n):
        ph = 2 * np.pi * (0.003 * (ch + 1)) * t + 3.0 * np.sin(2 * np.pi * t / (200.0 + ch))
        amp = 70.0 * (1 + 0.4 * np.sin(2 * np.pi * t / 3000.0))
        iq[ch, 0::2] = np.clip(np.round(128 + amp * np.cos(ph) + rng.normal(0, 3, t.size)), 1, 255)
        iq[ch, 1::2] = np.clip(np.round(128 + amp * np.sin(ph) + rng.normal(0, 3, t.size)), 1, 255)
    return iq


@pytest.mark.parametrize("tma", [10, 12])
def test_tensor_core_stage1_and_its_clipping_fallback(tma):
    """Clean carriers (tensor-core path throughout), then the same with single 0x00 bytes at the
    edges: the last and first rotation groups of a tile (the next tile's history), the first and
    last tile of a call, negated and non-negated byte positions."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 12, 2 * 32768
    modes = np.array([(1, 4, 5)[ch % 3] for ch in range(n)], dtype=np.uint8)
    iq = _carrier(n, 3 * nbytes, seed=tma)
    # channel 0-2 stay clean; the others get zeros at chosen byte offsets of the stream
    spots = {3: [2048 - 5], 4: [2048 - 4, 2048 + 3], 5: [3], 6: [nbytes - 2], 7: [nbytes - 5, nbytes + 4],
             8: [2048 * 7 + 1000], 9: [0, 1, 2, 3, 4, 5, 6, 7], 10: [2048 * 5 - 1, 2048 * 5 - 2, 2048 * 5 - 3],
             11: list(range(2048 * 3 - 8, 2048 * 3 + 8))}
    for ch, offs in spots.items():
        for o in offs:
            iq[ch, o] = 0
            iq[ch, nbytes + o] = 0
    e = R.Engine(n, 0, nbytes)
    e.set_modes(modes)
    e.debug_set_tile_loader(tma)
    pcm, _ = e.demodulate(iq)
    exp = _oracle_rows(modes, iq)
    for ch in range(n):
        assert np.array_equal(pcm[ch], exp[ch]), "channel %d mode %d" % (ch, modes[ch])
    e.close()


@pytest.mark.parametrize("tma", [0, 2, 12])
def test_signed_rotated_input(tma):
    """The .iq file / IQ dump format (already signed and rotated) through every loader."""
    import rtlsdrdiags_b200 as R
    n, nbytes = 9, 32768 + 4096
    modes = np.array([(1, 4, 5)[ch % 3] for ch in range(n)], dtype=np.uint8)
    rng = np.random.default_rng(8)
    iq = rng.integers(-128, 128, size=(n, 2 * nbytes), dtype=np.int8)
    iq[1] = np.clip(rng.normal(0, 30, size=2 * nbytes), -127, 127).astype(np.int8)
    e = R.Engine(n, 0, nbytes)
    e.set_modes(modes)
    e.debug_set_tile_loader(tma)
    pcm, _ = e.demodulate(iq, fmt=R.IQ_S8_ROTATED)
    for ch in range(n):
        c = O.OracleChain()
        exp = c.accept_s8(int(modes[ch]), iq[ch])
        assert np.array_equal(pcm[ch], exp), "channel %d mode %d" % (ch, modes[ch])
    e.close()
