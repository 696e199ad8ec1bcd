"""The AM / SSB FIR kernel fetches full tiles with TMA (cp.async.bulk.tensor through a tensor map
of the caller's IQ array, 128-byte hardware swizzle) and a partial last tile with cp.async. Both
loaders against the oracle: strided and offset inputs, ragged lengths, device-resident input."""
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


@pytest.mark.parametrize("tma", [0, 2, 3, 4])
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


@pytest.mark.parametrize("tma", [0, 2, 4])
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
