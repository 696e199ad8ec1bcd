"""CPU check of the tensor-core formulation of the WBFM pre-filter (WbMma::prefilter): the taps table
the engine uploads (wb_mma_table), unpacked from its mma.m16n8k32 A-fragment layout and
multiplied with raw input bytes (and the constant granule that carries the accumulator starts) exactly as
the GEMM does, must reproduce the reference's arithmetic
-- u8 -> s8, Fs/4 rotation (IqDataProcessor.cc:735-738, 567-611), then the 16-tap FirFilter_int16 on
both arms with the result truncated to int8 (WbFmDemodulator.cc:17-35, 389-398) -- for every sample
of a stream, except where a raw byte 0 sits in a position the rotation negates (the kernel detects
those tiles and takes the CUDA-core path). No GPU needed: the table is host code."""
import ctypes as C

import numpy as np

import _oracle as O


def _table():
    import rtlsdrdiags_b200 as R
    L = R.load_library()
    L.sdr_debug_wb_mma_table.argtypes = [C.c_void_p]
    out = np.zeros(4 * 32 * 4 + 4, dtype=np.uint32)
    assert L.sdr_debug_wb_mma_table(out.ctypes.data_as(C.c_void_p)) == out.size
    return out[:512].reshape(4, 32, 4), out[512:]


def _s8(v):
    return v - 256 if v >= 128 else v


def _unpack(tab, const):
    """E[part][row 0..15][K 0..63] (int8 values): K 0..15 meets the raw bytes kb 0..15, K 16..31 the constant
    granule, K 32..63 the raw bytes kb 16..47; and the constant granule's 16 bytes (u8)."""
    E = np.zeros((2, 16, 64), dtype=np.int64)
    for lane in range(32):
        g, tq = lane >> 2, lane & 3
        for h in range(2):
            for s2 in range(2):
                for r in range(4):                               # mma.m16n8k32 A fragment
                    w = int(tab[2 * s2 + h, lane, r])
                    for b in range(4):
                        E[h, g + 8 * (r & 1), 32 * s2 + 16 * (r >> 1) + 4 * tq + b] = _s8((w >> (8 * b)) & 0xFF)
    cbytes = np.array([(int(const[i >> 2]) >> (8 * (i & 3))) & 0xFF for i in range(16)], dtype=np.int64)
    return E, cbytes


def _prefilter_reference(s8_rotated):
    """FirFilter_int16 with the WBFM pre-filter taps on both arms of a signed, rotated stream that starts
    from silence, each output truncated to int8."""
    q = O.q15_taps(6).astype(np.int64)           # WB_PRE: -515 -1068 305 2036 ...
    assert q.size == 16 and np.abs(q).sum() == 54924
    x = s8_rotated.astype(np.int64)
    arms = []
    for arm in range(2):
        v = np.concatenate([np.zeros(15, np.int64), x[arm::2]])
        acc = np.full(v.size - 15, 16384, dtype=np.int64)
        for k in range(16):
            acc += q[k] * v[15 - k: v.size - k]
        arms.append(((acc >> 15) & 0xFF).astype(np.uint8).view(np.int8))
    return arms


def _gemm(E, cbytes, padded, n_granules):
    """What the kernel computes: per 16-byte granule G of the stream the 48 bytes that end with it, the
    constant granule spliced in after the first 16."""
    out_i = np.empty(8 * n_granules, dtype=np.int8)
    out_q = np.empty(8 * n_granules, dtype=np.int8)
    for G in range(n_granules):
        raw = padded[16 * G: 16 * G + 48]                         # bytes 16 G - 32 .. 16 G + 15
        col = np.concatenate([raw[:16], cbytes, raw[16:]])
        acc = 256 * (E[0] @ col) + E[1] @ col                     # doubled accumulator: int8 result = byte 2
        got = ((acc >> 16) & 0xFF).astype(np.uint8).view(np.int8)
        out_i[8 * G: 8 * G + 8] = got[:8]
        out_q[8 * G: 8 * G + 8] = got[8:]
    return out_i, out_q


def test_gemm_with_the_table_is_the_reference_prefilter():
    E, cb = _unpack(*_table())
    rng = np.random.default_rng(6)
    n_bytes = 2048 * 3
    raw = rng.integers(1, 256, size=n_bytes, dtype=np.uint8)      # no clipping byte
    raw[100:180] = 255
    raw[300:380] = 1
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    padded = np.concatenate([np.full(32, 128, np.int64), raw.astype(np.int64)])  # the stream starts from silence
    got_i, got_q = _gemm(E, cb, padded, n_bytes // 16)
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_q, ref_q)
    # the history really is 15 samples: nothing meets kb 0, 1; the constant granule: twelve 255 and four 1
    assert not E[:, :, :2].any() and E[:, :, 2:4].any()
    assert list(cb) == [255] * 12 + [1] * 4


def test_the_clipping_byte_is_what_the_gemm_cannot_do():
    """Raw 0x00 where the rotation negates: the reference keeps -128, a linear map gives +128."""
    E, cb = _unpack(*_table())
    raw = np.full(2048, 140, dtype=np.uint8)
    raw[64 + 3] = 0                                               # Q1 of a rotation group: negated
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    padded = np.concatenate([np.full(32, 128, np.int64), raw.astype(np.int64)])
    got_i, got_q = _gemm(E, cb, padded, 2048 // 16)
    assert not (np.array_equal(got_i, ref_i) and np.array_equal(got_q, ref_q))
    raw[64 + 3] = 1                                               # -127: no wrap, linear again
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    padded = np.concatenate([np.full(32, 128, np.int64), raw.astype(np.int64)])
    got_i, got_q = _gemm(E, cb, padded, 2048 // 16)
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_q, ref_q)
