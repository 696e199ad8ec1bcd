"""CPU check of the tensor-core formulation of AM / SSB stage 1 (AmSsbTile::stage1_mma): the taps
table the engine uploads (am_mma_table), unpacked from its mma.m16n8k32 B-fragment layout and
multiplied with raw input bytes exactly as the GEMM does, must reproduce the reference's
arithmetic -- u8 -> s8, Fs/4 rotation (IqDataProcessor.cc:735-738, 567-611), then the 8-tap 4:1
Decimator_int16 (Decimator_int16.cc:176-238, 310-351; AmDemodulator.cc:349-374) -- for every
output of a stream, except where a raw byte 0 sits in a position the rotation negates (the kernel
detects those tiles and takes the CUDA-core path). No GPU needed: the table is host code."""
import ctypes as C

import numpy as np

import _oracle as O


def _table():
    import rtlsdrdiags_b200 as R
    L = R.load_library()
    L.sdr_debug_am_mma_table.argtypes = [C.c_void_p]
    out = np.zeros(2 * 32 * 12, dtype=np.uint32)
    assert L.sdr_debug_am_mma_table(out.ctypes.data_as(C.c_void_p)) == out.size
    return out.reshape(2, 32, 12)


def _unpack(tab_f):
    """B[part][kb 0..63][col 0..7] (int8 values) and the accumulator starts [part][col]."""
    B = np.zeros((2, 64, 8), dtype=np.int64)
    init = np.zeros((2, 8), dtype=np.int64)
    for lane in range(32):
        g, tq = lane >> 2, lane & 3
        for h in range(2):
            for s2 in range(2):
                for r in range(2):
                    w = int(tab_f[lane, (h * 2 + s2) * 2 + r])
                    for b in range(4):
                        v = (w >> (8 * b)) & 0xFF
                        B[h, 32 * s2 + 16 * r + 4 * tq + b, g] = v - 256 if v >= 128 else v
        for i in range(4):
            v = int(tab_f[lane, 8 + i])
            init[i >> 1, 2 * tq + (i & 1)] = v - (1 << 32) if v >= (1 << 31) else v
    return B, init


def _stage1_reference(s8_rotated):
    """Decimator_int16 with the AM stage-1 taps on both arms of a signed, rotated stream: int8 out."""
    q = O.q15_taps(0).astype(np.int64)           # AM1, SURVEY A.1: 795 2511 4776 6419 ...
    assert q.sum() == 29002
    x = s8_rotated.astype(np.int64)
    arms = []
    for arm in range(2):
        v = np.concatenate([np.zeros(8, np.int64), x[arm::2]])
        n_out = (v.size - 8) // 4
        out = np.empty(n_out, dtype=np.int64)
        for m in range(n_out):
            newest = 8 + 4 * m + 3
            acc = 16384 + sum(int(q[k]) * int(v[newest - k]) for k in range(8))
            out[m] = acc >> 15
        arms.append(out)
    return arms


def test_gemm_with_the_table_is_the_reference_stage1():
    tab = _table()
    rng = np.random.default_rng(4)
    n_bytes = 2048 * 3
    for fmt in (0, 1):
        B, init = _unpack(tab[fmt])
        if fmt == 0:
            raw = rng.integers(1, 256, size=n_bytes, dtype=np.uint8)      # no clipping byte
            raw[100:140] = 255
            signed = O.front_end(raw)
            data = raw.astype(np.int64)
            hist_fill = 128
        else:
            signed = rng.integers(-128, 128, size=n_bytes, dtype=np.int8)
            data = signed.astype(np.int64)
            hist_fill = 0
        ref_i, ref_q = _stage1_reference(signed)
        padded = np.concatenate([np.full(32, hist_fill, np.int64), data])  # the stream starts from silence
        for h in range(n_bytes // 32):
            row = padded[32 * h: 32 * h + 64]                 # bytes 32h-32 .. 32h+31 of the stream
            hi = init[0] + row @ B[0]
            lo = init[1] + row @ B[1]
            acc = 256 * hi + lo                                 # doubled accumulator: int8 result = byte 2
            got = ((acc >> 16) & 0xFF).astype(np.uint8).view(np.int8)
            for p in range(4):
                assert got[p] == np.int8(ref_i[4 * h + p]), (fmt, h, p)
                assert got[4 + p] == np.int8(ref_q[4 * h + p]), (fmt, h, p)
        assert np.abs(B).max() <= 127 and np.abs(B[0]).max() <= 51


def test_the_clipping_byte_is_what_the_gemm_cannot_do():
    """Raw 0x00 where the rotation negates: the reference keeps -128, a linear map gives +128."""
    tab = _table()
    B, init = _unpack(tab[0])
    raw = np.full(2048, 140, dtype=np.uint8)
    raw[64 + 3] = 0                                           # Q1 of a rotation group: negated
    signed = O.front_end(raw)
    ref_i, _ = _stage1_reference(signed)
    padded = np.concatenate([np.full(32, 128, np.int64), raw.astype(np.int64)])
    h = 2
    row = padded[32 * h: 32 * h + 64]
    acc = 256 * (init[0] + row @ B[0]) + init[1] + row @ B[1]
    got = ((acc >> 16) & 0xFF).astype(np.uint8).view(np.int8)
    assert any(got[p] != np.int8(ref_i[4 * h + p]) for p in range(4))
