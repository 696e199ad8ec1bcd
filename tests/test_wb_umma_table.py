"""CPU check of the tcgen05 formulation of the WBFM pre-filter (wbfm_tile4_kernel, WbUmma): the tap
matrices the engine uploads (wb_umma_table), unpacked from their no-swizzle K-major layout, multiplied
with raw input bytes exactly as the kernel's MMAs do -- per window and group of 16 samples two K = 32
steps, the first from the row above for the window's first group -- plus the accumulator starts the kernel
stores into tensor memory, must reproduce the reference's arithmetic -- u8 -> s8, Fs/4 rotation
(IqDataProcessor.cc:735-738, 567-611), then the 16-tap FirFilter_int16 on both arms with the result
truncated to int8 (WbFmDemodulator.cc:17-35, 389-398) -- for every sample of a stream, except where a raw
byte 0 sits in a position the rotation negates (the kernel detects those tiles and takes the CUDA-core
path). No GPU needed: the table is host code. (The descriptor semantics the kernel relies on are pinned on
the GPU by tools/micro/umma_toeplitz.cu.)"""
import ctypes as C

import numpy as np

import _oracle as O
from test_wb_mma_table import _prefilter_reference


def _table():
    import rtlsdrdiags_b200 as R
    L = R.load_library()
    L.sdr_debug_wb_umma_table.argtypes = [C.c_void_p, C.c_void_p]
    taps = np.zeros(4096, dtype=np.uint8)
    starts = np.zeros(8, dtype=np.int32)
    assert L.sdr_debug_wb_umma_table(taps.ctypes.data_as(C.c_void_p), starts.ctypes.data_as(C.c_void_p)) == taps.size
    B = np.zeros((2, 2, 32, 32), dtype=np.int64)                  # [k-step][part][column n][k]
    t = taps.view(np.int8).astype(np.int64)
    for ks in range(2):
        for part in range(2):
            for n in range(32):
                for k in range(32):
                    off = (n & 7) * 16 + (n >> 3) * 256 + (k & 15) + (k >> 4) * 128
                    B[ks, part, n, k] = t[(ks * 2 + part) * 1024 + off]
    return B, starts.astype(np.int64)


def _gemm(B, starts, raw, hist):
    """raw: the stream's bytes (a multiple of 64); hist: the 32 bytes before it. What the MMAs compute."""
    n_win = raw.size // 64
    rows = raw.reshape(n_win, 64).astype(np.int64)
    above = np.vstack([np.concatenate([np.zeros(32, np.int64), hist.astype(np.int64)])[None, :], rows[:-1]])
    out_i = np.empty(n_win * 32, dtype=np.int8)
    out_q = np.empty(n_win * 32, dtype=np.int8)
    start_cols = np.array([starts[2 * (p & 3) + arm] for p in range(16) for arm in range(2)], dtype=np.int64)
    for jj in range(2):
        a0 = above[:, 32:] if jj == 0 else rows[:, :32]
        a1 = rows[:, :32] if jj == 0 else rows[:, 32:]
        hi = a0 @ B[0, 0].T + a1 @ B[1, 0].T
        lo = a0 @ B[0, 1].T + a1 @ B[1, 1].T + start_cols[None, :]
        acc = 256 * hi + lo                                        # doubled accumulator: int8 result = byte 2
        got = ((acc >> 16) & 0xFF).astype(np.uint8).view(np.int8)  # [window][2 p + arm]
        for w in range(n_win):
            out_i[32 * w + 16 * jj: 32 * w + 16 * jj + 16] = got[w, 0::2]
            out_q[32 * w + 16 * jj: 32 * w + 16 * jj + 16] = got[w, 1::2]
    return out_i, out_q


def test_mma_with_the_table_is_the_reference_prefilter():
    B, starts = _table()
    rng = np.random.default_rng(8)
    n_bytes = 2048 * 3
    raw = rng.integers(1, 256, size=n_bytes, dtype=np.uint8)      # no clipping byte
    raw[100:180] = 255
    raw[300:380] = 1
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    got_i, got_q = _gemm(B, starts, raw, np.full(32, 128, np.uint8))   # the stream starts from silence
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_q, ref_q)
    # a stream cut in two: the second piece's history record is the first piece's last 32 bytes
    cut = 2048 + 64 * 5
    a_i, a_q = _gemm(B, starts, raw[:cut], np.full(32, 128, np.uint8))
    b_i, b_q = _gemm(B, starts, raw[cut:], raw[cut - 32:cut])
    assert np.array_equal(np.concatenate([a_i, b_i]), ref_i) and np.array_equal(np.concatenate([a_q, b_q]), ref_q)
    assert np.abs(B).max() <= 128 and B.max() <= 127
    # the history really is 15 samples: nothing meets the K step's first two bytes
    assert not B[0, :, :, :2].any() and B[0, :, :, 2:4].any()


def test_the_clipping_byte_is_what_the_mma_cannot_do():
    """Raw 0x00 where the rotation negates: the reference keeps -128, a linear map gives +128."""
    B, starts = _table()
    raw = np.full(2048, 140, dtype=np.uint8)
    raw[64 + 3] = 0                                               # Q1 of a rotation group: negated
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    got_i, got_q = _gemm(B, starts, raw, np.full(32, 128, np.uint8))
    assert not (np.array_equal(got_i, ref_i) and np.array_equal(got_q, ref_q))
    raw[64 + 3] = 1                                               # -127: no wrap, linear again
    ref_i, ref_q = _prefilter_reference(O.front_end(raw))
    got_i, got_q = _gemm(B, starts, raw, np.full(32, 128, np.uint8))
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_q, ref_q)
