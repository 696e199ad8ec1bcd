"""ctypes bindings for the CHECKERS: oracle/liboracle.so (our C restatement) and,
when present, oracle/_ref/libref_*.so (the unmodified reference compiled in place).

Test infrastructure only: imported by tests/ (through tests/_oracle.py),
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under rtlsdrdiags_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)

MODE_NONE, MODE_AM, MODE_FM, MODE_WBFM, MODE_LSB, MODE_USB = range(6)
KIND_AM, KIND_FM, KIND_WBFM, KIND_SSB = 1, 2, 3, 4
VARIANT_RADIODIAGS, VARIANT_RESEARCH = 0, 1
MODE_TO_KIND = {MODE_AM: KIND_AM, MODE_FM: KIND_FM, MODE_WBFM: KIND_WBFM,
                MODE_LSB: KIND_SSB, MODE_USB: KIND_SSB}

_vp, _u32, _i32, _f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
_pf = C.POINTER(C.c_float)
_pi16 = C.POINTER(C.c_int16)
_pi8 = C.POINTER(C.c_int8)
_pu8 = C.POINTER(C.c_uint8)


def build_oracle():
    """(Re)build liboracle.so and, if /root/reference is mounted, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL)


def _ptr(a, t):
    return a.ctypes.data_as(t)


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.sdro_dec16_new.restype = _vp
    L.sdro_dec16_new.argtypes = [_i32, _pf, _i32]
    L.sdro_dec16_free.argtypes = [_vp]
    L.sdro_dec16_reset.argtypes = [_vp]
    L.sdro_dec16_run.restype = _u32
    L.sdro_dec16_run.argtypes = [_vp, _pi16, _u32, _pi16]
    L.sdro_fir16_new.restype = _vp
    L.sdro_fir16_new.argtypes = [_i32, _pf]
    L.sdro_fir16_free.argtypes = [_vp]
    L.sdro_fir16_run.argtypes = [_vp, _pi16, _u32, _pi16]
    L.sdro_fir_new.restype = _vp
    L.sdro_fir_new.argtypes = [_i32, _pf]
    L.sdro_fir_free.argtypes = [_vp]
    L.sdro_fir_run.argtypes = [_vp, _pf, _u32, _pf]
    L.sdro_iir_new.restype = _vp
    L.sdro_iir_new.argtypes = [_i32, _pf, _i32, _pf]
    L.sdro_iir_free.argtypes = [_vp]
    L.sdro_iir_run.argtypes = [_vp, _pf, _u32, _pf]
    L.sdro_chain_new.restype = _vp
    L.sdro_chain_new.argtypes = [_i32]
    L.sdro_chain_free.argtypes = [_vp]
    L.sdro_chain_set_mode.argtypes = [_vp, _i32]
    L.sdro_chain_set_gain.argtypes = [_vp, _i32, _f32]
    L.sdro_chain_reset.argtypes = [_vp, _i32]
    L.sdro_chain_accept_u8.restype = _u32
    L.sdro_chain_accept_u8.argtypes = [_vp, _pu8, _u32, _pi16, _u32]
    L.sdro_chain_accept_s8.restype = _u32
    L.sdro_chain_accept_s8.argtypes = [_vp, _i32, _pi8, _u32, _pi16, _u32]
    L.sdro_chain_set_threshold.argtypes = [_vp, C.c_int32]
    L.sdro_chain_set_rx_gain.argtypes = [_vp, _u32]
    L.sdro_chain_signal.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    L.sdro_front_end.argtypes = [_pu8, C.c_uint32]
    L.sdro_front_end.restype = None
    L.sdro_dump_datagrams.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
    L.sdro_dump_datagrams.restype = C.c_uint32
    L.sdro_q15_taps.restype = _i32
    L.sdro_q15_taps.argtypes = [_i32, _pi16]
    L.sdro_atan2f.restype = _f32
    L.sdro_atan2f.argtypes = [_i32, _i32]
    L.sdro_bank_run.restype = C.c_double
    L.sdro_bank_run.argtypes = [_pu8, _u32, _pu8, C.c_uint64, _u32, _pi16, _u32, _i32]
    L.sdro_mr_new.restype = _vp
    L.sdro_mr_new.argtypes = [_i32, _i32, _pf, _i32]
    L.sdro_mr_free.argtypes = [_vp]
    L.sdro_mr_reset.argtypes = [_vp]
    L.sdro_mr_run.restype = C.c_uint64
    L.sdro_mr_run.argtypes = [_vp, _vp, C.c_uint64, _vp]
    _oracle = L
    return L


_refs = {}


def ref(tree="radiodiags"):
    """The compiled reference, or None when oracle/_ref was never built."""
    if tree in _refs:
        return _refs[tree]
    path = os.path.join(ORACLE_DIR, "_ref", "libref_%s.so" % tree)
    if not os.path.exists(path):
        _refs[tree] = None
        return None
    L = C.CDLL(path)
    L.ref_demod_new.restype = _vp
    L.ref_demod_new.argtypes = [_i32]
    L.ref_demod_free.argtypes = [_vp]
    L.ref_demod_set_gain.argtypes = [_vp, _f32]
    L.ref_demod_reset.argtypes = [_vp]
    L.ref_ssb_set_lsb.argtypes = [_vp, _i32]
    L.ref_demod_accept.restype = _u32
    L.ref_demod_accept.argtypes = [_vp, _pi8, _u32, _pi16, _u32]
    L.ref_dec16_new.restype = _vp
    L.ref_dec16_new.argtypes = [_i32, _pf, _i32]
    L.ref_dec16_free.argtypes = [_vp]
    L.ref_dec16_run.restype = _u32
    L.ref_dec16_run.argtypes = [_vp, _pi16, _u32, _pi16]
    L.ref_fir16_new.restype = _vp
    L.ref_fir16_new.argtypes = [_i32, _pf]
    L.ref_fir16_free.argtypes = [_vp]
    L.ref_fir16_run.argtypes = [_vp, _pi16, _u32, _pi16]
    L.ref_fir_new.restype = _vp
    L.ref_fir_new.argtypes = [_i32, _pf]
    L.ref_fir_free.argtypes = [_vp]
    L.ref_fir_run.argtypes = [_vp, _pf, _u32, _pf]
    L.ref_iir_new.restype = _vp
    L.ref_iir_new.argtypes = [_i32, _pf, _i32, _pf]
    L.ref_iir_free.argtypes = [_vp]
    L.ref_iir_run.argtypes = [_vp, _pf, _u32, _pf]
    if tree == "radiodiags":
        L.ref_iqp_new.restype = _vp
        L.ref_iqp_new_port.argtypes = [C.c_int]
        L.ref_iqp_new_port.restype = _vp
        L.ref_iqp_set_dump.argtypes = [_vp, C.c_int]
        L.ref_iqp_free.argtypes = [_vp]
        L.ref_iqp_set_mode.argtypes = [_vp, _i32]
        L.ref_iqp_set_gain.argtypes = [_vp, _i32, _f32]
        L.ref_iqp_reset.argtypes = [_vp, _i32]
        L.ref_iqp_accept.restype = _u32
        L.ref_iqp_accept.argtypes = [_vp, _pu8, _u32, _pi16, _u32]
        L.ref_iqp_set_threshold.argtypes = [_vp, C.c_int32]
        L.ref_set_rx_gain.argtypes = [C.c_int32]
        L.ref_iqp_signal.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
        L.ref_bank_run.restype = C.c_double
        L.ref_bank_run.argtypes = [_pu8, _u32, _pu8, C.c_uint64, _u32, _pi16, _u32]
        L.ref_mr_new.restype = _vp
        L.ref_mr_new.argtypes = [_i32, _i32, _pf, _i32]
        L.ref_mr_free.argtypes = [_vp]
        L.ref_mr_reset.argtypes = [_vp]
        L.ref_mr_run.restype = C.c_uint64
        L.ref_mr_run.argtypes = [_vp, _vp, C.c_uint64, _vp]
    _refs[tree] = L
    return L


# ---------------------------------------------------------------- wrappers
class OracleChain:
    """One channel of the restated path (IqDataProcessor + four demodulators)."""

    def __init__(self, variant=VARIANT_RADIODIAGS):
        self.L = oracle()
        self.h = self.L.sdro_chain_new(variant)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sdro_chain_free(self.h)
            self.h = None

    def set_mode(self, mode):
        self.L.sdro_chain_set_mode(self.h, mode)

    def set_gain(self, kind, gain):
        self.L.sdro_chain_set_gain(self.h, kind, gain)

    def reset(self, kind):
        self.L.sdro_chain_reset(self.h, kind)

    def set_threshold(self, dbfs):
        self.L.sdro_chain_set_threshold(self.h, int(dbfs))

    def set_rx_gain(self, gain_db):
        self.L.sdro_chain_set_rx_gain(self.h, int(gain_db))

    def signal(self):
        a, m = C.c_int(), C.c_uint32()
        self.L.sdro_chain_signal(self.h, C.byref(a), C.byref(m))
        return bool(a.value), int(m.value)

    def accept_u8(self, iq):
        buf = np.array(iq, dtype=np.uint8, copy=True)
        out = np.empty(buf.size // 64 + 2, dtype=np.int16)
        n = self.L.sdro_chain_accept_u8(self.h, _ptr(buf, _pu8), buf.size, _ptr(out, _pi16), out.size)
        return out[:n].copy()

    def accept_s8(self, mode, iq):
        buf = np.array(iq, dtype=np.int8, copy=True)
        out = np.empty(buf.size // 64 + 2, dtype=np.int16)
        n = self.L.sdro_chain_accept_s8(self.h, mode, _ptr(buf, _pi8), buf.size, _ptr(out, _pi16), out.size)
        return out[:n].copy()


def front_end(iq_u8):
    """Oracle: the signed, Fs/4-rotated block acceptIqData leaves in the caller's buffer."""
    buf = np.array(iq_u8, dtype=np.uint8, copy=True)
    oracle().sdro_front_end(_ptr(buf, _pu8), buf.size)
    return buf.view(np.int8)


def dump_datagrams(nbytes):
    """Oracle: the datagram sizes UdpClient::sendData cuts nbytes into."""
    sizes = (C.c_uint32 * (nbytes // 2048 + 2))()
    k = oracle().sdro_dump_datagrams(nbytes, sizes, len(sizes))
    return [int(sizes[i]) for i in range(k)]


class RefChain:
    """One channel of the compiled reference product path (radioDiags tree)."""

    def __init__(self, dump_port=None):
        self.L = ref("radiodiags")
        if self.L is None:
            raise RuntimeError("oracle/_ref not built")
        self.h = self.L.ref_iqp_new() if dump_port is None else self.L.ref_iqp_new_port(int(dump_port))

    def set_dump(self, on):
        """enableIqDump / disableIqDump: UDP datagrams to 127.0.0.1:dump_port."""
        self.L.ref_iqp_set_dump(self.h, int(bool(on)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_iqp_free(self.h)
            self.h = None

    def set_mode(self, mode):
        self.L.ref_iqp_set_mode(self.h, mode)

    def set_gain(self, kind, gain):
        self.L.ref_iqp_set_gain(self.h, kind, gain)

    def reset(self, kind):
        self.L.ref_iqp_reset(self.h, kind)

    def set_threshold(self, dbfs):
        self.L.ref_iqp_set_threshold(self.h, int(dbfs))

    def set_rx_gain(self, gain_db):
        """Process-wide in the reference: affects every RefChain."""
        self.L.ref_set_rx_gain(int(gain_db))

    def signal(self):
        a, m = C.c_int(), C.c_uint32()
        self.L.ref_iqp_signal(self.h, C.byref(a), C.byref(m))
        return bool(a.value), int(m.value)

    def accept_u8(self, iq, block=32768):
        buf = np.array(iq, dtype=np.uint8, copy=True)
        out = np.empty(buf.size // 64 + 2, dtype=np.int16)
        done = 0
        for off in range(0, buf.size, block):
            piece = buf[off:off + block]
            done += self.L.ref_iqp_accept(self.h, _ptr(piece, _pu8), piece.size,
                                          _ptr(out[done:], _pi16), out.size - done)
        return out[:done].copy()


class RefDemod:
    """One reference demodulator object (either tree), fed signed rotated IQ."""

    def __init__(self, kind, tree="radiodiags"):
        self.L = ref(tree)
        if self.L is None:
            raise RuntimeError("oracle/_ref not built")
        self.h = self.L.ref_demod_new(kind)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_demod_free(self.h)
            self.h = None

    def set_gain(self, g):
        self.L.ref_demod_set_gain(self.h, g)

    def reset(self):
        self.L.ref_demod_reset(self.h)

    def set_lsb(self, lsb):
        self.L.ref_ssb_set_lsb(self.h, int(lsb))

    def accept(self, iq, block=32768):
        buf = np.array(iq, dtype=np.int8, copy=True)
        out = np.empty(buf.size // 64 + 2, dtype=np.int16)
        done = 0
        for off in range(0, buf.size, block):
            piece = buf[off:off + block]
            done += self.L.ref_demod_accept(self.h, _ptr(piece, _pi8), piece.size,
                                            _ptr(out[done:], _pi16), out.size - done)
        return out[:done].copy()


def oracle_bank(modes, iq_u8, block_bytes=32768, nthreads=1, variant=VARIANT_RADIODIAGS,
                want_pcm=True):
    """Run the restated path over a [n_channels][bytes] u8 bank. Returns (pcm, seconds)."""
    L = oracle()
    iq = np.array(iq_u8, dtype=np.uint8, copy=True)
    n_ch, nbytes = iq.shape
    modes = np.ascontiguousarray(modes, dtype=np.uint8)
    pcm = np.zeros((n_ch, nbytes // 64), dtype=np.int16) if want_pcm else None
    secs = L.sdro_bank_run(_ptr(modes, _pu8), n_ch, _ptr(iq, _pu8), nbytes, block_bytes,
                           _ptr(pcm, _pi16) if want_pcm else None, nthreads, variant)
    return pcm, secs


def ref_bank(modes, iq_u8, block_bytes=32768, nthreads=1, want_pcm=True):
    L = ref("radiodiags")
    if L is None:
        raise RuntimeError("oracle/_ref not built")
    iq = np.array(iq_u8, dtype=np.uint8, copy=True)
    n_ch, nbytes = iq.shape
    modes = np.ascontiguousarray(modes, dtype=np.uint8)
    pcm = np.zeros((n_ch, nbytes // 64), dtype=np.int16) if want_pcm else None
    secs = L.ref_bank_run(_ptr(modes, _pu8), n_ch, _ptr(iq, _pu8), nbytes, block_bytes,
                          _ptr(pcm, _pi16) if want_pcm else None, nthreads)
    return pcm, secs


def q15_taps(filter_id):
    q = np.zeros(64, dtype=np.int16)
    n = oracle().sdro_q15_taps(filter_id, _ptr(q, _pi16))
    return q[:n].copy()


# ---- generic multirate classes ----
MR_DECIMATOR_F32, MR_INTERPOLATOR_F32, MR_DECIMATOR_I16, MR_INTERPOLATOR_I16 = 1, 2, 3, 4


class Multirate:
    """One Decimator / Interpolator / Decimator_int16 / Interpolator_int16 object: the C
    restatement (impl="oracle") or the compiled reference class itself (impl="ref")."""

    def __init__(self, kind, taps, factor, impl="oracle"):
        self.kind, self.factor = int(kind), int(factor)
        self.dtype = np.float32 if kind <= 2 else np.int16
        self.interp = kind in (MR_INTERPOLATOR_F32, MR_INTERPOLATOR_I16)
        h = np.ascontiguousarray(taps, dtype=np.float32)
        if impl == "oracle":
            self.L, self.pre = oracle(), "sdro_mr_"
        else:
            self.L, self.pre = ref("radiodiags"), "ref_mr_"
            if self.L is None:
                raise RuntimeError("oracle/_ref is not built")
        self.h = getattr(self.L, self.pre + "new")(self.kind, h.size, _ptr(h, _pf), self.factor)
        if not self.h:
            raise ValueError("bad multirate parameters")

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.L, self.pre + "free")(self.h)
            self.h = None

    def reset(self):
        getattr(self.L, self.pre + "reset")(self.h)

    def run(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        cap = x.size * self.factor if self.interp else x.size // self.factor + 2
        out = np.zeros(cap, dtype=self.dtype)
        n = getattr(self.L, self.pre + "run")(self.h, x.ctypes.data_as(_vp), x.size, out.ctypes.data_as(_vp))
        return out[:n].copy()
