"""ctypes driver for tests/emu/libemu.so -- the serial host emulation of the CUDA
pipelines (test harness only; see tests/emu/emu.cc)."""
import ctypes as C
import os
import subprocess

import numpy as np

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "rtlsdrdiags_b200", "csrc")

KIND_AM, KIND_FM, KIND_WBFM, KIND_SSB = 1, 2, 3, 4
FMT_U8, FMT_S8 = 0, 1

_lib = None


def build(defines=()):
    out = os.path.join(EMU_DIR, "libemu.so")
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu.cc", "sdr_phase.cuh", "emu_config.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(out) and not defines and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-DSDR_EMU", "-fPIC",
           "-shared", "-I", CSRC, "-I", EMU_DIR, "-o", out, os.path.join(EMU_DIR, "emu.cc")] + ["-D" + d for d in defines]
    subprocess.run(cmd, check=True)
    return out


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_state_bytes.argtypes = [C.c_int]
        L.emu_run.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32,
                              C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_uint64, C.c_void_p, C.c_int]
        _lib = L
    return _lib


_luts = {}


def lut(kind):
    """atan2 tables exactly as the engine builds them (libm atan2 -> float)."""
    if kind not in _luts:
        L = O.oracle()
        if kind == KIND_FM:
            t = np.array([[L.sdro_atan2f(q, i) for i in range(-140, 140)] for q in range(-140, 140)],
                         dtype=np.float32)
        else:
            t = np.array([[L.sdro_atan2f(q - 128, i - 128) for i in range(256)] for q in range(256)],
                         dtype=np.float32)
        _luts[kind] = np.ascontiguousarray(t)
    return _luts[kind]


def scale_for(kind, gain, variant=0):
    g = np.float32(gain)
    if kind in (KIND_AM, KIND_SSB) or variant == 1:
        return g
    dev = np.float32(15000 if kind == KIND_FM else 75000)
    return np.float32(np.float32(g / dev) * np.float32(32767))


class EmuBank:
    """A bank of channels of ONE mode run through the emulated kernel, with state."""

    def __init__(self, kind, n_channels, G=4, NT=64):
        self.L = lib()
        self.kind, self.n, self.G, self.NT = kind, n_channels, G, NT
        self.sb = self.L.emu_state_bytes(kind)
        self.state = np.zeros((n_channels, self.sb), dtype=np.uint8)
        self.scale = np.zeros(n_channels, dtype=np.float32)
        self.lsb = np.ones(n_channels, dtype=np.uint8)
        self.chan_ids = np.arange(n_channels, dtype=np.uint32)
        self.lut = lut(kind) if kind in (KIND_FM, KIND_WBFM) else np.zeros(1, dtype=np.float32)

    def run(self, iq, fmt):
        iq = np.ascontiguousarray(iq).view(np.uint8)
        n_ch, nbytes = iq.shape
        assert n_ch == self.n and nbytes % 64 == 0
        pcm = np.zeros((n_ch, nbytes // 64 + (nbytes // 64) % 2), dtype=np.int16)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.L.emu_run(self.kind, p(iq), iq.strides[0], nbytes // 2, fmt, p(self.chan_ids), self.n,
                            self.G, p(self.state), self.state.strides[0], p(self.scale), p(self.lsb), p(pcm),
                            pcm.strides[0] // 2, p(self.lut), self.NT)
        assert rc == 0
        return pcm[:, :nbytes // 64]
