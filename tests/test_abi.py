"""The C-ABI shared library builds, loads and exports every symbol include/sdr_b200.h
declares; without a GPU it refuses to create an engine (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import rtlsdrdiags_b200 as R
    lib = R.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libsdr_b200.so does not export %s" % n
    assert sorted(R.ABI_SYMBOLS) == names
    assert b"sm_100a" in lib.sdr_version()


def test_state_sizes_and_argument_checks_without_a_device():
    import rtlsdrdiags_b200 as R
    lib = R.load_library()
    sizes = [lib.sdr_state_bytes(k) for k in (1, 2, 3, 4)]
    assert all(s > 0 and s % 16 == 0 for s in sizes)
    assert lib.sdr_state_bytes(9) == -1
    h = ctypes.c_void_p()
    assert lib.sdr_engine_create(0, 0, 32768, ctypes.byref(h)) == -1      # SDR_E_ARG
    assert lib.sdr_engine_create(4, 0, 1000, ctypes.byref(h)) == -1       # not a multiple of 64
    assert lib.sdr_set_mode(None, 0, 1) == -1
    assert lib.sdr_accept_iq(None, None, 64, 64, 0) == -1


def test_no_cpu_fallback():
    import torch
    import rtlsdrdiags_b200 as R
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.SdrError, match="no CPU path"):
        R.Engine(4)


def test_product_sources_never_touch_the_oracle():
    """The oracle is a checker: nothing the product builds from may include or import it."""
    pkg = os.path.join(ROOT, "rtlsdrdiags_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "sdr_oracle" not in txt and "oracle_binding" not in txt and "liboracle" not in txt, f
                assert "libemu" not in txt, f
