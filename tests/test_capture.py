"""The reference's own capture. tests/golden/golden_capture_v1.npz holds a 64 KiB excerpt of
demodulatorResearch/yoyo.iq -- the only IQ capture in the reference tree -- and the PCM the
unmodified reference compiled in place makes of it, for both trees (demod.cc:200-323 feeding the
research demodulators; IqDataProcessor::acceptIqData feeding the radioDiags ones). The oracle is
pinned to it on the CPU; on the GPU the engine (both entry formats) and the offline driver
b200_demod must reproduce it bit for bit. With the reference mounted (the build container) the
whole 2 MiB file is compared by md5 as well (tests/golden/golden_v1.npz)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import _oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}
YOYO = "/root/reference/demodulatorResearch/yoyo.iq"


def _golden():
    return np.load(os.path.join(HERE, "golden", "golden_capture_v1.npz"))


def _unrotate(s8):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import unrotate_to_u8
    return unrotate_to_u8(s8)


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_oracle_reproduces_the_capture_golden(mode):
    g = _golden()
    s8 = g["iq_s8"]
    c = O.OracleChain(O.VARIANT_RESEARCH)
    assert np.array_equal(c.accept_s8(mode, s8), g["research_" + NAMES[mode]])
    c = O.OracleChain()
    c.set_mode(mode)
    u8 = _unrotate(s8)
    got = np.concatenate([c.accept_u8(u8[o:o + 32768]) for o in range(0, u8.size, 32768)])
    assert np.array_equal(got, g["radiodiags_" + NAMES[mode]])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_engine_reproduces_the_capture_golden(mode):
    import rtlsdrdiags_b200 as R
    g = _golden()
    s8 = g["iq_s8"]
    # research tree: signed, rotated input straight into the demodulators, 16384-byte reads
    e = R.Engine(1, 0, 16384)
    e.set_scaling(R.SCALING_RESEARCH)
    e.set_mode(0, mode)
    pcm, _ = e.demodulate(s8.reshape(1, -1), fmt=R.IQ_S8_ROTATED)
    assert np.array_equal(pcm[0], g["research_" + NAMES[mode]])
    e.close()
    # radioDiags tree: the dongle's u8 format through the IqDataProcessor entry, 32768-byte blocks
    e = R.Engine(1, 0, 32768)
    e.set_mode(0, mode)
    pcm, _ = e.demodulate(_unrotate(s8).reshape(1, -1))
    assert np.array_equal(pcm[0], g["radiodiags_" + NAMES[mode]])
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_offline_driver_reproduces_the_capture_golden(mode):
    """b200_demod, the counterpart of demod.cc: `-r` links the research scaling, `-u` takes the
    dongle's u8 format through the IqDataProcessor drop-in; `-l` asks for the real LSB (demod.cc's
    switch falls through to USB for -d 4, demod.cc:233-241)."""
    from rtlsdrdiags_b200 import _build
    _build.build()
    exe = _build.build_host()
    g = _golden()
    s8 = g["iq_s8"]

    def run(args, data):
        r = subprocess.run([exe] + args, input=data.tobytes(), capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        return np.frombuffer(r.stdout, dtype=np.int16)

    lsb = ["-l"] if mode == 4 else []
    assert np.array_equal(run(["-d", str(mode), "-r"] + lsb, s8), g["research_" + NAMES[mode]])
    assert np.array_equal(run(["-d", str(mode), "-u"] + lsb, _unrotate(s8)), g["radiodiags_" + NAMES[mode]])
    if mode == 4:  # the fall-through: -d 4 without -l is USB
        assert np.array_equal(run(["-d", "4", "-r"], s8), g["research_usb"])


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(YOYO), reason="the reference tree is not mounted on this box")
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_whole_capture_md5(mode):
    import rtlsdrdiags_b200 as R
    g1 = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    yoyo = np.fromfile(YOYO, dtype=np.int8)
    assert hashlib.md5(yoyo.tobytes()).hexdigest() == str(g1["yoyo_md5_input"])
    e = R.Engine(1, 0, 16384)
    e.set_scaling(R.SCALING_RESEARCH)
    e.set_mode(0, mode)
    pcm, _ = e.demodulate(yoyo.reshape(1, -1), fmt=R.IQ_S8_ROTATED)
    assert hashlib.md5(pcm[0].tobytes()).hexdigest() == str(g1["yoyo_md5_research_" + NAMES[mode]])
    e.close()


def test_whole_capture_md5_oracle():
    """CPU side of the same pin (runs where the reference is mounted)."""
    if not os.path.exists(YOYO):
        pytest.skip("the reference tree is not mounted on this box")
    g1 = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    yoyo = np.fromfile(YOYO, dtype=np.int8)
    c = O.OracleChain(O.VARIANT_RESEARCH)
    assert hashlib.md5(c.accept_s8(2, yoyo).tobytes()).hexdigest() == str(g1["yoyo_md5_research_fm"])
