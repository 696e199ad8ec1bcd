"""The C++ drop-in classes (rtlsdrdiags_b200/host: AmDemodulator, FmDemodulator,
WbFmDemodulator, SsbDemodulator, IqDataProcessor) and the offline driver b200_demod,
the counterpart of the reference's demod.cc."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
import _signals as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rtlsdrdiags_b200", "host")


def _demod():
    from rtlsdrdiags_b200 import _build
    _build.build()
    return _build.build_host()


def test_facade_keeps_the_reference_surface():
    """Same class names, constructor and method signatures as the reference headers
    (FmDemodulator.h:23-31, SsbDemodulator.h:24-34, IqDataProcessor.h:16-58)."""
    _demod()
    base = open(os.path.join(HOST, "B200Demodulator.h")).read()
    for sig in ["void resetDemodulator(void);", "void setDemodulatorGain(float gain);",
                "void acceptIqData(int8_t *bufferPtr, uint32_t bufferLength);"]:
        assert sig in base
    for cls in ["Am", "Fm", "WbFm", "Ssb"]:
        h = open(os.path.join(HOST, cls + "Demodulator.h")).read()
        assert "class %sDemodulator" % cls in h
        assert "%sDemodulator(void (*pcmCallbackPtr)(int16_t *bufferPtr, uint32_t bufferLength))" % cls in h
        assert "void displayInternalInformation(void)" in h
    ssb = open(os.path.join(HOST, "SsbDemodulator.h")).read()
    assert "void setLsbDemodulationMode(void)" in ssb and "void setUsbDemodulationMode(void)" in ssb
    iqp = open(os.path.join(HOST, "IqDataProcessor.h")).read()
    assert "enum demodulatorType {None = 0, Am = 1, Fm = 2, WbFm = 3, Lsb = 4, Usb = 5};" in iqp
    assert "void acceptIqData(unsigned long timeStamp, unsigned char *bufferPtr, unsigned long byteCount);" in iqp
    assert "IqDataProcessor(char *hostIpAddress, int hostPort);" in iqp


def test_driver_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([_demod(), "-d", "2"], input=b"", capture_output=True)
    assert r.returncode == 2 and b"no CPU path" in r.stderr


def _run(args, data):
    r = subprocess.run([_demod()] + args, input=data.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    return np.frombuffer(r.stdout, dtype=np.int16)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_driver_signed_input_matches_research_tree(mode):
    """b200_demod -r -d N == demod.cc -d N (research scaling, 16384-byte reads, -d 4 == USB)."""
    s8 = np.random.default_rng(mode).integers(-128, 128, size=16384 * 9 + 4096 + 40, dtype=np.int8)
    got = _run(["-d", str(mode), "-r"], s8)
    c = O.OracleChain(O.VARIANT_RESEARCH)
    exp = c.accept_s8(5 if mode == 4 else mode, s8)
    assert np.array_equal(got, exp)
    if mode == 4:
        got = _run(["-d", "4", "-r", "-l"], s8)
        assert np.array_equal(got, O.OracleChain(O.VARIANT_RESEARCH).accept_s8(4, s8))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3, 5])
@pytest.mark.parametrize("block", [32768, 4096 + 8, 1000])
def test_driver_raw_u8_input_matches_product_path(mode, block):
    """b200_demod -u: IqDataProcessor -> demodulator, any block size that is a multiple of 8."""
    u8 = S.noise(1, block * 7, seed=mode)[0]
    got = _run(["-d", str(mode), "-u", "-b", str(block)], u8)
    c = O.OracleChain()
    c.set_mode(mode)
    exp = np.concatenate([c.accept_u8(u8[o:o + block]) for o in range(0, u8.size, block)])
    assert np.array_equal(got, exp)


@pytest.mark.gpu
@pytest.mark.parametrize("first_quiet", [False, True])
def test_driver_squelch_and_signal_callbacks(first_quiet):
    """b200_demod -u -s T: IqDataProcessor::setSignalDetectThreshold plus the signal-state
    and signal-magnitude callbacks, block by block, against the oracle."""
    block = 32768
    rng = np.random.default_rng(3)
    amps = ([0.6] if first_quiet else []) + [40.0, 1.0, 0.7, 0.7, 25.0, 0.7, 0.7, 0.7, 50.0]
    u8 = np.concatenate([np.clip(np.round(128 + a * rng.standard_normal(block)), 0, 255).astype(np.uint8)
                         for a in amps])
    r = subprocess.run([_demod(), "-d", "2", "-u", "-s", "-25"], input=u8.tobytes(),
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(r.stdout, dtype=np.int16)
    c = O.OracleChain()
    c.set_mode(2)
    c.set_threshold(-25)
    exp, lines = [], []
    for o in range(0, u8.size, block):
        exp.append(c.accept_u8(u8[o:o + block]))
        a, m = c.signal()
        lines.append("signal %d magnitude %d" % (int(a), m))
    assert [l for l in r.stderr.decode().splitlines() if l.startswith("signal")] == lines
    assert any(e.size == 0 for e in exp) and any(e.size for e in exp)
    assert np.array_equal(got, np.concatenate(exp))


@pytest.mark.gpu
def test_driver_iq_dump_datagrams():
    """b200_demod -u -p PORT: enableIqDump; the datagrams the drop-in sends equal the ones the
    reference sends (tests/test_iq_dump.py pins the oracle to those), squelched or not."""
    import socket
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sock.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 22)
    sock.bind(("127.0.0.1", 0))
    sock.settimeout(5.0)
    try:
        block = 4096 + 64
        u8 = S.noise(1, block * 3, seed=8)[0]
        r = subprocess.run([_demod(), "-d", "3", "-u", "-b", str(block), "-p", str(sock.getsockname()[1]), "-s", "0"],
                           input=u8.tobytes(), capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        assert len(r.stdout) == 0   # threshold 0 dBFS: everything squelched
        for o in range(0, u8.size, block):
            sizes = O.dump_datagrams(block)
            got = [sock.recv(4096) for _ in sizes]
            assert [len(g) for g in got] == sizes
            assert np.array_equal(np.frombuffer(b"".join(got), dtype=np.int8), O.front_end(u8[o:o + block]))
    finally:
        sock.close()


def _helpers_exe():
    from rtlsdrdiags_b200 import _build
    _build.build()
    _build.build_host()
    exe = os.path.join(ROOT, "tests", "host", "iqdp_helpers")
    src = os.path.join(ROOT, "tests", "host", "iqdp_helpers_main.cc")
    pkg = os.path.join(ROOT, "rtlsdrdiags_b200")
    deps = [src, os.path.join(HOST, "IqDataProcessor.cc"), os.path.join(HOST, "B200Demodulator.cc"), _build.LIB]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++11", "-I", HOST, "-o", exe, src, os.path.join(HOST, "IqDataProcessor.cc"),
                        os.path.join(HOST, "B200Demodulator.cc"), "-L", pkg, "-lsdr_b200", "-Wl,-rpath," + pkg, "-lm"],
                       check=True)
    return exe


def test_public_fs4_helpers_match_the_front_end():
    """upconvertByFsOver4 / downconvertByFsOver4 (hdr_diags/IqDataProcessor.h:32-33) on the host:
    up is what the oracle's front end does after the offset, down is its inverse (int8 negation
    wraps, so -128 stays), and acceptIqData leaves the caller's buffer signed and translated as the
    reference does (IqDataProcessor.cc:735-749)."""
    exe = _helpers_exe()
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, size=4096, dtype=np.uint8)
    u8[:64] = 0
    s8 = (u8.astype(np.int16) - 128).astype(np.int8)

    def run(what, data):
        r = subprocess.run([exe, what], input=data.tobytes(), capture_output=True, timeout=60)
        assert r.returncode == 0, r.stderr.decode()
        return np.frombuffer(r.stdout, dtype=np.int8)

    up = run("up", s8)
    assert np.array_equal(up, O.front_end(u8))
    down = run("down", up)
    assert np.array_equal(down, s8)
    # mode None and the default squelch: no engine is needed, the buffer is converted all the same
    assert np.array_equal(run("accept", u8), O.front_end(u8))


def _filters_exe():
    from rtlsdrdiags_b200 import _build
    _build.build()
    exe = os.path.join(ROOT, "tests", "host", "filters_main")
    src = os.path.join(ROOT, "tests", "host", "filters_main.cc")
    pkg = os.path.join(ROOT, "rtlsdrdiags_b200")
    deps = [src, os.path.join(HOST, "B200Filters.h"), _build.LIB]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++11", "-I", HOST, "-o", exe, src, "-L", pkg, "-lsdr_b200",
                        "-Wl,-rpath," + pkg, "-lm"], check=True)
    return exe


@pytest.mark.gpu
@pytest.mark.parametrize("what,kind,factor,n_taps", [("dec32", O.MR_DECIMATOR_F32, 4, 80), ("dec16", O.MR_DECIMATOR_I16, 2, 40),
                                                    ("int32", O.MR_INTERPOLATOR_F32, 2, 64), ("int16", O.MR_INTERPOLATOR_I16, 4, 32),
                                                    ("fir32", O.MR_DECIMATOR_F32, 1, 7), ("fir16", O.MR_DECIMATOR_I16, 1, 16)])
@pytest.mark.parametrize("how", ["block", "sample"])
def test_filter_class_drop_ins(tmp_path, what, kind, factor, n_taps, how):
    """Decimator / Interpolator / FirFilter and their _int16 twins with the reference's class names
    (Filters/Decimator.h:28-42, Interpolator.h:35-49, FirFilter.h:21-27, Int16/*.h) over a one-row
    filter bank: block calls cut unevenly and sample-at-a-time calls against the oracle's classes."""
    exe = _filters_exe()
    rng = np.random.default_rng(n_taps)
    h = (rng.standard_normal(n_taps) / n_taps).astype(np.float32)
    tf = tmp_path / "taps.f32"
    h.tofile(tf)
    n = 600 if how == "block" else 96
    if what.endswith("16"):
        x = rng.integers(-20000, 20000, size=n).astype(np.int16)
    else:
        x = rng.standard_normal(n).astype(np.float32)
    r = subprocess.run([exe, what, str(factor), str(tf), how], input=x.tobytes(), capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(r.stdout, dtype=x.dtype)
    exp = O.Multirate(kind, h, factor).run(x)
    assert np.array_equal(got, exp)
