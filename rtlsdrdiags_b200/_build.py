"""Builds rtlsdrdiags_b200/libsdr_b200.so (CUDA kernels + engine + C ABI) in-tree
with nvcc for sm_100a. nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsdr_b200.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(ROOT, "include", "sdr_b200.h")]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, defines=(), verbose=False, out=None):
    """Compile the shared library if it is missing or older than its sources. `out`: another file
    name for an A/B build with `defines` (load it with SDR_B200_LIB=<path>)."""
    if not force and not defines and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D%s" % d for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out or LIB, os.path.join(CSRC, "sdr_engine.cu"), os.path.join(CSRC, "sdr_filter_bank.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        print(r.stderr)
    return out or LIB


HOST = os.path.join(PKG, "host")
DEMOD = os.path.join(PKG, "b200_demod")
BANK = os.path.join(PKG, "b200_bank")
H2D_PROBE = os.path.join(ROOT, "tools", "h2d_probe")


def build_host(force=False):
    """Compile the C++ drop-in classes and the offline driver (g++, links libsdr_b200.so)."""
    srcs = [os.path.join(HOST, f) for f in ("B200Demodulator.cc", "IqDataProcessor.cc", "demod_main.cc")]
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [LIB]
    if not force and os.path.exists(DEMOD) and all(os.path.getmtime(d) <= os.path.getmtime(DEMOD) for d in deps):
        return DEMOD
    cmd = ["g++", "-O2", "-std=c++11", "-Wall", "-I", HOST, "-o", DEMOD] + srcs + \
          ["-L", PKG, "-lsdr_b200", "-Wl,-rpath,$ORIGIN", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    # the multi-GPU C++ driver (one host thread, sdr_bank_*) and the host->device copy probe
    cmd = ["g++", "-O2", "-std=c++11", "-Wall", "-o", BANK, os.path.join(HOST, "bank_demo.cc"),
           "-L", PKG, "-lsdr_b200", "-Wl,-rpath,$ORIGIN", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    cmd = [_nvcc(), "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-o", H2D_PROBE, os.path.join(ROOT, "tools", "h2d_probe.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    return DEMOD


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
