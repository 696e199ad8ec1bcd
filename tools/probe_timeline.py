#!/usr/bin/env python
"""Launch timeline of the AM step (SDR_TRACE=1): when do the FIR kernel and the recurrence kernel
of consecutive calls start and end? Runs the bench loop for N steps with and without bench.py's
nvidia-smi sampler and prints statistics of the last 4096 calls plus a dozen calls in full."""
import ctypes as C
import os
import subprocess
import sys
import time

os.environ["SDR_TRACE"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtlsdrdiags_b200 as R  # noqa: E402
from rtlsdrdiags_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
channels, nbytes = 1024, 16 * R.BLOCK_BYTES
modes = synth.modes_for("am", channels, first_channel=0)
iq = synth.make_bank("tone", modes, nbytes, 0xB200, dev)
L = R.load_library()
L.sdr_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]


def run(steps, sampler, tma=4, seg=0, warm=28):
    eng = R.Engine(channels, 0, nbytes)
    eng.set_modes(modes.numpy())
    eng.debug_set_tile_loader(tma)
    eng.debug_set_dc_shape(seg, warm)
    print("loader %s, recurrence segments %s, warm-up rows %d" % ("TMA x%d, stage 1 on %s" % (tma & 7, "tensor cores" if tma & 8 else "CUDA cores") if tma & 7 else "cp.async", seg or "auto", warm))
    stream = torch.cuda.Stream(dev)
    eng.set_stream(stream.cuda_stream)
    for _ in range(5):
        eng.accept_iq_device(iq)
    eng.join()
    torch.cuda.synchronize()
    p = None
    if sampler:
        p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        eng.accept_iq_device(iq)
    eng.join()
    b.record(stream)
    torch.cuda.synchronize()
    if p:
        p.terminate()
        p.wait()
    ms = a.elapsed_time(b) / steps
    n = min(steps, 4096)
    tr = np.zeros((4096, 4), dtype=np.uint64)
    assert L.sdr_debug_read_trace(eng.h, tr.ctypes.data_as(C.c_void_p), 4096) == 0
    total = steps + 5
    order = [(total - n + i) % 4096 for i in range(n)]  # oldest .. newest of the last n calls
    t = tr[order].astype(np.int64)
    t -= t[0, 0]
    fir_s, fir_e, dc_s, dc_e = t[:, 0] / 1e3, t[:, 1] / 1e3, t[:, 2] / 1e3, t[:, 3] / 1e3
    step = np.diff(fir_s)
    print("steps %d sampler %s: %.4f ms/step | FIR %.1f us, dc_block %.1f us, FIR start to next FIR start %.1f (min %.1f max %.1f)"
          % (steps, sampler, ms, (fir_e - fir_s).mean(), (dc_e - dc_s).mean(), step.mean(), step.min(), step.max()))
    print("   dc_block(k) starts %.1f us after FIR(k) ends (min %.1f max %.1f); FIR(k+1) starts %.1f us after FIR(k) ends; "
          "dc_block(k) ends %.1f us after FIR(k+1) starts"
          % ((dc_s - fir_e).mean(), (dc_s - fir_e).min(), (dc_s - fir_e).max(), (fir_s[1:] - fir_e[:-1]).mean(),
             (dc_e[:-1] - fir_s[1:]).mean()))
    k0 = n // 2
    for k in range(k0, k0 + (10 if os.environ.get('SDR_TIMELINE_ROWS') else 0)):
        print("   call %d: FIR %9.1f .. %9.1f   dc %9.1f .. %9.1f" % (k, fir_s[k] - fir_s[k0], fir_e[k] - fir_s[k0],
                                                                      dc_s[k] - fir_s[k0], dc_e[k] - fir_s[k0]))
    eng.close()
    return ms


if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for tma, seg, warm in ((0, 0, 28), (2, 0, 28), (3, 0, 28), (4, 0, 28), (10, 0, 28), (12, 0, 28),
                           (2, 4, 28), (2, 16, 28), (2, 8, 24)):
        run(2000, False, tma, seg, warm)
        time.sleep(0.3)
else:
    for steps, sampler in ((3000, False), (3000, True), (3000, False)):
        run(steps, sampler)
        time.sleep(0.5)
