#!/usr/bin/env python
"""How long do the first steps after an idle gap take? Per-step device times of a burst of AM steps
after warm-up + synchronize (what bench.py's timed region starts with)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtlsdrdiags_b200 as R  # noqa: E402
from rtlsdrdiags_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
channels, nbytes = 1024, 16 * R.BLOCK_BYTES
modes = synth.modes_for("am", channels, first_channel=0)
eng = R.Engine(channels, 0, nbytes)
eng.set_modes(modes.numpy())
stream = torch.cuda.Stream(dev)
eng.set_stream(stream.cuda_stream)
iq = synth.make_bank("tone", modes, nbytes, 0xB200, dev)
for _ in range(5):
    eng.accept_iq_device(iq)
eng.join()
torch.cuda.synchronize()
for idle in (0.0, 0.001, 0.05, 1.0):
    time.sleep(idle)
    n = 60
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record(stream)
    host = []
    for i in range(n):
        eng.accept_iq_device(iq)
        ev[i + 1].record(stream)
        host.append(time.perf_counter() - t0)
    eng.join()
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print("idle %.3f s: first steps ms %s ... steps 20-59 mean %.4f; total %.3f ms; host enqueue of 60 steps took %.3f ms (first call %.3f ms)" % (
        idle, " ".join("%.3f" % x for x in per[:12]), sum(per[20:]) / 40, sum(per), host[-1] * 1e3, host[0] * 1e3), flush=True)

# ---- the time profile of a long burst: mean ms per step over groups of 20 steps, with the SM clock NVML reports ----
import threading
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for label, grp, n_grp in (("burst of 4000 steps", 20, 200),):
    torch.cuda.synchronize()
    time.sleep(0.5)
    clk, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            clk.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.005)
    th = threading.Thread(target=poll)
    th.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_grp + 1)]
    t0 = time.perf_counter()
    ev[0].record(stream)
    for g in range(n_grp):
        for _ in range(grp):
            eng.accept_iq_device(iq)
        ev[g + 1].record(stream)
    eng.join()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    per = [ev[g].elapsed_time(ev[g + 1]) / grp for g in range(n_grp)]
    print(label, "ms/step per group of %d:" % grp)
    print(" ".join("%.3f" % x for x in per))
    print("sm MHz / W every 5 ms:", " ".join("%d/%d" % (c, p) for (_, c, p) in clk[:120]))
