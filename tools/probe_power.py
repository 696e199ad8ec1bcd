#!/usr/bin/env python
"""Why does a long AM run go faster while nvidia-smi polls the GPU? Runs the same 15000-step
timed region (a) unobserved, (b) with an in-process NVML poll every 100 ms, (c) with NVML polled
once a second, and prints ms/step with the SM clock / power NVML saw."""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtlsdrdiags_b200 as R  # noqa: E402
from rtlsdrdiags_b200 import synth  # noqa: E402
import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
channels, nbytes = 1024, 16 * R.BLOCK_BYTES
modes = synth.modes_for("am", channels, first_channel=0)
eng = R.Engine(channels, 0, nbytes)
eng.set_modes(modes.numpy())
stream = torch.cuda.Stream(dev)
eng.set_stream(stream.cuda_stream)
iq = synth.make_bank("tone", modes, nbytes, 0xB200, dev)


def timed(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        eng.accept_iq_device(iq)
    eng.join()
    b.record(stream)
    b.synchronize()
    return a.elapsed_time(b) / n


def poll(period, stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(period)


timed(50)
for label, period in (("unobserved", None), ("nvml 100 ms", 0.1), ("nvml 1 s", 1.0), ("unobserved", None), ("nvml 100 ms", 0.1)):
    stop, out = threading.Event(), []
    t = None
    if period:
        t = threading.Thread(target=poll, args=(period, stop, out))
        t.start()
    t0 = time.time()
    ms = timed(15000)
    wall = time.time() - t0
    stop.set()
    if t:
        t.join()
    after = (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
    clk = sorted(c for c, _ in out)
    pw = sorted(p for _, p in out)
    print("%-12s ms/step %.4f  wall %.2f s  sm MHz median %s  power W median %s  right after: %s" % (
        label, ms, wall, clk[len(clk) // 2] if clk else None, pw[len(pw) // 2] if pw else None, after), flush=True)
