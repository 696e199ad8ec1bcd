#!/usr/bin/env python
"""Sustained AM loop (15,000 steps) a few times per setting, with and without an nvidia-smi poller:
how often does the pipeline fall out of its fast regime? Settings come from the environment
(SDR_RING, SDR_PACE, ...). Prints ms/step of every run."""
import os
import subprocess
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
iq = synth.make_bank("tone", modes, nbytes, 0xB200, dev)
out = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for sampler in (False, True):
        eng = R.Engine(channels, 0, nbytes)
        eng.set_modes(modes.numpy())
        stream = torch.cuda.Stream(dev)
        eng.set_stream(stream.cuda_stream)
        for _ in range(5):
            eng.accept_iq_device(iq)
        eng.join()
        torch.cuda.synchronize()
        p = None
        if sampler:
            p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw",
                                  "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(15000):
            eng.accept_iq_device(iq)
        eng.join()
        b.record(stream)
        torch.cuda.synchronize()
        if p:
            p.terminate()
            p.wait()
        out.append("%s%.4f" % ("s" if sampler else "-", a.elapsed_time(b) / 15000))
        eng.close()
print(" ".join(out), flush=True)
