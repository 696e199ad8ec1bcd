#!/usr/bin/env python
"""Generates tests/golden/golden_capture_v1.npz: a 64 KiB excerpt of the only IQ capture the
reference ships (demodulatorResearch/yoyo.iq: signed, Fs/4-rotated int8 IQ, the format demod.cc
reads, demod.cc:8-11) and what the UNMODIFIED reference compiled in place (oracle/_ref) makes of it:

  iq_s8                  [65536] i8    bytes [OFFSET, OFFSET + 65536) of yoyo.iq
  research_<mode>        [1024] i16    demodulatorResearch tree, the demodulator classes fed the
                                       signed bytes in 16384-byte reads like demod.cc:250
                                       (-d 4 and -d 5 through set{Lsb,Usb}DemodulationMode)
  radiodiags_<mode>      [1024] i16    radioDiags tree: the excerpt un-rotated and re-offset to the
                                       dongle's u8 format, through IqDataProcessor::acceptIqData in
                                       32768-byte blocks
  offset                               OFFSET

BASELINE.json's config 1 names demodulatorResearch/f135_4.iq, which is absent from the reference
mount (.MISSING_LARGE_BLOBS); this excerpt is the stand-in that can travel to the GPU box.

Run in the build container, where /root/reference is mounted:
    make -C oracle && python tests/golden/make_golden_capture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _oracle as O  # noqa: E402
from make_golden import NAMES, unrotate_to_u8  # noqa: E402

OFFSET = 640 * 1024   # well inside the transmission
LENGTH = 64 * 1024


def main():
    assert O.ref("radiodiags") is not None and O.ref("research") is not None, "build oracle/_ref first (make -C oracle)"
    yoyo = np.fromfile("/root/reference/demodulatorResearch/yoyo.iq", dtype=np.int8)
    s8 = yoyo[OFFSET:OFFSET + LENGTH].copy()
    out = {"iq_s8": s8, "offset": np.array(OFFSET)}
    for m, name in NAMES.items():
        d = O.RefDemod(O.MODE_TO_KIND[m], "research")
        if m in (4, 5):
            d.set_lsb(m == 4)
        out["research_" + name] = d.accept(s8, block=16384)
        r = O.RefChain()
        r.set_mode(m)
        out["radiodiags_" + name] = r.accept_u8(unrotate_to_u8(s8))
    path = os.path.join(HERE, "golden_capture_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in sorted(out):
        if k != "iq_s8" and k != "offset":
            print(k, out[k].shape, int(np.abs(out[k].astype(int)).max()))


if __name__ == "__main__":
    main()
