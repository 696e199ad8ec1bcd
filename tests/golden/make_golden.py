#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz with the UNMODIFIED reference compiled in
place (oracle/_ref, built by oracle/Makefile from /root/reference).

Run in the build container, where /root/reference is mounted:
    make -C oracle && python tests/golden/make_golden.py

Contents (all inputs are synthetic; no reference data file is copied):
  iq_u8            [6][98304] u8   three 32768-byte blocks per channel:
                                   ch0 noise + quirk runs (0x00 / 0xFF / alternating),
                                   ch1..5 a modulated carrier for AM, FM, WBFM, LSB, USB
  pcm_<mode>       [6][1536] i16   radioDiags tree: IqDataProcessor -> demodulator,
                                   every channel through mode 1..5
  iq_s8            [2][65536] i8   signed, rotated IQ (the demodulator classes' own input)
  research_<mode>  [2][1024] i16   demodulatorResearch tree, mode 1..5 (4 and 5 via
                                   set{Lsb,Usb}DemodulationMode)
  yoyo_md5_*                       md5 of the PCM of demodulatorResearch/yoyo.iq (the only
                                   capture shipped with the reference) for both trees
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402
import _signals as S  # noqa: E402

NAMES = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}


def main():
    assert O.ref("radiodiags") is not None, "build oracle/_ref first (make -C oracle)"
    nbytes = 3 * 32768
    iq = np.zeros((6, nbytes), dtype=np.uint8)
    iq[0] = S.noise(1, nbytes, seed=20261017)[0]
    iq[0, 0:2048] = 0x00
    iq[0, 2048:4096] = 0xFF
    iq[0, 4096:6144:2] = 0x00
    iq[0, 4097:6144:2] = 0xFF
    for ch, m in enumerate([1, 2, 3, 4, 5], start=1):
        iq[ch] = S.tone(m, nbytes // 2, seed=ch)
    out = {"iq_u8": iq}
    for m, name in NAMES.items():
        rows = []
        for ch in range(6):
            r = O.RefChain()
            r.set_mode(m)
            rows.append(r.accept_u8(iq[ch]))
        out["pcm_" + name] = np.stack(rows)

    s8 = np.random.default_rng(424242).integers(-128, 128, size=(2, 65536), dtype=np.int8)
    s8[1] = (S.tone(2, 32768, seed=9).astype(np.int16) - 128).astype(np.int8)
    out["iq_s8"] = s8
    for m, name in NAMES.items():
        rows = []
        for ch in range(2):
            d = O.RefDemod(O.MODE_TO_KIND[m], "research")
            if m in (4, 5):
                d.set_lsb(m == 4)
            rows.append(d.accept(s8[ch], block=16384))
        out["research_" + name] = np.stack(rows)

    yoyo_path = "/root/reference/demodulatorResearch/yoyo.iq"
    yoyo = np.fromfile(yoyo_path, dtype=np.int8)
    out["yoyo_md5_input"] = np.array(hashlib.md5(yoyo.tobytes()).hexdigest())
    # research tree, signed input as shipped (demod.cc reads 16384-byte pieces)
    for m, name in NAMES.items():
        d = O.RefDemod(O.MODE_TO_KIND[m], "research")
        if m in (4, 5):
            d.set_lsb(m == 4)
        out["yoyo_md5_research_" + name] = np.array(hashlib.md5(d.accept(yoyo, block=16384).tobytes()).hexdigest())
    # radioDiags tree: undo the rotation and the offset so IqDataProcessor redoes them
    u8 = unrotate_to_u8(yoyo)
    for m, name in NAMES.items():
        r = O.RefChain()
        r.set_mode(m)
        out["yoyo_md5_radiodiags_" + name] = np.array(hashlib.md5(r.accept_u8(u8).tobytes()).hexdigest())
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in sorted(out):
        if k.startswith("yoyo"):
            print(k, out[k])


def unrotate_to_u8(signed_rotated):
    """Inverse of IqDataProcessor's offset + Fs/4 rotation (values that would need +128 wrap)."""
    s = signed_rotated.astype(np.int16).reshape(-1, 4, 2)
    u = s.copy()
    u[:, 1, 0] = s[:, 1, 1]
    u[:, 1, 1] = -s[:, 1, 0]
    u[:, 2, 0] = -s[:, 2, 0]
    u[:, 2, 1] = -s[:, 2, 1]
    u[:, 3, 0] = -s[:, 3, 1]
    u[:, 3, 1] = s[:, 3, 0]
    return ((u.reshape(-1) + 128) & 0xFF).astype(np.uint8)


if __name__ == "__main__":
    main()
