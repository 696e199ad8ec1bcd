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

and tests/golden/golden_squelch_v1.npz (the squelch gate, Squelch.cc:227-273):
  blocks           [12][4096] u8   noise blocks of scheduled amplitude (one radio, FM mode)
  cfg              [7][2] i32      (threshold dBFS, tuner gain dB) per configuration
  allowed          [7][12] u8      what the signal-state callback was handed per block
  magnitude        [7][12] u32     what the signal-magnitude callback was handed per block
  counts           [7][12] u32     PCM samples that came out per block
  pcm_<i>          i16             the concatenated PCM of configuration i
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


SQUELCH_CFG = [(-200, 0), (-30, 0), (-20, 0), (-12, 3), (-6, 0), (-25, 10), (0, 0)]
SQUELCH_AMPS = [0.5, 70.0, 1.0, 0.5, 10.0, 33.0, 0.6, 0.6, 100.0, 3.0, 127.0, 0.0]


def squelch_golden():
    rng = np.random.default_rng(77)
    blocks = np.stack([np.clip(np.round(128 + a * rng.standard_normal(4096)), 0, 255).astype(np.uint8)
                       for a in SQUELCH_AMPS])
    out = {"blocks": blocks, "cfg": np.array(SQUELCH_CFG, dtype=np.int32)}
    allowed = np.zeros((len(SQUELCH_CFG), len(blocks)), dtype=np.uint8)
    magnitude = np.zeros(allowed.shape, dtype=np.uint32)
    counts = np.zeros(allowed.shape, dtype=np.uint32)
    for i, (thr, gain) in enumerate(SQUELCH_CFG):
        r = O.RefChain()
        r.set_mode(2)
        r.set_threshold(thr)
        r.set_rx_gain(gain)
        pcm = []
        for b, blk in enumerate(blocks):
            pcm.append(r.accept_u8(blk, block=4096))
            a, m = r.signal()
            allowed[i, b], magnitude[i, b], counts[i, b] = a, m, pcm[-1].size
        r.set_rx_gain(0)
        out["pcm_%d" % i] = np.concatenate(pcm)
    out.update(allowed=allowed, magnitude=magnitude, counts=counts)
    path = os.path.join(HERE, "golden_squelch_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


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
    if "--squelch-only" not in sys.argv:
        main()
    squelch_golden()
