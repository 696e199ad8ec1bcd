#!/usr/bin/env python
"""How fast do two trajectories of the reference's one-pole recurrences become bit-identical?

    y[n] = fl(d[n] - fl(a1 * y[n-1]))       (IirFilter.cc:161-176; a1 = -0.95 for the AM / SSB
                                             DC-removal filter, -0.9492274 for WBFM de-emphasis)

Two copies fed the same numerators from different start states differ by |a1|^n times the initial
difference until that drops below the rounding step; a few steps later they are equal bit for bit
and stay so. dc_block_kernel relies on it: a segment warms up from y = 0 and is accepted only if
its state at the boundary equals its predecessor's final state, so this script sizes the warm-up
(the verification makes a miss cost time, never correctness). CPU only, numpy float32.
Prints, per input class, the merge step's median / 99.9 % / maximum over the trials.
"""
import sys

import numpy as np


def merge_steps(rng, trials, amp, y0_scale, a1, steps=2000, kind="int"):
    a1 = np.float32(a1)
    if kind == "int":      # integer-valued magnitude, iid: a noisy channel
        x = rng.integers(0, amp + 1, size=(steps + 1, trials)).astype(np.float32)
    elif kind == "tone":   # slowly varying magnitude plus a little noise: a modulated carrier
        t = np.arange(steps + 1)[:, None]
        ph = rng.uniform(0, 2 * np.pi, size=(1, trials))
        x = np.round(amp * (1 + 0.5 * np.sin(2 * np.pi * t / 8 + ph)) + rng.normal(0, 1, size=(steps + 1, trials)))
        x = x.astype(np.float32)
    else:                  # float numerators of the WBFM de-emphasis filter's size
        x = np.cumsum(rng.normal(0, amp, size=(steps + 1, trials)), axis=0).astype(np.float32)
    d = (x[1:] - x[:-1]).astype(np.float32)
    ya = rng.normal(0, y0_scale, size=trials).astype(np.float32)
    yb = np.zeros(trials, np.float32)
    merged = np.full(trials, -1)
    for n in range(steps):
        ya = (d[n] - (a1 * ya).astype(np.float32)).astype(np.float32)
        yb = (d[n] - (a1 * yb).astype(np.float32)).astype(np.float32)
        eq = ya.view(np.uint32) == yb.view(np.uint32)
        merged = np.where((merged < 0) & eq, n, merged)
        merged = np.where(~eq, -1, merged)  # (never observed: once equal they stay equal)
    return merged


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    rng = np.random.default_rng(1)
    cases = [("AM/SSB noise, |x| <= 190", "int", 190, 2000.0, -0.95),
             ("AM/SSB, |x| <= 20", "int", 20, 2000.0, -0.95),
             ("AM/SSB, |x| <= 3", "int", 3, 2000.0, -0.95),
             ("AM tone", "tone", 60, 2000.0, -0.95),
             ("AM/SSB noise, start 1e6 away", "int", 190, 1e6, -0.95),
             ("constant input (d = 0): y sticks on a denormal, never merges", "int", 0, 100.0, -0.95),
             ("WBFM de-emphasis", "float", 300.0, 30000.0, -0.9492274)]
    for name, kind, amp, y0, a1 in cases:
        m = merge_steps(rng, trials, amp, y0, a1, kind=kind)
        ok = m[m >= 0]
        if ok.size:
            print("%-62s median %4d  99.9%% %4d  max %4d  never %d" % (name, np.median(ok), np.percentile(ok, 99.9),
                                                                      ok.max(), (m < 0).sum()))
        else:
            print("%-62s never merges (%d trials)" % (name, trials))


if __name__ == "__main__":
    main()
