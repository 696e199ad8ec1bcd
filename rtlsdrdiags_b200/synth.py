"""Synthetic IQ banks of the shapes BASELINE.json names, built
with torch so they can be generated directly in HBM. Plumbing, not product.

Every generator returns a uint8 tensor [n_channels][n_bytes] of offset-binary
interleaved I,Q -- the input of IqDataProcessor::acceptIqData.
"""
import math

import torch

FS = 256000.0
MODE_NONE, MODE_AM, MODE_FM, MODE_WBFM, MODE_LSB, MODE_USB = range(6)

WORKLOADS = {
    # name: (channels per GPU, mode pattern, BASELINE.json config it stands for)
    "am": (1024, [MODE_AM], "AM envelope demod, batch of 1024 synthetic 256 kS/s IQ channels on 1xB200"),
    "fm": (8192, [MODE_FM], "NBFM, 8192 channels per GPU (north-star real-time target mode)"),
    "wbfm": (8192, [MODE_WBFM], "WBFM broadcast (discriminator + de-emphasis), 8192 channels on 1xB200"),
    "ssb": (8192, [MODE_LSB, MODE_USB], "LSB/USB SSB demod with the phasing network and sideband selection (BASELINE config 4: 16384 channels across 2/4 B200)"),
    "mixed": (8192, [MODE_AM, MODE_FM, MODE_WBFM, MODE_LSB, MODE_USB],
              "mixed-mode AM/FM/WBFM/LSB/USB bank (BASELINE config 5: 65536 streams across 8 B200)"),
}


def modes_for(workload, n_channels, first_channel=0):
    pattern = WORKLOADS[workload][1]
    idx = torch.arange(first_channel, first_channel + n_channels)
    return torch.tensor(pattern, dtype=torch.uint8)[idx % len(pattern)]


def noise_bank(n_channels, n_bytes, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 256, (n_channels, n_bytes), dtype=torch.uint8, device=device, generator=g)


def tone_bank(modes, n_bytes, seed, device, chunk=256):
    """A modulated carrier per channel at -Fs/4: AM 1 kHz m=0.5; NBFM 1 kHz +-5 kHz;
    WBFM 1 kHz +-75 kHz; SSB 1.5 kHz tone in the wanted sideband; amplitude ~0.52 FS
    plus N(0, 2^2) noise, rounded and clipped to u8."""
    n_ch = int(modes.numel())
    n = n_bytes // 2
    out = torch.empty((n_ch, n_bytes), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64, device=device) / FS
    w1k = (2 * math.pi * 1000.0 * t)
    carrier = -2 * math.pi * (FS / 4) * t
    modes = modes.to(device)
    for c0 in range(0, n_ch, chunk):
        m = modes[c0:c0 + chunk].view(-1, 1)
        k = m.shape[0]
        # a small per-channel audio-phase offset keeps channels distinct
        ph0 = (torch.arange(c0, c0 + k, dtype=torch.float64, device=device) * 0.37).view(-1, 1)
        s1k = torch.sin(w1k.view(1, -1) + ph0)
        env = torch.where(m == MODE_AM, 1.0 + 0.5 * s1k, torch.ones_like(s1k))
        dev = torch.where(m == MODE_FM, 5.0, torch.where(m == MODE_WBFM, 75.0, 0.0)).to(torch.float64)
        ph = dev * s1k
        ssb = torch.where(m == MODE_LSB, -1.0, torch.where(m == MODE_USB, 1.0, 0.0)).to(torch.float64)
        ph = ph + ssb * (2 * math.pi * 1500.0 * t).view(1, -1) + carrier.view(1, -1)
        amp = (100.0 / 1.5) * env
        i = 128.0 + amp * torch.cos(ph)
        q = 128.0 + amp * torch.sin(ph)
        i = i + 2.0 * torch.randn(i.shape, dtype=torch.float64, device=device, generator=g)
        q = q + 2.0 * torch.randn(q.shape, dtype=torch.float64, device=device, generator=g)
        iq = torch.stack([i, q], dim=2).round_().clamp_(0, 255).to(torch.uint8)
        out[c0:c0 + k] = iq.view(k, n_bytes)
    return out


def make_bank(signal, modes, n_bytes, seed, device):
    if signal == "noise":
        return noise_bank(int(modes.numel()), n_bytes, seed, device)
    if signal == "tone":
        return tone_bank(modes, n_bytes, seed, device)
    raise ValueError("signal must be 'noise' or 'tone'")
