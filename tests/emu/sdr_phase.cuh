// TEST HARNESS -- phase-structured kernels (the engine does not launch them).
//
// The first working version of the path: one CTA owns a group of channels and runs
// every reference stage as a "phase" over shared-memory arrays, phases separated by a
// CTA barrier. Because a phase is a plain function of (thread id, shared memory), the
// whole pipeline can be executed serially on the host (emu.cc), which is how the
// arithmetic building blocks in csrc/sdr_device.cuh -- shared with the warp-tile
// kernels the engine does launch -- are checked against the oracle without a GPU.
// On a B200 these kernels reached 0.11 (AM), 0.15 (FM), 0.03 (WBFM) of the HBM roofline
// (profiles/r01v1_*); the warp-tile kernels replaced them.
#pragma once
#include "sdr_device.cuh"

namespace sdr {

// Per-thread view of the CTA. Holds no state that must survive a phase.
struct Ctx {
  const LaunchParams *p;
  char *hdr;   // CTA header: channel ids, scales, flags
  char *smem;  // first channel's region
  int tid, nt;
  int Gc;            // channels this CTA owns
  uint32_t sample0;  // first sample of the current sub-block
  int cur;           // samples in the current sub-block (multiple of 32)
  int parity;
};

// header layout
constexpr int HDR_CHAN = 0, HDR_SCALE = 128, HDR_LSB = 256, HDR_FLAG_A = 384, HDR_FLAG_B = 512;
constexpr int HDR_BYTES = 640;

SDR_DEV uint32_t hdr_chan(const Ctx &t, int c) { return lds<uint32_t>(t.hdr + HDR_CHAN + 4 * c); }
SDR_DEV float hdr_scale(const Ctx &t, int c) { return lds<float>(t.hdr + HDR_SCALE + 4 * c); }
SDR_DEV uint32_t hdr_u32(const Ctx &t, int off, int c) { return lds<uint32_t>(t.hdr + off + 4 * c); }
SDR_DEV void hdr_set(const Ctx &t, int off, int c, uint32_t v) { sts<uint32_t>(t.hdr + off + 4 * c, v); }

SDR_HD constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }
// per-channel stride == 16 (mod 128): 128-bit lane-per-channel accesses are
// bank-conflict free, and every channel base stays 16-byte aligned.
SDR_HD constexpr int channel_stride(int end) { return round_up(end, 128) + 16; }

#define SDR_FOR_ITEMS(t, ITEMS, c, j)                                                 \
  for (int _idx = (t).tid, _tot = (t).Gc * (ITEMS); _idx < _tot; _idx += (t).nt)      \
    if (int c = _idx / (ITEMS), j = _idx % (ITEMS); true)

// 8 complex samples (16 input bytes) per item -> 8 bytes into each plane.
template <int S, int HX>
SDR_DEV void phase_front_end(const Ctx &t, int XI, int XQ, int STRIDE) {
  constexpr int ITEMS = S / 8;
  const int lim = t.cur >> 3;
  const int fmt = t.p->fmt;
  SDR_FOR_ITEMS(t, ITEMS, c, j) {
    if (j >= lim) continue;
    const uint8_t *src = t.p->iq + (uint64_t)hdr_chan(t, c) * t.p->ch_stride +
                         ((uint64_t)t.sample0 + (uint64_t)j * 8) * 2;
    u32x4 w = ld_stream_u4(src);
    u32x2 a, b;
    front_end_group(fmt, w.x, w.y, a.x, b.x);
    front_end_group(fmt, w.z, w.w, a.y, b.y);
    char *cb = t.smem + c * STRIDE;
    sts_u2(cb + XI + HX + j * 8, a);
    sts_u2(cb + XQ + HX + j * 8, b);
  }
}

// ---------------------------------------------------------------------------
// History handling. Every array is [H bytes of history][data]; after a
// sub-block that produced `n` data bytes the last H bytes move to the front.
// ---------------------------------------------------------------------------
template <int H>
SDR_DEV void shift_history(char *arr, int n) {
  if ((n & 15) == 0 && n >= H) {
#pragma unroll
    for (int i = 0; i < H; i += 16) sts_u4(arr + i, lds_u4(arr + n + i));
  } else {
    for (int i = 0; i < H; ++i) arr[i] = arr[n + i];  // ascending: safe for overlap
  }
}
// threads from the top of the CTA do the moves (warp 0 hosts the recurrence lanes)
#define SDR_FOR_CHANNELS_HI(t, c) for (int c = (t).nt - 1 - (t).tid; c < (t).Gc; c += (t).nt)

// Interleaved PCM copy-out: smem int16[cur/32] per channel -> global row.
template <int S>
SDR_DEV void phase_pcm_out(const Ctx &t, int PCM, int STRIDE) {
  constexpr int ITEMS = S / 64;  // one 32-bit word = two PCM samples
  const int n = t.cur >> 5;
  SDR_FOR_ITEMS(t, ITEMS, c, j) {
    if (2 * j >= n) continue;
    const char *cb = t.smem + c * STRIDE;
    int16_t *dst = t.p->pcm + (uint64_t)hdr_chan(t, c) * t.p->pcm_stride + (t.sample0 >> 5) + 2 * j;
    if (2 * j + 1 < n) {
      stg_u32(dst, lds<uint32_t>(cb + PCM + 4 * j));
    } else {
      *dst = lds<int16_t>(cb + PCM + 4 * j);
    }
  }
}

// Sequential one-pole recurrences, lane == channel, eight outputs per trip:
//   y[n] = fl( fl(x[n] - x[n-1]) - fl(-0.95f * y[n-1]) ),  pcm = (int16)(gain * y)
// (IirFilter.cc:161-176 with b = {1,-1}, a = {-0.95}; AmDemodulator.cc:461-467).
template <int S, bool FLOAT_IN>
SDR_DEV void phase_dc_block(const Ctx &t, int IN, int PCM, int IIR, int STRIDE) {
  if (t.tid >= t.Gc) return;
  char *cb = t.smem + t.tid * STRIDE;
  const float gain = hdr_scale(t, t.tid);
  const float a1 = (float)(-0.95);
  float x1 = lds<float>(cb + IIR), y1 = lds<float>(cb + IIR + 4);
  const int n = t.cur >> 5;
  for (int g = 0; g < n; g += 8) {
    float x[8];
    if constexpr (FLOAT_IN) {
      u32x4 a = lds_u4(cb + IN + 4 * g), b = lds_u4(cb + IN + 4 * g + 16);
      x[0] = u2f(a.x); x[1] = u2f(a.y); x[2] = u2f(a.z); x[3] = u2f(a.w);
      x[4] = u2f(b.x); x[5] = u2f(b.y); x[6] = u2f(b.z); x[7] = u2f(b.w);
    } else {
      u32x4 a = lds_u4(cb + IN + 2 * g);
      const uint32_t r[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = i2f((i & 1) ? ((int)r[i / 2] >> 16) : (int)(int16_t)(r[i / 2] & 0xffffu));
    }
    int o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (g + i < n) {
        float y = fsub(fadd(x[i], fmul(-1.0f, x1)), fmul(a1, y1));
        x1 = x[i];
        y1 = y;
        o[i] = f2i16_wrap(fmul(gain, y));
      } else {
        o[i] = 0;
      }
    }
    sts_u4(cb + PCM + 2 * g, u32x4{pack_i16x2(o[0], o[1]), pack_i16x2(o[2], o[3]),
                                   pack_i16x2(o[4], o[5]), pack_i16x2(o[6], o[7])});
  }
  sts<float>(cb + IIR, x1);
  sts<float>(cb + IIR + 4, y1);
}

// ===========================================================================
// AM and SSB: three Q15 decimators per arm (8/4:1, 12/4:1, 16/2:1), then
//   AM : max+min/2 magnitude estimate -> DC block -> gain   (AmDemodulator.cc)
//   SSB: -I delayed 15, Hilbert(Q), I -/+ Q -> DC block -> gain (SsbDemodulator.cc)
// Stage 1 and 2 outputs are bounded by |x| <= 114 and 122 for 8-bit input, so they
// are kept as int8 and every stage runs on IDP.2A with exact int16 taps.
// ===========================================================================
template <int S_, bool SSB>
struct AmSsb {
  static constexpr int S = S_;
  static_assert(S % 256 == 0, "sub-block must keep every history tail 16-byte aligned");
  static constexpr int HX = 16, H1 = 16, H2 = 16;  // bytes of history (4, 8, 14 used)
  static constexpr int H3I = 32, H3Q = 64;         // int16 histories: 15 and 30 used
  static constexpr int XI = 0;
  static constexpr int XQ = XI + HX + S;
  static constexpr int S1I = XQ + HX + S;
  static constexpr int S1Q = S1I + H1 + S / 4;
  static constexpr int S2I = S1Q + H1 + S / 4;
  static constexpr int S2Q = S2I + H2 + S / 16;
  static constexpr int S3I = S2Q + H2 + S / 16;                  // SSB: int16 [16 + S/32]
  static constexpr int S3Q = S3I + (SSB ? H3I + S / 16 : 0);     // SSB: int16 [32 + S/32]
  static constexpr int DEM = S3Q + (SSB ? H3Q + S / 16 : 0);     // AM: int16 mag, SSB: float
  static constexpr int PCM = DEM + (SSB ? S / 8 : S / 16);
  static constexpr int IIR = PCM + S / 16;
  static constexpr int END = IIR + 16;
  static constexpr int STRIDE = channel_stride(END);
  static constexpr int NPHASES = SSB ? 7 : 6;

  template <class Fn>
  SDR_DEVM static void for_each_persistent(Fn &&f) {
    f(XI, HX); f(XQ, HX); f(S1I, H1); f(S1Q, H1); f(S2I, H2); f(S2Q, H2);
    if constexpr (SSB) { f(S3I, H3I); f(S3Q, H3Q); }
    f(IIR, 16);
  }
  static constexpr int STATE_BYTES = HX * 2 + H1 * 2 + H2 * 2 + (SSB ? H3I + H3Q : 0) + 16;

  SDR_DEVM static void init_flags(const Ctx &, int) {}
  SDR_DEVM static void finalize(const Ctx &, int) {}

  // stage 1: y[m] = q15(sum h[k] x[4m+3-k]), 8 taps; four outputs per item
  SDR_DEVM static void stage1(const Ctx &t) {
    constexpr int ITEMS = S / 16 * 2;
    const int lim = t.cur >> 4;
    SDR_FOR_ITEMS(t, ITEMS, c, jj) {
      const int arm = jj & 1, j = jj >> 1;
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      const char *x = cb + (arm ? XQ : XI) + HX + 16 * j;
      const uint32_t w0 = lds<uint32_t>(x - 4);
      const u32x4 v = lds_u4(x);
      using F = taps::AM1;
      const uint32_t a[2] = {w0, v.x}, b[2] = {v.x, v.y}, cc[2] = {v.y, v.z}, d[2] = {v.z, v.w};
      const int y0 = fir_s8<F, 7, 2>(a) >> 15, y1 = fir_s8<F, 7, 2>(b) >> 15;
      const int y2 = fir_s8<F, 7, 2>(cc) >> 15, y3 = fir_s8<F, 7, 2>(d) >> 15;
      sts<uint32_t>(cb + (arm ? S1Q : S1I) + H1 + 4 * j, pack_i8x4(y0, y1, y2, y3));
    }
  }

  // stage 2: 12 taps, 4:1; four outputs per item
  SDR_DEVM static void stage2(const Ctx &t) {
    constexpr int ITEMS = S / 64 * 2;
    const int lim = (t.cur + 63) >> 6;
    SDR_FOR_ITEMS(t, ITEMS, c, jj) {
      const int arm = jj & 1, j = jj >> 1;
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      const char *x = cb + (arm ? S1Q : S1I) + H1 + 16 * j;
      const u32x2 u = lds_u2(x - 8);
      const u32x4 v = lds_u4(x);
      using F = taps::AM2;
      const uint32_t a[3] = {u.x, u.y, v.x}, b[3] = {u.y, v.x, v.y};
      const uint32_t cc[3] = {v.x, v.y, v.z}, d[3] = {v.y, v.z, v.w};
      const int y0 = fir_s8<F, 11, 3>(a) >> 15, y1 = fir_s8<F, 11, 3>(b) >> 15;
      const int y2 = fir_s8<F, 11, 3>(cc) >> 15, y3 = fir_s8<F, 11, 3>(d) >> 15;
      sts<uint32_t>(cb + (arm ? S2Q : S2I) + H2 + 4 * j, pack_i8x4(y0, y1, y2, y3));
    }
  }

  // stage 3: 16 taps, 2:1; outputs 2j and 2j+1 of both arms per item
  SDR_DEVM static void stage3(const Ctx &t) {
    constexpr int ITEMS = S / 64;
    const int n = t.cur >> 5;
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (2 * j >= n) continue;
      char *cb = t.smem + c * STRIDE;
      int y[2][2];
#pragma unroll
      for (int arm = 0; arm < 2; ++arm) {
        const char *x = cb + (arm ? S2Q : S2I) + H2 + 4 * j;  // x[4j]; window starts at x[4j-16]
        uint32_t w[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) w[i] = lds<uint32_t>(x - 16 + 4 * i);
        using F = taps::AM3;
        y[arm][0] = (int)(int16_t)(fir_s8<F, 17, 5>(w) >> 15);  // newest = x[4j+1]
        y[arm][1] = (int)(int16_t)(fir_s8<F, 19, 5>(w) >> 15);  // newest = x[4j+3]
      }
      if constexpr (SSB) {
        sts<uint32_t>(cb + S3I + H3I + 4 * j, pack_i16x2(y[0][0], y[0][1]));
        sts<uint32_t>(cb + S3Q + H3Q + 4 * j, pack_i16x2(y[1][0], y[1][1]));
      } else {
        int m[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) {  // AmDemodulator.cc:441-458, tie -> q branch
          const int im = (int)(int16_t)iabs(y[0][o]), qm = (int)(int16_t)iabs(y[1][o]);
          m[o] = (int)(int16_t)(im > qm ? im + (qm >> 1) : qm + (im >> 1));
        }
        sts<uint32_t>(cb + DEM + 4 * j, pack_i16x2(m[0], m[1]));
      }
    }
  }

  // SSB phasing network at 8 kS/s (SsbDemodulator.cc:569-590)
  SDR_DEVM static void phasing(const Ctx &t) {
    constexpr int ITEMS = S / 32;
    const int n = t.cur >> 5;
    SDR_FOR_ITEMS(t, ITEMS, c, m) {
      if (m >= n) continue;
      char *cb = t.smem + c * STRIDE;
      const int16_t *si = reinterpret_cast<const int16_t *>(cb + S3I + H3I) + m;
      const int16_t *sq = reinterpret_cast<const int16_t *>(cb + S3Q + H3Q) + m;
      // delay line: taps {0 x15, -32768}: acc = 16384 + (-32768) * x[m-15]
      int acc = (1 << 14) + taps::SSB_DELAY::tap(15) * (int)si[-15];
      acc = acc > 0x3fffffff ? 0x3fffffff : acc;
      acc = acc < -0x40000000 ? -0x40000000 : acc;
      const int iDelayed = (int)(int16_t)(acc >> 15);
      // Hilbert transformer: 31 taps, odd taps are zero; |x| <= 179 so no clamp can fire
      int h = 1 << 14;
#pragma unroll
      for (int k = 0; k < 31; k += 2) h += taps::SSB_HILBERT::tap(k) * (int)sq[-k];
      const int qShifted = (int)(int16_t)(h >> 15);
      const bool lsb = hdr_u32(t, HDR_LSB, c) != 0;
      sts<float>(cb + DEM + 4 * m, i2f(lsb ? iDelayed - qShifted : iDelayed + qShifted));
    }
  }

  template <int PH>
  SDR_DEVM static void phase(const Ctx &t) {
    if constexpr (PH == 0) {
      phase_front_end<S, HX>(t, XI, XQ, STRIDE);
    } else if constexpr (PH == 1) {
      stage1(t);
    } else if constexpr (PH == 2) {
      stage2(t);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<HX>(t.smem + c * STRIDE + XI, t.cur);
        shift_history<HX>(t.smem + c * STRIDE + XQ, t.cur);
      }
    } else if constexpr (PH == 3) {
      stage3(t);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<H1>(t.smem + c * STRIDE + S1I, t.cur >> 2);
        shift_history<H1>(t.smem + c * STRIDE + S1Q, t.cur >> 2);
      }
    } else if constexpr (!SSB && PH == 4) {
      phase_dc_block<S, false>(t, DEM, PCM, IIR, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<H2>(t.smem + c * STRIDE + S2I, t.cur >> 4);
        shift_history<H2>(t.smem + c * STRIDE + S2Q, t.cur >> 4);
      }
    } else if constexpr (SSB && PH == 4) {
      phasing(t);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<H2>(t.smem + c * STRIDE + S2I, t.cur >> 4);
        shift_history<H2>(t.smem + c * STRIDE + S2Q, t.cur >> 4);
      }
    } else if constexpr (SSB && PH == 5) {
      phase_dc_block<S, true>(t, DEM, PCM, IIR, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<H3I>(t.smem + c * STRIDE + S3I, t.cur >> 4);
        shift_history<H3Q>(t.smem + c * STRIDE + S3Q, t.cur >> 4);
      }
    } else {
      phase_pcm_out<S>(t, PCM, STRIDE);
    }
  }
};

// ===========================================================================
// Shared tail of FM and WBFM: 40-tap 2:1 audio decimator on int16 input,
// two outputs per item. Input array: int16 [40 history][data].
// ===========================================================================
constexpr int H_AUDIO = 80;  // bytes: 40 int16 of history, 38 used

template <int S>
SDR_DEV void phase_audio40(const Ctx &t, int IN, int PCM, int STRIDE) {
  constexpr int ITEMS = S / 64;
  const int n = t.cur >> 5;
  SDR_FOR_ITEMS(t, ITEMS, c, j) {
    if (2 * j >= n) continue;
    char *cb = t.smem + c * STRIDE;
    // outputs m = 2j, 2j+1 use x[4j-38 .. 4j+3]; data starts at element 40
    const char *x = cb + IN + 2 * (4 * j + 2);
    uint32_t w[21];
    w[0] = lds<uint32_t>(x);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const u32x2 v = lds_u2(x + 4 + 8 * i);
      w[1 + 2 * i] = v.x;
      w[2 + 2 * i] = v.y;
    }
    const bool exact = hdr_u32(t, HDR_FLAG_B, c) != 0;
    using F = taps::AUDIO40;
    const int y0 = (int)(int16_t)(fir_s16<F, 39, 21>(w, exact) >> 15);
    const int y1 = (int)(int16_t)(fir_s16<F, 41, 21>(w, exact) >> 15);
    sts<uint32_t>(cb + PCM + 4 * j, pack_i16x2(y0, y1));
  }
}

// 12-tap 4:1 post-demodulation decimator on int16 input (FM stage 2, WBFM stage 2),
// two outputs per item; raises FLAG_B when an output could make the audio
// decimator's clamp reachable.
template <int ITEMS, int HIN_BYTES, bool CAN_CLAMP>
SDR_DEV void phase_post12(const Ctx &t, int IN, int OUT, int STRIDE, int n_out) {
  SDR_FOR_ITEMS(t, ITEMS, c, j) {
    if (2 * j >= n_out) continue;
    char *cb = t.smem + c * STRIDE;
    // outputs m = 2j, 2j+1 use x[8j-8 .. 8j+7]
    const char *x = cb + IN + HIN_BYTES + 16 * j - 16;
    const u32x4 a = lds_u4(x), b = lds_u4(x + 16);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const bool exact = CAN_CLAMP && hdr_u32(t, HDR_FLAG_A, c) != 0;
    using F = taps::FM_POST;
    int y0, y1;
    if constexpr (CAN_CLAMP) {
      y0 = (int)(int16_t)(fir_s16<F, 11, 8>(w, exact) >> 15);
      y1 = (int)(int16_t)(fir_s16<F, 15, 8>(w, exact) >> 15);
    } else {
      y0 = (int)(int16_t)(fir_s16_fast<F, 11, 8>(w) >> 15);
      y1 = (int)(int16_t)(fir_s16_fast<F, 15, 8>(w) >> 15);
    }
    if (iabs(y0) > taps::AUDIO40::SAFE || iabs(y1) > taps::AUDIO40::SAFE) hdr_set(t, HDR_FLAG_B, c, 1);
    sts<uint32_t>(cb + OUT + H_AUDIO + 4 * j, pack_i16x2(y0, y1));
  }
}

// scan a history region for values that make a clamp reachable
SDR_DEV bool any_above(const char *arr, int n_i16, int limit) {
  bool hit = false;
  for (int i = 0; i < n_i16; ++i) hit |= iabs((int)lds<int16_t>(arr + 2 * i)) > limit;
  return hit;
}

// ===========================================================================
// Narrow-band FM (FmDemodulator.cc): I,Q 32-tap 4:1 -> theta = atan2 (table) ->
// theta[n-2] - theta[n-4] -> wrap -> * k -> int16 -> 12-tap 4:1 -> 40-tap 2:1.
// ===========================================================================
template <int S_>
struct Fm {
  static constexpr int S = S_;
  static_assert(S % 256 == 0, "");
  static constexpr int HX = 32;   // 28 used
  static constexpr int HT = 16;   // 4 floats of theta history
  static constexpr int HD = 16;   // 8 int16 of discriminator history
  static constexpr int XI = 0;
  static constexpr int XQ = XI + HX + S;
  static constexpr int TH = XQ + HX + S;           // float [4 + S/4]
  static constexpr int D = TH + HT + S;            // int16 [8 + S/4]
  static constexpr int E2 = D + HD + S / 2;        // int16 [40 + S/16]
  static constexpr int PCM = E2 + H_AUDIO + S / 8; // int16 [S/32]
  static constexpr int END = PCM + S / 16;
  static constexpr int STRIDE = channel_stride(END);
  static constexpr int NPHASES = 6;
  static constexpr int STATE_BYTES = HX * 2 + HT + HD + H_AUDIO;

  template <class Fn>
  SDR_DEVM static void for_each_persistent(Fn &&f) {
    f(XI, HX); f(XQ, HX); f(TH, HT); f(D, HD); f(E2, H_AUDIO);
  }
  SDR_DEVM static void init_flags(const Ctx &t, int c) {
    const char *cb = t.smem + c * STRIDE;
    hdr_set(t, HDR_FLAG_A, c, any_above(cb + D, 8, taps::FM_POST::SAFE));
    hdr_set(t, HDR_FLAG_B, c, any_above(cb + E2, 40, taps::AUDIO40::SAFE));
  }
  SDR_DEVM static void finalize(const Ctx &, int) {}

  // tuner decimators + atan2: four 64 kS/s samples per item
  SDR_DEVM static void tuner(const Ctx &t) {
    constexpr int ITEMS = S / 16;
    const int lim = t.cur >> 4;
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      int y[2][4];
#pragma unroll
      for (int arm = 0; arm < 2; ++arm) {
        const char *x = cb + (arm ? XQ : XI) + HX + 16 * j;  // window x[16j-28 .. 16j+15]
        uint32_t w[11];
        w[0] = lds<uint32_t>(x - 28);
        w[1] = lds<uint32_t>(x - 24);
        w[2] = lds<uint32_t>(x - 20);
        const u32x4 a = lds_u4(x - 16), b = lds_u4(x);
        w[3] = a.x; w[4] = a.y; w[5] = a.z; w[6] = a.w;
        w[7] = b.x; w[8] = b.y; w[9] = b.z; w[10] = b.w;
        using F = taps::FM_TUNER;
        y[arm][0] = (int)(int16_t)(fir_s8<F, 31, 11>(w) >> 15);
        y[arm][1] = (int)(int16_t)(fir_s8<F, 35, 11>(w) >> 15);
        y[arm][2] = (int)(int16_t)(fir_s8<F, 39, 11>(w) >> 15);
        y[arm][3] = (int)(int16_t)(fir_s8<F, 43, 11>(w) >> 15);
      }
      float th[4];
#pragma unroll
      for (int o = 0; o < 4; ++o)  // theta = (float)atan2((double)q,(double)i), FmDemodulator.cc:476
        th[o] = ld_lut(t.p->lut + (y[1][o] - FM_LUT_MIN) * FM_LUT_DIM + (y[0][o] - FM_LUT_MIN));
      sts_u4(cb + TH + HT + 16 * j, u32x4{f2u(th[0]), f2u(th[1]), f2u(th[2]), f2u(th[3])});
    }
  }

  // discriminator: the 7-tap "differentiator" has taps {0,0,1,0,-1,0,0}
  // (-1/16 and 1/16 are integer divisions, FmDemodulator.cc:113-122)
  SDR_DEVM static void discriminator(const Ctx &t) {
    constexpr int ITEMS = S / 8;
    const int lim = t.cur >> 3;
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      const float k = hdr_scale(t, c);
      const char *x = cb + TH + 8 * j;  // theta[2j-4 .. 2j-1] (history is 4 floats)
      const u32x2 a = lds_u2(x), b = lds_u2(x + 8);
      const float t0 = u2f(a.x), t1 = u2f(a.y), t2 = u2f(b.x), t3 = u2f(b.y);
      const int d0 = f2i16_wrap(fmul(k, wrap_pi(fsub(t2, t0))));  // n = 2j:   th[n-2]-th[n-4]
      const int d1 = f2i16_wrap(fmul(k, wrap_pi(fsub(t3, t1))));  // n = 2j+1
      if (iabs(d0) > taps::FM_POST::SAFE || iabs(d1) > taps::FM_POST::SAFE) hdr_set(t, HDR_FLAG_A, c, 1);
      sts<uint32_t>(cb + D + HD + 4 * j, pack_i16x2(d0, d1));
    }
  }

  template <int PH>
  SDR_DEVM static void phase(const Ctx &t) {
    if constexpr (PH == 0) {
      phase_front_end<S, HX>(t, XI, XQ, STRIDE);
    } else if constexpr (PH == 1) {
      tuner(t);
    } else if constexpr (PH == 2) {
      discriminator(t);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<HX>(t.smem + c * STRIDE + XI, t.cur);
        shift_history<HX>(t.smem + c * STRIDE + XQ, t.cur);
      }
    } else if constexpr (PH == 3) {
      phase_post12<S / 32, HD, true>(t, D, E2, STRIDE, t.cur >> 4);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<HT>(t.smem + c * STRIDE + TH, t.cur);
    } else if constexpr (PH == 4) {
      phase_audio40<S>(t, E2, PCM, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<HD>(t.smem + c * STRIDE + D, t.cur >> 1);
    } else {
      phase_pcm_out<S>(t, PCM, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<H_AUDIO>(t.smem + c * STRIDE + E2, t.cur >> 3);
    }
  }
};

// ===========================================================================
// Wide-band FM (WbFmDemodulator.cc): I,Q 16-tap FIR at 256 kS/s truncated to
// int8 -> theta = atan2 table -> first difference -> wrap -> * k -> de-emphasis
// IIR (sequential, lane == channel) -> int16 -> 8-tap 4:1 -> 12-tap 4:1 -> 40-tap 2:1.
// ===========================================================================
template <int S_>
struct WbFm {
  static constexpr int S = S_;
  static_assert(S % 256 == 0, "");
  static constexpr int HX = 16;  // 15 used
  static constexpr int HT = 16;  // 4 floats, 2 used
  static constexpr int HD = 16;  // 8 int16, 4 used
  static constexpr int H1 = 16;  // 8 int16, 8 used
  static constexpr int XI = 0;
  static constexpr int XQ = XI + HX + S;
  static constexpr int TH = XQ + HX + S;             // float [4 + S]
  static constexpr int U = TH + HT + 4 * S;          // float [S]
  static constexpr int D = U + 4 * S;                // int16 [8 + S]
  static constexpr int E1 = D + HD + 2 * S;          // int16 [8 + S/4]
  static constexpr int E2 = E1 + H1 + S / 2;         // int16 [40 + S/16]
  static constexpr int PCM = E2 + H_AUDIO + S / 8;   // int16 [S/32]
  static constexpr int SC = PCM + S / 16;            // float vprev[2], y1, pad
  static constexpr int END = SC + 16;
  static constexpr int STRIDE = channel_stride(END);
  static constexpr int NPHASES = 8;
  static constexpr int STATE_BYTES = HX * 2 + HT + HD + H1 + H_AUDIO + 16;

  template <class Fn>
  SDR_DEVM static void for_each_persistent(Fn &&f) {
    f(XI, HX); f(XQ, HX); f(TH, HT); f(D, HD); f(E1, H1); f(E2, H_AUDIO); f(SC, 16);
  }
  SDR_DEVM static void init_flags(const Ctx &t, int c) {
    const char *cb = t.smem + c * STRIDE;
    hdr_set(t, HDR_FLAG_A, c, 0);
    hdr_set(t, HDR_FLAG_B, c, any_above(cb + E2, 40, taps::AUDIO40::SAFE));
  }
  // the carried v[n-1] alternates between two slots; the next launch reads slot 0
  SDR_DEVM static void finalize(const Ctx &t, int c) {
    char *cb = t.smem + c * STRIDE;
    if (t.parity) sts<float>(cb + SC, lds<float>(cb + SC + 4));
  }

  // one output of the 16-tap pre-demodulation filter; T = offset inside the
  // 16-sample item, window words w[0..7] cover x[16j-16 .. 16j+15]
  template <int T>
  SDR_DEVM static int prefilter_one(const uint32_t (&w)[8]) {
    return fir_s8<taps::WB_PRE, 16 + T, 8>(w) >> 15;
  }
  // four consecutive outputs of both arms -> four thetas
  template <int T0>
  SDR_DEVM static u32x4 prefilter_quad(const Ctx &t, const uint32_t (&wi)[8], const uint32_t (&wq)[8]) {
    // (int8_t)sample truncation, then table[(uint8)(q+128)][(uint8)(i+128)]
    // (WbFmDemodulator.cc:393-397, 458-462)
    const float *lut = t.p->lut;
    const int i0 = (prefilter_one<T0 + 0>(wi) + 128) & 255, q0 = (prefilter_one<T0 + 0>(wq) + 128) & 255;
    const int i1 = (prefilter_one<T0 + 1>(wi) + 128) & 255, q1 = (prefilter_one<T0 + 1>(wq) + 128) & 255;
    const int i2 = (prefilter_one<T0 + 2>(wi) + 128) & 255, q2 = (prefilter_one<T0 + 2>(wq) + 128) & 255;
    const int i3 = (prefilter_one<T0 + 3>(wi) + 128) & 255, q3 = (prefilter_one<T0 + 3>(wq) + 128) & 255;
    return u32x4{f2u(ld_lut(lut + q0 * 256 + i0)), f2u(ld_lut(lut + q1 * 256 + i1)),
                 f2u(ld_lut(lut + q2 * 256 + i2)), f2u(ld_lut(lut + q3 * 256 + i3))};
  }
  SDR_DEVM static void prefilter(const Ctx &t) {
    constexpr int ITEMS = S / 16;
    const int lim = t.cur >> 4;
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      uint32_t wi[8], wq[8];
      {
        const char *x = cb + XI + HX + 16 * j;
        const u32x4 a = lds_u4(x - 16), b = lds_u4(x);
        wi[0] = a.x; wi[1] = a.y; wi[2] = a.z; wi[3] = a.w;
        wi[4] = b.x; wi[5] = b.y; wi[6] = b.z; wi[7] = b.w;
      }
      {
        const char *x = cb + XQ + HX + 16 * j;
        const u32x4 a = lds_u4(x - 16), b = lds_u4(x);
        wq[0] = a.x; wq[1] = a.y; wq[2] = a.z; wq[3] = a.w;
        wq[4] = b.x; wq[5] = b.y; wq[6] = b.z; wq[7] = b.w;
      }
      char *dst = cb + TH + HT + 64 * j;
      sts_u4(dst, prefilter_quad<0>(t, wi, wq));
      sts_u4(dst + 16, prefilter_quad<4>(t, wi, wq));
      sts_u4(dst + 32, prefilter_quad<8>(t, wi, wq));
      sts_u4(dst + 48, prefilter_quad<12>(t, wi, wq));
    }
  }

  // discriminator and the numerator of the de-emphasis filter; eight samples per item.
  //   v[n] = k * wrap(theta[n] - theta[n-1]);  u[n] = fl(fl(b0*v[n]) + fl(b1*v[n-1]))
  SDR_DEVM static void discriminator(const Ctx &t) {
    constexpr int ITEMS = S / 8;
    const int lim = t.cur >> 3;
    const float b0 = (float)(0.0253863), b1 = (float)(0.0253863);
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      const float k = hdr_scale(t, c);
      // theta[8j-2 .. 8j+7]; theta[n] lives at float index 4 + n
      const char *x = cb + TH + 4 * (8 * j + 2);
      const u32x2 h = lds_u2(x);
      const u32x4 a = lds_u4(x + 8), b = lds_u4(x + 24);
      const float th[10] = {u2f(h.x), u2f(h.y), u2f(a.x), u2f(a.y), u2f(a.z),
                            u2f(a.w), u2f(b.x), u2f(b.y), u2f(b.z), u2f(b.w)};
      float vprev;
      if (j == 0) {
        vprev = lds<float>(cb + SC + 4 * t.parity);  // carried: it was scaled with the old gain
      } else {
        vprev = fmul(k, wrap_pi(fsub(th[1], th[0])));
      }
      float u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = fmul(k, wrap_pi(fsub(th[2 + i], th[1 + i])));
        u[i] = fadd(fmul(b0, v), fmul(b1, vprev));
        vprev = v;
      }
      if (8 * j + 8 == t.cur) sts<float>(cb + SC + 4 * (t.parity ^ 1), vprev);
      sts_u4(cb + U + 32 * j, u32x4{f2u(u[0]), f2u(u[1]), f2u(u[2]), f2u(u[3])});
      sts_u4(cb + U + 32 * j + 16, u32x4{f2u(u[4]), f2u(u[5]), f2u(u[6]), f2u(u[7])});
    }
  }

  // y[n] = fl(u[n] - fl(a1 * y[n-1])), d[n] = (int16_t)y[n]   (IirFilter.cc:161-176)
  SDR_DEVM static void deemphasis(const Ctx &t) {
    if (t.tid >= t.Gc) return;
    char *cb = t.smem + t.tid * STRIDE;
    const float a1 = (float)(-0.9492274);
    float y1 = lds<float>(cb + SC + 8);
    for (int g = 0; g < t.cur; g += 4) {
      const u32x4 r = lds_u4(cb + U + 4 * g);
      const float u[4] = {u2f(r.x), u2f(r.y), u2f(r.z), u2f(r.w)};
      int o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        y1 = fsub(u[i], fmul(a1, y1));
        o[i] = f2i16_wrap(y1);
      }
      sts_u2(cb + D + HD + 2 * g, u32x2{pack_i16x2(o[0], o[1]), pack_i16x2(o[2], o[3])});
    }
    sts<float>(cb + SC + 8, y1);
  }

  // 8-tap 4:1 on int16 (clamp unreachable: sum|q| * 32768 + 16384 < 2^30); two outputs per item
  SDR_DEVM static void decim1(const Ctx &t) {
    constexpr int ITEMS = S / 8;
    const int lim = t.cur >> 3;
    SDR_FOR_ITEMS(t, ITEMS, c, j) {
      if (j >= lim) continue;
      char *cb = t.smem + c * STRIDE;
      const char *x = cb + D + HD + 16 * j;  // x[8j]; window x[8j-4 .. 8j+7]
      const u32x2 a = lds_u2(x - 8);
      const u32x4 b = lds_u4(x);
      const uint32_t w[6] = {a.x, a.y, b.x, b.y, b.z, b.w};
      using F = taps::WB_DEC1;
      static_assert(F::SAFE >= 32768, "");
      const int y0 = (int)(int16_t)(fir_s16_fast<F, 7, 6>(w) >> 15);
      const int y1 = (int)(int16_t)(fir_s16_fast<F, 11, 6>(w) >> 15);
      sts<uint32_t>(cb + E1 + H1 + 4 * j, pack_i16x2(y0, y1));
    }
  }

  template <int PH>
  SDR_DEVM static void phase(const Ctx &t) {
    if constexpr (PH == 0) {
      phase_front_end<S, HX>(t, XI, XQ, STRIDE);
    } else if constexpr (PH == 1) {
      prefilter(t);
    } else if constexpr (PH == 2) {
      discriminator(t);
      SDR_FOR_CHANNELS_HI(t, c) {
        shift_history<HX>(t.smem + c * STRIDE + XI, t.cur);
        shift_history<HX>(t.smem + c * STRIDE + XQ, t.cur);
      }
    } else if constexpr (PH == 3) {
      deemphasis(t);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<HT>(t.smem + c * STRIDE + TH, 4 * t.cur);
    } else if constexpr (PH == 4) {
      decim1(t);
    } else if constexpr (PH == 5) {
      // decimator 1 output is bounded by 29126 <= FM_POST::SAFE: clamp unreachable
      static_assert((16384 + (long long)taps::WB_DEC1::SUMABS * 32768) / 32768 <= taps::FM_POST::SAFE, "");
      phase_post12<S / 32, H1, false>(t, E1, E2, STRIDE, t.cur >> 4);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<HD>(t.smem + c * STRIDE + D, 2 * t.cur);
    } else if constexpr (PH == 6) {
      phase_audio40<S>(t, E2, PCM, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<H1>(t.smem + c * STRIDE + E1, t.cur >> 1);
    } else {
      phase_pcm_out<S>(t, PCM, STRIDE);
      SDR_FOR_CHANNELS_HI(t, c) shift_history<H_AUDIO>(t.smem + c * STRIDE + E2, t.cur >> 3);
    }
  }
};

// ===========================================================================
// CTA driver pieces shared by the kernel and the emulation
// ===========================================================================
template <class M>
SDR_DEV void cta_load_header(const Ctx &t, uint32_t list0) {
  for (int c = t.tid; c < t.Gc; c += t.nt) {
    const uint32_t ch = t.p->chan_ids[list0 + c];
    hdr_set(t, HDR_CHAN, c, ch);
    sts<float>(t.hdr + HDR_SCALE + 4 * c, t.p->scale[ch]);
    hdr_set(t, HDR_LSB, c, t.p->lsb ? t.p->lsb[ch] : 0);
  }
}

// one thread per channel moves that channel's persistent regions (16-byte units)
template <class M, bool LOAD>
SDR_DEV void cta_state(const Ctx &t) {
  for (int c = t.tid; c < t.Gc; c += t.nt) {
    char *cb = t.smem + c * M::STRIDE;
    uint8_t *blob = t.p->state + (uint64_t)hdr_chan(t, c) * t.p->state_stride;
    if (!LOAD) M::finalize(t, c);
    int pos = 0;
    M::for_each_persistent([&](int off, int bytes) {
      for (int i = 0; i < bytes; i += 16) {
        if (LOAD) sts_u4(cb + off + i, ldg_u4(blob + pos + i));
        else stg_u4(blob + pos + i, lds_u4(cb + off + i));
      }
      pos += bytes;
    });
    if (LOAD) M::init_flags(t, c);
  }
}

template <class M> SDR_HD constexpr int smem_bytes(int G) { return HDR_BYTES + G * M::STRIDE; }

#if SDR_DEVICE_BUILD
template <class M, int PH = 0>
__device__ __forceinline__ void run_phases(const Ctx &t) {
  if constexpr (PH < M::NPHASES) {
    M::template phase<PH>(t);
    __syncthreads();
    run_phases<M, PH + 1>(t);
  }
}

template <class M>
__global__ void __launch_bounds__(1024, 1) demod_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ uint4 smem_raw[];
  Ctx t;
  t.p = &p;
  t.hdr = reinterpret_cast<char *>(smem_raw);
  t.smem = t.hdr + HDR_BYTES;
  t.tid = threadIdx.x;
  t.nt = blockDim.x;
  const uint32_t list0 = blockIdx.x * p.G;
  t.Gc = (int)min(p.G, p.n_list - list0);
  t.sample0 = 0;
  t.cur = 0;
  t.parity = 0;
  cta_load_header<M>(t, list0);
  __syncthreads();
  cta_state<M, true>(t);
  __syncthreads();
  for (uint32_t s0 = 0; s0 < p.n_samples; s0 += M::S) {
    t.sample0 = s0;
    t.cur = (int)min((uint32_t)M::S, p.n_samples - s0);
    run_phases<M>(t);
    t.parity ^= 1;
  }
  cta_state<M, false>(t);
}
#endif

}  // namespace sdr
