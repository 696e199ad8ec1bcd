// Arithmetic building blocks of the IQ -> PCM demodulation kernels (sm_100a).
//
// Everything here is a pure function of its arguments: the packed-byte front end
// (u8 -> s8, Fs/4 rotation, de-interleave), the Q15 FIR cores on IDP.2A dot products
// (int8 samples x int16 taps, int16 samples x split taps, and the exact ordered
// MAC + clamp path), the pi wrap and the wrapping float -> int16 conversion. The
// warp-tile kernels in sdr_tile.cuh are built from them.
//
// Arithmetic contract: Q15 FIRs accumulate from the
// rounding constant 1<<14 in int32, clamp after every tap where the clamp is
// reachable, shift right by 15; float stages use single-rounded FP32 ops in the
// reference's order; float->int16 conversions wrap like x86 cvttss2si.
//
// This header also compiles under a host compiler with -DSDR_EMU (tests/emu), so the
// same functions are checked against the oracle on a CPU-only box; see sdr_platform.h.
#pragma once
#include "sdr_platform.h"
#include "sdr_q15_taps.h"

namespace sdr {

enum { FMT_U8_OFFSET_ROTATE = 0, FMT_S8_ROTATED = 1 };
enum { MODE_NONE = 0, MODE_AM = 1, MODE_FM = 2, MODE_WBFM = 3, MODE_LSB = 4, MODE_USB = 5 };

constexpr int MAX_G = 32;       // channels per CTA (one warp of recurrence lanes)
constexpr int FM_LUT_MIN = -140;  // tuner decimator output range for int8 input
constexpr int FM_LUT_DIM = 280;

// Everything one launch needs. Plain pointers: device memory under nvcc, host
// memory under the emulation.
struct LaunchParams {
  const uint8_t *iq;         // [n_channels][ch_stride] bytes, I,Q interleaved
  uint64_t ch_stride;        // bytes between channels (multiple of 16)
  uint32_t n_samples;        // complex samples per channel in this launch (multiple of 32)
  int fmt;                   // FMT_*
  const uint32_t *chan_ids;  // channels of this mode, grouped G per CTA
  uint32_t n_list;
  uint32_t G;
  uint8_t *state;            // this mode's state blobs, [n_channels][state_stride]
  uint32_t state_stride;
  const float *scale;        // [n_channels] AM/SSB: gain; FM/WBFM: gain/dev*32767
  const uint8_t *lsb;        // [n_channels] SSB only: 1 = LSB
  int16_t *pcm;              // [n_channels][pcm_stride]
  uint64_t pcm_stride;       // int16 elements between channels
  const float *lut;          // atan2 table of this mode (FM 280x280, WBFM 256x256)
  uint32_t aux;              // kernel-specific word (WBFM: scheduler sharing; AM/SSB FIR: worker warps; dc_block: tail offset)
  int16_t *scratch;          // AM/SSB: IIR numerators between the FIR and the recurrence kernel,
                             // [list index][row = tile][32 lanes] (exact small integers)
  const uint8_t *allowed;    // [n_channels] squelch gate of this call, or nullptr = all open
  uint32_t call_id;          // AM/SSB/FM: 31-bit id of this call, never 0 (versions the carry buffers)
  const uint32_t *tab;       // FM: tensor-core tuner tables (fm_mma_table in sdr_engine.cu)
  unsigned long long *trace; // SDR_TRACE=1 only: [0] = earliest CTA start, [1] = latest CTA end of this launch (globaltimer ns)
  // dc_block_kernel: a call's rows (32 PCM samples each) are cut into seg_count segments of
  // seg_rows rows per channel; a segment warms up on the warm_rows rows before it
  uint32_t seg_count, seg_rows, warm_rows;
  uint32_t *counters;        // [0] = segments dc_block_kernel had to redo serially (diagnostics); [1], [2] = WBFM
                             // (half-)tiles whose pre-filter ran on the tensor / CUDA cores (diagnostics; generation 4:
                             // when call_id != 0); [3] = generation 4: (half-)tiles a clipping byte kept off the tensor cores
};

#if SDR_DEVICE_BUILD
// launch timeline for tools/probe_timeline.py (engine built as usual, SDR_TRACE=1 in the environment)
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_begin(const LaunchParams &p) {
  if (p.trace && threadIdx.x == 0) atomicMin(p.trace, trace_now());
}
__device__ __forceinline__ void trace_end(const LaunchParams &p) {
  if (p.trace && threadIdx.x == 0) atomicMax(p.trace + 1, trace_now());
}
#endif

// ---------------------------------------------------------------------------
// C-semantics helpers
// ---------------------------------------------------------------------------
// (int16_t)floatValue the way x86-64 compiles it: cvttss2si (0x80000000 when out
// of range or NaN) then the low 16 bits. cvt.rzi saturates, so patch the range.
SDR_DEV int f2i16_wrap(float v) {
  int r = f2i_rz(v);
  if (!(fabsf(v) < 2147483648.0f)) r = 0;
  return (int)(int16_t)(uint16_t)(uint32_t)r;
}

// while (d > M_PI) d -= 2*M_PI; while (d < -M_PI) d += 2*M_PI; in the reference the
// compares and the additions are double, the store rounds to float
// (FmDemodulator.cc:486-494). For a float d, d > M_PI  <=>  d >= (float)M_PI. d is a
// difference of two atan2 values, |d| <= 2 pi, so each loop runs at most once; written
// as two predicated steps (no divergent branch) because in a warp some lane wraps on
// almost every sample.
SDR_DEV float wrap_pi(float d) {
  const float PI_F = 3.14159274101257324f;
  // adding 0.0 in double and rounding back returns d itself, so one FP64 add serves
  // all three cases
  const double off = d >= PI_F ? -6.283185307179586 : (d <= -PI_F ? 6.283185307179586 : 0.0);
  return dadd_to_f(d, off);
}

// The same wrap without FP64, for differences of two atan2-TABLE values only (both kernels take
// theta from a table: WBFM 256x256, WbFmDemodulator.cc:159-170; NBFM 280x280 over the tuner's
// output range): 2 pi = HI + LO with HI = (float)(2 pi); d - HI is exact (Sterbenz), and
// subtracting LO in float rounds to the value the reference's double subtraction rounds to for
// EVERY pair of distinct table values -- checked exhaustively, 39,920^2 = 1.59e9 pairs of the WBFM
// table and 47,808^2 = 2.29e9 of the NBFM table (tests/test_emu_parity.py). Not valid for
// arbitrary floats.
SDR_DEV float wrap_pi_table(float d) {
  const float PI_F = 3.14159274101257324f;
  // d -+ HI, then +- |LO|, each rounded once: with sgn = copysign(1, d) both steps are an FMA whose product is exact
  // (sgn * constant), i.e. the same two roundings as fsub(fsub(d, copysign(HI, d)), copysign(LO, -d)) -- and three
  // instructions instead of five
  const float sgn = u2f((f2u(d) & 0x80000000u) | 0x3F800000u);
  const float w = ffma(sgn, u2f(0x343BBD2Eu) /* (float)(2 pi) - 2 pi = 1.7484555e-7 */,
                       ffma(sgn, u2f(0xC0C90FDBu) /* -(float)(2 pi) */, d));
  return fabsf(d) >= PI_F ? w : d;
}

// ---------------------------------------------------------------------------
// Front end: u8 offset-binary -> s8, multiply by {1, j, -1, -j}, de-interleave
// (IqDataProcessor.cc:735-738 and 567-611). Works on packed bytes.
// ---------------------------------------------------------------------------
// bytes marked in `neg` become (0x80 - u) mod 256 = wrap-negated (u - 128);
// the others become u ^ 0x80 = u - 128. -(-128) stays -128, as int8 does.
SDR_DEV uint32_t offset_and_negate(uint32_t raw, uint32_t neg, uint32_t one) {
  uint32_t v = raw ^ neg;
  return ((v & 0x7f7f7f7fu) + one) ^ (~v & 0x80808080u);
}

// w0 = I0 Q0 I1 Q1, w1 = I2 Q2 I3 Q3 (one rotation period). a = I' plane word,
// b = Q' plane word, each holding four consecutive samples in time order.
SDR_DEV void front_end_group(int fmt, uint32_t w0, uint32_t w1, uint32_t &a, uint32_t &b) {
  if (fmt == FMT_U8_OFFSET_ROTATE) {
    // I' = I0, -Q1, -I2, Q3      Q' = Q0, I1, -Q2, -I3
    a = offset_and_negate(byte_perm(w0, w1, 0x7430), 0x00ffff00u, 0x00010100u);
    b = offset_and_negate(byte_perm(w0, w1, 0x6521), 0xffff0000u, 0x01010000u);
  } else {
    a = byte_perm(w0, w1, 0x6420);
    b = byte_perm(w0, w1, 0x7531);
  }
}

// The same front end with the bytes grouped by what has to be done to them (AM / SSB stage 1): of a rotation
// period's eight bytes the Fs/4 rotation negates four (Q1, I2, Q2, I3) and leaves four alone, so gathering
// the negated ones into one word makes the offset a single XOR for one word and the wrapping negation a
// three-instruction SWAR subtraction for the other -- six instructions per period instead of ten. The words:
//   a = (s0 of I', s3 of I', s0 of Q', s1 of Q')      b = (s1 of I', s2 of I', s2 of Q', s3 of Q')
// with s0..s3 the period's four samples in time order: I' lives in the low halves (a.b0, b.b0, b.b1, a.b1),
// Q' in the high halves (a.b2, a.b3, b.b2, b.b3), which is all a dot product with compile-time tap pairs needs.
//   u8 input:      a = (I0, Q3, Q0, I1) - 128            b = -(Q1, I2, Q2, I3) + 128 as int8 (so -(-128) = -128)
//   signed input:  a = (I0, I3, Q0, Q1)                  b = (I1, I2, Q2, Q3)
SDR_DEV void front_end_ab(int fmt, uint32_t w0, uint32_t w1, uint32_t &a, uint32_t &b) {
  if (fmt == FMT_U8_OFFSET_ROTATE) {
    a = byte_perm(w0, w1, 0x2170) ^ 0x80808080u;
    const uint32_t u = byte_perm(w0, w1, 0x6543);
    // per byte (0x80 - u) mod 256: 0x80808080 - (u & 0x7f..) cannot borrow across bytes; bit 7 comes from u's
    b = (0x80808080u - (u & 0x7f7f7f7fu)) ^ (u & 0x80808080u);
  } else {
    a = byte_perm(w0, w1, 0x3160);
    b = byte_perm(w0, w1, 0x7542);
  }
}
// and back: the raw bytes of a rotation period from (a, b) (both steps are involutions)
SDR_DEV void raw_from_ab(int fmt, uint32_t a, uint32_t b, uint32_t &w0, uint32_t &w1) {
  if (fmt == FMT_U8_OFFSET_ROTATE) {
    const uint32_t ar = a ^ 0x80808080u;
    const uint32_t br = (0x80808080u - (b & 0x7f7f7f7fu)) ^ (b & 0x80808080u);
    w0 = byte_perm(ar, br, 0x4320);  // I0 Q0 I1 Q1
    w1 = byte_perm(ar, br, 0x1765);  // I2 Q2 I3 Q3
  } else {
    w0 = byte_perm(a, b, 0x3420);
    w1 = byte_perm(a, b, 0x7165);
  }
}

// ---------------------------------------------------------------------------
// Q15 FIR cores
// ---------------------------------------------------------------------------
SDR_HD constexpr uint32_t pack16(int lo, int hi) {
  return ((uint32_t)lo & 0xffffu) | (((uint32_t)hi & 0xffffu) << 16);
}
// int16 tap split into a signed high byte and an unsigned low byte: t = 256*hi + lo
SDR_HD constexpr uint32_t tap_bytes(int t0, int t1) {
  return ((uint32_t)(t0 >> 8) & 0xffu) | (((uint32_t)(t1 >> 8) & 0xffu) << 8) |
         (((uint32_t)t0 & 0xffu) << 16) | (((uint32_t)t1 & 0xffu) << 24);
}

// --- int8 samples, four per word, byte 0 oldest. Sample at window position
// `pos` meets tap P - pos; taps outside the filter are zero and fold away.
template <class F, int P, int I>
SDR_DEV int fir_s8_word(int acc, uint32_t w) {
  constexpr int t0 = F::tap(P - 4 * I), t1 = F::tap(P - 4 * I - 1);
  constexpr int t2 = F::tap(P - 4 * I - 2), t3 = F::tap(P - 4 * I - 3);
  constexpr uint32_t lo = pack16(t0, t1), hi = pack16(t2, t3);
  if constexpr (t0 != 0 || t1 != 0) acc = dp2a_lo_ss(lo, w, acc);
  if constexpr (t2 != 0 || t3 != 0) acc = dp2a_hi_ss(hi, w, acc);
  return acc;
}
template <class F, int P, int NW, int I = 0>
SDR_DEV int fir_s8(const uint32_t (&w)[NW], int acc = 1 << 14) {
  if constexpr (I < NW) {
    return fir_s8<F, P, NW, I + 1>(w, fir_s8_word<F, P, I>(acc, w[I]));
  } else {
    return acc;
  }
}

// --- int16 samples, two per word. Fast path: valid only while the clamp is
// unreachable (every |x| in the window <= F::SAFE). acc = 16384 + 256*H + L.
template <class F, int P, int I>
SDR_DEV void fir_s16_word(int &h, int &l, uint32_t w) {
  constexpr int t0 = F::tap(P - 2 * I), t1 = F::tap(P - 2 * I - 1);
  if constexpr (t0 != 0 || t1 != 0) {
    constexpr uint32_t tb = tap_bytes(t0, t1);
    h = dp2a_lo_ss(w, tb, h);
    l = dp2a_hi_su(w, tb, l);
  }
}
template <class F, int P, int NW, int I = 0>
SDR_DEV void fir_s16_acc(const uint32_t (&w)[NW], int &h, int &l) {
  if constexpr (I < NW) {
    fir_s16_word<F, P, I>(h, l, w[I]);
    fir_s16_acc<F, P, NW, I + 1>(w, h, l);
  }
}
template <class F, int P, int NW>
SDR_DEV int fir_s16_fast(const uint32_t (&w)[NW]) {
  int h = 0, l = 1 << 14;
  fir_s16_acc<F, P, NW>(w, h, l);
  return (h << 8) + l;
}

// --- exact path: taps in the reference's order (k = 0 is the newest sample),
// clamp to [-2^30, 2^30-1] after every tap (Decimator_int16.cc:201-219).
template <class F, int P, int NW, int K = 0>
SDR_DEV int fir_s16_exact(const uint32_t (&w)[NW], int acc = 1 << 14) {
  if constexpr (K < F::N) {
    constexpr int pos = P - K;
    constexpr int t = F::tap(K);
    if constexpr (pos >= 0 && pos < 2 * NW) {
      const int x = (pos & 1) ? ((int)w[pos / 2] >> 16) : (int)(int16_t)(w[pos / 2] & 0xffffu);
      acc += t * x;
      acc = acc > 0x3fffffff ? 0x3fffffff : acc;
      acc = acc < -0x40000000 ? -0x40000000 : acc;
    }
    return fir_s16_exact<F, P, NW, K + 1>(w, acc);
  } else {
    return acc;
  }
}
// --- guarded path for windows in which the clamp IS reachable (some |x| > F::SAFE). A clamp
// needs a prefix sum beyond 2^30, and the filters here carry most of their weight in a few
// centre taps. So: the first HEAD taps, whose absolute sum is at most 8191, cannot clamp
// whatever the data (16384 + 32768 * 8191 < 2^30) and are summed with IDP.2A; the middle taps
// run in the reference's order with the clamp; and if the sum then leaves room for the worst
// the remaining taps could add (|acc| <= 2^30 - 1 - 32768 * their absolute sum) the tail is
// summed plainly too -- otherwise it takes the clamped loop. Bit-identical to fir_s16_exact.
template <class F, int K0, int K1>
struct TapRange {
  static constexpr int N = F::N;
  SDR_HD static constexpr int tap(int k) { return (k >= K0 && k < K1) ? F::tap(k) : 0; }
};
template <class F>
struct Guard {
  SDR_HD static constexpr int aabs(int v) { return v < 0 ? -v : v; }
  SDR_HD static constexpr int head() {
    int s = 0, k = 0;
    while (k < F::N && s + aabs(F::tap(k)) <= 8191) { s += aabs(F::tap(k)); ++k; }
    return k;
  }
  SDR_HD static constexpr int tail() {
    int s = 0, k = F::N;
    while (k > head() && s + aabs(F::tap(k - 1)) <= 6144) { s += aabs(F::tap(k - 1)); --k; }
    return k;
  }
  SDR_HD static constexpr int tail_sum() {
    int s = 0;
    for (int k = tail(); k < F::N; ++k) s += aabs(F::tap(k));
    return s;
  }
  static constexpr int HEAD = head(), TAIL = tail();
  static constexpr int LIMIT = 0x3fffffff - 32768 * tail_sum();
};
template <class F, int P, int NW, int K, int KEND>
SDR_DEV int fir_s16_exact_range(const uint32_t (&w)[NW], int acc) {
  if constexpr (K < KEND) {
    constexpr int pos = P - K;
    constexpr int t = F::tap(K);
    if constexpr (pos >= 0 && pos < 2 * NW) {
      const int x = (pos & 1) ? ((int)w[pos / 2] >> 16) : (int)(int16_t)(w[pos / 2] & 0xffffu);
      acc += t * x;
      acc = acc > 0x3fffffff ? 0x3fffffff : acc;
      acc = acc < -0x40000000 ? -0x40000000 : acc;
    }
    return fir_s16_exact_range<F, P, NW, K + 1, KEND>(w, acc);
  } else {
    return acc;
  }
}
// The middle taps WITHOUT the clamp, keeping the largest and smallest prefix sum: if every prefix
// stayed inside [-2^30, 2^30-1] the reference's clamps changed nothing and the plain sum is its
// result. One dependent IMAD per tap instead of IMAD -> min -> max. (No int32 overflow can hide a
// violation: up to the first prefix outside the range every |prefix| <= 2^30 and a term is below
// 2^29, and hi / lo never forget that first violation.)
template <class F, int P, int NW, int K, int KEND>
SDR_DEV void fir_s16_prefix_range(const uint32_t (&w)[NW], int &acc, int &hi, int &lo) {
  if constexpr (K < KEND) {
    constexpr int pos = P - K;
    constexpr int t = F::tap(K);
    if constexpr (pos >= 0 && pos < 2 * NW) {
      const int x = (pos & 1) ? ((int)w[pos / 2] >> 16) : (int)(int16_t)(w[pos / 2] & 0xffffu);
      acc += t * x;
      hi = acc > hi ? acc : hi;
      lo = acc < lo ? acc : lo;
    }
    fir_s16_prefix_range<F, P, NW, K + 1, KEND>(w, acc, hi, lo);
  }
}
// head + unclamped middle; clean = no prefix of the middle left the clamp range
template <class F, int P, int NW>
SDR_DEV int fir_s16_guard_mid_plain(const uint32_t (&w)[NW], bool &clean) {
  static_assert(F::N <= 64, "a term t * x must stay below 2^29");
  int h = 0, l = 1 << 14;
  fir_s16_acc<TapRange<F, 0, Guard<F>::HEAD>, P, NW>(w, h, l);
  int acc = (h << 8) + l, hi = acc, lo = acc;
  fir_s16_prefix_range<F, P, NW, Guard<F>::HEAD, Guard<F>::TAIL>(w, acc, hi, lo);
  clean = hi <= 0x3fffffff && lo >= -0x40000000;
  return acc;
}
// head + clamped middle; the caller decides about the tail (warp-uniformly on the GPU)
template <class F, int P, int NW>
SDR_DEV int fir_s16_guard_mid(const uint32_t (&w)[NW]) {
  int h = 0, l = 1 << 14;
  fir_s16_acc<TapRange<F, 0, Guard<F>::HEAD>, P, NW>(w, h, l);
  return fir_s16_exact_range<F, P, NW, Guard<F>::HEAD, Guard<F>::TAIL>(w, (h << 8) + l);
}
template <class F>
SDR_DEV bool fir_s16_guard_tail_is_free(int acc) {
  return acc <= Guard<F>::LIMIT && acc >= -Guard<F>::LIMIT;
}
template <class F, int P, int NW>
SDR_DEV int fir_s16_guard_tail(const uint32_t (&w)[NW], int acc, bool clamped) {
  if (clamped) return fir_s16_exact_range<F, P, NW, Guard<F>::TAIL, F::N>(w, acc);
  int h = 0, l = 0;
  fir_s16_acc<TapRange<F, Guard<F>::TAIL, F::N>, P, NW>(w, h, l);
  return acc + (h << 8) + l;
}

template <class F, int P, int NW>
SDR_DEV int fir_s16(const uint32_t (&w)[NW], bool exact) {
  if constexpr (F::SAFE >= 32768) {
    return fir_s16_fast<F, P, NW>(w);
  } else {
    return exact ? fir_s16_exact<F, P, NW>(w) : fir_s16_fast<F, P, NW>(w);
  }
}

SDR_DEV uint32_t pack_i8x4(int a, int b, int c, int d) {
  return ((uint32_t)a & 0xffu) | (((uint32_t)b & 0xffu) << 8) | (((uint32_t)c & 0xffu) << 16) |
         ((uint32_t)d << 24);
}
SDR_DEV uint32_t pack_i16x2(int a, int b) { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); }
SDR_DEV int iabs(int v) { return v < 0 ? -v : v; }

}  // namespace sdr
