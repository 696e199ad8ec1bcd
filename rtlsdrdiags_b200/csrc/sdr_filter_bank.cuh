// Batched multirate filter banks: the reference's generic filter classes
// that the IQ->PCM path does not instantiate itself, as one kernel over [row][sample].
//
//   Filters/Decimator.cc:168-209, 281-322          float FIR + M:1 compressor (M = 1: FirFilter.cc:144-185)
//   Filters/Interpolator.cc (filterData, interpolate, createPolyphaseCoefficients)   float polyphase 1:L
//   Filters/Int16/Decimator_int16.cc:176-238, 310-351     Q15 twin (M = 1: FirFilter_int16.cc:151-213)
//   Filters/Int16/Interpolator_int16.cc (filterData, interpolate)                    Q15 polyphase 1:L
//
// Every row is an independent filter object with the same taps. The arithmetic is the
// reference's, tap by tap in its order: float `y = y + h[k]*x` as a single-rounded multiply and a
// single-rounded add (never an FMA; the reference is built for baseline x86-64), Q15 with the
// rounding constant 1<<14, the clamp after every tap and the arithmetic >>15.
//
// Layout. One CTA computes FB_TILE consecutive outputs of one row. It stages the input span it
// needs in shared memory, split by phase (sample s of the span sits at X[s % M][s / M]) so that a
// warp whose lanes own consecutive outputs reads consecutive words for every tap, and keeps the
// taps in shared memory in prototype order (lanes of an interpolator read h[i + kL], consecutive
// in i). Samples before the call's first input come from the row's carry buffer: the last
// C = N-1 + M-1 samples the object had seen (decimator; q-1 for an interpolator), of which the
// newest `pending` were not consumed yet (Decimator.cc:296-303 keeps them in decimationBuffer).
// The carry is double-buffered: CTA 0 of a row writes the next call's carry into the other
// buffer while the rest still read this call's.
#pragma once
#include "sdr_platform.h"

namespace sdr {

enum { FB_DEC_F32 = 1, FB_INT_F32 = 2, FB_DEC_I16 = 3, FB_INT_I16 = 4 };
constexpr int FB_THREADS = 256;

struct FilterBankParams {
  const void *in;         // [rows][in_stride] elements
  void *out;              // [rows][out_stride]
  const void *carry_in;   // [rows][C]
  void *carry_out;        // [rows][C]
  const void *taps;       // N taps, prototype order (float or int16 widened to int32)
  uint64_t in_stride, out_stride;
  uint64_t n_in, n_out;   // per row, this call
  uint32_t N, F, q;       // taps, factor (M or L), taps per output (N for a decimator, N / L)
  uint32_t C;             // carry length
  uint32_t pending;       // decimator: unconsumed samples at the end of the carry
  uint32_t tile_out;      // outputs per CTA
  uint32_t span;          // input samples a full tile needs
  uint32_t pitch;         // words per phase row of X
};

template <class T> struct FbAcc;
template <> struct FbAcc<float> {
  using acc_t = float;
  using tap_t = float;
  __device__ __forceinline__ static float init() { return 0.f; }
  __device__ __forceinline__ static float mac(float acc, float h, float x) { return fadd(acc, fmul(h, x)); }
  __device__ __forceinline__ static float done(float acc) { return acc; }
};
template <> struct FbAcc<int16_t> {
  using acc_t = int32_t;
  using tap_t = int32_t;
  __device__ __forceinline__ static int32_t init() { return 1 << 14; }
  __device__ __forceinline__ static int32_t mac(int32_t acc, int32_t h, int16_t x) {
    // |acc| <= 2^30 and |h x| <= 2^30: the sum cannot leave int32 before the clamp
    acc += h * (int32_t)x;
    return max(min(acc, 0x3fffffff), -0x40000000);
  }
  __device__ __forceinline__ static int16_t done(int32_t acc) { return (int16_t)(acc >> 15); }
};

// element s of the row's logical stream, s counted from the call's first input (s < 0: carry)
template <class T>
__device__ __forceinline__ T fb_sample(const T *in_row, const T *carry_row, uint32_t C, int64_t s, uint64_t n_in) {
  if (s >= 0) return s < (int64_t)n_in ? in_row[s] : (T)0;
  return s >= -(int64_t)C ? carry_row[(int64_t)C + s] : (T)0;
}

template <class T, bool INTERP>
__global__ void __launch_bounds__(FB_THREADS) filter_bank_kernel(const __grid_constant__ FilterBankParams p) {
  using A = FbAcc<T>;
  using tap_t = typename A::tap_t;
  extern __shared__ uint4 fb_smem_raw[];
  tap_t *h = reinterpret_cast<tap_t *>(fb_smem_raw);
  T *X = reinterpret_cast<T *>(h + p.N);
  const uint32_t row = blockIdx.y;
  const T *in_row = reinterpret_cast<const T *>(p.in) + (uint64_t)row * p.in_stride;
  const T *carry_row = reinterpret_cast<const T *>(p.carry_in) + (uint64_t)row * p.C;
  const uint64_t o0 = (uint64_t)blockIdx.x * p.tile_out;  // first output of this tile
  const uint32_t M = INTERP ? 1u : p.F;

  // first stream sample the tile touches: the oldest tap of its first output
  //   decimator: newest input of output o is M o + M-1 - pending (Decimator.cc:296-316)
  //   interpolator: outputs nL .. nL+L-1 filter inputs n, n-1, .. n-q+1 (Interpolator.cc interpolate)
  const int64_t s0 = INTERP ? (int64_t)(o0 / p.F) - (int64_t)(p.q - 1)
                            : (int64_t)(o0 * M) + (int64_t)(M - 1) - (int64_t)p.pending - (int64_t)(p.N - 1);
  for (uint32_t i = threadIdx.x; i < p.N; i += FB_THREADS) h[i] = reinterpret_cast<const tap_t *>(p.taps)[i];
  for (uint32_t i = threadIdx.x; i < p.span; i += FB_THREADS)
    X[(i % M) * p.pitch + i / M] = fb_sample(in_row, carry_row, p.C, s0 + i, p.n_in);
  __syncthreads();

  T *out_row = reinterpret_cast<T *>(p.out) + (uint64_t)row * p.out_stride;
  for (uint32_t t = threadIdx.x; t < p.tile_out; t += FB_THREADS) {
    const uint64_t o = o0 + t;
    if (o >= p.n_out) break;
    typename A::acc_t acc = A::init();
    if (INTERP) {
      const uint32_t n = (uint32_t)(o / p.F - o0 / p.F) + (p.q - 1);  // newest input, span index
      const uint32_t i = (uint32_t)(o % p.F);
      for (uint32_t k = 0; k < p.q; ++k) acc = A::mac(acc, h[i + k * p.F], X[n - k]);
    } else {
      // newest input at span index M t + N-1; tap k reads index M t + (N-1-k)
      uint32_t ph = (p.N - 1) % M, base = t + (p.N - 1) / M;
      for (uint32_t k = 0; k < p.N; ++k) {
        acc = A::mac(acc, h[k], X[ph * p.pitch + base]);
        if (ph == 0) { ph = M - 1; --base; } else --ph;
      }
    }
    out_row[o] = A::done(acc);
  }

  // the row's next carry: the last C samples of (carry ++ input)
  if (blockIdx.x == 0) {
    T *carry_next = reinterpret_cast<T *>(p.carry_out) + (uint64_t)row * p.C;
    for (uint32_t i = threadIdx.x; i < p.C; i += FB_THREADS)
      carry_next[i] = fb_sample(in_row, carry_row, p.C, (int64_t)p.n_in - (int64_t)p.C + i, p.n_in);
  }
}

}  // namespace sdr
