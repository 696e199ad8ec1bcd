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
  uint32_t sum_abs_taps;  // Q15 kinds: sum |q[k]|
  uint32_t hoff, np;      // taps sit at word hoff of the tap area (so that tap blocks are 16-byte aligned), X at word np
  uint32_t kp;            // float decimators with M = 2, 4: taps before the first full block of 4 M
};

// CLAMP = false: the tile's samples are small enough that no partial sum can reach the Q15
// clamp (sum|q| * max|x| + 2^14 < 2^30), so the per-tap clamp is skipped -- same result
template <class T, bool CLAMP> struct FbAcc;
template <bool CLAMP> struct FbAcc<float, CLAMP> {
  using acc_t = float;
  using tap_t = float;
  __device__ __forceinline__ static float init() { return 0.f; }
  __device__ __forceinline__ static float mac(float acc, float h, float x) { return fadd(acc, fmul(h, x)); }
  __device__ __forceinline__ static float done(float acc) { return acc; }
};
template <bool CLAMP> struct FbAcc<int16_t, CLAMP> {
  using acc_t = int32_t;
  using tap_t = int32_t;
  __device__ __forceinline__ static int32_t init() { return 1 << 14; }
  __device__ __forceinline__ static int32_t mac(int32_t acc, int32_t h, int16_t x) {
    // |acc| <= 2^30 and |h x| <= 2^30: the sum cannot leave int32 before the clamp
    acc += h * (int32_t)x;
    return CLAMP ? max(min(acc, 0x3fffffff), -0x40000000) : acc;
  }
  __device__ __forceinline__ static int16_t done(int32_t acc) { return (int16_t)(acc >> 15); }
};

// element s of the row's logical stream, s counted from the call's first input (s < 0: carry)
template <class T>
__device__ __forceinline__ T fb_sample(const T *in_row, const T *carry_row, uint32_t C, int64_t s, uint64_t n_in) {
  if (s >= 0) return s < (int64_t)n_in ? in_row[s] : (T)0;
  return s >= -(int64_t)C ? carry_row[(int64_t)C + s] : (T)0;
}

// the outputs of one tile from the staged span
template <class T, bool INTERP, int MT, bool CLAMP>
__device__ __forceinline__ void fb_compute(const FilterBankParams &p, const typename FbAcc<T, CLAMP>::tap_t *h, const T *X,
                                           uint32_t row, uint64_t o0, uint32_t M) {
  using A = FbAcc<T, CLAMP>;
  using tap_t = typename A::tap_t;
  using acc_t = typename A::acc_t;
  T *out_row = reinterpret_cast<T *>(p.out) + (uint64_t)row * p.out_stride;
  if constexpr (INTERP && MT > 0) {
    // Interpolator with the factor known at compile time: a thread owns INPUT samples (n, n + 256)
    // and all L outputs of each. Per tap index k it loads the L taps h[kL .. kL+L-1] with one or two
    // vector loads (the same for every lane) and one input per owned sample: 2 L multiply-adds for
    // 3 shared-memory loads instead of 2 loads per multiply-add. Every output still accumulates
    // its own taps h[i], h[i+L], ... in that order (Interpolator.cc filterData).
    constexpr int L = MT, RN = 2;
    const uint32_t n_tile = p.tile_out / L;  // inputs per tile
    for (uint32_t n0 = threadIdx.x; n0 < n_tile; n0 += RN * FB_THREADS) {
      if (o0 + (uint64_t)n0 * L >= p.n_out) break;
      acc_t acc[RN][L];
      const T *xp[RN];
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        xp[j] = X + min(n0 + j * FB_THREADS, n_tile - 1) + (p.q - 1);  // newest input, span index
#pragma unroll
        for (int i = 0; i < L; ++i) acc[j][i] = A::init();
      }
      for (uint32_t k = 0; k < p.q; ++k) {
        tap_t hk[L];
        if constexpr (L % 4 == 0) {
#pragma unroll
          for (int v = 0; v < L / 4; ++v) {
            const uint4 t4 = *reinterpret_cast<const uint4 *>(h + k * L + 4 * v);
            memcpy(&hk[4 * v], &t4, 16);
          }
        } else if constexpr (L % 2 == 0) {
#pragma unroll
          for (int v = 0; v < L / 2; ++v) {
            const uint2 t2 = *reinterpret_cast<const uint2 *>(h + k * L + 2 * v);
            memcpy(&hk[2 * v], &t2, 8);
          }
        } else {
#pragma unroll
          for (int i = 0; i < L; ++i) hk[i] = h[k * L + i];
        }
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          const T x = *(xp[j] - k);
#pragma unroll
          for (int i = 0; i < L; ++i) acc[j][i] = A::mac(acc[j][i], hk[i], x);
        }
      }
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        const uint32_t n = n0 + j * FB_THREADS;
        const uint64_t o = o0 + (uint64_t)n * L;
        if (n < n_tile && o < p.n_out) {
#pragma unroll
          for (int i = 0; i < L; ++i) out_row[o + i] = A::done(acc[j][i]);
        }
      }
    }
    return;
  }
  // (filters shorter than one block of 4 M taps stay on the generic path below)
  if constexpr (!INTERP && (MT == 2 || MT == 4) && sizeof(T) == 4 && sizeof(tap_t) == 4) if (p.N > p.kp) {
    // Float decimator, M = 2 or 4: a thread owns FOUR CONSECUTIVE outputs t0 .. t0+3. Tap k meets the
    // sample at X[ph][t + a] with a = (N-1-k) / M, ph = (N-1-k) % M; four consecutive values of a
    // for four consecutive outputs touch seven consecutive words of a phase row, i.e. two aligned
    // 16-byte loads (t0 and the block's lowest a are multiples of 4), and the block's 4 M taps are
    // consecutive in k: 2 M + M loads of 16 bytes feed 16 M multiply-adds, against one load per
    // multiply-add. Each output still adds its taps in the reference's order k = 0, 1, ...
    // (Decimator.cc:168-209): the kp taps above the highest full block one by one, then the blocks
    // from the highest a down, within a block a descending and ph descending.
    constexpr int M_ = MT;
    const int n_blocks = (int)((p.N - p.kp) / (4 * M_));
    for (uint32_t t0 = 4 * threadIdx.x; t0 < p.tile_out; t0 += 4 * FB_THREADS) {
      if (o0 + t0 >= p.n_out) break;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (uint32_t k = 0; k < p.kp; ++k) {
        const uint32_t r = p.N - 1 - k;
        const float hk = h[k];
        const float *xr = X + (r % M_) * p.pitch + t0 + r / M_;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fadd(acc[j], fmul(hk, xr[j]));
      }
      const float *hb = h + p.kp;
      for (int m = n_blocks - 1; m >= 0; --m, hb += 4 * M_) {
        float tp[4 * M_];
#pragma unroll
        for (int v = 0; v < M_; ++v) {
          const uint4 t4 = *reinterpret_cast<const uint4 *>(hb + 4 * v);
          memcpy(&tp[4 * v], &t4, 16);
        }
        float W[M_][8];
#pragma unroll
        for (int ph = 0; ph < M_; ++ph) {
          const uint4 lo = *reinterpret_cast<const uint4 *>(X + ph * p.pitch + t0 + 4 * m);
          const uint4 hi = *reinterpret_cast<const uint4 *>(X + ph * p.pitch + t0 + 4 * m + 4);
          memcpy(&W[ph][0], &lo, 16);
          memcpy(&W[ph][4], &hi, 16);
        }
#pragma unroll
        for (int g = 3; g >= 0; --g)
#pragma unroll
          for (int ph = M_ - 1; ph >= 0; --ph) {
            const float hk = tp[(3 - g) * M_ + (M_ - 1 - ph)];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fadd(acc[j], fmul(hk, W[ph][j + g]));
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (t0 + j < p.tile_out && o0 + t0 + j < p.n_out) out_row[o0 + t0 + j] = acc[j];
    }
    return;
  }
  if constexpr (!INTERP && (MT == 2 || MT == 4) && sizeof(T) == 2 && !CLAMP) if (p.N > p.kp) {
    // Q15 decimator, M = 2 or 4, in a tile whose samples cannot reach the clamp: the same blocking as
    // the float path above (four consecutive outputs per thread, blocks of four a-values), the window of
    // a phase row as two 8-byte loads of four int16 each. Without the clamp the sum is an exact integer,
    // so the order of the multiply-adds is free.
    constexpr int M_ = MT;
    const int n_blocks = (int)((p.N - p.kp) / (4 * M_));
    for (uint32_t t0 = 4 * threadIdx.x; t0 < p.tile_out; t0 += 4 * FB_THREADS) {
      if (o0 + t0 >= p.n_out) break;
      int acc[4] = {1 << 14, 1 << 14, 1 << 14, 1 << 14};
      for (uint32_t k = 0; k < p.kp; ++k) {
        const uint32_t r = p.N - 1 - k;
        const int hk = h[k];
        const T *xr = X + (r % M_) * p.pitch + t0 + r / M_;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += hk * (int)xr[j];
      }
      const int *hb = h + p.kp;
      for (int m = n_blocks - 1; m >= 0; --m, hb += 4 * M_) {
        int tp[4 * M_];
#pragma unroll
        for (int v = 0; v < M_; ++v) {
          const uint4 t4 = *reinterpret_cast<const uint4 *>(hb + 4 * v);
          memcpy(&tp[4 * v], &t4, 16);
        }
#pragma unroll
        for (int ph = M_ - 1; ph >= 0; --ph) {
          const uint2 lo = *reinterpret_cast<const uint2 *>(X + ph * p.pitch + t0 + 4 * m);
          const uint2 hi = *reinterpret_cast<const uint2 *>(X + ph * p.pitch + t0 + 4 * m + 4);
          const int W[8] = {(int)(int16_t)lo.x, (int)lo.x >> 16, (int)(int16_t)lo.y, (int)lo.y >> 16,
                            (int)(int16_t)hi.x, (int)hi.x >> 16, (int)(int16_t)hi.y, (int)hi.y >> 16};
#pragma unroll
          for (int g = 3; g >= 0; --g) {
            const int hk = tp[(3 - g) * M_ + (M_ - 1 - ph)];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += hk * W[j + g];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (t0 + j < p.tile_out && o0 + t0 + j < p.n_out) out_row[o0 + t0 + j] = (int16_t)(acc[j] >> 15);
    }
    return;
  }
  // R outputs per thread and pass (t, t + 256, ..): one tap load serves R accumulators. Outputs
  // past the tile or the row read staged zeros / neighbours and are simply not stored.
  constexpr int R = 4;
  for (uint32_t t0 = threadIdx.x; t0 < p.tile_out; t0 += R * FB_THREADS) {
    if (o0 + t0 >= p.n_out) break;
    acc_t acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = A::init();
    if (INTERP) {
      // lanes own consecutive outputs: the same input for L lanes, taps h[i + kL] consecutive in i
      const T *xp[R];
      const tap_t *hp[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const uint32_t t = min(t0 + j * FB_THREADS, p.tile_out - 1);
        xp[j] = X + t / p.F + (p.q - 1);  // newest input of this output, span index
        hp[j] = h + t % p.F;
      }
      for (uint32_t k = 0; k < p.q; ++k) {
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = A::mac(acc[j], hp[j][k * p.F], *(xp[j] - k));
      }
    } else {
      // newest input of output t at span index M t + N-1; tap k reads index M t + (N-1-k),
      // which sits at X[(N-1-k) % M][t + (N-1-k) / M]: consecutive words across lanes
      const uint32_t last = min(t0 + (R - 1) * FB_THREADS, p.tile_out - 1) - t0;  // clamp the last pass
      uint32_t k = 0;
      int32_t ph = (int32_t)((p.N - 1) % M);
      const T *xb = X + t0 + (p.N - 1) / M;
      for (; ph >= 0; --ph, ++k) {  // the partial group of the newest samples
        const tap_t hk = h[k];
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = A::mac(acc[j], hk, xb[(uint32_t)ph * p.pitch + min((uint32_t)j * FB_THREADS, last)]);
      }
      while (k < p.N) {  // full groups of M, one word further back each
        --xb;
#pragma unroll
        for (int32_t f = (int32_t)M - 1; f >= 0; --f, ++k) {
          const tap_t hk = h[k];
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = A::mac(acc[j], hk, xb[(uint32_t)f * p.pitch + min((uint32_t)j * FB_THREADS, last)]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint32_t t = t0 + j * FB_THREADS;
      if (t < p.tile_out && o0 + t < p.n_out) out_row[o0 + t] = A::done(acc[j]);
    }
  }
}

// MT = decimation factor known at compile time (0: use p.F), so the phase loop unrolls
template <class T, bool INTERP, int MT>
__global__ void __launch_bounds__(FB_THREADS) filter_bank_kernel(const __grid_constant__ FilterBankParams p) {
  using tap_t = typename FbAcc<T, true>::tap_t;
  extern __shared__ uint4 fb_smem_raw[];
  tap_t *h = reinterpret_cast<tap_t *>(fb_smem_raw) + p.hoff;
  T *X = reinterpret_cast<T *>(reinterpret_cast<tap_t *>(fb_smem_raw) + p.np);
  const uint32_t row = blockIdx.y;
  const T *in_row = reinterpret_cast<const T *>(p.in) + (uint64_t)row * p.in_stride;
  const T *carry_row = reinterpret_cast<const T *>(p.carry_in) + (uint64_t)row * p.C;
  const uint64_t o0 = (uint64_t)blockIdx.x * p.tile_out;  // first output of this tile
  const uint32_t M = INTERP ? 1u : (MT ? (uint32_t)MT : p.F);

  // first stream sample the tile touches: the oldest tap of its first output
  //   decimator: newest input of output o is M o + M-1 - pending (Decimator.cc:296-316)
  //   interpolator: outputs nL .. nL+L-1 filter inputs n, n-1, .. n-q+1 (Interpolator.cc interpolate)
  const int64_t s0 = INTERP ? (int64_t)(o0 / p.F) - (int64_t)(p.q - 1)
                            : (int64_t)(o0 * M) + (int64_t)(M - 1) - (int64_t)p.pending - (int64_t)(p.N - 1);
  for (uint32_t i = threadIdx.x; i < p.N; i += FB_THREADS) h[i] = reinterpret_cast<const tap_t *>(p.taps)[i];
  if (s0 >= 0 && (uint64_t)s0 + p.span <= p.n_in) {
    // interior tile (all but the first and last of a row): 16-byte loads from the first aligned
    // element on, scalar head and tail
    constexpr uint32_t VEC = 16 / sizeof(T);
    const T *src = in_row + s0;
    const uint32_t head = min(p.span, (uint32_t)(((16 - (uint32_t)((uintptr_t)src & 15)) & 15) / sizeof(T)));
    const uint32_t nvec = (p.span - head) / VEC;
    const uint4 *vsrc = reinterpret_cast<const uint4 *>(src + head);
    for (uint32_t v = threadIdx.x; v < nvec; v += FB_THREADS) {
      const uint4 raw = __ldg(vsrc + v);
      T e[VEC];
      memcpy(e, &raw, 16);
#pragma unroll
      for (uint32_t j = 0; j < VEC; ++j) {
        const uint32_t i = head + v * VEC + j;
        X[(i % M) * p.pitch + i / M] = e[j];
      }
    }
    for (uint32_t i = threadIdx.x; i < head; i += FB_THREADS) X[(i % M) * p.pitch + i / M] = src[i];
    for (uint32_t i = head + nvec * VEC + threadIdx.x; i < p.span; i += FB_THREADS)
      X[(i % M) * p.pitch + i / M] = src[i];
  } else {
    for (uint32_t i = threadIdx.x; i < p.span; i += FB_THREADS)
      X[(i % M) * p.pitch + i / M] = fb_sample(in_row, carry_row, p.C, s0 + i, p.n_in);
  }
  __syncthreads();

  if (sizeof(T) == 2) {
    // largest |x| of the staged span decides whether the clamp can fire anywhere in this tile
    __shared__ uint32_t tile_max;
    if (threadIdx.x == 0) tile_max = 0;
    __syncthreads();
    uint32_t m = 0;
    for (uint32_t i = threadIdx.x; i < p.span; i += FB_THREADS) m = max(m, (uint32_t)abs((int)X[(i % M) * p.pitch + i / M]));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(&tile_max, m);
    __syncthreads();
    if ((uint64_t)p.sum_abs_taps * tile_max + (1u << 14) < (1u << 30))
      fb_compute<T, INTERP, MT, false>(p, h, X, row, o0, M);
    else
      fb_compute<T, INTERP, MT, true>(p, h, X, row, o0, M);
  } else {
    fb_compute<T, INTERP, MT, true>(p, h, X, row, o0, M);
  }

  // the row's next carry: the last C samples of (carry ++ input)
  if (blockIdx.x == 0) {
    T *carry_next = reinterpret_cast<T *>(p.carry_out) + (uint64_t)row * p.C;
    for (uint32_t i = threadIdx.x; i < p.C; i += FB_THREADS)
      carry_next[i] = fb_sample(in_row, carry_row, p.C, (int64_t)p.n_in - (int64_t)p.C + i, p.n_in);
  }
}

}  // namespace sdr
