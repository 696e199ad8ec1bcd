// WBFM kernel, third generation: TWO channels per worker warp.
//
// wbfm_tile2_kernel's round is bound by the recurrence warp: 1024 dependent FMUL -> FSUB steps
// per round whatever the number of channels, and only 14-15 of its 32 lanes have a channel,
// because a channel costs 6 KB of shared memory (2 KB input slot + 4 KB ring for 1024 samples)
// beside the 129 KB atan2 table and a worker warp 128 registers per thread
// (profiles/r01v7_wbfm_ncu.txt: 0.79 barrier stalls per issued instruction).
//
// Here a worker warp serves two channels, lanes 0-15 one and lanes 16-31 the other, each
// advancing HALF a tile (512 samples, 16 lanes x 32) per round. Per round a worker does the same
// work as before, its two channels need the same 6 KB (2 x (1 KB input + 2 KB ring)), but the
// recurrence warp now runs 512 steps for up to 28 channels instead of 1024 steps for 14: the
// serial chain per sample is halved and the round becomes worker-bound.
//
// What changes in the per-lane code is the width of a "tile": the lanes below are the lanes of
// the same 16-lane half (shfl_prev16), the previous half-tile's registers serve lanes near the
// half's first, and the one stage that reaches further back than 16 lanes -- the 40-tap audio
// decimator, 19 lanes -- reads its window from a 48-word ring in shared memory (like NBFM's).
// A, B, C are those of WbTile / WbTile2; the carry blob is the same, so all three kernels are
// interchangeable between calls.
#pragma once
#include "sdr_wbfm2.cuh"

#if SDR_DEVICE_BUILD
namespace sdr {

// value of "virtual lane (l - j)" of the lane's own 16-lane half: lanes below j of the half read the
// previous half-tile's register
template <class T>
__device__ __forceinline__ T shfl_prev16(T cur, T prev, int j, int lane) {
  return __shfl_sync(FULL, (lane & 15) >= 16 - j ? prev : cur, ((lane - j) & 15) | (lane & 16));
}
// after a half-tile with r valid lanes (1..16): the registers of the last 16 lanes of the stream
template <class T>
__device__ __forceinline__ T roll_prev16(T cur, T prev, int r, int lane) {
  return __shfl_sync(FULL, (lane & 15) >= r ? prev : cur, ((lane + r) & 15) | (lane & 16));
}

struct WbTile3 {
  using T1 = WbTile;
  using T2 = WbTile2;
  static constexpr int HALF = 512;                 // samples per channel and round
  static constexpr int MAX_WORKERS = 14;           // 15 warps x 128 registers; 28 channels per CTA
  static constexpr int RING_BYTES = 2048 + 16;     // 16 rows of 32 floats + pad: channel stride == 16 (mod 128)
  static constexpr int ERING_WORDS = 48;           // the last 48 decimator-2 output pairs of a channel
  // mma: the pre-filter on the tensor cores (WbMma): + each channel's raw history and the taps table
  __host__ __device__ static constexpr int smem_bytes(int workers, bool mma = false) {
    return T2::LUT_BYTES + workers * ((mma ? WbMma::area_bytes<true>() : TILE_BYTES) + 2 * RING_BYTES + 2 * ERING_WORDS * 4) + 64 +
           (mma ? WB_TAB_WORDS * 4 : 0);
  }
  static constexpr int REC_WARP = 3;  // as in WbTile2: the recurrence warp's scheduler carries fewer workers

  // A(k) for both halves of the warp: w = the lane's 64 input bytes -> its 32 numerators u
  __device__ __forceinline__ static void part_a(const uint32_t (&w)[16], int fmt, float k, uint32_t lut_s, WbCarry &pv,
                                                float &v_boundary, uint32_t (&u)[32], int lane, int r) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) front_end_group(fmt, w[2 * g], w[2 * g + 1], a[g], b[g]);
    uint32_t ea[12], eb[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ea[i] = shfl_prev16(a[4 + i], pv.a[i], 1, lane);
      eb[i] = shfl_prev16(b[4 + i], pv.b[i], 1, lane);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { ea[4 + i] = a[i]; eb[4 + i] = b[i]; }

    // the lane's LAST four samples first: theta[31] and v[31] depend on this lane's data only,
    // and the lane above needs them before it can start
    float th_last[4];
    T2::theta4<28>(ea, eb, lut_s, th_last);
    const float my_th31 = th_last[3];
    const float my_v31 = fmul(k, wrap_pi_table(fsub(th_last[3], th_last[2])));
    float th_prev = shfl_prev16(my_th31, pv.th31, 1, lane);
    float v_prev = __shfl_up_sync(FULL, my_v31, 1, 16);
    if ((lane & 15) == 0) v_prev = v_boundary;
    T2::u_chunks<0>(ea, eb, lut_s, k, th_prev, v_prev, u);
    T2::u4(th_last, k, th_prev, v_prev, &u[28]);

    v_boundary = __shfl_sync(FULL, my_v31, (lane & 16) | (r - 1));
    if (r == 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { pv.a[i] = a[4 + i]; pv.b[i] = b[4 + i]; }
      pv.th31 = my_th31;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pv.a[i] = roll_prev16(a[4 + i], pv.a[i], r, lane);
        pv.b[i] = roll_prev16(b[4 + i], pv.b[i], r, lane);
      }
      pv.th31 = roll_prev16(my_th31, pv.th31, r, lane);
    }
  }

  // C(j): dW = the lane's 32 de-emphasised samples; er = the channel's 48-word ring (its first
  // 32 words: the decimator-2 outputs before this half-tile). Returns the lane's PCM sample.
  __device__ __forceinline__ static int part_c(const uint32_t (&dW)[16], WbCarry &pv, uint32_t *er, int lane, int r,
                                               bool &big_b) {
    const int l16 = lane & 15;
    // decimator 1: 8 taps, 4:1; output m uses d[4m-4 .. 4m+3]
    uint32_t ext[18];
    ext[0] = shfl_prev16(dW[14], pv.dw[0], 1, lane);
    ext[1] = shfl_prev16(dW[15], pv.dw[1], 1, lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) ext[2 + i] = dW[i];
    uint32_t e1w[4];
    e1w[0] = pack_i16x2(T1::dec1_one<0>(ext), T1::dec1_one<1>(ext));
    e1w[1] = pack_i16x2(T1::dec1_one<2>(ext), T1::dec1_one<3>(ext));
    e1w[2] = pack_i16x2(T1::dec1_one<4>(ext), T1::dec1_one<5>(ext));
    e1w[3] = pack_i16x2(T1::dec1_one<6>(ext), T1::dec1_one<7>(ext));
    // decimator 2: 12 taps, 4:1, clamp-free (decimator 1 output <= 29126 <= FM_POST::SAFE)
    uint32_t de[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      de[i] = shfl_prev16(e1w[i], pv.e1w[i], 1, lane);
      de[4 + i] = e1w[i];
    }
    const uint32_t w0[6] = {de[0], de[1], de[2], de[3], de[4], de[5]};
    const uint32_t w1[6] = {de[2], de[3], de[4], de[5], de[6], de[7]};
    const int e0 = (int)(int16_t)(fir_s16_fast<taps::FM_POST, 11, 6>(w0) >> 15);
    const int e1 = (int)(int16_t)(fir_s16_fast<taps::FM_POST, 11, 6>(w1) >> 15);
    const uint32_t ew = pack_i16x2(e0, e1);
    // the clamp-path flag is per channel: votes stay inside the 16-lane half
    const unsigned half_mask = 0xffffu << (lane & 16);
    const bool cur_b = (__ballot_sync(FULL, (iabs(e0) > taps::AUDIO40::SAFE || iabs(e1) > taps::AUDIO40::SAFE) && l16 < r) &
                        half_mask) != 0;
    const bool exact_b = cur_b || big_b;
    // audio decimator: 40 taps, 2:1: the lane's pair and the 19 pairs before it, from the ring
    er[32 + l16] = ew;
    __syncwarp();
    uint32_t ee[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) ee[i] = er[13 + l16 + i];
    int acc;
    if (!__any_sync(FULL, exact_b)) {
      acc = fir_s16_fast<taps::AUDIO40, 39, 20>(ee);
    } else {
      acc = fir_s16<taps::AUDIO40, 39, 20>(ee, exact_b);
    }
    const int pcm = (int)(int16_t)(acc >> 15);
    big_b = cur_b || (r < 16 && big_b);
    // the ring's first 32 words become the 32 pairs that end with this half-tile's last valid one
    __syncwarp();
    const uint32_t m0 = er[l16 + r], m1 = er[16 + l16 + r];
    __syncwarp();
    er[l16] = m0;
    er[16 + l16] = m1;
    if (r == 16) {
      pv.dw[0] = dW[14]; pv.dw[1] = dW[15];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv.e1w[i] = e1w[i];
    } else {
      pv.dw[0] = roll_prev16(dW[14], pv.dw[0], r, lane);
      pv.dw[1] = roll_prev16(dW[15], pv.dw[1], r, lane);
#pragma unroll
      for (int i = 0; i < 4; ++i) pv.e1w[i] = roll_prev16(e1w[i], pv.e1w[i], r, lane);
    }
    return pcm;
  }

  // The carry blob keeps the registers of the last 32 lanes of the stream (WbTile::load_carry).
  // Of those, everything but the decimator-2 pairs is only ever read one lane back, so a half
  // takes the blob's lanes 16-31 as its "previous half-tile" and the pairs of all 32 lanes
  // into its ring.
  __device__ __forceinline__ static void load_carry(WbCarry &c, const uint32_t *blob, uint32_t *er, int lane) {
    const int l16 = lane & 15, src = 16 + l16;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c.a[i] = blob[i * 32 + src]; c.b[i] = blob[(4 + i) * 32 + src]; }
    c.th31 = u2f(blob[8 * 32 + src]);
    c.v31 = 0.f;
    c.dw[0] = blob[10 * 32 + src]; c.dw[1] = blob[11 * 32 + src];
#pragma unroll
    for (int i = 0; i < 4; ++i) c.e1w[i] = blob[(12 + i) * 32 + src];
    c.ew = 0;
    er[l16] = blob[16 * 32 + l16];
    er[16 + l16] = blob[16 * 32 + 16 + l16];
  }
  __device__ __forceinline__ static void store_carry(const WbCarry &c, uint32_t *blob, const uint32_t *er, int lane) {
    const int l16 = lane & 15, dst = 16 + l16;
#pragma unroll
    for (int i = 0; i < 4; ++i) { blob[i * 32 + dst] = c.a[i]; blob[(4 + i) * 32 + dst] = c.b[i]; }
    blob[8 * 32 + dst] = f2u(c.th31);
    blob[9 * 32 + dst] = 0;
    blob[10 * 32 + dst] = c.dw[0]; blob[11 * 32 + dst] = c.dw[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) blob[(12 + i) * 32 + dst] = c.e1w[i];
    // lanes 0-15 of the blob: nothing reads them but the pairs
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i != 9) blob[i * 32 + l16] = 0;
    blob[9 * 32 + l16] = 0;
    blob[16 * 32 + l16] = er[l16];
    blob[16 * 32 + 16 + l16] = er[16 + l16];
  }
};

// blockDim = 32 * (workers + 1), p.G = channels per CTA = 2 * workers. Warp p.aux runs the
// recurrences (lane == channel slot), the others are workers (two channel slots each). All roles
// share one round loop and meet the same two barrier instructions.
// MMA: the pre-filter of full half-tiles of u8 input runs on the tensor cores (sdr_wbfm_mma.cuh); the
// pre-filter history then lives as raw bytes in shared memory instead of WbCarry::a, b.
template <bool MMA>
__global__ void __launch_bounds__(32 * (WbTile3::MAX_WORKERS + 1), 1) wbfm_tile3_kernel(const __grid_constant__ LaunchParams p) {
  using T = WbTile3;
  using T1 = WbTile;
  using T2 = WbTile2;
  extern __shared__ uint4 smem_raw[];
  char *smem = reinterpret_cast<char *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (int)p.G / 2;  // worker warps
  const int rec = (int)p.aux;
  const bool is_iir = warp == rec;
  const int widx = warp < rec ? warp : warp - 1;
  const bool is_worker = !is_iir && widx < nw;
  const uint32_t list0 = blockIdx.x * p.G;
  const int n_here = (int)min(p.G, p.n_list - list0);
  const uint32_t n_half = (p.n_samples + T::HALF - 1) / T::HALF;
  // a worker's input area: 2 KB of windows, or (MMA) WbMma's layout with a history area in front of each channel's
  constexpr int AREA = MMA ? WbMma::area_bytes<true>() : TILE_BYTES;
  char *in_base = smem + T2::LUT_BYTES;
  char *ring_base = in_base + nw * AREA;
  uint32_t *ering_base = reinterpret_cast<uint32_t *>(ring_base + 2 * nw * T::RING_BYTES);
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(smem);
  char *tab_base = reinterpret_cast<char *>(ering_base + 2 * nw * T::ERING_WORDS);  // MMA: wb_mma_table()

  // the table: 129 KB from L2 once per CTA
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.lut);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < T2::LUT_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    if constexpr (MMA) {
      const uint4 *tsrc = reinterpret_cast<const uint4 *>(p.tab);
      uint4 *tdst = reinterpret_cast<uint4 *>(tab_base);
      for (int i = threadIdx.x; i < WB_TAB_WORDS / 4; i += blockDim.x) tdst[i] = __ldg(tsrc + i);
      // (the history areas' 32 bytes nobody writes are never read either)
    }
  }

  // channel slot of this lane: a worker's lanes 0-15 serve slot 2 widx, lanes 16-31 slot 2 widx + 1;
  // the recurrence warp's lane l serves slot l
  const int slot_id = is_iir ? lane : 2 * widx + (lane >> 4);
  const bool owned = (is_iir || is_worker) && slot_id < n_here;
  const uint32_t ch = owned ? p.chan_ids[list0 + slot_id] : 0;
  const bool active = owned && !(p.allowed && !p.allowed[ch]);  // a squelched channel is skipped
  uint32_t *blob = reinterpret_cast<uint32_t *>(p.state + (uint64_t)ch * p.state_stride);
  // a worker's two slots always exist (an unowned or squelched half computes on whatever its slot
  // holds and stores nothing); the recurrence warp only touches the rings of its active lanes
  char *ring = ring_base + (is_worker || owned ? slot_id : 0) * T::RING_BYTES;
  uint32_t *er = ering_base + (is_worker ? 2 * widx + (lane >> 4) : 0) * T::ERING_WORDS;
  const int l16 = lane & 15;
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab_base);

  // ---- worker state ----
  WbCarry pv;
  const uint8_t *src = p.iq + (uint64_t)ch * p.ch_stride;
  int16_t *out = p.pcm + (uint64_t)ch * p.pcm_stride;
  char *area = in_base + (is_worker ? widx : 0) * AREA;
  // where tile_read's "window = lane" numbering starts for this lane, and (MMA) the lane's channel's raw
  // pre-filter history in front of the channel's windows
  char *in_slot = MMA ? area + WbMma::window_base<true>(0) + (lane >> 4) * WB_HIST_AREA : area;
  char *hist = area + (lane >> 4) * (WB_HIST_AREA + TILE_BYTES / 2);
  float k = 0.f, v_boundary = 0.f;
  bool big_b = false, no_patch = true;
  uint32_t u[32] = {};
  // ---- recurrence state ----
  float y1 = 0.f;
  const float a1 = (float)(-0.9492274);

  // the warp's 2 KB input slot: chunks 0-63 from the first channel's stream, 64-127 from the
  // second's, so that tile_read hands lanes 0-15 and 16-31 their channel's 64 bytes each
  auto fetch_half = [&](uint32_t t) {
    const uint32_t s0 = t * T::HALF;
    const int valid = active ? (int)min((uint32_t)T::HALF, p.n_samples - s0) >> 3 : 0;  // 16-byte chunks of this channel
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      // chunk (32 j + 2 l16 + {0, 1}) would split a lane's pieces over both halves; instead each lane
      // fills four chunks of its own channel: c = l16 + 16 i, i = 0..3
      const int c0 = l16 + 32 * j, c1 = l16 + 16 + 32 * j;
      const int q0 = 64 * (lane >> 4) + c0, q1 = 64 * (lane >> 4) + c1;
      if (c0 < valid) cp_async16(in_slot + 16 * tile_slot(q0), src + (uint64_t)s0 * 2 + 16 * c0);
      if (c1 < valid) cp_async16(in_slot + 16 * tile_slot(q1), src + (uint64_t)s0 * 2 + 16 * c1);
    }
  };

  if (is_worker) {
    // the ring of a slot nobody owns is never read for output, but keep it defined
    er[l16] = 0; er[16 + l16] = 0; er[32 + l16] = 0;
    if (active) {
      T::load_carry(pv, blob, er, lane);
      v_boundary = u2f(blob[T1::NREG * 32 + 1]);
      big_b = blob[T1::NREG * 32 + 2] != 0;
      k = p.scale[ch];
      // |y| <= max(|y[-1]|, |u|max / (1 - |a1|)) < 3.2 |k|: with |k| < 1e8 and |y[-1]| < 1e9 no
      // value can reach 2^31, where cvt.rzi (saturating) and x86 cvttss2si (wrapping) differ
      no_patch = fabsf(k) < 1e8f && fabsf(u2f(blob[T1::NREG * 32])) < 1e9f;
    } else {
      pv = WbCarry{};
    }
    no_patch = __all_sync(FULL, no_patch);
    if constexpr (MMA) {
      if (l16 == 15) WbMma::history_from_planes(hist, p.fmt, pv);  // a dead channel's: 0x80 bytes
    }
    fetch_half(0);
    cp_async_commit();
  } else if (is_iir && active) {
    y1 = u2f(blob[T1::NREG * 32]);
  }
  __syncthreads();  // table in place

  for (uint32_t kk = 0; kk < n_half + 2; ++kk) {
    // ---- phase 1: hand-over through the channel's slot: y(kk-2) out, u(kk-1) in ----
    if (is_worker) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const u32x4 v = lds_u4(ring + T1::u_off(l16, j));
        sts_u4(ring + T1::u_off(l16, j), u32x4{u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]});
        u[4 * j] = v.x; u[4 * j + 1] = v.y; u[4 * j + 2] = v.z; u[4 * j + 3] = v.w;
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    if (is_worker) {
      if (kk >= 2) {
        const uint32_t t = kk - 2;
        const int r = (int)min((uint32_t)T::HALF, p.n_samples - t * T::HALF) >> 5;
        uint32_t dW[16];  // (int16_t)y of the half-tile the recurrence finished last round
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (no_patch)
            dW[j] = __byte_perm((uint32_t)f2i_rz(u2f(u[2 * j])), (uint32_t)f2i_rz(u2f(u[2 * j + 1])), 0x5410);
          else
            dW[j] = f2i16x2_wrap(u2f(u[2 * j]), u2f(u[2 * j + 1]));
        }
        const int pcm = T::part_c(dW, pv, er, lane, r, big_b);
        if (active && l16 < r) out[(uint64_t)t * 16 + l16] = (int16_t)pcm;
      }
      if (kk < n_half) {
        cp_async_wait<0>();
        __syncwarp();
        const int r = (int)min((uint32_t)T::HALF, p.n_samples - kk * T::HALF) >> 5;
        bool pairs = false;  // the slot holds the pre-filter's outputs instead of raw samples
        if constexpr (MMA) {
          if (r == 16 && p.fmt == FMT_U8_OFFSET_ROTATE)
            pairs = WbMma::prefilter<true>((uint32_t)__cvta_generic_to_shared(area), tab_s, lane, active);
          __syncwarp();
          // diagnostics: [1] = tiles whose pre-filter ran on the tensor cores, [2] = on the CUDA cores
          const bool warp_live = __any_sync(FULL, active);
          if (lane == 0 && p.counters && warp_live) atomicAdd(p.counters + (pairs ? 1 : 2), 1u);
        }
        uint32_t w[16];
        tile_read(in_slot, lane, w);
        __syncwarp();
        if (kk + 1 < n_half) fetch_half(kk + 1);  // the input slot is free again
        cp_async_commit();
        if (MMA && pairs) {
          WbMma::part_a<true>(w, k, lut_s, pv, v_boundary, u, lane);
        } else {
          if constexpr (MMA) WbMma::planes_from_history(hist, p.fmt, pv);
          T::part_a(w, p.fmt, k, lut_s, pv, v_boundary, u, lane, r);
          if constexpr (MMA) {
            __syncwarp();
            if (l16 == r - 1) WbMma::history_from_window(hist, w);
            __syncwarp();
          }
        }
      }
    } else if (is_iir && active && kk >= 1 && kk <= n_half) {
      // B(kk-1): y[n] = fl(u[n] - fl(a1 * y[n-1])) in place, lane == channel (IirFilter.cc:161-176)
      const uint32_t t = kk - 1;
      const int r = (int)min((uint32_t)T::HALF, p.n_samples - t * T::HALF) >> 5;
      if (r == 16) {
        // full half-tile: eight rows per iteration, so the swizzle (row & 7) is a compile-time
        // constant and every shared address is base + immediate; only a row's FIRST chunk is
        // fetched ahead (see wbfm_tile2_kernel)
        u32x4 first = lds_u4(ring);
        for (int row0 = 0; row0 < 16; row0 += 8) {
          char *base = ring + 128 * row0;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            u32x4 v[8];
            char *row = base + 128 * rr;
            v[0] = first;
#pragma unroll
            for (int j = 1; j < 8; ++j) v[j] = lds_u4(row + 16 * (j ^ rr));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float y0 = fsub(u2f(v[j].x), fmul(a1, y1));
              const float y2 = fsub(u2f(v[j].y), fmul(a1, y0));
              const float y3 = fsub(u2f(v[j].z), fmul(a1, y2));
              y1 = fsub(u2f(v[j].w), fmul(a1, y3));
              sts_u4(row + 16 * (j ^ rr), u32x4{f2u(y0), f2u(y2), f2u(y3), f2u(y1)});
              // next row's chunk 0 (row 16 = the pad behind the slot for the last row: read, never used)
              if (j == 3) first = lds_u4(row + 128 + 16 * ((rr + 1) & 7));
            }
          }
        }
      } else {
        for (int row = 0; row < r; ++row) {
          u32x4 v[8];
          T2::chain_load(ring + 128 * row, row & 7, v);
          T2::chain_run(ring + 128 * row, row & 7, v, a1, y1);
        }
      }
    }
    __syncthreads();
  }

  if (active) {
    if (is_worker) {
      if constexpr (MMA) WbMma::planes_from_history(hist, p.fmt, pv);
      T::store_carry(pv, blob, er, lane);  // a lane reads back the ring words it wrote itself
      if (l16 == 0) {
        blob[T1::NREG * 32 + 1] = f2u(v_boundary);
        blob[T1::NREG * 32 + 2] = big_b;
      }
    } else {
      blob[T1::NREG * 32] = f2u(y1);
    }
  }
}

}  // namespace sdr
#endif
