// WBFM kernel, second generation: the atan2 table lives in SHARED memory.
//
// wbfm_tile_kernel (sdr_tile.cuh) is bound by the scattered atan2 gathers: one lookup per
// sample into a 256 KB table that cannot stay in the few KB of L1 left beside 195 KB of rings,
// ~26 L1 tag cycles per warp-wide gather, a round of 19 x 1024 samples = 19 x 32 x 26 =
// 15,800 tag cycles against a measured round of ~19,000 cycles (profiles/r01v3_wbfm_ncu.txt).
// Shared memory serves a 32-address gather in ~3-4 cycles (bank conflicts only).
//
// To fit:
//   * The table is the half plane q >= 0: 129 rows x 256 columns of float = 129 KB. atan2 is
//     odd in q for every table entry ((float)atan2(-q, i) == -(float)atan2(q, i), checked on
//     the host when the table is built), so theta = sign(q) * T[|q|][i]. Column c holds
//     i = (int8_t)c, so the pre-filter's output byte indexes it without the +128 of
//     WbFmDemodulator.cc:458-459.
//   * One 4 KB ring slot per channel instead of two: the worker keeps the tile's 32 IIR
//     numerators per lane in REGISTERS until the slot is free. A round has two phases:
//       1 (short)  worker: y(k-2) slot -> registers (as int16 pairs), then u(k-1) registers -> slot
//       2          recurrence warp: B(k-1) in place, lane == channel (IirFilter.cc:161-176);
//                  worker: C(k-2) from its registers -> PCM, then A(k) -> registers
//     with a CTA barrier after each. 15 channels per CTA (2 KB input slot + 4 KB ring each).
//   * The +-pi wrap uses wrap_pi_table (no FP64, exact for table values).
// A, B, C are those of WbTile; the carry blob is the same, so the two kernels are interchangeable
// between calls.
#pragma once
#include "sdr_tile.cuh"
#include "sdr_wbfm_mma.cuh"

#if SDR_DEVICE_BUILD
#ifndef SDR_WB_FP32_PREFILTER
#define SDR_WB_FP32_PREFILTER 0  // 1: the pre-filter on FFMA2 (A/B builds; exact, measured slower, see theta_fp32)
#endif
#ifndef SDR_WB_FP32_SPLIT
#define SDR_WB_FP32_SPLIT 2      // partial sums per pre-filter output (1, 2 or 4)
#endif
namespace sdr {

struct WbTile2 {
  using T1 = WbTile;
  static constexpr int LUT_ROWS = 129, LUT_BYTES = LUT_ROWS * 256 * 4;
  static constexpr int MAX_WORKERS = 15;   // 512 threads x 128 registers; 14 is the default (see warps_for)
  static constexpr int RING_BYTES = 4096 + 16;  // one slot + pad: channel stride == 16 (mod 128)
  // mma: the pre-filter on the tensor cores (WbMma): + each channel's raw history and the taps table
  __host__ __device__ static constexpr int smem_bytes(int nw, bool mma = false) {
    return LUT_BYTES + nw * ((mma ? WbMma::area_bytes<false>() : TILE_BYTES) + RING_BYTES) + 64 + (mma ? WB_TAB_WORDS * 4 : 0);
  }
  // The recurrence warp's dependent FMUL -> FSUB chain is the round's critical path (measured:
  // the round time does not depend on the number of workers between 12 and 14). It is warp
  // REC_WARP = 3, so that with 14 workers (15 warps) its scheduler carries two workers and the
  // other three carry four each; as the first or the last warp it shares a scheduler with three
  // workers and the kernel is 9-13 % slower (profiles/r01v7_wbfm2_sweep.txt).
  static constexpr int REC_WARP = 3;
  __host__ __device__ static constexpr int warps_for(int workers) { return workers + 1; }

  // theta of sample N from the shared half-plane table. lut_s = shared byte address of T[0][0].
  //   x   byte 0 = i (int8), byte 1 = q (int8): byte 2 of the doubled accumulators (WbTile::Pre2)
  //   sg  0xFF00FF00 where q < 0
  //   row |q| = (q ^ s) - s: the xor on byte 1, the +1 as one more row (+0x400 bytes)
  template <int N>
  __device__ __forceinline__ static float theta(const uint32_t (&ea)[12], const uint32_t (&eb)[12], uint32_t lut_s) {
    const uint32_t ai = (uint32_t)fir_s8<T1::Pre2, 16 + N, 12>(ea, 1 << 15);
    const uint32_t aq = (uint32_t)fir_s8<T1::Pre2, 16 + N, 12>(eb, 1 << 15);
    const uint32_t x = __byte_perm(ai, aq, 0x7762);
    const uint32_t sg = prmt_sx(aq, 0, 0xA4A4);
    const uint32_t v = __byte_perm(x ^ (sg & 0xff00u), 0, 0x4410);
    const uint32_t addr = lut_s + (v << 2) + (sg & 0x400u);
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
    return u2f(t ^ (sg & 0x80000000u));
  }
  // ---- the same pre-filter on the FP32 pipe (-DSDR_WB_FP32_PREFILTER=1; NOT the default) ----
  // Measured (profiles/r02_wbfm_fp32_prefilter.txt): bit-exact, and slower -- WBFM x8192 343 -> 290 G S/s
  // with one chain of 16 FFMA2 per output, 284 / 274 with two / four partial sums (so it is not the chains'
  // latency): 132 I2F per tile queue on the quarter-rate XU pipe behind the shared-memory traffic (ncu:
  // mio_throttle 0.48 -> 0.84 per issue, wait 0.88 -> 1.64) and the kernel executes 7 % more instructions.
  // IDP.2A runs on the integer datapath, half the FP32 rate, and the workers are bound by it: 634 of
  // 2,068 warp-instructions per 1024 samples and a third of all stall samples (profiles/r02_wbfm3_ncu.txt).
  // The 16-tap pre-filter is exact in FP32: with the taps scaled by 2^-15 (exact) every partial sum is
  // a multiple of 2^-16 below 256 -- 24 bits -- so FFMA2 (two FP32 FMAs per lane and instruction: the I
  // and the Q arm as one register pair, the tap an immediate) adds up without rounding whatever the
  // order. The accumulator starts at 2^-16 instead of the rounding constant 1/2: then
  // floor(1/2 + sum) == round-to-nearest(2^-16 + sum) (the fraction is an odd multiple of 2^-16: never
  // a tie), and one FADD2 of 1.5 * 2^23 leaves that integer in the low mantissa bits, its low byte
  // being the reference's (int8_t) truncation (WbFmDemodulator.cc:393-397). The samples become floats
  // with I2F.S8 straight from the packed bytes (XU pipe, otherwise idle here).
  template <int S>  // sample S of the lane's window, -16 .. 31, as the pair (I', Q')
  __device__ __forceinline__ static unsigned long long xpair(const uint32_t (&ea)[12], const uint32_t (&eb)[12]) {
    constexpr int w = (S + 16) >> 2, by = (S + 16) & 3;
    const float fi = (float)(int8_t)(ea[w] >> (8 * by)), fq = (float)(int8_t)(eb[w] >> (8 * by));
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(fi), "f"(fq));
    return r;
  }
  template <int N, int K, int KEND>
  __device__ __forceinline__ static void pre_fp32(const uint32_t (&ea)[12], const uint32_t (&eb)[12], unsigned long long &acc) {
    if constexpr (K < KEND) {
      constexpr float h = (float)taps::WB_PRE::tap(K) * (1.0f / 32768.0f);
      unsigned long long hh;
      asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xpair<N - K>(ea, eb)), "l"(hh));
      pre_fp32<N, K + 1, KEND>(ea, eb, acc);
    }
  }
  template <int N>
  __device__ __forceinline__ static float theta_fp32(const uint32_t (&ea)[12], const uint32_t (&eb)[12], uint32_t lut_s) {
    // SDR_WB_FP32_SPLIT partial sums per output (every order of summation is exact): shorter dependent chains
    unsigned long long acc, part[SDR_WB_FP32_SPLIT];
    asm("mov.b64 %0, {%1, %1};" : "=l"(part[0]) : "f"(1.52587890625e-05f));   // 2^-16
    constexpr int STEP = 16 / SDR_WB_FP32_SPLIT;
    pre_fp32<N, 0, STEP>(ea, eb, part[0]);
#pragma unroll
    for (int i = 1; i < SDR_WB_FP32_SPLIT; ++i) asm("mov.b64 %0, {%1, %1};" : "=l"(part[i]) : "f"(0.0f));
    if constexpr (SDR_WB_FP32_SPLIT >= 2) pre_fp32<N, STEP, 2 * STEP>(ea, eb, part[1]);
    if constexpr (SDR_WB_FP32_SPLIT >= 4) {
      pre_fp32<N, 2 * STEP, 3 * STEP>(ea, eb, part[2]);
      pre_fp32<N, 3 * STEP, 4 * STEP>(ea, eb, part[3]);
      asm("add.rn.f32x2 %0, %0, %1;" : "+l"(part[0]) : "l"(part[2]));
      asm("add.rn.f32x2 %0, %0, %1;" : "+l"(part[1]) : "l"(part[3]));
    }
    acc = part[0];
    if constexpr (SDR_WB_FP32_SPLIT >= 2) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(part[1]));
    unsigned long long magic;
    asm("mov.b64 %0, {%1, %1};" : "=l"(magic) : "f"(12582912.0f));        // 1.5 * 2^23
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(magic));
    uint32_t ri, rq;  // low byte = i, q of the pre-filter's int8 output
    asm("mov.b64 {%0, %1}, %2;" : "=r"(ri), "=r"(rq) : "l"(acc));
    const uint32_t x = __byte_perm(ri, rq, 0x0040);
    const uint32_t sg = prmt_sx(rq, 0, 0x8484);  // 0xFF00FF00 where q < 0
    const uint32_t v = __byte_perm(x ^ (sg & 0xff00u), 0, 0x4410);
    const uint32_t addr = lut_s + (v << 2) + (sg & 0x400u);
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
    return u2f(t ^ (sg & 0x80000000u));
  }
  template <int N0>
  __device__ __forceinline__ static void theta4(const uint32_t (&ea)[12], const uint32_t (&eb)[12], uint32_t lut_s,
                                                float (&th)[4]) {
#if SDR_WB_FP32_PREFILTER
    th[0] = theta_fp32<N0>(ea, eb, lut_s);
    th[1] = theta_fp32<N0 + 1>(ea, eb, lut_s);
    th[2] = theta_fp32<N0 + 2>(ea, eb, lut_s);
    th[3] = theta_fp32<N0 + 3>(ea, eb, lut_s);
#else
    th[0] = theta<N0>(ea, eb, lut_s);
    th[1] = theta<N0 + 1>(ea, eb, lut_s);
    th[2] = theta<N0 + 2>(ea, eb, lut_s);
    th[3] = theta<N0 + 3>(ea, eb, lut_s);
#endif
  }
  __device__ __forceinline__ static void u4(const float (&th)[4], float k, float &th_prev, float &v_prev, uint32_t *u) {
    wb_u4(th, k, th_prev, v_prev, u);
  }
  template <int J>
  __device__ __forceinline__ static void u_chunks(const uint32_t (&ea)[12], const uint32_t (&eb)[12], uint32_t lut_s,
                                                  float k, float &th_prev, float &v_prev, uint32_t (&u)[32]) {
    if constexpr (J < 7) {
      float th[4];
      theta4<4 * J>(ea, eb, lut_s, th);
      u4(th, k, th_prev, v_prev, &u[4 * J]);
      u_chunks<J + 1>(ea, eb, lut_s, k, th_prev, v_prev, u);
    }
  }

  // B, one row of one channel per lane: y[n] = fl(u[n] - fl(a1 * y[n-1])) in place
  // (IirFilter.cc:161-176). row_base = the row's first byte, x = row & 7 (chunk j sits at j ^ x).
  __device__ __forceinline__ static void chain_load(const char *row_base, int x, u32x4 (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = lds_u4(row_base + 16 * (j ^ x));
  }
  __device__ __forceinline__ static void chain_run(char *row_base, int x, const u32x4 (&v)[8], float a1, float &y1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y0 = fsub(u2f(v[j].x), fmul(a1, y1));
      const float y2 = fsub(u2f(v[j].y), fmul(a1, y0));
      const float y3 = fsub(u2f(v[j].z), fmul(a1, y2));
      y1 = fsub(u2f(v[j].w), fmul(a1, y3));
      sts_u4(row_base + 16 * (j ^ x), u32x4{f2u(y0), f2u(y2), f2u(y3), f2u(y1)});
    }
  }

  // A(k): w = the lane's 64 input bytes -> the lane's 32 numerators u (registers)
  __device__ __forceinline__ static void part_a(const uint32_t (&w)[16], int fmt, float k, uint32_t lut_s, WbCarry &pv,
                                                float &v_boundary, uint32_t (&u)[32], int lane, int r) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) front_end_group(fmt, w[2 * g], w[2 * g + 1], a[g], b[g]);
    uint32_t ea[12], eb[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ea[i] = shfl_prev(a[4 + i], pv.a[i], 1, lane);
      eb[i] = shfl_prev(b[4 + i], pv.b[i], 1, lane);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { ea[4 + i] = a[i]; eb[4 + i] = b[i]; }

    // the lane's LAST four samples first: theta[31] and v[31] depend on this lane's data only,
    // and the lane above needs them before it can start
    float th_last[4];
    theta4<28>(ea, eb, lut_s, th_last);
    const float my_th31 = th_last[3];
    const float my_v31 = fmul(k, wrap_pi_table(fsub(th_last[3], th_last[2])));
    float th_prev = shfl_prev(my_th31, pv.th31, 1, lane);
    float v_prev = __shfl_up_sync(FULL, my_v31, 1);
    if (lane == 0) v_prev = v_boundary;
    u_chunks<0>(ea, eb, lut_s, k, th_prev, v_prev, u);
    u4(th_last, k, th_prev, v_prev, &u[28]);

    v_boundary = __shfl_sync(FULL, my_v31, r - 1);
    if (r == 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { pv.a[i] = a[4 + i]; pv.b[i] = b[4 + i]; }
      pv.th31 = my_th31;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pv.a[i] = roll_prev(a[4 + i], pv.a[i], r, lane);
        pv.b[i] = roll_prev(b[4 + i], pv.b[i], r, lane);
      }
      pv.th31 = roll_prev(my_th31, pv.th31, r, lane);
    }
  }
};

// blockDim = 32 * WbTile2::warps_for(p.G). The last warp runs the recurrences (lane == channel
// slot), warps 0 .. G-1 are workers. All roles share one round loop and meet the same two barrier
// instructions.
// MMA: the pre-filter of full tiles of u8 input runs on the tensor cores (sdr_wbfm_mma.cuh); the pre-filter
// history then lives as raw bytes in shared memory instead of WbCarry::a, b.
template <bool MMA>
__global__ void __launch_bounds__(32 * (WbTile2::MAX_WORKERS + 1), 1) wbfm_tile2_kernel(const __grid_constant__ LaunchParams p) {
  using T = WbTile2;
  using T1 = WbTile;
  extern __shared__ uint4 smem_raw[];
  char *smem = reinterpret_cast<char *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (int)p.G;
  const int rec = (int)p.aux;  // which warp runs the recurrences
  const bool is_iir = warp == rec;
  const int widx = warp < rec ? warp : warp - 1;  // worker index
  const bool is_worker = !is_iir && widx < nw;
  const uint32_t list0 = blockIdx.x * (uint32_t)nw;
  const int n_here = (int)min((uint32_t)nw, p.n_list - list0);
  const uint32_t n_tiles = (p.n_samples + TILE - 1) / TILE;
  // a worker's input area: 2 KB of windows, or (MMA) WbMma's layout with the history area in front
  constexpr int AREA = MMA ? WbMma::area_bytes<false>() : TILE_BYTES;
  char *in_base = smem + T::LUT_BYTES;
  char *ring_base = in_base + nw * AREA;
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(smem);
  char *tab_base = ring_base + nw * T::RING_BYTES;  // MMA: wb_mma_table() (every size before it is a multiple of 16)

  // the table: 129 KB from L2 once per CTA
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.lut);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < T::LUT_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    if constexpr (MMA) {
      const uint4 *tsrc = reinterpret_cast<const uint4 *>(p.tab);
      uint4 *tdst = reinterpret_cast<uint4 *>(tab_base);
      for (int i = threadIdx.x; i < WB_TAB_WORDS / 4; i += blockDim.x) tdst[i] = __ldg(tsrc + i);
    }
  }

  const int slot_id = is_iir ? lane : widx;  // channel slot in this CTA
  const bool owned = (is_iir || is_worker) && slot_id < n_here;
  const uint32_t ch = owned ? p.chan_ids[list0 + slot_id] : 0;
  const bool active = owned && !(p.allowed && !p.allowed[ch]);  // a squelched channel is skipped
  uint32_t *blob = reinterpret_cast<uint32_t *>(p.state + (uint64_t)ch * p.state_stride);
  char *ring = ring_base + (active ? slot_id : 0) * T::RING_BYTES;

  // ---- worker state ----
  WbCarry pv;
  const uint8_t *src = p.iq + (uint64_t)ch * p.ch_stride;
  int16_t *out = p.pcm + (uint64_t)ch * p.pcm_stride;
  char *area = in_base + (active && is_worker ? slot_id : 0) * AREA;
  char *in_slot = MMA ? area + WbMma::window_base<false>(0) : area;
  char *hist = area;  // MMA: the channel's raw pre-filter history, in front of its windows
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab_base);
  float k = 0.f, v_boundary = 0.f;
  bool big_b = false, no_patch = false;
  // phase 2 -> phase 1: numerators of the tile computed this round, waiting for the slot;
  // phase 1 -> phase 2: y of the tile the recurrence finished last round (float bits)
  uint32_t u[32] = {};
  // ---- recurrence state ----
  float y1 = 0.f;
  const float a1 = (float)(-0.9492274);

  if (is_worker) {
    if (active) {
      T1::load_carry(pv, blob, lane);
      v_boundary = u2f(blob[T1::NREG * 32 + 1]);
      big_b = blob[T1::NREG * 32 + 2] != 0;
      k = p.scale[ch];
      // |y| <= max(|y[-1]|, |u|max / (1 - |a1|)) < 3.2 |k|: with |k| < 1e8 and |y[-1]| < 1e9 no
      // value can reach 2^31, where cvt.rzi (saturating) and x86 cvttss2si (wrapping) differ
      no_patch = fabsf(k) < 1e8f && fabsf(u2f(blob[T1::NREG * 32])) < 1e9f;
      if constexpr (MMA) {
        if (lane == 31) WbMma::history_from_planes(hist, p.fmt, pv);
      }
      tile_fill(in_slot, src, lane, (int)min((uint32_t)TILE, p.n_samples) >> 3);
    }
    cp_async_commit();
  } else if (is_iir && active) {
    y1 = u2f(blob[T1::NREG * 32]);
  }
  __syncthreads();  // table in place

  for (uint32_t kk = 0; kk < n_tiles + 2; ++kk) {
    // ---- phase 1: hand-over through the channel's single slot (kept short: the recurrence
    //      warp idles meanwhile; the float -> int16 conversions wait for phase 2) ----
    if (is_worker && active) {
      // chunk by chunk: y(kk-2) out of the slot, u(kk-1) into it, through the same 32 registers
      // (rounds without a finished tile / a new tile move don't-care values nobody reads)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const u32x4 v = lds_u4(ring + T1::u_off(lane, j));
        sts_u4(ring + T1::u_off(lane, j), u32x4{u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]});
        u[4 * j] = v.x; u[4 * j + 1] = v.y; u[4 * j + 2] = v.z; u[4 * j + 3] = v.w;
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    if (is_worker && active) {
      if (kk >= 2) {
        const uint32_t t = kk - 2;
        const int r = (int)min((uint32_t)TILE, p.n_samples - t * TILE) >> 5;
        uint32_t dW[16];  // (int16_t)y of the tile the recurrence finished last round
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (no_patch)
            dW[j] = __byte_perm((uint32_t)f2i_rz(u2f(u[2 * j])), (uint32_t)f2i_rz(u2f(u[2 * j + 1])), 0x5410);
          else
            dW[j] = f2i16x2_wrap(u2f(u[2 * j]), u2f(u[2 * j + 1]));
        }
        const int pcm = T1::part_c(dW, pv, lane, r, big_b);
        if (lane < r) out[(uint64_t)t * 32 + lane] = (int16_t)pcm;
      }
      if (kk < n_tiles) {
        cp_async_wait<0>();
        __syncwarp();
        const int r = (int)min((uint32_t)TILE, p.n_samples - kk * TILE) >> 5;
        bool pairs = false;  // the slot holds the pre-filter's outputs instead of raw samples
        if constexpr (MMA) {
          if (r == 32 && p.fmt == FMT_U8_OFFSET_ROTATE)
            pairs = WbMma::prefilter<false>((uint32_t)__cvta_generic_to_shared(area), tab_s, lane, true);
          __syncwarp();
          // diagnostics: [1] = tiles whose pre-filter ran on the tensor cores, [2] = on the CUDA cores
          if (lane == 0 && p.counters) atomicAdd(p.counters + (pairs ? 1 : 2), 1u);
        }
        uint32_t w[16];
        tile_read(in_slot, lane, w);
        __syncwarp();
        if (kk + 1 < n_tiles) {  // the input slot is free again: fetch the next tile now
          const uint32_t s1 = (kk + 1) * TILE;
          tile_fill(in_slot, src + (uint64_t)s1 * 2, lane, (int)min((uint32_t)TILE, p.n_samples - s1) >> 3);
        }
        cp_async_commit();
        if (MMA && pairs) {
          WbMma::part_a<false>(w, k, lut_s, pv, v_boundary, u, lane);
        } else {
          if constexpr (MMA) WbMma::planes_from_history(hist, p.fmt, pv);
          T::part_a(w, p.fmt, k, lut_s, pv, v_boundary, u, lane, r);
          if constexpr (MMA) {
            __syncwarp();
            if (lane == r - 1) WbMma::history_from_window(hist, w);
            __syncwarp();
          }
        }
      }
    } else if (is_iir && active && kk >= 1 && kk <= n_tiles) {
      // B(kk-1): y[n] = fl(u[n] - fl(a1 * y[n-1])) in place, lane == channel (IirFilter.cc:161-176)
      const uint32_t t = kk - 1;
      const int r = (int)min((uint32_t)TILE, p.n_samples - t * TILE) >> 5;
      if (r == 32) {
        // full tile: eight rows per iteration, so the swizzle (row & 7) is a compile-time constant
        // and every shared address is base + immediate -- the warp issues nothing but the loads,
        // the chain and the stores. (Fetching the next row early was measured twice and lost 4-13 %:
        // the extra loads in flight land between the chain's dependent instructions.)
        u32x4 first = lds_u4(ring);  // row 0, chunk 0 ^ 0
        for (int row0 = 0; row0 < 32; row0 += 8) {
          char *base = ring + 128 * row0;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            // only the row's FIRST chunk is fetched ahead (in the middle of the row before): it is
            // what the chain needs first, the other seven arrive while its four steps run
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
              // next row's chunk 0 (row 32 = the pad behind the slot for the last row: read, never used)
              if (j == 3) first = lds_u4(row + 128 + 16 * ((rr + 1) & 7));
            }
          }
        }
      } else {
        for (int row = 0; row < r; ++row) {
          u32x4 v[8];
          T::chain_load(ring + 128 * row, row & 7, v);
          T::chain_run(ring + 128 * row, row & 7, v, a1, y1);
        }
      }
    }
    __syncthreads();
  }

  if (active) {
    if (is_worker) {
      if constexpr (MMA) WbMma::planes_from_history(hist, p.fmt, pv);
      T1::store_carry(pv, blob, lane);
      if (lane == 0) {
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
