// WBFM pre-demodulation filter on the LEGACY tensor path (mma.sync): an option of generations 2 and 3, measured and
// not faster. The version that pays is generation 4 (tcgen05, sdr_wbfm4.cuh).
//
// The 16-tap pre-filter (WbFmDemodulator.cc:17-35, 389-398) runs at the full 256 kS/s on both arms:
// 32 multiply-adds per complex sample, 512 IDP.2A per 1024-sample tile and lane on the half-rate
// integer datapath -- a quarter of the WBFM kernels' instructions and more than half of their ALU-pipe
// cycles (profiles/r02_final_ncu.txt). Here it is a warp-private int8 Toeplitz GEMM on the RAW bytes of
// the input slot, like the NBFM tuner (FmTile::theta_mma):
//   D[16 x 8] = A[16 x 48] * B[48 x 8]   per "M-tile" j = 0..3 of an "N-tile" of eight windows
//   B column n = window n's raw bytes [16 j - 32, 16 j + 16): the 16-byte granule that holds the eight
//     complex samples the M-tile's outputs belong to and the two granules before it (15 samples of
//     history), fetched with ldmatrix straight from the cp.async slot -- no offset, no rotation, no
//     de-interleave; a window is one lane's 32 samples = 64 bytes = four granules, and the granules
//     before a window are the window before it (the channel's 32-byte history for its first window);
//   A row o (o < 8) = the taps that turn those bytes into I' output 8 j + o of the window, row 8 + o the
//     same for Q': the Fs/4 rotation and the de-interleave are in where the taps sit and which sign they
//     carry, and because a granule is two rotation periods long A is the same for every j. The doubled
//     int16 taps (WbTile::Pre2) are split as 256 * hi + lo, two int8 matrices; the u8 offset and the
//     doubled rounding constant are the accumulator starts, and they ride in the K dimension: the first of
//     an accumulator's two m16n8k32 meets [granule j - 2 | a constant granule], so every IMMA chain starts
//     from RZ (as a C operand they cost four MOVs per IMMA: ptxas accumulates in place).
//   A lane ends up with (I', Q') of output 8 j + (lane >> 2) of windows 2 (lane & 3), + 1: byte 2 of the
//     two accumulators, stored as a 16-bit pair OVER the raw bytes of that sample -- the slot turns from
//     1024 raw samples into 1024 pre-filter outputs in place, in the same layout, and each lane then
//     reads its own window exactly as it read the raw bytes before. The N-tiles go from the last to the
//     first, so that the window tail an N-tile reads as history is still raw. One thing is not linear:
//     int8 negation leaves -128 alone (IqDataProcessor.cc:594-607), so a raw byte 0 in a position the
//     rotation negates would come out as +128; every lane checks its own window first and such a tile
//     goes down the CUDA-core path (WbTile2 / WbTile3::part_a).
// 16 LDSM + 64 IMMA + 64 LEA + ~130 other instructions per tile replace the front end, the window shuffles
// and 512 IDP.2A: 15 % fewer instructions for the kernel. MEASURED: NOT FASTER (WBFM x8192 0.802 ms against
// 0.785, profiles/r02_wbfm_mma_prefilter.txt): on B200 the legacy int8 mma.sync occupies the integer datapath
// (~13 ALU-pipe cycles per IMMA.16832 by ncu's pipe counters), the same pipe the IDP.2A it replaces runs on,
// and this Toeplitz matrix is sparse (32,768 useful MACs in 262,144 MAC slots per tile). So it is an OPTION
// (sdr_debug_set_wbfm_kernel generation + 16, SDR_WB_MMA=1), bit-exact and tested, not the default.
// Table: wb_mma_table() in the engine; checked on the CPU by tests/test_wb_mma_table.py.
#pragma once
#include "sdr_tile.cuh"

#if SDR_DEVICE_BUILD
namespace sdr {

constexpr int WB_TAB_A_WORDS = 4 * 32 * 4;        // wb_mma_table(): four A fragments of [lane][4 words] ...
constexpr int WB_TAB_WORDS = WB_TAB_A_WORDS + 4;  // ... and the constant granule C
constexpr int WB_HIST_AREA = 64;  // bytes in front of a channel's windows; the first 32 = its raw pre-filter history

// four samples from their thetas -> u[0..3], advances (th_prev, v_prev)  (WbFmDemodulator.cc:463-486;
// numerator of the de-emphasis IIR, IirFilter.cc:161-176 with b0 == b1)
__device__ __forceinline__ void wb_u4(const float (&th)[4], float k, float &th_prev, float &v_prev, uint32_t *u) {
  const float b0 = (float)(0.0253863), b1 = (float)(0.0253863);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v = fmul(k, wrap_pi_table(fsub(th[i], th_prev)));
    u[i] = f2u(fadd(fmul(b0, v), fmul(b1, v_prev)));
    th_prev = th[i];
    v_prev = v;
  }
}

struct WbMma {
  // Shared-memory layout of a worker's input area when the pre-filter runs here: every channel's data is
  // preceded by WB_HIST_AREA bytes that hold, in their first 32, granules 3 and 2 of the window before the
  // channel's first (= the physical order of a last window, whose swizzle is 3): ldmatrix rows that reach
  // back from window 0 need no special case.
  //   one channel per worker:  [hist | 32 windows]
  //   two channels per worker: [hist | 16 windows | hist | 16 windows]
  template <bool TWO>
  __host__ __device__ static constexpr int area_bytes() { return TWO ? 2 * (WB_HIST_AREA + TILE_BYTES / 2) : WB_HIST_AREA + TILE_BYTES; }
  // byte offset in the area of window w (0..31, in tile_read's numbering: lane = window)
  template <bool TWO>
  __host__ __device__ static constexpr int window_base(int w) {
    return TWO ? WB_HIST_AREA + (w >> 4) * (WB_HIST_AREA + TILE_BYTES / 2) + 64 * (w & 15) : WB_HIST_AREA + 64 * w;
  }

  __device__ __forceinline__ static void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
  }
  __device__ __forceinline__ static void lds128(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
  }
  // signed taps (A), unsigned raw bytes (B); the first MMA of an accumulator starts from zero
  __device__ __forceinline__ static void imma_first(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0));
  }
  __device__ __forceinline__ static void imma_more(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }

  // area_s = shared address of the worker's input area (see above), raw u8 bytes in tile_fill's layout
  // (window w's granule i at 16 (i ^ ((w >> 1) & 3))). On success the windows hold the pre-filter's outputs as
  // (i', q') byte pairs in place of the samples and the history areas are those of the next tile; returns
  // false, with nothing changed, if a raw byte 0 sits where the rotation negates.
  // live: whether the lane's channel is (a dead one's stale bytes must not force the slow path).
  template <bool TWO>
  __device__ __forceinline__ static bool prefilter(uint32_t area_s, uint32_t tab_s, int lane, bool live) {
    const uint32_t own = area_s + (uint32_t)window_base<TWO>(0) + (TWO ? (uint32_t)((lane >> 4) * WB_HIST_AREA) : 0u) + 64u * lane;
    const bool hist_lane = (TWO ? (lane & 15) : lane) < 8;
    // the lane's channel's history area
    const uint32_t hist = area_s + (TWO ? (uint32_t)((lane >> 4) * (WB_HIST_AREA + TILE_BYTES / 2)) : 0u);
    // ---- raw byte 0 where the rotation negates: Q1 (byte 3 of an even word), I2 Q2 I3 (bytes 0-2 of an odd
    //      word); each lane looks at its own window, lanes 0-7 of a channel at a word of the history ----
    {
      uint32_t ze = 0, zo = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // the chunks in physical order: word parity is what matters
        uint32_t v[4];
        lds128(own + 16u * i, v);
        ze |= ((v[0] - 0x01010101u) & ~v[0]) | ((v[2] - 0x01010101u) & ~v[2]);
        zo |= ((v[1] - 0x01010101u) & ~v[1]) | ((v[3] - 0x01010101u) & ~v[3]);
      }
      if (hist_lane) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(hist + 4u * (lane & 7)) : "memory");
        const uint32_t zz = (v - 0x01010101u) & ~v;
        if (lane & 1) zo |= zz; else ze |= zz;
      }
      const uint32_t z = (ze & 0x80000000u) | (zo & 0x00808080u);
      if (__any_sync(FULL, live && z != 0)) return false;
    }
    // the next tile's history: granules 2, 3 of the channel's last window = its first 32 bytes (swizzle 3),
    // in the order the history keeps. Read now, before the window is overwritten; stored at the end, after
    // the N-tiles that read the old history.
    uint32_t next_hist = 0;
    if (hist_lane)
      asm volatile("ld.shared.u32 %0, [%1];"
                   : "=r"(next_hist)
                   : "r"(hist + (uint32_t)(WB_HIST_AREA + (TWO ? TILE_BYTES / 2 : TILE_BYTES) - 64) + 4u * (lane & 7))
                   : "memory");

    const int g = lane >> 2, tq = lane & 3;
    const int row = lane & 7, m = lane >> 3;
    // Four ldmatrix.x4 per N-tile; matrix m of each is a granule of the eight windows (negative: of the
    // window before) or C, the constant granule behind the tables that carries the accumulator starts:
    //   LA = (-2, C, -1, C)   LB = (0, C, 1, C)   LC = (-1, 0, 1, 2)   LD = (0, 1, 2, 3)
    // so that every operand pair of an IMMA -- (granule j - 2, C) and (granule j - 1, granule j) -- is an
    // aligned register pair as loaded. Offsets are relative to the N-tile's first window.
    auto rel = [&](int gran) -> uint32_t {
      const int wr = gran < 0 ? row - 1 : row, gr = gran & 3;
      return (uint32_t)(64 * wr + 16 * (gr ^ ((wr >> 1) & 3)));
    };
    const bool is_c = (m & 1) != 0;
    const uint32_t c_s = tab_s + 4u * WB_TAB_A_WORDS;
    const uint32_t off_a = is_c ? 0u : rel(m == 0 ? -2 : -1), off_b = is_c ? 0u : rel(m == 0 ? 0 : 1);
    const uint32_t off_c = rel(m - 1), off_d = rel(m);

    uint32_t a0h[4], a0l[4], a1h[4], a1l[4];
    lds128(tab_s + 16u * lane, a0h);
    lds128(tab_s + 16u * lane + 512u, a0l);
    lds128(tab_s + 16u * lane + 1024u, a1h);
    lds128(tab_s + 16u * lane + 1536u, a1l);
    // pair (i', q') of output 8 j + g of window 8 n + 2 tq + e: bytes 2 g, 2 g + 1 of granule j (at j ^ tq)
    const uint32_t soff = (uint32_t)(128 * tq + 2 * g);
    // N-tiles from the last to the first: each overwrites its own eight windows and reads, beside them, only
    // the tail of the window before -- which belongs to the N-tile done next (or is the history)
#pragma unroll 1
    for (int n = 3; n >= 0; --n) {
      const uint32_t nb = area_s + (uint32_t)(TWO ? window_base<true>(0) + (n >> 1) * (WB_HIST_AREA + TILE_BYTES / 2) + 512 * (n & 1)
                                                  : window_base<false>(0) + 512 * n);
      const uint32_t nbc = is_c ? c_s : nb;
      uint32_t la[4], lb[4], lc[4], ld[4];
      ldsm4(nbc + off_a, la);
      ldsm4(nb + off_c, lc);
      ldsm4(nb + off_d, ld);
      ldsm4(nbc + off_b, lb);
      int hi[4][4], lo[4][4];
      imma_first(hi[0], a0h, la[0], la[1]); imma_first(lo[0], a0l, la[0], la[1]);
      imma_first(hi[1], a0h, la[2], la[3]); imma_first(lo[1], a0l, la[2], la[3]);
      imma_first(hi[2], a0h, lb[0], lb[1]); imma_first(lo[2], a0l, lb[0], lb[1]);
      imma_first(hi[3], a0h, lb[2], lb[3]); imma_first(lo[3], a0l, lb[2], lb[3]);
      imma_more(hi[0], a1h, lc[0], lc[1]); imma_more(lo[0], a1l, lc[0], lc[1]);
      imma_more(hi[1], a1h, ld[0], ld[1]); imma_more(lo[1], a1l, ld[0], ld[1]);
      imma_more(hi[2], a1h, lc[2], lc[3]); imma_more(lo[2], a1l, lc[2], lc[3]);
      imma_more(hi[3], a1h, ld[2], ld[3]); imma_more(lo[3], a1l, ld[2], ld[3]);
      __syncwarp();  // every lane has its fragments: the windows may be overwritten
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t acc_i = (uint32_t)(hi[j][e] * 256 + lo[j][e]);
          const uint32_t acc_q = (uint32_t)(hi[j][2 + e] * 256 + lo[j][2 + e]);
          const uint32_t pair = __byte_perm(acc_i, acc_q, 0x0062);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(nb + soff + (uint32_t)(16 * (tq ^ j)) + (uint32_t)(64 * e)), "h"((uint16_t)pair)
                       : "memory");
        }
    }
    __syncwarp();
    if (hist_lane) asm volatile("st.shared.u32 [%0], %1;" ::"r"(hist + 4u * (lane & 7)), "r"(next_hist) : "memory");
    return true;
  }

  // The history <-> the planes the CUDA-core path and the carry blob keep (WbCarry::a, b: the last 16
  // samples after the front end, four per word). Every lane gets / gives the same words.
  __device__ __forceinline__ static void planes_from_history(const char *hist, int fmt, WbCarry &pv) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // sample group i = words 2 i, 2 i + 1 of [granule 2 | granule 3]; granule 3 is stored first
      const uint32_t *w = reinterpret_cast<const uint32_t *>(hist + 16 * ((i >> 1) ^ 1) + 8 * (i & 1));
      front_end_group(fmt, w[0], w[1], pv.a[i], pv.b[i]);
    }
  }
  __device__ __forceinline__ static void history_from_planes(char *hist, int fmt, const WbCarry &pv) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w0, w1;
      if (fmt == FMT_U8_OFFSET_ROTATE) {
        // front_end_group's inverse (wrapping int8 negation is its own inverse):
        // I0 = I'0, Q0 = Q'0, I1 = Q'1, Q1 = -I'1  |  I2 = -I'2, Q2 = -Q'2, I3 = -Q'3, Q3 = I'3
        w0 = offset_and_negate(byte_perm(pv.a[i], pv.b[i], 0x1540), 0xff000000u, 0x01000000u);
        w1 = offset_and_negate(byte_perm(pv.a[i], pv.b[i], 0x3762), 0x00ffffffu, 0x00010101u);
      } else {
        w0 = byte_perm(pv.a[i], pv.b[i], 0x5140);
        w1 = byte_perm(pv.a[i], pv.b[i], 0x7362);
      }
      uint32_t *w = reinterpret_cast<uint32_t *>(hist + 16 * ((i >> 1) ^ 1) + 8 * (i & 1));
      w[0] = w0;
      w[1] = w1;
    }
  }
  // after a tile on the CUDA-core path: the raw history from the lane's own raw window (its last 32 bytes)
  __device__ __forceinline__ static void history_from_window(char *hist, const uint32_t (&w)[16]) {
    sts_u4(hist, u32x4{w[12], w[13], w[14], w[15]});       // granule 3
    sts_u4(hist + 16, u32x4{w[8], w[9], w[10], w[11]});    // granule 2
  }

  // ---- the rest of A from the pairs: w = the lane's window, word i = samples 2 i (bytes 0, 1) and 2 i + 1 ----
  template <int N>
  __device__ __forceinline__ static float theta(const uint32_t (&w)[16], uint32_t lut_s) {
    // WbTile2::theta with the pair already in place: row |q| = (q ^ s) - s, theta = sign(q) * T[|q|][i]
    const uint32_t x = w[N >> 1];
    uint32_t sg, v;
    if constexpr ((N & 1) == 0) {
      sg = prmt_sx(x, 0, 0x9494);  // 0xFF00FF00 where q < 0
      v = __byte_perm(x ^ (sg & 0xff00u), 0, 0x4410);
    } else {
      sg = prmt_sx(x, 0, 0xB4B4);
      v = __byte_perm(x ^ (sg & 0xff000000u), 0, 0x4432);
    }
    const uint32_t addr = lut_s + (v << 2) + (sg & 0x400u);
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
    return u2f(t ^ (sg & 0x80000000u));
  }
  template <int N0>
  __device__ __forceinline__ static void theta4(const uint32_t (&w)[16], uint32_t lut_s, float (&th)[4]) {
    th[0] = theta<N0>(w, lut_s);
    th[1] = theta<N0 + 1>(w, lut_s);
    th[2] = theta<N0 + 2>(w, lut_s);
    th[3] = theta<N0 + 3>(w, lut_s);
  }
  template <int J>
  __device__ __forceinline__ static void u_chunks(const uint32_t (&w)[16], uint32_t lut_s, float k, float &th_prev,
                                                  float &v_prev, uint32_t (&u)[32]) {
    if constexpr (J < 7) {
      float th[4];
      theta4<4 * J>(w, lut_s, th);
      wb_u4(th, k, th_prev, v_prev, &u[4 * J]);
      u_chunks<J + 1>(w, lut_s, k, th_prev, v_prev, u);
    }
  }
  // A after the pre-filter for a FULL tile (H16 = false: 32 lanes, one channel) or two full half-tiles
  // (H16 = true: lanes 0-15 and 16-31): WbTile2 / WbTile3::part_a from theta on
  template <bool H16>
  __device__ __forceinline__ static void part_a(const uint32_t (&w)[16], float k, uint32_t lut_s, WbCarry &pv, float &v_boundary,
                                                uint32_t (&u)[32], int lane) {
    float th_last[4];
    theta4<28>(w, lut_s, th_last);
    const float my_th31 = th_last[3];
    const float my_v31 = fmul(k, wrap_pi_table(fsub(th_last[3], th_last[2])));
    float th_prev, v_prev;
    if constexpr (H16) {
      th_prev = __shfl_sync(FULL, (lane & 15) == 15 ? pv.th31 : my_th31, ((lane - 1) & 15) | (lane & 16));
      v_prev = __shfl_up_sync(FULL, my_v31, 1, 16);
      if ((lane & 15) == 0) v_prev = v_boundary;
    } else {
      th_prev = shfl_prev(my_th31, pv.th31, 1, lane);
      v_prev = __shfl_up_sync(FULL, my_v31, 1);
      if (lane == 0) v_prev = v_boundary;
    }
    u_chunks<0>(w, lut_s, k, th_prev, v_prev, u);
    wb_u4(th_last, k, th_prev, v_prev, &u[28]);
    v_boundary = __shfl_sync(FULL, my_v31, H16 ? ((lane & 16) | 15) : 31);
    pv.th31 = my_th31;
  }
};

}  // namespace sdr
#endif
