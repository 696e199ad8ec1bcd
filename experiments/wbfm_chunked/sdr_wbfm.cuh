// Wide-band FM (WbFmDemodulator.cc), chunked kernel (device only).
//
// What bounds WBFM is not arithmetic but the atan2 table: 1024 scattered 4-byte lookups per
// 1024 samples cost ~26 L1 tag cycles per warp-wide gather (ncu, profiles/r01v3_wbfm_ncu.txt:
// the L1 tag stage 80 % busy, long_scoreboard the top stall). atan2 is odd in q -- exactly,
// for every one of the 256 x 256 entries (checked when the engine builds the table) -- so the
// half plane q >= 0 (129 x 256 floats = 129 KB) is enough, and that fits in shared memory next
// to the per-channel buffers once the de-emphasis recurrence is handed over in CHUNKS of 256
// samples instead of tiles of 1024. Shared-memory gathers cost a few bank-conflict cycles.
//
// One CTA per SM: up to 21 WORKER warps, one channel each, and one CHAIN warp with lane ==
// channel. Round r, one CTA barrier per round:
//   worker : C1(r-2), C23 if that chunk closed a tile of 1024 samples, then A(r)
//   chain  : B(r-1)
// A: the 16-tap pre-filter of both arms at the full 256 kS/s as an int8 Toeplitz GEMM on the
//    tensor cores (mma.sync m16n8k32, operands straight from the raw cp.async chunk with
//    ldmatrix; the Fs/4 rotation, de-interleave and u8 offset live in the tap matrices, see
//    wbfm_mma_table() in the engine), int8 truncation, table lookup, first difference, wrap,
//    gain, numerator of the de-emphasis filter -> u[0..255] into ring slot r & 1. All in the
//    GEMM's own thread layout. A raw byte 0 where the rotation negates (-(-128) = -128 is not
//    linear) sends the chunk through a plain scalar pre-filter instead.
// B: y[n] = fl(u[n] - fl(a1 * y[n-1])) in place (IirFilter.cc:161-176), nothing else.
// C1: (int16_t)y -> 8-tap 4:1 decimator, 64 outputs per chunk into the channel's tile buffer.
// C23: per 1024 samples, lane == PCM sample: 12-tap 4:1 -> 40-tap 2:1 -> PCM.
#pragma once
#include "sdr_tile.cuh"

#if SDR_DEVICE_BUILD
namespace sdr {

constexpr int WB_CHUNK = 256;                    // samples per channel and round
constexpr int WB_CHUNK_BYTES = 2 * WB_CHUNK;
constexpr int WB_MAX_CH = 21;                    // channels = worker warps per CTA
constexpr int WB_LUT_ROWS = 129;                 // |q| = 0..128
constexpr int WB_LUT_BYTES = WB_LUT_ROWS * 256 * 4;
// wbfm_mma_table(): per format nine entries of [lane][4 words]: A fragments [hi/lo][config],
// config = (8 P - 16 s + 16) / 8 for phase group P and k-step s, then the accumulator starts
constexpr int WB_TAB_WORDS_PER_FMT = 9 * 32 * 4;
// per-channel shared memory
constexpr int WB_IN_STRIDE = 64 + WB_CHUNK_BYTES;  // [history: window 7 of the chunk before | chunk]
constexpr int WB_OFF_IN = 0;                        // two input buffers
constexpr int WB_OFF_RING = 2 * WB_IN_STRIDE;       // two ring slots of 256 floats
constexpr int WB_OFF_E1 = WB_OFF_RING + 2 * WB_CHUNK * 4;  // 128 words: decimator-1 outputs of the tile
constexpr int WB_OFF_ER = WB_OFF_E1 + 512;          // 64 words: audio-decimator ring
constexpr int WB_CH_SMEM = WB_OFF_ER + 256 + 16;    // + pad: stride / 16 odd -> lane == channel LDS.128 conflict free
static_assert((WB_CH_SMEM / 16) % 2 == 1 && WB_CH_SMEM % 16 == 0, "channel stride");
constexpr int WB_SMEM_BYTES = WB_LUT_BYTES + WB_TAB_WORDS_PER_FMT * 4 + WB_MAX_CH * WB_CH_SMEM;

struct WbChunk {
  // state blob: 7 words per lane (C1 carry 2, decimator-1 outputs 4, decimator-2 outputs 1), then
  // 16 words: theta of the last sample, clamp flag, 8 words of raw history (16 samples, in the
  // format-independent signed/rotated form), pad; then the 16 bytes a reset leaves alone
  // (WbFmDemodulator.cc:304-320): y[n-1] and v[n-1] of the de-emphasis filter.
  static constexpr int NREG = 7;
  static constexpr int SCAL = NREG * 32;
  static constexpr int STATE_BYTES = (SCAL + 16 + 4) * 4;
  static constexpr int IIR_WORD = SCAL + 16;

  // sample n of a chunk sits at word perm(n) of its ring slot: bits 3-4 are XORed with bits 6-7,
  // which makes the workers' scattered stores conflict free and keeps groups of four together
  __device__ __forceinline__ static int perm(int n) { return n ^ (((n >> 6) & 3) << 3); }

  // ---- scalar pre-filter for one window of a chunk with clipping bytes (lanes 0-7) ----
  __device__ __noinline__ static void prefilter_scalar(const char *buf, int win, int fmt, uint16_t *scratch) {
    int xi[48], xq[48];  // samples -16 .. 31 of the window
#pragma unroll 1
    for (int j = -16; j < 32; ++j) {
      const int w = win + (j >> 5);  // j < 0: the window before (-1 = the history)
      const int jj = j & 31;
      const int piece = jj >> 3;  // 16-byte piece of the window
      const uint8_t *p8 = reinterpret_cast<const uint8_t *>(buf + 64 * w + 16 * (piece ^ ((w >> 1) & 3)) + 2 * (jj & 7));
      int I, Q;
      if (fmt == FMT_U8_OFFSET_ROTATE) {
        const int si = (int)(int8_t)(uint8_t)(p8[0] - 128u), sq = (int)(int8_t)(uint8_t)(p8[1] - 128u);
        const int ni = (int)(int8_t)(-si), nq = (int)(int8_t)(-sq);  // -(-128) stays -128
        switch (jj & 3) {
          case 0: I = si; Q = sq; break;
          case 1: I = nq; Q = si; break;
          case 2: I = ni; Q = nq; break;
          default: I = sq; Q = ni; break;
        }
      } else {
        I = (int)(int8_t)p8[0];
        Q = (int)(int8_t)p8[1];
      }
      xi[j + 16] = I;
      xq[j + 16] = Q;
    }
#pragma unroll 1
    for (int pp = 0; pp < 32; ++pp) {
      int ai = 1 << 14, aq = 1 << 14;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        ai += taps::WB_PRE::tap(t) * xi[16 + pp - t];
        aq += taps::WB_PRE::tap(t) * xq[16 + pp - t];
      }
      // (int8_t)(acc >> 15): the clamp is out of reach for int8 input (WbFmDemodulator.cc:389-398)
      scratch[win * 32 + pp] = (uint16_t)(((ai >> 15) & 0xff) | (((aq >> 15) & 0xff) << 8));
    }
  }

  __device__ __forceinline__ static void lds128(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
  }
  __device__ __forceinline__ static void ldmatrix4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
  }
  __device__ __forceinline__ static void ldmatrix2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
  }
};

// blockDim = 32 * (workers + 1): warps 0 .. workers-1 are workers (slot = warp), the last warp
// is the chain warp. Every thread runs the same round loop and meets the same barrier.
__global__ void __launch_bounds__(32 * (WB_MAX_CH + 1), 1) wbfm_chunk_kernel(const __grid_constant__ LaunchParams p) {
  using T = WbChunk;
  extern __shared__ uint4 smem_raw[];
  char *smem = reinterpret_cast<char *>(smem_raw);
  float *lut = reinterpret_cast<float *>(smem);                           // [129][256]
  uint32_t *tab = reinterpret_cast<uint32_t *>(smem + WB_LUT_BYTES);      // A fragments + starts
  char *chan_base = smem + WB_LUT_BYTES + WB_TAB_WORDS_PER_FMT * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (int)p.G;  // channels of a full CTA = worker warps
  const bool is_chain = warp == nw;
  const uint32_t list0 = blockIdx.x * (uint32_t)nw;
  const int n_here = (int)min((uint32_t)nw, p.n_list - list0);
  const int fmt = p.fmt;

  // tables into shared memory
  {
    const uint4 *src4 = reinterpret_cast<const uint4 *>(p.lut);
    uint4 *dst4 = reinterpret_cast<uint4 *>(lut);
    for (int i = threadIdx.x; i < WB_LUT_BYTES / 16; i += blockDim.x) dst4[i] = __ldg(src4 + i);
    const uint32_t *t = p.tab + (fmt == FMT_U8_OFFSET_ROTATE ? 0 : WB_TAB_WORDS_PER_FMT);
    for (int i = threadIdx.x; i < WB_TAB_WORDS_PER_FMT; i += blockDim.x) tab[i] = t[i];
  }

  const int slot_id = is_chain ? lane : warp;
  const bool owned = slot_id < n_here;
  const uint32_t ch = owned ? p.chan_ids[list0 + slot_id] : 0;
  const bool active = owned && !(p.allowed && !p.allowed[ch]);  // a squelched channel is skipped
  uint32_t *blob = reinterpret_cast<uint32_t *>(p.state + (uint64_t)ch * p.state_stride);
  char *cs = chan_base + (active ? slot_id : 0) * WB_CH_SMEM;  // this channel's shared memory
  const uint32_t n_chunks = (p.n_samples + WB_CHUNK - 1) / WB_CHUNK;

  // ---- worker state ----
  const uint8_t *src = p.iq + (uint64_t)ch * p.ch_stride;
  int16_t *out = p.pcm + (uint64_t)ch * p.pcm_stride;
  const int g = lane >> 2, tq = lane & 3;
  float k = 0.f, th_carry = 0.f, v_carry = 0.f;
  bool big_b = false, no_patch = false;
  uint32_t c1p[2] = {0, 0}, e1c[4] = {0, 0, 0, 0}, ewc = 0;  // C carries
  const uint32_t cs_s = (uint32_t)__cvta_generic_to_shared(cs);
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab) + 16u * (uint32_t)lane;
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
  // ldmatrix rows: window = lane & 7, 16-byte piece lane >> 3 of the 64 bytes that start 32
  // bytes before the window (x4), and pieces 2, 3 of the window itself (x2)
  uint32_t relA, relB;
  {
    const int w = lane & 7, i = lane >> 3;
    // x4 piece i: i < 2 -> piece 2 + i of window w - 1; else piece i - 2 of window w
    const int wa = i < 2 ? w - 1 : w, ja = i < 2 ? 2 + i : i - 2;
    relA = (uint32_t)(64 * wa + 16 * (ja ^ ((wa >> 1) & 3)));
    relB = (uint32_t)(64 * w + 16 * ((2 + (i & 1)) ^ ((w >> 1) & 3)));
  }
  const uint32_t zmask = (lane & 1) ? 0x00808080u : 0x80000000u;
  const int srcA = g >= 1 ? lane - 4 : (lane + 28) & 31;  // lane holding the sample before mine
  const int srcB = g >= 1 ? lane - 4 : (lane + 27) & 31;  // the same for (P, e) = (0, 0)
  // ---- chain state ----
  float y1 = 0.f;
  const float a1 = (float)(-0.9492274);

  if (!is_chain) {
    if (active) {
      c1p[0] = blob[0 * 32 + lane]; c1p[1] = blob[1 * 32 + lane];
#pragma unroll
      for (int i = 0; i < 4; ++i) e1c[i] = blob[(2 + i) * 32 + lane];
      ewc = blob[6 * 32 + lane];
      th_carry = u2f(blob[T::SCAL + 0]);
      big_b = blob[T::SCAL + 1] != 0;
      v_carry = u2f(blob[T::IIR_WORD + 1]);
      k = p.scale[ch];
      // |y| <= max(|y[-1]|, |u|max / (1 - |a1|)) < 3.2 |k|: with |k| < 1e8 and |y[-1]| < 1e9 no
      // value can reach 2^31, where cvt.rzi (saturating) and x86 cvttss2si (wrapping) differ
      no_patch = fabsf(k) < 1e8f && fabsf(u2f(blob[T::IIR_WORD])) < 1e9f;
      // raw history of the block's head: 16 samples = pieces 2, 3 of "window -1", physical order
      if (lane < 8) {
        const uint32_t st = blob[T::SCAL + 2 + lane];
        const int piece = 2 + (lane >> 2);
        *reinterpret_cast<uint32_t *>(cs + WB_OFF_IN + 64 - 64 + 16 * (piece ^ 3) + 4 * (lane & 3)) =
            FmTile::hist_from_state(st, fmt, (lane & 1) == 0);
      }
      reinterpret_cast<uint32_t *>(cs + WB_OFF_ER)[lane] = ewc;
      // first chunk
      if (lane < (int)(min((uint32_t)WB_CHUNK, p.n_samples) >> 3))
        cp_async16(cs + WB_OFF_IN + 64 + 16 * (lane ^ ((lane >> 3) & 3)), src + 16 * lane);
    }
    cp_async_commit();
  } else if (active) {
    y1 = u2f(blob[T::IIR_WORD]);
  }
  __syncthreads();  // tables are in place

  for (uint32_t r = 0; r < n_chunks + 2; ++r) {
    if (!is_chain && active) {
      float *ring = reinterpret_cast<float *>(cs + WB_OFF_RING) + (r & 1) * WB_CHUNK;
      // ------------------------------ C1(r-2) and C23 ------------------------------
      if (r >= 2) {
        const uint32_t c = r - 2;
        const int wv = (int)min((uint32_t)WB_CHUNK, p.n_samples - c * WB_CHUNK) >> 5;  // valid windows
        const int rv = 4 * wv;                                                          // valid lanes
        // the lane's eight y: samples 8 lane .. 8 lane + 7
        const u32x4 v0 = lds_u4(ring + T::perm(8 * lane)), v1 = lds_u4(ring + T::perm(8 * lane + 4));
        uint32_t dw[4];
        if (no_patch) {
          dw[0] = __byte_perm((uint32_t)f2i_rz(u2f(v0.x)), (uint32_t)f2i_rz(u2f(v0.y)), 0x5410);
          dw[1] = __byte_perm((uint32_t)f2i_rz(u2f(v0.z)), (uint32_t)f2i_rz(u2f(v0.w)), 0x5410);
          dw[2] = __byte_perm((uint32_t)f2i_rz(u2f(v1.x)), (uint32_t)f2i_rz(u2f(v1.y)), 0x5410);
          dw[3] = __byte_perm((uint32_t)f2i_rz(u2f(v1.z)), (uint32_t)f2i_rz(u2f(v1.w)), 0x5410);
        } else {
          dw[0] = f2i16x2_wrap(u2f(v0.x), u2f(v0.y));
          dw[1] = f2i16x2_wrap(u2f(v0.z), u2f(v0.w));
          dw[2] = f2i16x2_wrap(u2f(v1.x), u2f(v1.y));
          dw[3] = f2i16x2_wrap(u2f(v1.z), u2f(v1.w));
        }
        // decimator 1: 8 taps, 4:1 (clamp-free); output m uses d[4m-4 .. 4m+3]
        const uint32_t ext[6] = {shfl_prev(dw[2], c1p[0], 1, lane), shfl_prev(dw[3], c1p[1], 1, lane),
                                 dw[0], dw[1], dw[2], dw[3]};
        static_assert(taps::WB_DEC1::SAFE >= 32768, "decimator 1 must be clamp-free");
        const uint32_t wlo[4] = {ext[0], ext[1], ext[2], ext[3]}, whi[4] = {ext[2], ext[3], ext[4], ext[5]};
        const int o0 = (int)(int16_t)(fir_s16_fast<taps::WB_DEC1, 7, 4>(wlo) >> 15);
        const int o1 = (int)(int16_t)(fir_s16_fast<taps::WB_DEC1, 7, 4>(whi) >> 15);
        uint32_t *e1buf = reinterpret_cast<uint32_t *>(cs + WB_OFF_E1);
        e1buf[(c & 3) * 32 + lane] = pack_i16x2(o0, o1);
        if (rv == 32) {
          c1p[0] = dw[2]; c1p[1] = dw[3];
        } else {
          c1p[0] = roll_prev(dw[2], c1p[0], rv, lane);
          c1p[1] = roll_prev(dw[3], c1p[1], rv, lane);
        }
        if ((c & 3) == 3 || c + 1 == n_chunks) {
          // C23: the tile's decimator-1 outputs, eight per lane
          const uint32_t tile = c >> 2;
          const int rt = (int)min((uint32_t)TILE, p.n_samples - tile * TILE) >> 5;  // valid lanes = PCM samples
          __syncwarp();
          const u32x4 ev = lds_u4(e1buf + 4 * lane);
          const uint32_t e1w[4] = {ev.x, ev.y, ev.z, ev.w};
          // decimator 2: 12 taps, 4:1, clamp-free (decimator 1 output <= 29126 <= FM_POST::SAFE)
          uint32_t de[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            de[i] = shfl_prev(e1w[i], e1c[i], 1, lane);
            de[4 + i] = e1w[i];
          }
          const uint32_t w0[6] = {de[0], de[1], de[2], de[3], de[4], de[5]};
          const uint32_t w1[6] = {de[2], de[3], de[4], de[5], de[6], de[7]};
          const int e0 = (int)(int16_t)(fir_s16_fast<taps::FM_POST, 11, 6>(w0) >> 15);
          const int e1 = (int)(int16_t)(fir_s16_fast<taps::FM_POST, 11, 6>(w1) >> 15);
          const uint32_t ew = pack_i16x2(e0, e1);
          const bool cur_b =
              __any_sync(FULL, (iabs(e0) > taps::AUDIO40::SAFE || iabs(e1) > taps::AUDIO40::SAFE) && lane < rt);
          const bool exact_b = cur_b || big_b;
          // audio decimator: 40 taps, 2:1, through the 64-word ring (previous tile | this tile)
          uint32_t *ering = reinterpret_cast<uint32_t *>(cs + WB_OFF_ER);
          ering[32 + lane] = ew;
          __syncwarp();
          uint32_t ee[20];
#pragma unroll
          for (int i = 0; i < 20; ++i) ee[i] = ering[13 + lane + i];
          int acc;
          if (!exact_b) {
            acc = fir_s16_fast<taps::AUDIO40, 39, 20>(ee);
          } else {
            acc = fir_s16_guard_mid<taps::AUDIO40, 39, 20>(ee);
            const bool clamped = __any_sync(FULL, !fir_s16_guard_tail_is_free<taps::AUDIO40>(acc));
            acc = fir_s16_guard_tail<taps::AUDIO40, 39, 20>(ee, acc, clamped);
          }
          if (lane < rt) out[(uint64_t)tile * 32 + lane] = (int16_t)(acc >> 15);
          big_b = cur_b || (rt < 32 && big_b);
          if (rt == 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) e1c[i] = e1w[i];
            ewc = ew;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) e1c[i] = roll_prev(e1w[i], e1c[i], rt, lane);
            ewc = roll_prev(ew, ewc, rt, lane);
          }
          __syncwarp();
          ering[lane] = ewc;
        }
        __syncwarp();  // every lane has read y before A overwrites the slot
      }
      // ------------------------------------ A(r) ------------------------------------
      if (r < n_chunks) {
        const int wv = (int)min((uint32_t)WB_CHUNK, p.n_samples - r * WB_CHUNK) >> 5;  // valid windows
        char *buf = cs + WB_OFF_IN + (r & 1) * WB_IN_STRIDE + 64;
        char *nbuf = cs + WB_OFF_IN + ((r + 1) & 1) * WB_IN_STRIDE + 64;
        cp_async_wait<0>();
        __syncwarp();
        if (r + 1 < n_chunks) {  // the other buffer was read by A(r-1): fetch the next chunk now
          const uint32_t s1 = (r + 1) * WB_CHUNK;
          if (lane < (int)(min((uint32_t)WB_CHUNK, p.n_samples - s1) >> 3))
            cp_async16(nbuf + 16 * (lane ^ ((lane >> 3) & 3)), src + (uint64_t)s1 * 2 + 16 * lane);
        }
        cp_async_commit();

        const uint32_t buf_s = cs_s + WB_OFF_IN + (r & 1) * WB_IN_STRIDE + 64;
        uint32_t b[6];  // B fragments: (b0, b1) of k-steps 0, 1, 2
        T::ldmatrix4(buf_s + relA, b[0], b[1], b[2], b[3]);
        T::ldmatrix2(buf_s + relB, b[4], b[5]);
        bool gemm = wv == 8;  // a partial chunk (stale bytes behind the data) takes the scalar path
        if (gemm && fmt == FMT_U8_OFFSET_ROTATE) {
          uint32_t z = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) z |= (b[i] - 0x01010101u) & ~b[i];
          gemm = !__any_sync(FULL, (z & zmask) != 0);
        }
        uint32_t iqw[4][2];  // per (P, e): byte 0 = I8, byte 1 = Q8 of phase 8P+g, window 2tq+e
        if (gemm) {
          uint32_t c0[4];
          T::lds128(tab_s + 8 * 512, c0);
#pragma unroll
          for (int P = 0; P < 4; ++P) {
            int hi[4] = {(int)c0[0], (int)c0[0], (int)c0[1], (int)c0[1]};
            int lo[4] = {(int)c0[2], (int)c0[2], (int)c0[3], (int)c0[3]};
#pragma unroll
            for (int ss = 0; ss < 2; ++ss) {
              const int s = (P >> 1) + ss;
              const int cfg = (8 * P - 16 * s + 16) / 8;
              uint32_t ahi[4], alo[4];
              T::lds128(tab_s + cfg * 512, ahi);
              T::lds128(tab_s + (4 + cfg) * 512, alo);
              if (fmt == FMT_U8_OFFSET_ROTATE) {
                FmTile::imma<true>(hi, ahi, b[2 * s], b[2 * s + 1]);
                FmTile::imma<true>(lo, alo, b[2 * s], b[2 * s + 1]);
              } else {
                FmTile::imma<false>(hi, ahi, b[2 * s], b[2 * s + 1]);
                FmTile::imma<false>(lo, alo, b[2 * s], b[2 * s + 1]);
              }
            }
            // doubled taps: acc' = 2 acc = 256 hi + lo, (int8_t)(acc >> 15) = byte 2 of acc'
#pragma unroll
            for (int e = 0; e < 2; ++e)
              iqw[P][e] = __byte_perm((uint32_t)(256 * hi[e] + lo[e]), (uint32_t)(256 * hi[2 + e] + lo[2 + e]), 0x7762);
          }
        } else {
          uint16_t *scratch = reinterpret_cast<uint16_t *>(ring);  // the slot is free until u is stored
          if (lane < 8) T::prefilter_scalar(buf, lane, fmt, scratch);
          __syncwarp();
#pragma unroll
          for (int P = 0; P < 4; ++P)
#pragma unroll
            for (int e = 0; e < 2; ++e) iqw[P][e] = scratch[(2 * tq + e) * 32 + 8 * P + g];
          __syncwarp();
        }
        // history for the next chunk: this chunk's last valid window, physical order kept
        if (lane < 16) {
          const int last = wv - 1;
          const uint32_t wv32 = *reinterpret_cast<const uint32_t *>(buf + 64 * last + 16 * ((lane >> 2) ^ ((last >> 1) & 3)) +
                                                                    4 * (lane & 3));
          *reinterpret_cast<uint32_t *>(nbuf - 64 + 16 * ((lane >> 2) ^ 3) + 4 * (lane & 3)) = wv32;
          if (r + 1 == n_chunks && lane >= 8)  // the last 16 samples, for the next call
            blob[T::SCAL + 2 + lane - 8] = FmTile::hist_to_state(wv32, fmt, (lane & 1) == 0);
        }
        // theta = table[(uint8)(q+128)][(uint8)(i+128)] (WbFmDemodulator.cc:458-462), odd in q
        float th[4][2];
#pragma unroll
        for (int P = 0; P < 4; ++P)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t w = iqw[P][e];
            int q8;  // sign-extended byte 1 (prmt's replicate-sign mode; __byte_perm ignores that bit)
            asm("prmt.b32 %0, %1, %1, 0x9991;" : "=r"(q8) : "r"(w));
            const uint32_t col = (w ^ 0x80u) & 0xffu;
            const uint32_t addr = lut_s + 4u * ((uint32_t)iabs(q8) * 256u + col);
            uint32_t t;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
            th[P][e] = u2f(t ^ ((uint32_t)q8 & 0x80000000u));
          }
        // theta[n-1]: the lane that holds phase g-1 (or, for g = 0, the g = 7 lane of the phase
        // group / window before); lanes with g = 7 offer that other register
        float thp[4][2];
#pragma unroll
        for (int P = 0; P < 4; ++P)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float alt = P >= 1 ? th[P - 1][e] : th[3][e ^ 1];
            const float sv = g == 7 ? alt : th[P][e];
            thp[P][e] = __shfl_sync(FULL, sv, (P == 0 && e == 0) ? srcB : srcA);
          }
        if (lane == 0) thp[0][0] = th_carry;
        float d[4][2], dmax = 0.f;
#pragma unroll
        for (int P = 0; P < 4; ++P)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            d[P][e] = fsub(th[P][e], thp[P][e]);
            dmax = fmaxf(dmax, fabsf(d[P][e]));
          }
        float v[4][2];
        if (__any_sync(FULL, dmax >= 3.14159274101257324f)) {
#pragma unroll
          for (int P = 0; P < 4; ++P)
#pragma unroll
            for (int e = 0; e < 2; ++e) v[P][e] = fmul(k, wrap_pi(d[P][e]));
        } else {
#pragma unroll
          for (int P = 0; P < 4; ++P)
#pragma unroll
            for (int e = 0; e < 2; ++e) v[P][e] = fmul(k, d[P][e]);
        }
        // numerator of the de-emphasis filter: u = b0 v[n] + b1 v[n-1] (IirFilter.cc:164)
        const float b0 = (float)(0.0253863), b1 = (float)(0.0253863);
#pragma unroll
        for (int P = 0; P < 4; ++P)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float alt = P >= 1 ? v[P - 1][e] : v[3][e ^ 1];
            const float sv = g == 7 ? alt : v[P][e];
            float vp = __shfl_sync(FULL, sv, (P == 0 && e == 0) ? srcB : srcA);
            if (P == 0 && e == 0 && lane == 0) vp = v_carry;
            ring[T::perm(32 * (2 * tq + e) + 8 * P + g)] = fadd(fmul(b0, v[P][e]), fmul(b1, vp));
          }
        // carries: the chunk's last valid sample = window wv-1, phase 31 = lane (7, (wv-1)>>1), e = (wv-1)&1
        {
          const int last = wv - 1, ll = 28 + (last >> 1);
          th_carry = __shfl_sync(FULL, (last & 1) ? th[3][1] : th[3][0], ll);
          v_carry = __shfl_sync(FULL, (last & 1) ? v[3][1] : v[3][0], ll);
        }
      }
    } else if (is_chain && active && r >= 1 && r <= n_chunks) {
      // B(r-1): y[n] = fl(u[n] - fl(a1 * y[n-1])) in place, lane == channel (IirFilter.cc:161-176)
      const uint32_t c = r - 1;
      const int nv = (int)min((uint32_t)WB_CHUNK, p.n_samples - c * WB_CHUNK);  // valid samples
      float *ring = reinterpret_cast<float *>(cs + WB_OFF_RING) + (c & 1) * WB_CHUNK;
      for (int n0 = 0; n0 < nv; n0 += 32) {
        u32x4 vv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) vv[j] = lds_u4(ring + T::perm(n0 + 4 * j));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y0 = fsub(u2f(vv[j].x), fmul(a1, y1));
          const float y2 = fsub(u2f(vv[j].y), fmul(a1, y0));
          const float y3 = fsub(u2f(vv[j].z), fmul(a1, y2));
          y1 = fsub(u2f(vv[j].w), fmul(a1, y3));
          sts_u4(ring + T::perm(n0 + 4 * j), u32x4{f2u(y0), f2u(y2), f2u(y3), f2u(y1)});
        }
      }
    }
    __syncthreads();
  }

  if (active) {
    if (!is_chain) {
      blob[0 * 32 + lane] = c1p[0]; blob[1 * 32 + lane] = c1p[1];
#pragma unroll
      for (int i = 0; i < 4; ++i) blob[(2 + i) * 32 + lane] = e1c[i];
      blob[6 * 32 + lane] = ewc;
      if (lane == 0) {
        blob[T::SCAL + 0] = f2u(th_carry);
        blob[T::SCAL + 1] = big_b;
        blob[T::IIR_WORD + 1] = f2u(v_carry);
      }
    } else {
      blob[T::IIR_WORD] = f2u(y1);
    }
  }
}

}  // namespace sdr
#endif
