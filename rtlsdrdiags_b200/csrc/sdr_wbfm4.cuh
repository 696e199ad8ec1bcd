// WBFM kernel, fourth generation: the pre-demodulation filter on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, accumulators in tensor memory).
//
// The 16-tap pre-filter (WbFmDemodulator.cc:17-35, 389-398) is 32 multiply-adds per complex sample at the
// full 256 kS/s: 512 IDP.2A per 1024 samples and lane on the half-rate integer datapath, a quarter of the
// earlier generations' instructions and more than half of their ALU-pipe cycles. The legacy mma.sync int8
// path does not help -- on B200 it occupies the same integer datapath (profiles/r02_wbfm_mma_prefilter.txt).
// tcgen05 does: the MMA is issued by ONE thread of a warp that does nothing else, reads its operands from
// shared memory by descriptor and leaves the result in tensor memory, beside the CUDA cores, not on them.
//
// The GEMM (a Toeplitz product on the RAW bytes, as in WbMma):
//   D[128 rows x 64 columns] per "M-block" and tap half, s32 in TMEM
//   row     = one window = one lane's 32 samples = 64 raw bytes. The input slots as cp.async fills them ARE
//             the canonical K-major SWIZZLE_64B operand (64-byte rows, 16-byte chunk c of row r at
//             c ^ ((r >> 1) & 3), 512 bytes per 8-row group): four worker warps' slots = 128 rows, no copy,
//             no conversion. tools/micro/umma_toeplitz.cu pins the descriptor semantics this relies on.
//   column  = 2 p + arm: I' (arm 0) or Q' (arm 1) of sample p of the window, in two groups of 16 samples.
//             A group needs the 64 bytes that end with its own 32: two K = 32 steps,
//               samples  0..15:  [row above, bytes 32..63] * B0  +  [row, bytes  0..31] * B1
//               samples 16..31:  [row,       bytes  0..31] * B0  +  [row, bytes 32..63] * B1
//             with the SAME two 32 x 32 tap matrices (a group is four rotation periods long): B0, B1 for the
//             high and the low byte of the doubled Q15 taps = 4 KB of shared memory in all. The Fs/4
//             rotation and the I/Q de-interleave are in where the taps sit and which sign they carry.
//   "row above" is the A descriptor started 64 bytes early. For a channel's FIRST window the row above is
//             somebody else's: that MMA runs with those lanes disabled (disable-output-lane mask) and a
//             second one, with every other lane disabled, takes the same K step from the channel's 32-byte
//             history record through a no-swizzle descriptor whose 8-row groups alias (SBO = LBO = 16 bytes).
//   starts  the u8 offset and the doubled rounding constant, 256 * (hi start) + (lo start) as one s32 per
//             column, are stored into the low half's accumulator with tcgen05.st before the MMAs, which then
//             accumulate onto them.
// A worker warp is a TMEM lane quarter: lane = row = window, so tcgen05.ld hands every lane the 32 (I', Q')
// accumulator pairs of its own window -- the layout the rest of the chain wants, no transposition. Per 1024
// samples a warp issues 8 tcgen05.ld + 64 LEA instead of the front end, the window shuffles and 512 IDP.2A.
// One thing is not linear: int8 negation leaves -128 alone (IqDataProcessor.cc:594-607), so a raw byte 0 in
// a position the rotation negates would come out as +128; every lane checks its own window and such a
// (half-)tile -- like partial tiles and input that is already signed -- takes the CUDA-core path of the
// earlier generations (its MMA result is simply not used).
//
// Roles (all share one round loop and the same two CTA barriers, as in wbfm_tile2 / tile3_kernel):
//   worker warps   TWO: two channels each, half a tile (16 windows) per channel and round (WbTile3);
//                  else one channel, a tile per round (WbTile2). Slot, TMEM quarter and M-block by WARP index.
//   warp p.aux     the de-emphasis recurrences, lane = channel.
//   the last warp  allocates TMEM and issues the round's MMAs after the first barrier; the workers run
//                  C(k-2) meanwhile and then wait on the mbarrier the MMAs commit to.
// The carry blob is the earlier generations': the kernels are interchangeable between calls.
#pragma once
#include "sdr_wbfm3.cuh"

#if SDR_DEVICE_BUILD
namespace sdr {

constexpr int WB4_B_TILE = 32 * 32;           // one tap matrix: 32 columns x K = 32, int8
constexpr int WB4_TAB_BYTES = 4 * WB4_B_TILE;  // wb_umma_table(): [k-step][hi / lo]
constexpr int WB4_MAX_WARPS = 16;
// history records, 64 bytes per warp, + what the aliased descriptor of the last M-block reads behind them
// (8-row group 15 at 16 * 15, its rows 16 bytes apart, the second K chunk 16 further)
constexpr int WB4_HIST_BYTES = WB4_MAX_WARPS * 64 + 256;

struct WbUmma {
  // ---- what the host table and the kernel must agree on ----
  // byte of element (column n, k) inside a tap matrix: no-swizzle K-major, core matrix = 8 columns x 16 bytes
  __host__ __device__ static constexpr int b_offset(int n, int k) { return (n & 7) * 16 + (n >> 3) * 256 + (k & 15) + (k >> 4) * 128; }
  // sign with which the raw component that feeds `arm` enters at rotation phase cm (IqDataProcessor.cc:567-611:
  // I' = I0, -Q1, -I2, Q3; Q' = Q0, I1, -Q2, -I3)
  __host__ __device__ static constexpr int sign_of(int cm, int arm) {
    return arm == 0 ? ((cm == 0 || cm == 3) ? 1 : -1) : ((cm == 0 || cm == 1) ? 1 : -1);
  }
  // accumulator start of column 2 p + arm (depends on p mod 4 only): the doubled rounding constant minus
  // 128 x the sum of the column's doubled, signed taps (u = s + 128)
  __host__ __device__ static constexpr int start_of(int p, int arm) {
    int sum = 0;
    for (int k = 0; k < taps::WB_PRE::N; ++k) sum += 2 * taps::WB_PRE::tap(k) * sign_of((p - k) & 3, arm);
    return (1 << 15) - 128 * sum;
  }

  // ---- descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor) ----
  // A: K-major SWIZZLE_64B, 512 bytes between 8-row groups; the swizzle is a function of the absolute shared
  // address, so a start 64 bytes early or 32 bytes into the row needs no base_offset (micro test)
  __device__ __forceinline__ static uint64_t desc_rows(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
  }
  // the history records: no swizzle, SBO = LBO = 16 bytes: row 0 of 8-row group g is the 32 bytes at 16 g
  __device__ __forceinline__ static uint64_t desc_history(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | (1ull << 32) | (1ull << 46);
  }
  // B: no swizzle, K-major: next 16 K bytes 128 bytes on, next 8 columns 256 bytes on (b_offset)
  __device__ __forceinline__ static uint64_t desc_taps(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
  }
  // D s32; A = raw bytes, unsigned; B = taps, signed; both K-major; N = 32, M = 128
  static constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  // lanes_off: TMEM lanes (rows) the MMA must not write, the same 32-bit pattern for every quarter
  __device__ __forceinline__ static void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t accumulate, uint32_t lanes_off) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %5, {%4, %4, %4, %4}, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(accumulate), "r"(lanes_off), "r"(IDESC)
        : "memory");
  }
  // One M-block (four warps' slots at rows_s, their history records at hist_s, D at TMEM columns d .. d + 127):
  // ten MMAs. first_rows: the lanes of a quarter that are a channel's first window.
  __device__ __forceinline__ static void issue_block(uint32_t d, uint32_t rows_s, uint32_t hist_s, uint32_t taps_s, uint32_t first_rows) {
#pragma unroll
    for (int part = 0; part < 2; ++part) {  // 0: high tap bytes, overwrites; 1: low tap bytes, onto the stored starts
      const uint64_t b0 = desc_taps(taps_s + (0 * 2 + part) * WB4_B_TILE), b1 = desc_taps(taps_s + (1 * 2 + part) * WB4_B_TILE);
      const uint32_t dd = d + 64 * part, acc = part;
      mma(dd, desc_rows(rows_s - 64 + 32), b0, acc, first_rows);
      mma(dd, desc_history(hist_s), b0, acc, ~first_rows);
      mma(dd, desc_rows(rows_s), b1, 1, 0);
      mma(dd + 32, desc_rows(rows_s), b0, acc, 0);
      mma(dd + 32, desc_rows(rows_s + 32), b1, 1, 0);
    }
  }

  __device__ __forceinline__ static void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
  }
  // the lane's row of the low half's accumulator <- the starts (eight values, period eight columns)
  __device__ __forceinline__ static void store_starts(uint32_t taddr_lo) {
    constexpr int s0 = start_of(0, 0), s1 = start_of(0, 1), s2 = start_of(1, 0), s3 = start_of(1, 1);
    constexpr int s4 = start_of(2, 0), s5 = start_of(2, 1), s6 = start_of(3, 0), s7 = start_of(3, 1);
#pragma unroll
    for (int c = 0; c < 64; c += 16)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr_lo + c),
                   "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7)
                   : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  // ---- the raw history record of a channel: bytes 32..63 of the window before its first, logical order ----
  __device__ __forceinline__ static void planes_from_history(const char *hist, int fmt, WbCarry &pv) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(hist + 8 * i);
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
      uint32_t *w = reinterpret_cast<uint32_t *>(hist + 8 * i);
      w[0] = w0;
      w[1] = w1;
    }
  }
  __device__ __forceinline__ static void history_from_window(char *hist, const uint32_t (&w)[16]) {
    sts_u4(hist, u32x4{w[8], w[9], w[10], w[11]});
    sts_u4(hist + 16, u32x4{w[12], w[13], w[14], w[15]});
  }
  // raw byte 0 where the rotation negates: Q1 (byte 3 of an even word), I2 Q2 I3 (bytes 0-2 of an odd word)
  // w: the lane's window; hist: its channel's history record, of which lanes lw = 0..3 look at a rotation group each
  __device__ __forceinline__ static bool clipping_byte(const uint32_t (&w)[16], const char *hist, int lw) {
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t c = __byte_perm(w[2 * i], w[2 * i + 1], 0x6543);  // Q1 I2 Q2 I3
      z |= (c - 0x01010101u) & ~c;
    }
    if (lw < 4) {
      const uint32_t *h = reinterpret_cast<const uint32_t *>(hist + 8 * lw);
      const uint32_t c = __byte_perm(h[0], h[1], 0x6543);
      z |= (c - 0x01010101u) & ~c;
    }
    return (z & 0x80808080u) != 0;
  }

  // theta of one sample from its two doubled accumulators (WbTile2::theta from the PRMT on)
  __device__ __forceinline__ static float theta(uint32_t ai, uint32_t aq, uint32_t lut_s) {
    const uint32_t x = __byte_perm(ai, aq, 0x7762);
    const uint32_t sg = prmt_sx(aq, 0, 0xA4A4);
    const uint32_t v = __byte_perm(x ^ (sg & 0xff00u), 0, 0x4410);
    const uint32_t addr = lut_s + (v << 2) + (sg & 0x400u);
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
    return u2f(t ^ (sg & 0x80000000u));
  }
  // theta of samples 8 g .. 8 g + 7 of the lane's window from tensor memory
  __device__ __forceinline__ static void theta8(uint32_t taddr, int g, uint32_t lut_s, float (&th)[8]) {
    uint32_t hi[16], lo[16];
    tmem_ld16(taddr + 16 * g, hi);
    tmem_ld16(taddr + 64 + 16 * g, lo);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i)
      th[i] = theta((hi[2 * i] << 8) + lo[2 * i], (hi[2 * i + 1] << 8) + lo[2 * i + 1], lut_s);
  }
  // A after the pre-filter for a FULL tile (H16 = false: 32 lanes, one channel) or two full half-tiles (H16 = true):
  // WbTile2 / WbTile3::part_a from theta on. taddr: the lane's row, column 0 of the M-block's accumulators.
  template <bool H16>
  __device__ __forceinline__ static void part_a(uint32_t taddr, float k, uint32_t lut_s, WbCarry &pv, float &v_boundary,
                                                uint32_t (&u)[32], int lane) {
    // the lane's LAST samples first: theta[31] and v[31] depend on this lane's data only, and the lane above
    // needs them before it can start
    float th3[8];
    theta8(taddr, 3, lut_s, th3);
    const float my_th31 = th3[7];
    const float my_v31 = fmul(k, wrap_pi_table(fsub(th3[7], th3[6])));
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
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float th[8];
      theta8(taddr, g, lut_s, th);
      const float a[4] = {th[0], th[1], th[2], th[3]}, b[4] = {th[4], th[5], th[6], th[7]};
      wb_u4(a, k, th_prev, v_prev, &u[8 * g]);
      wb_u4(b, k, th_prev, v_prev, &u[8 * g + 4]);
    }
    const float a[4] = {th3[0], th3[1], th3[2], th3[3]}, b[4] = {th3[4], th3[5], th3[6], th3[7]};
    wb_u4(a, k, th_prev, v_prev, &u[24]);
    wb_u4(b, k, th_prev, v_prev, &u[28]);
    v_boundary = __shfl_sync(FULL, my_v31, H16 ? ((lane & 16) | 15) : 31);
    pv.th31 = my_th31;
  }

  // ---- shared memory ----
  // [atan2 table][slots by warp index][rings by worker][TWO: audio rings][tap matrices][history: 64 B per warp + pad][barrier]
  template <bool TWO>
  __host__ __device__ static constexpr int smem_bytes(int workers, int slots) {
    return WbTile2::LUT_BYTES + slots * TILE_BYTES +
           (TWO ? workers * (2 * WbTile3::RING_BYTES + 2 * WbTile3::ERING_WORDS * 4) : workers * WbTile2::RING_BYTES) + WB4_TAB_BYTES +
           WB4_HIST_BYTES + 64;  // + the barrier and the TMEM base
  }
};

// blockDim = 32 * (workers + 2): worker warps, the recurrence warp p.aux among them, the MMA warp last.
// TWO: p.G = 2 * workers channels per CTA (WbTile3's geometry); else p.G = workers (WbTile2's).
template <bool TWO>
__global__ void __launch_bounds__(32 * WB4_MAX_WARPS, 1) wbfm_tile4_kernel(const __grid_constant__ LaunchParams p) {
  using T1 = WbTile;
  using T2 = WbTile2;
  using T3 = WbTile3;
  using U = WbUmma;
  constexpr int STEP = TWO ? T3::HALF : TILE;          // samples per channel and round
  constexpr int ROWS = STEP / 32;                        // windows per channel and round
  constexpr int RING = TWO ? T3::RING_BYTES : T2::RING_BYTES;
  constexpr uint32_t FIRST_ROWS = TWO ? 0x00010001u : 0x00000001u;
  extern __shared__ __align__(1024) uint4 smem_raw[];
  char *smem = reinterpret_cast<char *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = TWO ? (int)p.G / 2 : (int)p.G;  // worker warps
  const int rec = (int)p.aux;
  const int mma_warp = nw + 1;
  const int n_slots = rec < nw ? nw + 1 : nw;    // slots are indexed by warp: the recurrence warp's is a hole
  const bool is_iir = warp == rec, is_mma = warp == mma_warp;
  const int widx = warp < rec ? warp : warp - 1;
  const bool is_worker = !is_iir && !is_mma && widx < nw;
  const uint32_t list0 = blockIdx.x * p.G;
  const int n_here = (int)min(p.G, p.n_list - list0);
  const uint32_t n_rounds = (p.n_samples + STEP - 1) / STEP;
  char *slot_base = smem + T2::LUT_BYTES;
  char *ring_base = slot_base + n_slots * TILE_BYTES;
  uint32_t *ering_base = reinterpret_cast<uint32_t *>(ring_base + (TWO ? 2 : 1) * nw * RING);
  char *tab_base = reinterpret_cast<char *>(ering_base) + (TWO ? 2 * nw * T3::ERING_WORDS * 4 : 0);
  char *hist_base = tab_base + WB4_TAB_BYTES;
  uint64_t *bar = reinterpret_cast<uint64_t *>(hist_base + WB4_HIST_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);

  // the atan2 table (129 KB from L2 once per CTA) and the tap matrices
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.lut);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < T2::LUT_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    const uint4 *tsrc = reinterpret_cast<const uint4 *>(p.tab);
    uint4 *tdst = reinterpret_cast<uint4 *>(tab_base);
    for (int i = threadIdx.x; i < WB4_TAB_BYTES / 16; i += blockDim.x) tdst[i] = __ldg(tsrc + i);
    // history records nobody owns (the hole, dead channels) must still be defined bytes
    uint32_t *h = reinterpret_cast<uint32_t *>(hist_base);
    for (int i = threadIdx.x; i < WB4_HIST_BYTES / 4; i += blockDim.x) h[i] = 0x80808080u;
  }
  if (is_mma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    if (lane == 0) {
      mbar_init(bar_s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }

  // channel slot of this lane (WbTile3 / WbTile2): a worker's lanes serve one or two slots, the recurrence warp's lane l slot l
  const int slot_id = is_iir ? lane : (TWO ? 2 * widx + (lane >> 4) : widx);
  const bool owned = (is_iir || is_worker) && slot_id < n_here;
  const uint32_t ch = owned ? p.chan_ids[list0 + slot_id] : 0;
  const bool active = owned && !(p.allowed && !p.allowed[ch]);  // a squelched channel is skipped
  uint32_t *blob = reinterpret_cast<uint32_t *>(p.state + (uint64_t)ch * p.state_stride);
  char *ring = ring_base + (is_worker || owned ? slot_id : 0) * RING;
  uint32_t *er = ering_base + (is_worker && TWO ? 2 * widx + (lane >> 4) : 0) * T3::ERING_WORDS;
  const int lw = TWO ? lane & 15 : lane;     // lane within the channel's windows
  const bool any_active = __any_sync(FULL, active);

  // ---- worker state ----
  WbCarry pv;
  const uint8_t *src = p.iq + (uint64_t)ch * p.ch_stride;
  int16_t *out = p.pcm + (uint64_t)ch * p.pcm_stride;
  char *in_slot = slot_base + (is_worker ? warp : 0) * TILE_BYTES;
  char *hist = hist_base + (is_worker ? warp : 0) * 64 + (TWO ? (lane >> 4) * 32 : 0);
  float k = 0.f, v_boundary = 0.f;
  bool big_b = false, no_patch = true;
  uint32_t u[32] = {};
  // ---- recurrence state ----
  float y1 = 0.f;
  const float a1 = (float)(-0.9492274);

  // the warp's 2 KB input slot, window = lane: (TWO) chunks 0-63 from the first channel's stream, 64-127 from the second's
  auto fetch = [&](uint32_t t) {
    const uint32_t s0 = t * STEP;
    const int valid = active ? (int)min((uint32_t)STEP, p.n_samples - s0) >> 3 : 0;  // 16-byte chunks of this channel
    if constexpr (TWO) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c0 = lw + 32 * j, c1 = lw + 16 + 32 * j;
        const int q0 = 64 * (lane >> 4) + c0, q1 = 64 * (lane >> 4) + c1;
        if (c0 < valid) cp_async16(in_slot + 16 * tile_slot(q0), src + (uint64_t)s0 * 2 + 16 * c0);
        if (c1 < valid) cp_async16(in_slot + 16 * tile_slot(q1), src + (uint64_t)s0 * 2 + 16 * c1);
      }
    } else {
      tile_fill(in_slot, src + (uint64_t)s0 * 2, lane, valid);
    }
  };

  __syncthreads();  // tables in place, TMEM allocated, barrier initialised
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  // the warp's lane quarter of its M-block's accumulators (M-block = warp / 4, quarter = warp % 4); tcgen05.ld / st
  // with shape 32x32b give thread l the quarter's lane l: thread = row = window
  const uint32_t taddr = tmem + 128u * (uint32_t)(warp >> 2) + ((uint32_t)(32 * (warp & 3)) << 16);

  if (is_worker) {
    if constexpr (TWO) { er[lw] = 0; er[16 + lw] = 0; er[32 + lw] = 0; }
    if (active) {
      if constexpr (TWO) T3::load_carry(pv, blob, er, lane); else T1::load_carry(pv, blob, lane);
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
    if (lw == ROWS - 1) U::history_from_planes(hist, p.fmt, pv);  // a dead channel's: 0x80 bytes
    U::store_starts(taddr + 64);
    fetch(0);
    cp_async_commit();
    cp_async_wait<0>();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // slot and history -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else if (is_iir && active) {
    y1 = u2f(blob[T1::NREG * 32]);
  }
  // The MMA warp's round: after the barrier between the phases it issues the round's MMAs (every input has landed,
  // every accumulator of the round before has been read and restarted by then); the workers run C(kk - 2) meanwhile.
  // (Issuing a round earlier, right after the barrier that ends the round before, was measured and lost: 0.776 ms
  // against 0.705 -- the issue then competes with the hand-over every warp is waiting for.)
  auto issue_round = [&]() {
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (lane == 0) {
      const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(slot_base), hist_s = (uint32_t)__cvta_generic_to_shared(hist_base);
      const uint32_t taps_s = (uint32_t)__cvta_generic_to_shared(tab_base);
      for (int b = 0; 4 * b < n_slots; ++b)
        U::issue_block(tmem + 128u * b, rows_s + 4u * TILE_BYTES * b, hist_s + 256u * b, taps_s, FIRST_ROWS);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s) : "memory");
    }
    __syncwarp();
  };

  for (uint32_t kk = 0; kk < n_rounds + 2; ++kk) {
    // ---- phase 1: hand-over through the channel's slot: y(kk-2) out, u(kk-1) in ----
    if (is_worker) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const u32x4 v = lds_u4(ring + T1::u_off(lw, j));
        sts_u4(ring + T1::u_off(lw, j), u32x4{u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]});
        u[4 * j] = v.x; u[4 * j + 1] = v.y; u[4 * j + 2] = v.z; u[4 * j + 3] = v.w;
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    if (is_mma) {
      if (kk < n_rounds) issue_round();
    } else if (is_worker) {
      if (kk >= 2) {
        const uint32_t t = kk - 2;
        const int r = (int)min((uint32_t)STEP, p.n_samples - t * STEP) >> 5;
        uint32_t dW[16];  // (int16_t)y of the (half-)tile the recurrence finished last round
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (no_patch)
            dW[j] = __byte_perm((uint32_t)f2i_rz(u2f(u[2 * j])), (uint32_t)f2i_rz(u2f(u[2 * j + 1])), 0x5410);
          else
            dW[j] = f2i16x2_wrap(u2f(u[2 * j]), u2f(u[2 * j + 1]));
        }
        int pcm;
        if constexpr (TWO) pcm = T3::part_c(dW, pv, er, lane, r, big_b); else pcm = T1::part_c(dW, pv, lane, r, big_b);
        if (active && lw < r) out[(uint64_t)t * ROWS + lw] = (int16_t)pcm;
      }
      if (kk < n_rounds) {
        const int r = (int)min((uint32_t)STEP, p.n_samples - kk * STEP) >> 5;
        mbar_wait(bar_s, kk & 1u);  // the round's MMAs are done: the accumulators are there, slot and history are free
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t w[16];
        tile_read(in_slot, lane, w);
        // the tensor-core result is good for full (half-)tiles of u8 input without a clipping byte
        bool umma_ok = r == ROWS && p.fmt == FMT_U8_OFFSET_ROTATE;
        if (umma_ok) umma_ok = !__any_sync(FULL, active && U::clipping_byte(w, hist, lw));
        if (!umma_ok) U::planes_from_history(hist, p.fmt, pv);  // the OLD history, before it is replaced
        __syncwarp();
        if (lw == r - 1) U::history_from_window(hist, w);
        if (kk + 1 < n_rounds) fetch(kk + 1);  // the input slot is free again
        cp_async_commit();
        if (lane == 0 && p.call_id && any_active) atomicAdd(p.counters + (umma_ok ? 1 : 2), 1u);  // diagnostics (tests only)
        // for the engine: how much of this launch the tensor cores could not do (it moves a clipping bank to generation 3)
        if (lane == 0 && !umma_ok && r == ROWS && p.fmt == FMT_U8_OFFSET_ROTATE && any_active) atomicAdd(p.counters + 3, 1u);
        if (umma_ok) {
          U::template part_a<TWO>(taddr, k, lut_s, pv, v_boundary, u, lane);
        } else {
          if constexpr (TWO) T3::part_a(w, p.fmt, k, lut_s, pv, v_boundary, u, lane, r);
          else T2::part_a(w, p.fmt, k, lut_s, pv, v_boundary, u, lane, r);
        }
        // the next round's MMAs accumulate onto the starts; its input must have landed and be visible to them
        U::store_starts(taddr + 64);
        cp_async_wait<0>();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
      }
    } else if (is_iir && active && kk >= 1 && kk <= n_rounds) {
      // B(kk-1): y[n] = fl(u[n] - fl(a1 * y[n-1])) in place, lane == channel (IirFilter.cc:161-176)
      const uint32_t t = kk - 1;
      const int r = (int)min((uint32_t)STEP, p.n_samples - t * STEP) >> 5;
      if (r == ROWS) {
        // full (half-)tile: eight rows per iteration, so the swizzle (row & 7) is a compile-time constant and
        // every shared address is base + immediate; only a row's FIRST chunk is fetched ahead (wbfm_tile2_kernel)
        u32x4 first = lds_u4(ring);
        for (int row0 = 0; row0 < ROWS; row0 += 8) {
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
              // next row's chunk 0 (the row behind the last: the pad behind the slot, read, never used)
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
      U::planes_from_history(hist, p.fmt, pv);
      if constexpr (TWO) T3::store_carry(pv, blob, er, lane); else T1::store_carry(pv, blob, lane);
      if (lw == 0) {
        blob[T1::NREG * 32 + 1] = f2u(v_boundary);
        blob[T1::NREG * 32 + 2] = big_b;
      }
    } else {
      blob[T1::NREG * 32] = f2u(y1);
    }
  }
  if (is_mma) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace sdr
#endif
