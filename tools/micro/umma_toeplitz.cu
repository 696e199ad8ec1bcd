// Microbenchmark / semantics check: tcgen05.mma kind::i8 reading raw bytes in the WBFM input-slot layout.
//
// Question it answers before the kernel is touched: can the input slots as cp.async leaves them (64-byte
// rows = windows, 16-byte chunk c of row r at c ^ ((r >> 1) & 3): the canonical K-major SWIZZLE_64B layout)
// be the A operand of a UMMA whose K span reaches back into the row ABOVE (start address 64 bytes below a
// 512-byte pattern boundary, K offset 32 within the row)? And with which base_offset encoding?
//   D[128 rows x 64 cols] per part = sum over two k-steps per column group jj of A_op(jj, ks) * B[ks]
//   A_op(0,0) = rows r-1, bytes 32..63   A_op(0,1) = rows r, bytes 0..31
//   A_op(1,0) = rows r,   bytes 0..31    A_op(1,1) = rows r, bytes 32..63
// B: random int8 [32 n x 32 k] per (k-step, part), no-swizzle K-major. Compared exactly with the host.
// Second question (mode 1): rows 0 and 16 of every 32 are a channel's FIRST window; the row above them is
// somebody else's. Can the MMA that uses the row above skip those rows (disable-output-lane mask) and a
// second MMA fill exactly them from a compact 32-byte-per-channel history array addressed through an
// ALIASED no-swizzle descriptor (SBO = LBO = 16 bytes: 8-row group g starts at 16 g, so the even groups'
// first rows are consecutive 32-byte records and every other row is garbage nobody keeps)? And can D start
// from values stored with tcgen05.st (the accumulator starts) instead of zero?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/umma_toeplitz tools/micro/umma_toeplitz.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROWS = 128, ROW_BYTES = 64;
constexpr int PRE = 512;                       // one 8-row atom in front: its last row is "row -1"
constexpr int A_BYTES = ROWS * ROW_BYTES;      // 8192
constexpr int B_TILE = 32 * 32;                // 1 KB
constexpr int SMEM = PRE + A_BYTES + 4 * B_TILE + 64;

__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, int base_offset) {
  // K-major SWIZZLE_64B: SBO = 512 B between 8-row groups, LBO = 1 (unused), version 1, layout type 4
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_offset & 7) << 49) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint64_t desc_interleave(uint32_t addr) {
  // K-major no swizzle: core matrix 8 rows x 16 B contiguous; LBO = 128 B (next 16 K bytes), SBO = 256 B (next 8 rows)
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void umma_i8_masked(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t mask) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(mask)
      : "memory");
}
__device__ __forceinline__ uint64_t desc_aliased(uint32_t addr) {
  // K-major no swizzle with SBO = LBO = 16 B: row 0 of 8-row group g = the 32 bytes at 16 g
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)1 << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) k(const uint8_t *raw /*[129][64] logical rows -1..127*/, const int8_t *bmat /*[2 ks][2 part][32 n][32 k]*/,
                                            int variant, int *out /*[2 part][128][64]*/, int mode, const uint8_t *hist /*[8][32]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) uint8_t hist_s[8 * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t *a_base = smem + PRE;
  uint8_t *b_base = a_base + A_BYTES;
  // rows -1 .. 127, swizzled as tile_fill does (relative to a 512-aligned base == absolute address bits)
  for (int i = tid; i < 129 * 4; i += 128) {
    const int r = i / 4 - 1, c = i & 3;
    const uint4 v = *reinterpret_cast<const uint4 *>(raw + (size_t)(r + 1) * 64 + 16 * c);
    *reinterpret_cast<uint4 *>(a_base + 64 * r + 16 * (c ^ ((r >> 1) & 3))) = v;
  }
  // B tiles: element (n, k) at (n & 7) * 16 + (n >> 3) * 256 + (k & 15) + (k >> 4) * 128
  for (int i = tid; i < 4 * B_TILE; i += 128) {
    const int t = i / B_TILE, n = (i / 32) & 31, kk = i & 31;
    b_base[t * B_TILE + (n & 7) * 16 + (n >> 3) * 256 + (kk & 15) + (kk >> 4) * 128] = (uint8_t)bmat[i];
  }
  if (tid < 64) reinterpret_cast<uint32_t *>(hist_s)[tid] = reinterpret_cast<const uint32_t *>(hist)[tid];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (mode == 1) {
    // accumulator starts: column c of part 1 starts at 1000 * (c & 7) - 3 (part 0 starts from zero: p = 0)
    for (int c0 = 0; c0 < 64; c0 += 8) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(64 + c0);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(-3), "r"(997), "r"(1997),
                   "r"(2997), "r"(3997), "r"(4997), "r"(5997), "r"(6997));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (tid == 0 && mode == 1) {
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a_base), b_s = (uint32_t)__cvta_generic_to_shared(b_base);
    const uint32_t h_s = (uint32_t)__cvta_generic_to_shared(hist_s);
    const uint32_t firsts = 0x00010001u;  // lanes 0 and 16 of every 32
    for (int part = 0; part < 2; ++part) {
      const uint64_t b0 = desc_interleave(b_s + (0 * 2 + part) * B_TILE), b1 = desc_interleave(b_s + (1 * 2 + part) * B_TILE);
      const uint32_t p0 = part;  // part 0 overwrites, part 1 accumulates onto the stored starts
      umma_i8_masked(tmem + 64 * part + 0, desc_sw64(a_s - 64 + 32, 0), b0, idesc, p0, variant == 0 ? firsts : ~firsts);
      umma_i8_masked(tmem + 64 * part + 0, desc_aliased(h_s), b0, idesc, p0, variant == 0 ? ~firsts : firsts);
      umma_i8(tmem + 64 * part + 0, desc_sw64(a_s, 0), b1, idesc, 1);
      umma_i8(tmem + 64 * part + 32, desc_sw64(a_s, 0), b0, idesc, p0);
      umma_i8(tmem + 64 * part + 32, desc_sw64(a_s + 32, 0), b1, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  if (tid == 0 && mode == 0) {
    // instruction descriptor: D s32, A u8 (raw bytes), B s8 (taps), both K-major, N = 32, M = 128
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a_base), b_s = (uint32_t)__cvta_generic_to_shared(b_base);
    const uint32_t up = a_s - 64 + 32;  // rows r-1, bytes 32..63
    const int bo_up = variant;  // tried: every value of the 3-bit field
    for (int part = 0; part < 2; ++part) {
      const uint64_t b0 = desc_interleave(b_s + (0 * 2 + part) * B_TILE), b1 = desc_interleave(b_s + (1 * 2 + part) * B_TILE);
      // column group 0 (cols 0..31 of the part) and 1 (cols 32..63)
      umma_i8(tmem + 64 * part + 0, desc_sw64(up, bo_up), b0, idesc, 0);
      umma_i8(tmem + 64 * part + 0, desc_sw64(a_s, 0), b1, idesc, 1);
      umma_i8(tmem + 64 * part + 32, desc_sw64(a_s, 0), b0, idesc, 0);
      umma_i8(tmem + 64 * part + 32, desc_sw64(a_s + 32, 0), b1, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  // everybody waits for the MMAs
  {
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(bar_s), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // warp w reads TMEM lanes 32 w .. 32 w + 31: thread = row
  for (int part = 0; part < 2; ++part)
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(64 * part + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) out[(part * 128 + 32 * warp + lane) * 64 + c0 + i] = (int)v[i];
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
  std::vector<uint8_t> raw(129 * 64);
  std::vector<int8_t> bm(2 * 2 * 32 * 32);
  srand(7);
  for (auto &v : raw) v = (uint8_t)(rand() & 255);
  for (auto &v : bm) v = (int8_t)((rand() & 255) - 128);
  // expected
  std::vector<int> exp(2 * 128 * 64);
  auto A = [&](int r, int byte) { return (int)raw[(size_t)(r + 1) * 64 + byte]; };
  for (int part = 0; part < 2; ++part)
    for (int r = 0; r < 128; ++r)
      for (int col = 0; col < 64; ++col) {
        const int jj = col >> 5, n = col & 31;
        long s = 0;
        for (int ks = 0; ks < 2; ++ks)
          for (int kk = 0; kk < 32; ++kk) {
            int a;
            if (jj == 0) a = ks == 0 ? A(r - 1, 32 + kk) : A(r, kk);
            else a = ks == 0 ? A(r, kk) : A(r, 32 + kk);
            s += (long)a * bm[((ks * 2 + part) * 32 + n) * 32 + kk];
          }
        exp[(part * 128 + r) * 64 + col] = (int)s;
      }
  uint8_t *d_raw, *d_hist; int8_t *d_b; int *d_out;
  std::vector<uint8_t> hist(8 * 32);
  for (auto &v : hist) v = (uint8_t)(rand() & 255);
  cudaMalloc(&d_raw, raw.size()); cudaMalloc(&d_b, bm.size()); cudaMalloc(&d_out, exp.size() * 4); cudaMalloc(&d_hist, hist.size());
  cudaMemcpy(d_raw, raw.data(), raw.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, bm.data(), bm.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_hist, hist.data(), hist.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  for (int variant = 0; variant < 8; ++variant) {
    cudaMemset(d_out, 0xff, exp.size() * 4);
    k<<<1, 128, SMEM>>>(d_raw, d_b, variant, d_out, 0, d_hist);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> got(exp.size());
    cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
    long bad_up = 0, bad_other = 0, first = -1;
    for (size_t i = 0; i < exp.size(); ++i)
      if (got[i] != exp[i]) {
        if (((i % 64) >> 5) == 0) ++bad_up; else ++bad_other;
        if (first < 0) first = (long)i;
      }
    printf("base_offset %d for the operand that starts in the row above: %s, mismatches: %ld in the columns that use it, %ld elsewhere",
           variant, cudaGetErrorString(e), bad_up, bad_other);
    if (first >= 0)
      printf("; first at part %ld row %ld col %ld: got %d expected %d", first / (128 * 64), (first / 64) % 128, first % 64, got[first], exp[first]);
    printf("\n");
    if (e != cudaSuccess) return 1;
  }
  // mode 1: rows 0 and 16 of every 32 take their "row above" from the history records; part 1 starts from stored values
  std::vector<int> exp1(exp.size());
  for (int part = 0; part < 2; ++part)
    for (int r = 0; r < 128; ++r)
      for (int col = 0; col < 64; ++col) {
        const int jj = col >> 5, n = col & 31;
        long s = part == 1 ? 1000 * (col & 7) - 3 : 0;
        for (int ks = 0; ks < 2; ++ks)
          for (int kk = 0; kk < 32; ++kk) {
            int a;
            if (jj == 0 && ks == 0) a = (r % 16 == 0) ? (int)hist[(r / 16) * 32 + kk] : A(r - 1, 32 + kk);
            else if (jj == 0) a = A(r, kk);
            else a = ks == 0 ? A(r, kk) : A(r, 32 + kk);
            s += (long)a * bm[((ks * 2 + part) * 32 + n) * 32 + kk];
          }
        exp1[(part * 128 + r) * 64 + col] = (int)s;
      }
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d_out, 0xff, exp.size() * 4);
    k<<<1, 128, SMEM>>>(d_raw, d_b, variant, d_out, 1, d_hist);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> got(exp.size());
    cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
    long bad_first = 0, bad_rest = 0, first = -1;
    for (size_t i = 0; i < exp1.size(); ++i)
      if (got[i] != exp1[i]) {
        if (((i / 64) % 16) == 0) ++bad_first; else ++bad_rest;
        if (first < 0) first = (long)i;
      }
    printf("history records + lane mask (a set bit %s the lane) + stored starts: %s, mismatches: %ld in first-window rows, %ld in the others",
           variant == 0 ? "DISABLES" : "ENABLES", cudaGetErrorString(e), bad_first, bad_rest);
    if (first >= 0)
      printf("; first at part %ld row %ld col %ld: got %d expected %d", first / (128 * 64), (first / 64) % 128, first % 64, got[first], exp1[first]);
    printf("\n");
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
