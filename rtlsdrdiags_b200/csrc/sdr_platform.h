// Platform layer for the demodulation kernels.
//
// Under nvcc this maps the handful of primitives the kernels need onto sm_100a
// instructions (IDP.2A dot products, PRMT byte permutes, 128-bit LDG/LDS/STS,
// single-rounded FP32 ops that the compiler may not contract into FMAs).
//
// Under a plain host compiler with -DSDR_EMU the same names are defined as
// scalar C++ so tests/emu can run every phase of a kernel serially, "thread" by
// "thread", and compare it with the oracle before any GPU time is spent. The
// emulation is a TEST harness: it is never built into libsdr_b200.so.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(SDR_EMU)
#define SDR_HD __host__ __device__
#define SDR_DEV __device__ __forceinline__
#define SDR_DEVM __device__ __forceinline__
#define SDR_DEVICE_BUILD 1
#else
#define SDR_HD
#define SDR_DEV static inline
#define SDR_DEVM inline
#define SDR_DEVICE_BUILD 0
#include <math.h>
#endif

namespace sdr {

struct u32x2 { uint32_t x, y; };
struct u32x4 { uint32_t x, y, z, w; };

#if SDR_DEVICE_BUILD
// ---- integer dot products: a = two 16-bit halves, b = four bytes (PTX dp2a) ----
SDR_DEV int dp2a_lo_ss(uint32_t a, uint32_t b, int c) { return __dp2a_lo((int)a, (int)b, c); }
SDR_DEV int dp2a_hi_ss(uint32_t a, uint32_t b, int c) { return __dp2a_hi((int)a, (int)b, c); }
// signed halves x UNSIGNED bytes (low tap bytes)
SDR_DEV int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
SDR_DEV uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// PTX prmt in its default mode: a selector nibble with bit 3 set replicates the SIGN of the
// selected byte (__byte_perm keeps only three bits per nibble)
SDR_DEV uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t s) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
  return d;
}

// ---- memory ----
SDR_DEV u32x4 ld_stream_u4(const void *p) {  // read-once input: bypass L1 allocation
  u32x4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
SDR_DEV float ld_lut(const float *p) { return __ldg(p); }
template <class T> SDR_DEV T lds(const void *p) { return *reinterpret_cast<const T *>(p); }
template <class T> SDR_DEV void sts(void *p, T v) { *reinterpret_cast<T *>(p) = v; }
SDR_DEV u32x4 ldg_u4(const void *p) { uint4 v = *reinterpret_cast<const uint4 *>(p); return {v.x, v.y, v.z, v.w}; }
SDR_DEV void stg_u4(void *p, u32x4 v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v.x, v.y, v.z, v.w); }
SDR_DEV u32x4 lds_u4(const void *p) { uint4 v = *reinterpret_cast<const uint4 *>(p); return {v.x, v.y, v.z, v.w}; }
SDR_DEV void sts_u4(void *p, u32x4 v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v.x, v.y, v.z, v.w); }
SDR_DEV u32x2 lds_u2(const void *p) { uint2 v = *reinterpret_cast<const uint2 *>(p); return {v.x, v.y}; }
SDR_DEV void sts_u2(void *p, u32x2 v) { *reinterpret_cast<uint2 *>(p) = make_uint2(v.x, v.y); }
SDR_DEV void stg_u32(void *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }

// ---- single-rounded FP32 (never contracted), and the one FP64 add the wrap needs ----
SDR_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
SDR_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
SDR_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
SDR_DEV float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }  // ONE rounding: only where that is what is wanted
SDR_DEV float dadd_to_f(float a, double b) { return __double2float_rn(__dadd_rn((double)a, b)); }
SDR_DEV int f2i_rz(float v) { return __float2int_rz(v); }
SDR_DEV float i2f(int v) { return __int2float_rn(v); }
SDR_DEV float u2f(uint32_t v) { return __uint_as_float(v); }
SDR_DEV uint32_t f2u(float v) { return __float_as_uint(v); }
#else
// ------------------------------ host emulation ------------------------------
SDR_DEV int dp2a_lo_ss(uint32_t a, uint32_t b, int c) {
  return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) +
         (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 8) & 0xff);
}
SDR_DEV int dp2a_hi_ss(uint32_t a, uint32_t b, int c) {
  return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)((b >> 16) & 0xff) +
         (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 24) & 0xff);
}
SDR_DEV int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  return c + (int)(int16_t)(a & 0xffff) * (int)((b >> 16) & 0xff) +
         (int)(int16_t)(a >> 16) * (int)((b >> 24) & 0xff);
}
SDR_DEV uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t sel = (s >> (4 * i)) & 0xf;
    uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
SDR_DEV uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t s) { return byte_perm(a, b, s); }
template <class T> SDR_DEV T lds(const void *p) { T v; memcpy(&v, p, sizeof(T)); return v; }
template <class T> SDR_DEV void sts(void *p, T v) { memcpy(p, &v, sizeof(T)); }
SDR_DEV u32x4 ld_stream_u4(const void *p) { return lds<u32x4>(p); }
SDR_DEV float ld_lut(const float *p) { return *p; }
SDR_DEV u32x4 ldg_u4(const void *p) { return lds<u32x4>(p); }
SDR_DEV void stg_u4(void *p, u32x4 v) { sts(p, v); }
SDR_DEV u32x4 lds_u4(const void *p) { return lds<u32x4>(p); }
SDR_DEV void sts_u4(void *p, u32x4 v) { sts(p, v); }
SDR_DEV u32x2 lds_u2(const void *p) { return lds<u32x2>(p); }
SDR_DEV void sts_u2(void *p, u32x2 v) { sts(p, v); }
SDR_DEV void stg_u32(void *p, uint32_t v) { sts(p, v); }
// the emulation is compiled with -ffp-contract=off, so plain ops round once
SDR_DEV float fmul(float a, float b) { volatile float r = a * b; return r; }
SDR_DEV float fadd(float a, float b) { volatile float r = a + b; return r; }
SDR_DEV float fsub(float a, float b) { volatile float r = a - b; return r; }
SDR_DEV float ffma(float a, float b, float c) { return fmaf(a, b, c); }
SDR_DEV float dadd_to_f(float a, double b) { return (float)((double)a + b); }
SDR_DEV int f2i_rz(float v) {  // saturating, NaN -> 0: what cvt.rzi.s32.f32 does
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 0x7fffffff;
  if (v <= -2147483648.0f) return (int)0x80000000;
  return (int)v;
}
SDR_DEV float i2f(int v) { return (float)v; }
SDR_DEV float u2f(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }
SDR_DEV uint32_t f2u(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
#endif

}  // namespace sdr
