// TEST HARNESS -- serial emulation of the CUDA pipelines on the host.
//
// Compiles the arithmetic building blocks of the product kernels
// (rtlsdrdiags_b200/csrc/sdr_device.cuh) together with the phase-structured test kernels
// (sdr_phase.cuh) with -DSDR_EMU: every phase of a CTA is run for tid = 0..NT-1 in turn, a
// phase boundary stands in for __syncthreads(). It exists so the front end, the FIR cores,
// the wrap and conversion helpers can be checked against the oracle on the CPU-only build box. It is not a
// fallback: nothing in rtlsdrdiags_b200/ builds, links or loads it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "emu_config.h"

namespace {
using namespace sdr;

template <class M, int PH = 0>
void emu_phases(Ctx base) {
  if constexpr (PH < M::NPHASES) {
    for (int tid = 0; tid < base.nt; ++tid) {
      Ctx t = base;
      t.tid = tid;
      M::template phase<PH>(t);
    }
    emu_phases<M, PH + 1>(base);
  }
}

template <class M>
void emu_launch(const LaunchParams &p, int NT) {
  const uint32_t n_cta = (p.n_list + p.G - 1) / p.G;
  std::vector<uint8_t> smem(smem_bytes<M>((int)p.G) + 64);
  for (uint32_t b = 0; b < n_cta; ++b) {
    // poison shared memory: nothing may depend on what a previous CTA left
    memset(smem.data(), 0xA5, smem.size());
    Ctx base;
    base.p = &p;
    base.hdr = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    base.smem = base.hdr + HDR_BYTES;
    base.tid = 0;
    base.nt = NT;
    const uint32_t list0 = b * p.G;
    base.Gc = (int)((p.n_list - list0 < p.G) ? p.n_list - list0 : p.G);
    base.sample0 = 0;
    base.cur = 0;
    base.parity = 0;
    auto each = [&](auto fn) {
      for (int tid = 0; tid < NT; ++tid) {
        Ctx t = base;
        t.tid = tid;
        fn(t);
      }
    };
    each([&](const Ctx &t) { cta_load_header<M>(t, list0); });
    each([&](const Ctx &t) { cta_state<M, true>(t); });
    for (uint32_t s0 = 0; s0 < p.n_samples; s0 += M::S) {
      base.sample0 = s0;
      base.cur = (int)((p.n_samples - s0 < (uint32_t)M::S) ? p.n_samples - s0 : (uint32_t)M::S);
      emu_phases<M>(base);
      base.parity ^= 1;
    }
    each([&](const Ctx &t) { cta_state<M, false>(t); });
  }
}
}  // namespace

extern "C" {

int emu_state_bytes(int kind) {
  switch (kind) {
    case sdr::KIND_AM: return sdr::AmPipe::STATE_BYTES;
    case sdr::KIND_FM: return sdr::FmPipe::STATE_BYTES;
    case sdr::KIND_WBFM: return sdr::WbFmPipe::STATE_BYTES;
    case sdr::KIND_SSB: return sdr::SsbPipe::STATE_BYTES;
  }
  return -1;
}

int emu_run(int kind, const uint8_t *iq, uint64_t ch_stride, uint32_t n_samples, int fmt,
            const uint32_t *chan_ids, uint32_t n_list, uint32_t G, uint8_t *state,
            uint32_t state_stride, const float *scale, const uint8_t *lsb, int16_t *pcm,
            uint64_t pcm_stride, const float *lut, int NT) {
  sdr::LaunchParams p;
  p.iq = iq; p.ch_stride = ch_stride; p.n_samples = n_samples; p.fmt = fmt;
  p.chan_ids = chan_ids; p.n_list = n_list; p.G = G; p.state = state;
  p.state_stride = state_stride; p.scale = scale; p.lsb = lsb; p.pcm = pcm;
  p.pcm_stride = pcm_stride; p.lut = lut; p.aux = 0;
  switch (kind) {
    case sdr::KIND_AM: emu_launch<sdr::AmPipe>(p, NT); return 0;
    case sdr::KIND_FM: emu_launch<sdr::FmPipe>(p, NT); return 0;
    case sdr::KIND_WBFM: emu_launch<sdr::WbFmPipe>(p, NT); return 0;
    case sdr::KIND_SSB: emu_launch<sdr::SsbPipe>(p, NT); return 0;
  }
  return -1;
}

// wrap_pi_table (the FP64-free +-pi wrap the second WBFM kernel uses) against wrap_pi (the
// reference's double arithmetic) for EVERY difference of two of the given values; returns the
// number of mismatching pairs. vals = the distinct entries of the WBFM atan2 table.
uint64_t emu_wrap_table_check(const float *vals, uint32_t n) {
  uint64_t bad = 0;
  for (uint32_t a = 0; a < n; ++a)
    for (uint32_t b = 0; b < n; ++b) {
      const float d = sdr::fsub(vals[a], vals[b]);
      const float r = sdr::wrap_pi(d), t = sdr::wrap_pi_table(d);
      bad += memcmp(&r, &t, 4) != 0;
    }
  return bad;
}

// front_end_ab (AM / SSB stage 1: a rotation period's bytes grouped by what is done to them) on one period of eight
// raw bytes: the eight samples in the order (I'0, I'1, I'2, I'3, Q'0, Q'1, Q'2, Q'3), picked from the low / high halves
// of (a, b) exactly where stage1_simt's tap pairs expect them; and raw_from_ab's round trip. Returns 0 if the round
// trip is the identity.
int emu_front_end_ab(int fmt, const uint8_t *raw8, int8_t *out8) {
  uint32_t w0, w1, a, b, r0, r1;
  memcpy(&w0, raw8, 4);
  memcpy(&w1, raw8 + 4, 4);
  sdr::front_end_ab(fmt, w0, w1, a, b);
  const uint32_t i_bytes[4] = {a & 0xff, b & 0xff, (b >> 8) & 0xff, (a >> 8) & 0xff};              // s0 s1 s2 s3 of I'
  const uint32_t q_bytes[4] = {(a >> 16) & 0xff, (a >> 24) & 0xff, (b >> 16) & 0xff, (b >> 24) & 0xff};  // of Q'
  for (int k = 0; k < 4; ++k) {
    out8[k] = (int8_t)i_bytes[k];
    out8[4 + k] = (int8_t)q_bytes[k];
  }
  sdr::raw_from_ab(fmt, a, b, r0, r1);
  return (r0 == w0 && r1 == w1) ? 0 : 1;
}
}
