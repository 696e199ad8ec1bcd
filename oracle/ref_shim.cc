// TEST INFRASTRUCTURE -- not part of the product path.
//
// C-callable harness around the UNMODIFIED reference classes. The reference
// sources are compiled where they lie under /root/reference (see
// oracle/Makefile); nothing from them is copied into this repository. This
// file only supplies what the reference expects its application to provide
// (nprintf, radio_adjustableReceiveGainInDb, a PCM callback) and flattens the
// C++ classes into extern "C" entry points that tests/ and bench.py's
// cpu_baseline leg can reach through ctypes.
//
// Built twice by oracle/Makefile:
//   oracle/_ref/libref_radiodiags.so  (radioDiags/ tree, -DREF_HAS_IQP)
//   oracle/_ref/libref_research.so    (demodulatorResearch/demodulators/ tree)
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "AmDemodulator.h"
#include "Decimator_int16.h"
#include "FirFilter.h"
#include "FirFilter_int16.h"
#include "FmDemodulator.h"
#include "IirFilter.h"
#include "SsbDemodulator.h"
#include "WbFmDemodulator.h"
#ifdef REF_HAS_IQP
#include "Decimator.h"
#include "Interpolator.h"
#include "Interpolator_int16.h"
#include "IqDataProcessor.h"
#endif

// ---- what the reference application normally provides -------------------
// radioDiags/src_diags/console.cc defines nprintf; IqDataProcessor.cc:8
// externs the tuner gain that Radio.cc owns.
void nprintf(FILE *s, const char *formatPtr, ...) {
  va_list ap;
  va_start(ap, formatPtr);
  vfprintf(s, formatPtr, ap);
  va_end(ap);
}
int32_t radio_adjustableReceiveGainInDb = 0;

namespace {

struct Sink {
  int16_t *out;
  uint32_t cap;
  uint32_t count;
  uint32_t calls;
};
thread_local Sink *g_sink = nullptr;

// The reference callback has no context argument (FmDemodulator.h:26), so the
// destination is a thread-local set around each acceptIqData call.
void pcmCallback(int16_t *bufferPtr, uint32_t bufferLength) {
  Sink *s = g_sink;
  if (!s) return;
  s->calls++;
  for (uint32_t i = 0; i < bufferLength; i++) {
    if (s->out && s->count < s->cap) s->out[s->count] = bufferPtr[i];
    s->count++;
  }
}

enum { KIND_AM = 1, KIND_FM = 2, KIND_WBFM = 3, KIND_SSB = 4 };

struct Demod {
  int kind;
  AmDemodulator *am;
  FmDemodulator *fm;
  WbFmDemodulator *wbfm;
  SsbDemodulator *ssb;
};

}  // namespace

extern "C" {

// ---------------- single demodulator objects ------------------------------
void *ref_demod_new(int kind) {
  Demod *d = new Demod();
  d->kind = kind;
  d->am = nullptr;
  d->fm = nullptr;
  d->wbfm = nullptr;
  d->ssb = nullptr;
  switch (kind) {
    case KIND_AM: d->am = new AmDemodulator(pcmCallback); break;
    case KIND_FM: d->fm = new FmDemodulator(pcmCallback); break;
    case KIND_WBFM: d->wbfm = new WbFmDemodulator(pcmCallback); break;
    case KIND_SSB: d->ssb = new SsbDemodulator(pcmCallback); break;
    default: delete d; return nullptr;
  }
  return d;
}

void ref_demod_free(void *h) {
  Demod *d = (Demod *)h;
  delete d->am;
  delete d->fm;
  delete d->wbfm;
  delete d->ssb;
  delete d;
}

void ref_demod_set_gain(void *h, float gain) {
  Demod *d = (Demod *)h;
  if (d->am) d->am->setDemodulatorGain(gain);
  if (d->fm) d->fm->setDemodulatorGain(gain);
  if (d->wbfm) d->wbfm->setDemodulatorGain(gain);
  if (d->ssb) d->ssb->setDemodulatorGain(gain);
}

void ref_demod_reset(void *h) {
  Demod *d = (Demod *)h;
  if (d->am) d->am->resetDemodulator();
  if (d->fm) d->fm->resetDemodulator();
  if (d->wbfm) d->wbfm->resetDemodulator();
  if (d->ssb) d->ssb->resetDemodulator();
}

void ref_ssb_set_lsb(void *h, int lsb) {
  Demod *d = (Demod *)h;
  if (!d->ssb) return;
  if (lsb) d->ssb->setLsbDemodulationMode();
  else d->ssb->setUsbDemodulationMode();
}

// buf is signed, already Fs/4-rotated IQ and MAY be modified in place (WBFM).
// Returns the number of PCM samples the callback delivered during this call.
uint32_t ref_demod_accept(void *h, int8_t *buf, uint32_t nbytes, int16_t *pcm,
                          uint32_t cap) {
  Demod *d = (Demod *)h;
  Sink s = {pcm, cap, 0, 0};
  g_sink = &s;
  if (d->am) d->am->acceptIqData(buf, nbytes);
  if (d->fm) d->fm->acceptIqData(buf, nbytes);
  if (d->wbfm) d->wbfm->acceptIqData(buf, nbytes);
  if (d->ssb) d->ssb->acceptIqData(buf, nbytes);
  g_sink = nullptr;
  return s.count;
}

// ---------------- filter classes -----------------------------------------
void *ref_dec16_new(int L, float *h, int M) { return new Decimator_int16(L, h, M); }
void ref_dec16_free(void *p) { delete (Decimator_int16 *)p; }
void ref_dec16_reset(void *p) { ((Decimator_int16 *)p)->resetFilterState(); }
uint32_t ref_dec16_run(void *p, const int16_t *in, uint32_t n, int16_t *out) {
  Decimator_int16 *d = (Decimator_int16 *)p;
  uint32_t c = 0;
  for (uint32_t i = 0; i < n; i++) {
    int16_t y;
    if (d->decimate(in[i], &y)) out[c++] = y;
  }
  return c;
}

void *ref_fir16_new(int L, float *h) { return new FirFilter_int16(L, h); }
void ref_fir16_free(void *p) { delete (FirFilter_int16 *)p; }
void ref_fir16_reset(void *p) { ((FirFilter_int16 *)p)->resetFilterState(); }
void ref_fir16_run(void *p, const int16_t *in, uint32_t n, int16_t *out) {
  FirFilter_int16 *f = (FirFilter_int16 *)p;
  for (uint32_t i = 0; i < n; i++) out[i] = f->filterData(in[i]);
}

void *ref_fir_new(int L, float *h) { return new FirFilter(L, h); }
void ref_fir_free(void *p) { delete (FirFilter *)p; }
void ref_fir_run(void *p, const float *in, uint32_t n, float *out) {
  FirFilter *f = (FirFilter *)p;
  for (uint32_t i = 0; i < n; i++) out[i] = f->filterData(in[i]);
}

void *ref_iir_new(int nb, float *b, int na, float *a) { return new IirFilter(nb, b, na, a); }
void ref_iir_free(void *p) { delete (IirFilter *)p; }
void ref_iir_run(void *p, const float *in, uint32_t n, float *out) {
  IirFilter *f = (IirFilter *)p;
  for (uint32_t i = 0; i < n; i++) out[i] = f->filterData(in[i]);
}

#ifdef REF_HAS_IQP
// ---------------- generic multirate classes ---------------
// kind 1 Decimator, 2 Interpolator, 3 Decimator_int16, 4 Interpolator_int16; same contract
// as sdro_mr_* in sdr_oracle.h. Samples go through the reference objects one at a time.
namespace {
struct Multirate {
  int kind, factor;
  void *obj;
};
}  // namespace
void *ref_mr_new(int kind, int N, float *h, int factor) {
  Multirate *m = new Multirate{kind, factor, nullptr};
  switch (kind) {
    case 1: m->obj = new Decimator(N, h, factor); break;
    case 2: m->obj = new Interpolator(N, h, factor); break;
    case 3: m->obj = new Decimator_int16(N, h, factor); break;
    case 4: m->obj = new Interpolator_int16(N, h, factor); break;
    default: delete m; return nullptr;
  }
  return m;
}
void ref_mr_free(void *p) {
  Multirate *m = (Multirate *)p;
  switch (m->kind) {
    case 1: delete (Decimator *)m->obj; break;
    case 2: delete (Interpolator *)m->obj; break;
    case 3: delete (Decimator_int16 *)m->obj; break;
    case 4: delete (Interpolator_int16 *)m->obj; break;
  }
  delete m;
}
void ref_mr_reset(void *p) {
  Multirate *m = (Multirate *)p;
  switch (m->kind) {
    case 1: ((Decimator *)m->obj)->resetFilterState(); break;
    case 2: ((Interpolator *)m->obj)->resetFilterState(); break;
    case 3: ((Decimator_int16 *)m->obj)->resetFilterState(); break;
    case 4: ((Interpolator_int16 *)m->obj)->resetFilterState(); break;
  }
}
uint64_t ref_mr_run(void *p, const void *in, uint64_t n, void *out) {
  Multirate *m = (Multirate *)p;
  uint64_t c = 0;
  for (uint64_t i = 0; i < n; i++) {
    switch (m->kind) {
      case 1: if (((Decimator *)m->obj)->decimate(((const float *)in)[i], (float *)out + c)) c++; break;
      case 2: ((Interpolator *)m->obj)->interpolate(((const float *)in)[i], (float *)out + c); c += m->factor; break;
      case 3: if (((Decimator_int16 *)m->obj)->decimate(((const int16_t *)in)[i], (int16_t *)out + c)) c++; break;
      case 4: ((Interpolator_int16 *)m->obj)->interpolate(((const int16_t *)in)[i], (int16_t *)out + c); c += m->factor; break;
    }
  }
  return c;
}

// ---------------- product path: IqDataProcessor -> demodulator ------------
namespace {
struct Chain {
  IqDataProcessor *iqp;
  AmDemodulator *am;
  FmDemodulator *fm;
  WbFmDemodulator *wbfm;
  SsbDemodulator *ssb;
  int signal_present;        // last value the signal-state callback delivered
  uint32_t signal_magnitude; // last value the signal-magnitude callback delivered
};
void onSignalState(bool signalPresent, void *contextPtr) { ((Chain *)contextPtr)->signal_present = signalPresent; }
void onSignalMagnitude(uint32_t magnitude, void *contextPtr) { ((Chain *)contextPtr)->signal_magnitude = magnitude; }
Chain *chainNew(int port = 8001) {
  // Mirrors the wiring in radioDiags/src_diags/Radio.cc:150-181.
  static char host[] = "127.0.0.1";
  Chain *c = new Chain();
  c->iqp = new IqDataProcessor(host, port);
  c->am = new AmDemodulator(pcmCallback);
  c->fm = new FmDemodulator(pcmCallback);
  c->wbfm = new WbFmDemodulator(pcmCallback);
  c->ssb = new SsbDemodulator(pcmCallback);
  c->iqp->setAmDemodulator(c->am);
  c->iqp->setFmDemodulator(c->fm);
  c->iqp->setWbFmDemodulator(c->wbfm);
  c->iqp->setSsbDemodulator(c->ssb);
  c->signal_present = -1;
  c->signal_magnitude = 0;
  c->iqp->registerSignalStateCallback(onSignalState, c);
  c->iqp->enableSignalNotification();
  c->iqp->registerSignalMagnitudeCallback(onSignalMagnitude, c);
  c->iqp->enableSignalMagnitudeNotification();
  return c;
}
void chainFree(Chain *c) {
  delete c->iqp;
  delete c->am;
  delete c->fm;
  delete c->wbfm;
  delete c->ssb;
  delete c;
}
}  // namespace

void *ref_iqp_new(void) { return chainNew(); }
// the IQ dump's link partner listens on 127.0.0.1:port
void *ref_iqp_new_port(int port) { return chainNew(port); }
void ref_iqp_set_dump(void *h, int on) {
  if (on) ((Chain *)h)->iqp->enableIqDump();
  else ((Chain *)h)->iqp->disableIqDump();
}
void ref_iqp_free(void *h) { chainFree((Chain *)h); }
void ref_iqp_set_mode(void *h, int mode) {
  ((Chain *)h)->iqp->setDemodulatorMode((IqDataProcessor::demodulatorType)mode);
}
void ref_iqp_set_gain(void *h, int kind, float gain) {
  Chain *c = (Chain *)h;
  switch (kind) {
    case KIND_AM: c->am->setDemodulatorGain(gain); break;
    case KIND_FM: c->fm->setDemodulatorGain(gain); break;
    case KIND_WBFM: c->wbfm->setDemodulatorGain(gain); break;
    case KIND_SSB: c->ssb->setDemodulatorGain(gain); break;
  }
}
void ref_iqp_reset(void *h, int kind) {
  Chain *c = (Chain *)h;
  switch (kind) {
    case KIND_AM: c->am->resetDemodulator(); break;
    case KIND_FM: c->fm->resetDemodulator(); break;
    case KIND_WBFM: c->wbfm->resetDemodulator(); break;
    case KIND_SSB: c->ssb->resetDemodulator(); break;
  }
}
void ref_iqp_set_threshold(void *h, int32_t threshold) { ((Chain *)h)->iqp->setSignalDetectThreshold(threshold); }
// the tuner gain is a process-wide global in the reference (Radio.cc owns it)
void ref_set_rx_gain(int32_t gain_db) { radio_adjustableReceiveGainInDb = gain_db; }
void ref_iqp_signal(void *h, int *present, uint32_t *magnitude) {
  *present = ((Chain *)h)->signal_present;
  *magnitude = ((Chain *)h)->signal_magnitude;
}
// buf is u8 offset-binary IQ and IS modified in place (IqDataProcessor.cc:735).
uint32_t ref_iqp_accept(void *h, uint8_t *buf, uint32_t nbytes, int16_t *pcm, uint32_t cap) {
  Chain *c = (Chain *)h;
  Sink s = {pcm, cap, 0, 0};
  g_sink = &s;
  c->iqp->acceptIqData(0, buf, nbytes);
  g_sink = nullptr;
  return s.count;
}

// Channel bank on host threads: one IqDataProcessor + four demodulators per
// channel, channels statically partitioned over
// nthreads. iq is [n_channels][bytes_per_channel] u8 and is consumed (modified
// in place, as the reference does). pcm is [n_channels][bytes_per_channel/64]
// or NULL. modes[ch] in 0..5. Returns elapsed seconds of the demodulation
// loops only (object construction excluded).
double ref_bank_run(const uint8_t *modes, uint32_t n_channels, uint8_t *iq,
                    uint64_t bytes_per_channel, uint32_t block_bytes, int16_t *pcm,
                    uint32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<Chain *> chains(n_channels);
  for (uint32_t ch = 0; ch < n_channels; ch++) {
    chains[ch] = chainNew();
    chains[ch]->iqp->setDemodulatorMode((IqDataProcessor::demodulatorType)modes[ch]);
  }
  const uint64_t pcm_per_channel = bytes_per_channel / 64;
  auto worker = [&](uint32_t t) {
    uint32_t lo = (uint64_t)n_channels * t / nthreads;
    uint32_t hi = (uint64_t)n_channels * (t + 1) / nthreads;
    for (uint32_t ch = lo; ch < hi; ch++) {
      Sink s = {pcm ? pcm + (uint64_t)ch * pcm_per_channel : nullptr,
                (uint32_t)pcm_per_channel, 0, 0};
      g_sink = &s;
      uint8_t *p = iq + (uint64_t)ch * bytes_per_channel;
      for (uint64_t off = 0; off < bytes_per_channel; off += block_bytes) {
        uint64_t n = bytes_per_channel - off;
        if (n > block_bytes) n = block_bytes;
        chains[ch]->iqp->acceptIqData(0, p + off, n);
      }
      g_sink = nullptr;
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < nthreads; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto &x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  for (uint32_t ch = 0; ch < n_channels; ch++) chainFree(chains[ch]);
  return std::chrono::duration<double>(t1 - t0).count();
}
#endif  // REF_HAS_IQP

}  // extern "C"
