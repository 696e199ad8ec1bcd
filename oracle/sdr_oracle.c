/* TEST INFRASTRUCTURE -- CPU restatement of the reference IQ->PCM path.
 * See sdr_oracle.h for who may use it and how it is pinned.
 *
 * Written from the reference's behaviour, not its text: histories are plain
 * shift registers (newest at [0]) instead of ring buffers, one generic Q15
 * FIR core serves both Decimator_int16 and FirFilter_int16, and every
 * C-semantics quirk the reference relies on implicitly (wrapping float->int16
 * casts, int8 negation of -128, tap 1.0 -> -32768) is spelled out.
 *
 * Build with -ffp-contract=off: every float operation below must round once,
 * in source order, exactly as the reference's x86-64 g++ -O3 build does.
 */
#include "sdr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ */
/* C-semantics helpers                                                 */
/* ------------------------------------------------------------------ */

/* (int16_t)floatValue as x86-64 g++ compiles it: cvttss2si to int32 (the
 * "integer indefinite" 0x80000000 when out of range or NaN), then keep the low
 * 16 bits. SURVEY A.5-1; reachable at FmDemodulator.cc:540 (k*pi ~ 69900). */
static int16_t f32_to_i16_wrap(float v) {
  int32_t r = (fabsf(v) < 2147483648.0f) ? (int32_t)v : INT32_MIN;
  return (int16_t)(uint16_t)(uint32_t)r;
}

/* int8 negate with wrap: -(-128) stays -128 (IqDataProcessor.cc:594-607). */
static int8_t neg_i8(int8_t v) { return (int8_t)(uint8_t)(0u - (uint8_t)v); }

/* ------------------------------------------------------------------ */
/* Q15 FIR / decimator: Decimator_int16.cc:41-79,176-238,310-351 and     */
/* FirFilter_int16.cc:33-64,151-213                                      */
/* ------------------------------------------------------------------ */
#define SDRO_MAX_TAPS 64

struct sdro_dec16 {
  int L, M;
  int16_t q[SDRO_MAX_TAPS];    /* quantised taps */
  int16_t hist[SDRO_MAX_TAPS]; /* hist[0] = newest sample in the pipeline */
  int16_t pend[16];            /* samples waiting for a full group of M */
  int npend;
};

/* Decimator_int16.cc:58-62: float product, round half away, truncating cast
 * (32768.0 -> -32768). */
static int16_t quantise_tap(float h) {
  float s = h * 32768;
  s = roundf(s);
  return (int16_t)(uint16_t)(uint32_t)(int32_t)s;
}

static void dec16_init(sdro_dec16 *d, int L, const float *h, int M) {
  memset(d, 0, sizeof(*d));
  d->L = L;
  d->M = M;
  for (int k = 0; k < L; k++) d->q[k] = quantise_tap(h[k]);
}

static void dec16_push(sdro_dec16 *d, int16_t x) {
  memmove(d->hist + 1, d->hist, (size_t)(d->L - 1) * sizeof(int16_t));
  d->hist[0] = x;
}

/* Decimator_int16.cc:193-234: rounding constant, ordered MAC with a clamp
 * after every tap, arithmetic shift, truncating cast. */
static int16_t dec16_mac(const sdro_dec16 *d) {
  int32_t acc = 1 << 14;
  for (int k = 0; k < d->L; k++) {
    acc += (int32_t)d->q[k] * (int32_t)d->hist[k];
    if (acc > 0x3fffffff) acc = 0x3fffffff;
    else if (acc < -0x40000000) acc = -0x40000000;
  }
  return (int16_t)(acc >> 15);
}

/* Decimator_int16.cc:310-351: buffer M samples, shift them all in, one MAC. */
static int dec16_step(sdro_dec16 *d, int16_t x, int16_t *y) {
  d->pend[d->npend++] = x;
  if (d->npend < d->M) return 0;
  for (int i = 0; i < d->M; i++) dec16_push(d, d->pend[i]);
  d->npend = 0;
  *y = dec16_mac(d);
  return 1;
}

static void dec16_reset(sdro_dec16 *d) {
  memset(d->hist, 0, sizeof(d->hist));
  memset(d->pend, 0, sizeof(d->pend));
  d->npend = 0;
}

sdro_dec16 *sdro_dec16_new(int L, const float *h, int M) {
  if (L < 1 || L > SDRO_MAX_TAPS || M < 1 || M > 16) return NULL;
  sdro_dec16 *d = (sdro_dec16 *)malloc(sizeof(*d));
  dec16_init(d, L, h, M);
  return d;
}
void sdro_dec16_free(sdro_dec16 *d) { free(d); }
void sdro_dec16_reset(sdro_dec16 *d) { dec16_reset(d); }
uint32_t sdro_dec16_run(sdro_dec16 *d, const int16_t *in, uint32_t n, int16_t *out) {
  uint32_t c = 0;
  for (uint32_t i = 0; i < n; i++) {
    int16_t y;
    if (dec16_step(d, in[i], &y)) out[c++] = y;
  }
  return c;
}
void sdro_dec16_taps(const sdro_dec16 *d, int16_t *q) { memcpy(q, d->q, (size_t)d->L * 2); }

sdro_fir16 *sdro_fir16_new(int L, const float *h) { return sdro_dec16_new(L, h, 1); }
void sdro_fir16_free(sdro_fir16 *f) { free(f); }
void sdro_fir16_reset(sdro_fir16 *f) { dec16_reset(f); }
void sdro_fir16_run(sdro_fir16 *f, const int16_t *in, uint32_t n, int16_t *out) {
  sdro_dec16_run(f, in, n, out);
}

/* ------------------------------------------------------------------ */
/* Generic multirate classes (not on the IQ->PCM path):     */
/*   Filters/Decimator.cc:168-209 (filterData), 281-322 (decimate)        */
/*   Filters/Interpolator.cc: createPolyphaseCoefficients, filterData,    */
/*     interpolate (p_i[k] = h[i + kL]; L outputs per input)               */
/*   Filters/Int16/Decimator_int16.cc, Interpolator_int16.cc: Q15 twins    */
/* Histories are shift registers, newest at [0].                          */
/* ------------------------------------------------------------------ */
struct sdro_mr {
  int kind, N, F, q; /* q = taps per output */
  float *hf;         /* float taps, prototype order */
  int16_t *hq;       /* Q15 taps, prototype order */
  float *xf;         /* float history, q entries */
  int16_t *xq;
  float *pf;         /* decimator: waiting samples */
  int16_t *pq;
  int npend;
};

sdro_mr *sdro_mr_new(int kind, int N, const float *h, int factor) {
  if (kind < 1 || kind > 4 || N < 1 || factor < 1) return NULL;
  int interp = (kind == 2 || kind == 4);
  if (interp && N % factor) return NULL;
  sdro_mr *m = (sdro_mr *)calloc(1, sizeof(*m));
  m->kind = kind;
  m->N = N;
  m->F = factor;
  m->q = interp ? N / factor : N;
  m->hf = (float *)calloc((size_t)N, sizeof(float));
  m->hq = (int16_t *)calloc((size_t)N, sizeof(int16_t));
  m->xf = (float *)calloc((size_t)m->q, sizeof(float));
  m->xq = (int16_t *)calloc((size_t)m->q, sizeof(int16_t));
  m->pf = (float *)calloc((size_t)factor, sizeof(float));
  m->pq = (int16_t *)calloc((size_t)factor, sizeof(int16_t));
  for (int k = 0; k < N; k++) { m->hf[k] = h[k]; m->hq[k] = quantise_tap(h[k]); }
  return m;
}
void sdro_mr_free(sdro_mr *m) {
  if (!m) return;
  free(m->hf); free(m->hq); free(m->xf); free(m->xq); free(m->pf); free(m->pq);
  free(m);
}
void sdro_mr_reset(sdro_mr *m) {
  memset(m->xf, 0, (size_t)m->q * sizeof(float));
  memset(m->xq, 0, (size_t)m->q * sizeof(int16_t));
  m->npend = 0;
}
/* taps h[first], h[first+step], ... against the history, newest first */
static float mr_mac_f(const sdro_mr *m, int first, int step) {
  float y = 0;
  for (int k = 0; k < m->q; k++) y = y + (m->hf[first + k * step] * m->xf[k]);
  return y;
}
static int16_t mr_mac_q(const sdro_mr *m, int first, int step) {
  int32_t acc = 1 << 14;
  for (int k = 0; k < m->q; k++) {
    acc += (int32_t)m->hq[first + k * step] * (int32_t)m->xq[k];
    if (acc > 0x3fffffff) acc = 0x3fffffff;
    else if (acc < -0x40000000) acc = -0x40000000;
  }
  return (int16_t)(acc >> 15);
}
uint64_t sdro_mr_run(sdro_mr *m, const void *in, uint64_t n, void *out) {
  const int is16 = m->kind >= 3, interp = (m->kind == 2 || m->kind == 4);
  uint64_t c = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (interp) {
      if (is16) {
        memmove(m->xq + 1, m->xq, (size_t)(m->q - 1) * sizeof(int16_t));
        m->xq[0] = ((const int16_t *)in)[i];
        for (int j = 0; j < m->F; j++) ((int16_t *)out)[c++] = mr_mac_q(m, j, m->F);
      } else {
        memmove(m->xf + 1, m->xf, (size_t)(m->q - 1) * sizeof(float));
        m->xf[0] = ((const float *)in)[i];
        for (int j = 0; j < m->F; j++) ((float *)out)[c++] = mr_mac_f(m, j, m->F);
      }
    } else if (is16) {
      m->pq[m->npend++] = ((const int16_t *)in)[i];
      if (m->npend < m->F) continue;
      for (int j = 0; j < m->F; j++) {
        memmove(m->xq + 1, m->xq, (size_t)(m->q - 1) * sizeof(int16_t));
        m->xq[0] = m->pq[j];
      }
      m->npend = 0;
      ((int16_t *)out)[c++] = mr_mac_q(m, 0, 1);
    } else {
      m->pf[m->npend++] = ((const float *)in)[i];
      if (m->npend < m->F) continue;
      for (int j = 0; j < m->F; j++) {
        memmove(m->xf + 1, m->xf, (size_t)(m->q - 1) * sizeof(float));
        m->xf[0] = m->pf[j];
      }
      m->npend = 0;
      ((float *)out)[c++] = mr_mac_f(m, 0, 1);
    }
  }
  return c;
}

/* ------------------------------------------------------------------ */
/* float FIR: FirFilter.cc:144-185. y = 0; y = y + h[k]*x[n-k], k up.    */
/* ------------------------------------------------------------------ */
struct sdro_fir {
  int L;
  float h[SDRO_MAX_TAPS];
  float hist[SDRO_MAX_TAPS];
};

static void fir_init(sdro_fir *f, int L, const float *h) {
  memset(f, 0, sizeof(*f));
  f->L = L;
  memcpy(f->h, h, (size_t)L * sizeof(float));
}
static float fir_step(sdro_fir *f, float x) {
  memmove(f->hist + 1, f->hist, (size_t)(f->L - 1) * sizeof(float));
  f->hist[0] = x;
  float y = 0;
  for (int k = 0; k < f->L; k++) y = y + (f->h[k] * f->hist[k]);
  return y;
}
sdro_fir *sdro_fir_new(int L, const float *h) {
  if (L < 1 || L > SDRO_MAX_TAPS) return NULL;
  sdro_fir *f = (sdro_fir *)malloc(sizeof(*f));
  fir_init(f, L, h);
  return f;
}
void sdro_fir_free(sdro_fir *f) { free(f); }
void sdro_fir_run(sdro_fir *f, const float *in, uint32_t n, float *out) {
  for (uint32_t i = 0; i < n; i++) out[i] = fir_step(f, in[i]);
}

/* ------------------------------------------------------------------ */
/* float IIR, Direct Form I: IirFilter.cc:161-176,199-229,250-266.        */
/* The recursive sum starts at the ring slot about to be overwritten, so  */
/* a[0] meets y[n-na] and a[k>=1] meets y[n-k]. With one   */
/* denominator tap -- the only case on the path -- that is a[0]*y[n-1].   */
/* ------------------------------------------------------------------ */
struct sdro_iir {
  sdro_fir num;
  int na;
  float a[SDRO_MAX_TAPS];
  float yhist[SDRO_MAX_TAPS]; /* yhist[0] = y[n-1] */
};

static void iir_init(sdro_iir *f, int nb, const float *b, int na, const float *a) {
  memset(f, 0, sizeof(*f));
  fir_init(&f->num, nb, b);
  f->na = na;
  memcpy(f->a, a, (size_t)na * sizeof(float));
}
static float iir_step(sdro_iir *f, float x) {
  float y = fir_step(&f->num, x);
  float r = 0;
  r = r + (f->a[0] * f->yhist[f->na - 1]);
  for (int k = 1; k < f->na; k++) r = r + (f->a[k] * f->yhist[k - 1]);
  y -= r;
  memmove(f->yhist + 1, f->yhist, (size_t)(f->na - 1) * sizeof(float));
  f->yhist[0] = y;
  return y;
}
static void iir_reset(sdro_iir *f) {
  memset(f->num.hist, 0, sizeof(f->num.hist));
  memset(f->yhist, 0, sizeof(f->yhist));
}
sdro_iir *sdro_iir_new(int nb, const float *b, int na, const float *a) {
  if (nb < 1 || nb > SDRO_MAX_TAPS || na < 1 || na > SDRO_MAX_TAPS) return NULL;
  sdro_iir *f = (sdro_iir *)malloc(sizeof(*f));
  iir_init(f, nb, b, na, a);
  return f;
}
void sdro_iir_free(sdro_iir *f) { free(f); }
void sdro_iir_run(sdro_iir *f, const float *in, uint32_t n, float *out) {
  for (uint32_t i = 0; i < n; i++) out[i] = iir_step(f, in[i]);
}

/* ------------------------------------------------------------------ */
/* Filter designs (values: the coefficient tables at the top of each      */
/* reference demodulator source; all symmetric, so half is listed).       */
/* ------------------------------------------------------------------ */
static void mirror(float *dst, const float *half, int nhalf, int odd_sign) {
  /* even-length linear phase: dst = half ++ reverse(half) (odd_sign=+1),
   * antisymmetric uses -1. */
  for (int i = 0; i < nhalf; i++) {
    dst[i] = half[i];
    dst[2 * nhalf - 1 - i] = (float)odd_sign * half[i];
  }
}

/* AmDemodulator.cc:14-68 == SsbDemodulator.cc:14-68 */
static const float H_AM1[4] = {0.0242683f, 0.0766338f, 0.1457589f, 0.1959036f};
static const float H_AM2[6] = {0.0057496f, 0.0263853f, 0.0605301f, 0.1074406f, 0.1523486f, 0.1804951f};
static const float H_AM3[8] = {0.0116487f, 0.0152694f, -0.0109804f, -0.0611915f,
                               -0.0736143f, 0.0187617f, 0.1988190f, 0.3481364f};
/* FmDemodulator.cc:14-47 */
static const float H_FM_TUNER[16] = {0.0041331f, 0.0054174f, 0.0076016f, 0.0115481f, 0.0151685f, 0.0203192f,
                                     0.0251608f, 0.0311322f, 0.0366372f, 0.0427168f, 0.0480527f, 0.0533425f,
                                     0.0575831f, 0.0611914f, 0.0635413f, 0.0648239f};
/* FmDemodulator.cc:50-64 == WbFmDemodulator.cc postDemodDecimator2 */
static const float H_FM_POST[6] = {0.0022977f, 0.0237042f, 0.0605386f, 0.1127073f, 0.1645167f, 0.1971107f};
/* FmDemodulator.cc:67-109 == WbFmDemodulator.cc audioDecimator */
static const float H_AUDIO40[20] = {0.0015969f, -0.0111080f, -0.0270501f, -0.0265610f, -0.0023190f,
                                    0.0180618f, 0.0065495f, -0.0183409f, -0.0133345f, 0.0184489f,
                                    0.0230891f, -0.0161248f, -0.0363745f, 0.0091343f, 0.0550219f,
                                    0.0070312f, -0.0862280f, -0.0497761f, 0.1793543f, 0.4145808f};
/* WbFmDemodulator.cc:17-35 */
static const float H_WB_PRE[8] = {-0.0157211f, -0.0325959f, 0.0092996f, 0.0621217f,
                                  -0.0148595f, -0.0989456f, 0.1182989f, 0.4862333f};
/* WbFmDemodulator.cc:38-48 */
static const float H_WB_DEC1[4] = {0.0243699f, 0.0769537f, 0.1463572f, 0.1967096f};
/* SsbDemodulator.cc:74-107, odd taps of the 31-tap Hilbert transformer
 * (even taps are 0; antisymmetric about tap 15) */
static const float H_HILBERT_ODD[8] = {-0.0033953f, -0.0058652f, -0.0134385f, -0.0281423f,
                                       -0.0534836f, -0.0980394f, -0.1935638f, -0.6302204f};

static void design(int id, float *h, int *L) {
  memset(h, 0, SDRO_MAX_TAPS * sizeof(float));
  switch (id) {
    case 0: mirror(h, H_AM1, 4, 1); *L = 8; break;
    case 1: mirror(h, H_AM2, 6, 1); *L = 12; break;
    case 2: mirror(h, H_AM3, 8, 1); *L = 16; break;
    case 3: mirror(h, H_FM_TUNER, 16, 1); *L = 32; break;
    case 4: mirror(h, H_FM_POST, 6, 1); *L = 12; break;
    case 5: mirror(h, H_AUDIO40, 20, 1); *L = 40; break;
    case 6: mirror(h, H_WB_PRE, 8, 1); *L = 16; break;
    case 7: mirror(h, H_WB_DEC1, 4, 1); *L = 8; break;
    case 8: h[15] = 1; *L = 16; break; /* SsbDemodulator.cc:71 delay line */
    case 9:
      for (int i = 0; i < 8; i++) {
        h[2 * i] = H_HILBERT_ODD[i];
        h[30 - 2 * i] = -H_HILBERT_ODD[i];
      }
      *L = 31;
      break;
    default: *L = 0; break;
  }
}

int sdro_q15_taps(int id, int16_t *q) {
  float h[SDRO_MAX_TAPS];
  int L;
  design(id, h, &L);
  for (int k = 0; k < L; k++) q[k] = quantise_tap(h[k]);
  return L;
}

float sdro_atan2f(int q, int i) { return (float)atan2((double)q, (double)i); }

/* ------------------------------------------------------------------ */
/* Demodulators                                                          */
/* ------------------------------------------------------------------ */
typedef struct { int16_t *out; uint32_t cap, count; } pcm_sink;
static void emit(pcm_sink *s, int16_t v) {
  if (s->out && s->count < s->cap) s->out[s->count] = v;
  s->count++;
}

static const float DC_B[2] = {1, -1};   /* AmDemodulator.cc:72-73 */
static const float DC_A[1] = {-0.95f};

/* magnitude estimate shared by AM (AmDemodulator.cc:441-458) */
static int16_t mag_estimate(int16_t i, int16_t q) {
  int16_t im = (int16_t)abs((int)i);
  int16_t qm = (int16_t)abs((int)q);
  if (im > qm) return (int16_t)(im + (qm >> 1));
  return (int16_t)(qm + (im >> 1));
}

typedef struct {
  sdro_dec16 s1i, s1q, s2i, s2q, s3i, s3q;
  sdro_iir dc;
  float gain;
} am_t;

static void front3_init(sdro_dec16 *s1i, sdro_dec16 *s1q, sdro_dec16 *s2i, sdro_dec16 *s2q,
                        sdro_dec16 *s3i, sdro_dec16 *s3q) {
  float h[SDRO_MAX_TAPS];
  int L;
  design(0, h, &L); dec16_init(s1i, L, h, 4); dec16_init(s1q, L, h, 4);
  design(1, h, &L); dec16_init(s2i, L, h, 4); dec16_init(s2q, L, h, 4);
  design(2, h, &L); dec16_init(s3i, L, h, 2); dec16_init(s3q, L, h, 2);
}

static void am_init(am_t *d) {
  front3_init(&d->s1i, &d->s1q, &d->s2i, &d->s2q, &d->s3i, &d->s3q);
  iir_init(&d->dc, 2, DC_B, 1, DC_A);
  d->gain = 300; /* AmDemodulator.cc:102 */
}
static void am_reset(am_t *d) { /* AmDemodulator.cc:232-247 */
  dec16_reset(&d->s1i); dec16_reset(&d->s1q); dec16_reset(&d->s2i);
  dec16_reset(&d->s2q); dec16_reset(&d->s3i); dec16_reset(&d->s3q);
  iir_reset(&d->dc);
}

/* three-stage 32:1 cascade of one arm (AmDemodulator.cc:349-374) */
static uint32_t cascade3(sdro_dec16 *s1, sdro_dec16 *s2, sdro_dec16 *s3, const int8_t *buf,
                         uint32_t nbytes, uint32_t first, int16_t *dst) {
  uint32_t n = 0;
  for (uint32_t i = first; i < nbytes + first; i += 2) {
    int16_t v;
    if (!dec16_step(s1, (int16_t)buf[i], &v)) continue;
    if (!dec16_step(s2, v, &v)) continue;
    if (!dec16_step(s3, v, &v)) continue;
    dst[n++] = v;
  }
  return n;
}

/* AmDemodulator::acceptIqData, AmDemodulator.cc:297-315 */
static void am_accept(am_t *d, const int8_t *buf, uint32_t nbytes, pcm_sink *s) {
  static _Thread_local int16_t id[16384], qd[16384];
  uint32_t n = cascade3(&d->s1i, &d->s2i, &d->s3i, buf, nbytes, 0, id);
  n = cascade3(&d->s1q, &d->s2q, &d->s3q, buf, nbytes, 1, qd);
  for (uint32_t m = 0; m < n; m++) {
    float y = iir_step(&d->dc, (float)mag_estimate(id[m], qd[m]));
    emit(s, f32_to_i16_wrap(d->gain * y)); /* AmDemodulator.cc:467 */
  }
}

typedef struct {
  sdro_dec16 s1i, s1q, s2i, s2q, s3i, s3q;
  sdro_fir16 delay, hilbert;
  sdro_iir dc;
  float gain;
  int lsb;
} ssb_t;

static void ssb_init(ssb_t *d) {
  float h[SDRO_MAX_TAPS];
  int L;
  front3_init(&d->s1i, &d->s1q, &d->s2i, &d->s2q, &d->s3i, &d->s3q);
  design(8, h, &L); dec16_init(&d->delay, L, h, 1);
  design(9, h, &L); dec16_init(&d->hilbert, L, h, 1);
  iir_init(&d->dc, 2, DC_B, 1, DC_A);
  d->gain = 300; /* SsbDemodulator.cc:146 */
  d->lsb = 1;    /* SsbDemodulator.cc:143 */
}
static void ssb_reset(ssb_t *d) { /* SsbDemodulator.cc:297-314 */
  dec16_reset(&d->s1i); dec16_reset(&d->s1q); dec16_reset(&d->s2i);
  dec16_reset(&d->s2q); dec16_reset(&d->s3i); dec16_reset(&d->s3q);
  dec16_reset(&d->delay); dec16_reset(&d->hilbert);
  iir_reset(&d->dc);
}
/* SsbDemodulator::acceptIqData, SsbDemodulator.cc:420-438,563-598 */
static void ssb_accept(ssb_t *d, const int8_t *buf, uint32_t nbytes, pcm_sink *s) {
  static _Thread_local int16_t id[16384], qd[16384];
  uint32_t n = cascade3(&d->s1i, &d->s2i, &d->s3i, buf, nbytes, 0, id);
  n = cascade3(&d->s1q, &d->s2q, &d->s3q, buf, nbytes, 1, qd);
  for (uint32_t m = 0; m < n; m++) {
    int16_t iDelayed, qShifted;
    dec16_step(&d->delay, id[m], &iDelayed);
    dec16_step(&d->hilbert, qd[m], &qShifted);
    float v = d->lsb ? (float)((int)iDelayed - (int)qShifted) : (float)((int)iDelayed + (int)qShifted);
    v = iir_step(&d->dc, v);
    emit(s, f32_to_i16_wrap(d->gain * v));
  }
}

/* while (d > M_PI) d -= 2*M_PI: the compare and the subtraction are in double,
 * the store rounds back to float (FmDemodulator.cc:486-494). */
static float wrap_pi(float d) {
  while (d > M_PI) d = (float)((double)d - (2 * M_PI));
  while (d < (-M_PI)) d = (float)((double)d + (2 * M_PI));
  return d;
}

typedef struct {
  sdro_dec16 ti, tq, post, audio;
  sdro_fir diff;
  float gain;
} fm_t;

static void fm_init(fm_t *d) {
  float h[SDRO_MAX_TAPS];
  int L;
  /* FmDemodulator.cc:113-122: -1/16 and 1/16 are integer divisions */
  const float diff[7] = {-1 / 16, 0, 1, 0, -1, 0, 1 / 16};
  design(3, h, &L); dec16_init(&d->ti, L, h, 4); dec16_init(&d->tq, L, h, 4);
  design(4, h, &L); dec16_init(&d->post, L, h, 4);
  design(5, h, &L); dec16_init(&d->audio, L, h, 2);
  fir_init(&d->diff, 7, diff);
  d->gain = 64000 / (2 * M_PI); /* FmDemodulator.cc:154 */
}
static void fm_reset(fm_t *d) { /* FmDemodulator.cc:271-283 */
  dec16_reset(&d->ti); dec16_reset(&d->tq); dec16_reset(&d->post); dec16_reset(&d->audio);
  memset(d->diff.hist, 0, sizeof(d->diff.hist));
}
/* FmDemodulator::acceptIqData, FmDemodulator.cc:334-352,376-423,460-504,526-560 */
static void fm_accept(fm_t *d, int variant, const int8_t *buf, uint32_t nbytes, pcm_sink *s) {
  static _Thread_local int16_t id[16384], qd[16384];
  uint32_t n = 0, nq = 0;
  for (uint32_t i = 0; i < nbytes; i += 2) {
    int16_t v;
    if (dec16_step(&d->ti, (int16_t)buf[i], &v)) id[n++] = v;
  }
  for (uint32_t i = 1; i < nbytes + 1; i += 2) {
    int16_t v;
    if (dec16_step(&d->tq, (int16_t)buf[i], &v)) qd[nq++] = v;
  }
  float k;
  if (variant == SDRO_VARIANT_RESEARCH) {
    k = d->gain;
  } else {
    k = d->gain / 15000; /* FmDemodulator.cc:465-471 */
    k *= 32767;
  }
  for (uint32_t m = 0; m < n; m++) {
    float theta = (float)atan2((double)qd[m], (double)id[m]);
    float dth = wrap_pi(fir_step(&d->diff, theta));
    float dem = k * dth;
    int16_t v;
    if (!dec16_step(&d->post, f32_to_i16_wrap(dem), &v)) continue;
    if (!dec16_step(&d->audio, v, &v)) continue;
    emit(s, v);
  }
}

typedef struct {
  sdro_fir16 pi, pq;
  sdro_dec16 d1, d2, audio;
  sdro_iir deemph;
  float prev_theta;
  float gain;
} wbfm_t;

static float g_atan_lut[256][256];
static pthread_once_t g_lut_once = PTHREAD_ONCE_INIT;
static void build_lut(void) { /* WbFmDemodulator.cc:159-170 */
  for (int x = 0; x < 256; x++)
    for (int y = 0; y < 256; y++)
      g_atan_lut[y][x] = (float)atan2((double)y - 128, (double)x - 128);
}

static void wbfm_init(wbfm_t *d, int variant) {
  float h[SDRO_MAX_TAPS];
  int L;
  const float b[2] = {0.0253863f, 0.0253863f}; /* WbFmDemodulator.cc:113-121 */
  const float a[1] = {-0.9492274f};
  pthread_once(&g_lut_once, build_lut);
  design(6, h, &L); dec16_init(&d->pi, L, h, 1); dec16_init(&d->pq, L, h, 1);
  design(7, h, &L); dec16_init(&d->d1, L, h, 4);
  design(4, h, &L); dec16_init(&d->d2, L, h, 4);
  design(5, h, &L); dec16_init(&d->audio, L, h, 2);
  iir_init(&d->deemph, 2, b, 1, a);
  d->prev_theta = 0;
  /* WbFmDemodulator.cc:173 vs the research tree's 64000 */
  d->gain = (variant == SDRO_VARIANT_RESEARCH) ? (float)(64000 / (2 * M_PI)) : (float)(256000 / (2 * M_PI));
}
static void wbfm_reset(wbfm_t *d) { /* WbFmDemodulator.cc:304-320: IIR NOT reset */
  dec16_reset(&d->pi); dec16_reset(&d->pq); dec16_reset(&d->d1);
  dec16_reset(&d->d2); dec16_reset(&d->audio);
  d->prev_theta = 0;
}
/* WbFmDemodulator::acceptIqData, WbFmDemodulator.cc:383-411,436-494,515-562.
 * The IQ buffer is overwritten with the pre-filter output, as the reference does. */
static void wbfm_accept(wbfm_t *d, int variant, int8_t *buf, uint32_t nbytes, pcm_sink *s) {
  for (uint32_t i = 0; i < nbytes; i += 2) {
    int16_t v;
    dec16_step(&d->pi, (int16_t)buf[i], &v);
    buf[i] = (int8_t)(uint8_t)(uint16_t)v;
    dec16_step(&d->pq, (int16_t)buf[i + 1], &v);
    buf[i + 1] = (int8_t)(uint8_t)(uint16_t)v;
  }
  float k;
  if (variant == SDRO_VARIANT_RESEARCH) {
    k = d->gain;
  } else {
    k = d->gain / 75000; /* WbFmDemodulator.cc:444-450 */
    k *= 32767;
  }
  uint32_t n = nbytes / 2;
  for (uint32_t m = 0; m < n; m++) {
    uint8_t ii = (uint8_t)((uint8_t)buf[2 * m] + 128);
    uint8_t qi = (uint8_t)((uint8_t)buf[2 * m + 1] + 128);
    float theta = g_atan_lut[qi][ii];
    float dth = wrap_pi(theta - d->prev_theta);
    float y = iir_step(&d->deemph, k * dth);
    d->prev_theta = theta;
    int16_t v;
    if (!dec16_step(&d->d1, f32_to_i16_wrap(y), &v)) continue;
    if (!dec16_step(&d->d2, v, &v)) continue;
    if (!dec16_step(&d->audio, v, &v)) continue;
    emit(s, v);
  }
}

/* ------------------------------------------------------------------ */
/* IqDataProcessor front end + mode switch                               */
/* ------------------------------------------------------------------ */
/* ------------------------------------------------------------------ */
/* Squelch: SignalDetector.cc:205-273, DbfsCalculator.cc:43-147,         */
/* SignalTracker.cc:104-145, Squelch.cc:227-273                          */
/* ------------------------------------------------------------------ */
typedef struct {
  int32_t threshold;   /* dBFS; IqDataProcessor.cc:41 starts at -200 = always open */
  uint32_t gain_db;    /* radio_adjustableReceiveGainInDb, owned by Radio.cc */
  int tracking;        /* SignalTracker state: 0 NoSignal, 1 Tracking */
  uint32_t magnitude;  /* mean magnitude of the last block */
  int allowed;         /* gate decision of the last block */
  int32_t db_table[257];
} squelch_t;

static void squelch_init(squelch_t *q) {
  memset(q, 0, sizeof(*q));
  q->threshold = -200;
  /* DbfsCalculator(7): full scale 127 -> 42 dB; table of (int32_t)(20*log10f(i)) */
  for (int i = 1; i <= 256; i++) {
    float db = 20 * log10f((float)i);
    q->db_table[i] = (int32_t)db;
  }
  q->db_table[0] = q->db_table[1];
}

/* buf: signed (already offset-corrected) IQ. Returns signalAllowed. */
static int squelch_run(squelch_t *q, const int8_t *buf, uint32_t nbytes) {
  uint32_t sum = 0, n = nbytes / 2;
  for (uint32_t i = 0; i < nbytes; i += 2) {
    uint8_t im = (uint8_t)abs((int)buf[i]), qm = (uint8_t)abs((int)buf[i + 1]);
    uint8_t m = (im > qm) ? (uint8_t)(im + (qm >> 1)) : (uint8_t)(qm + (im >> 1));
    sum += m;
  }
  uint32_t mag = n ? sum / n : 0;
  q->magnitude = mag;
  uint32_t idx = mag > 127u ? 127u : mag; /* clip to full scale; <= 256 so no halving loop */
  int32_t db = q->db_table[idx] + 0 - (int32_t)42u;
  db -= q->gain_db;
  int present = db >= q->threshold;
  int allowed;
  if (!q->tracking) { /* NoSignal */
    allowed = present;   /* START_OF_SIGNAL -> allowed, NOISE -> not */
    q->tracking = present;
  } else {              /* Tracking: SIGNAL_PRESENT or END_OF_SIGNAL (squelch tail) */
    allowed = 1;
    q->tracking = present;
  }
  q->allowed = allowed;
  return allowed;
}

struct sdro_chain {
  int variant;
  int mode;
  am_t am;
  fm_t fm;
  wbfm_t wbfm;
  ssb_t ssb;
  squelch_t sq;
};

sdro_chain *sdro_chain_new(int variant) {
  sdro_chain *c = (sdro_chain *)calloc(1, sizeof(*c));
  c->variant = variant;
  c->mode = SDRO_MODE_NONE; /* IqDataProcessor.cc:38 */
  am_init(&c->am);
  fm_init(&c->fm);
  wbfm_init(&c->wbfm, variant);
  ssb_init(&c->ssb);
  squelch_init(&c->sq);
  return c;
}
void sdro_chain_free(sdro_chain *c) { free(c); }

/* IqDataProcessor::setDemodulatorMode, IqDataProcessor.cc:236-262 */
void sdro_chain_set_mode(sdro_chain *c, int mode) {
  c->mode = mode;
  if (mode == SDRO_MODE_LSB) c->ssb.lsb = 1;
  if (mode == SDRO_MODE_USB) c->ssb.lsb = 0;
}
void sdro_chain_set_gain(sdro_chain *c, int kind, float gain) {
  switch (kind) {
    case SDRO_KIND_AM: c->am.gain = gain; break;
    case SDRO_KIND_FM: c->fm.gain = gain; break;
    case SDRO_KIND_WBFM: c->wbfm.gain = gain; break;
    case SDRO_KIND_SSB: c->ssb.gain = gain; break;
  }
}
void sdro_chain_reset(sdro_chain *c, int kind) {
  switch (kind) {
    case SDRO_KIND_AM: am_reset(&c->am); break;
    case SDRO_KIND_FM: fm_reset(&c->fm); break;
    case SDRO_KIND_WBFM: wbfm_reset(&c->wbfm); break;
    case SDRO_KIND_SSB: ssb_reset(&c->ssb); break;
  }
}

static void dispatch(sdro_chain *c, int mode, int8_t *buf, uint32_t nbytes, pcm_sink *s) {
  /* The reference's member arrays cap a call at 32768 bytes (FmDemodulator.h:50).
   * Every chain is invariant to how the stream is cut into calls,
   * so longer buffers are simply walked in reference-sized pieces. */
  for (uint32_t off = 0; off < nbytes; off += 32768) {
    uint32_t n = nbytes - off < 32768 ? nbytes - off : 32768;
    int8_t *p = buf + off;
    switch (mode) { /* IqDataProcessor.cc:793-835 */
      case SDRO_MODE_AM: am_accept(&c->am, p, n, s); break;
      case SDRO_MODE_FM: fm_accept(&c->fm, c->variant, p, n, s); break;
      case SDRO_MODE_WBFM: wbfm_accept(&c->wbfm, c->variant, p, n, s); break;
      case SDRO_MODE_LSB:
      case SDRO_MODE_USB: ssb_accept(&c->ssb, p, n, s); break;
      default: break;
    }
  }
}

/* IqDataProcessor::acceptIqData, IqDataProcessor.cc:722-840, with the squelch
 * at its default always-open threshold (IqDataProcessor.cc:41). */
/* The conversion half of acceptIqData (IqDataProcessor.cc:735-751), in place. What it leaves
 * in the buffer is what the IQ dump sends (IqDataProcessor.cc:756-760) and what a .iq file
 * holds (demod.cc:8-11). */
void sdro_front_end(uint8_t *ubuf, uint32_t nbytes) {
  int8_t *buf = (int8_t *)ubuf;
  for (uint32_t i = 0; i < nbytes; i++) buf[i] = (int8_t)(uint8_t)(ubuf[i] - 128u);
  for (uint32_t i = 0; i + 7 < nbytes; i += 8) {
    int8_t x, y;
    x = buf[i + 2]; y = buf[i + 3]; buf[i + 2] = neg_i8(y); buf[i + 3] = x;
    x = buf[i + 4]; y = buf[i + 5]; buf[i + 4] = neg_i8(x); buf[i + 5] = neg_i8(y);
    x = buf[i + 6]; y = buf[i + 7]; buf[i + 6] = y; buf[i + 7] = neg_i8(x);
  }
}

/* UdpClient::sendData (UdpClient.cc:173-241): n bytes leave as floor(n/2048) datagrams of
 * 2048 bytes and one of the remainder. Writes the sizes, returns how many. */
uint32_t sdro_dump_datagrams(uint32_t nbytes, uint32_t *sizes, uint32_t cap) {
  uint32_t k = 0;
  for (uint32_t off = 0; off + 2048 <= nbytes; off += 2048, k++)
    if (k < cap) sizes[k] = 2048;
  if (nbytes % 2048) {
    if (k < cap) sizes[k] = nbytes % 2048;
    k++;
  }
  return k;
}

uint32_t sdro_chain_accept_u8(sdro_chain *c, uint8_t *ubuf, uint32_t nbytes, int16_t *pcm,
                              uint32_t cap) {
  int8_t *buf = (int8_t *)ubuf;
  pcm_sink s = {pcm, cap, 0};
  for (uint32_t i = 0; i < nbytes; i++) buf[i] = (int8_t)(uint8_t)(ubuf[i] - 128u);
  /* upconvertByFsOver4, IqDataProcessor.cc:567-611: multiply by 1, j, -1, -j;
   * the phase restarts at the head of every call. */
  for (uint32_t i = 0; i + 7 < nbytes; i += 8) {
    int8_t x, y;
    x = buf[i + 2]; y = buf[i + 3]; buf[i + 2] = neg_i8(y); buf[i + 3] = x;
    x = buf[i + 4]; y = buf[i + 5]; buf[i + 4] = neg_i8(x); buf[i + 5] = neg_i8(y);
    x = buf[i + 6]; y = buf[i + 7]; buf[i + 6] = y; buf[i + 7] = neg_i8(x);
  }
  /* IqDataProcessor.cc:764-765, 793: the squelch sees the converted block; a closed
   * squelch skips the demodulator altogether (its state does not advance) */
  if (squelch_run(&c->sq, buf, nbytes)) dispatch(c, c->mode, buf, nbytes, &s);
  return s.count;
}

/* IqDataProcessor::setSignalDetectThreshold (IqDataProcessor.cc:284-295) */
void sdro_chain_set_threshold(sdro_chain *c, int32_t threshold) { c->sq.threshold = threshold; }
/* radio_adjustableReceiveGainInDb (IqDataProcessor.cc:8, 765) */
void sdro_chain_set_rx_gain(sdro_chain *c, uint32_t gain_db) { c->sq.gain_db = gain_db; }
/* what the signal-state and signal-magnitude callbacks deliver (IqDataProcessor.cc:771-790) */
void sdro_chain_signal(const sdro_chain *c, int *allowed, uint32_t *magnitude) {
  if (allowed) *allowed = c->sq.allowed;
  if (magnitude) *magnitude = c->sq.magnitude;
}

uint32_t sdro_chain_accept_s8(sdro_chain *c, int mode, int8_t *buf, uint32_t nbytes, int16_t *pcm,
                              uint32_t cap) {
  pcm_sink s = {pcm, cap, 0};
  if (mode == SDRO_MODE_LSB) c->ssb.lsb = 1;
  if (mode == SDRO_MODE_USB) c->ssb.lsb = 0;
  dispatch(c, mode, buf, nbytes, &s);
  return s.count;
}

/* ------------------------------------------------------------------ */
/* host-thread channel bank                                              */
/* ------------------------------------------------------------------ */
typedef struct {
  const uint8_t *modes;
  uint32_t lo, hi;
  uint8_t *iq;
  uint64_t bytes_per_channel;
  uint32_t block_bytes;
  int16_t *pcm;
  int variant;
} bank_job;

static void *bank_worker(void *arg) {
  bank_job *j = (bank_job *)arg;
  uint64_t ppc = j->bytes_per_channel / 64;
  for (uint32_t ch = j->lo; ch < j->hi; ch++) {
    sdro_chain *c = sdro_chain_new(j->variant);
    sdro_chain_set_mode(c, j->modes[ch]);
    uint8_t *p = j->iq + (uint64_t)ch * j->bytes_per_channel;
    int16_t *out = j->pcm ? j->pcm + (uint64_t)ch * ppc : NULL;
    uint64_t done = 0;
    for (uint64_t off = 0; off < j->bytes_per_channel; off += j->block_bytes) {
      uint64_t n = j->bytes_per_channel - off;
      if (n > j->block_bytes) n = j->block_bytes;
      uint32_t got = sdro_chain_accept_u8(c, p + off, (uint32_t)n, out ? out + done : NULL,
                                          (uint32_t)(ppc - done));
      done += got;
    }
    sdro_chain_free(c);
  }
  return NULL;
}

double sdro_bank_run(const uint8_t *modes, uint32_t n_channels, uint8_t *iq,
                     uint64_t bytes_per_channel, uint32_t block_bytes, int16_t *pcm,
                     uint32_t nthreads, int variant) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  bank_job *jobs = (bank_job *)calloc(nthreads, sizeof(bank_job));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (uint32_t t = 0; t < nthreads; t++) {
    jobs[t].modes = modes;
    jobs[t].lo = (uint32_t)((uint64_t)n_channels * t / nthreads);
    jobs[t].hi = (uint32_t)((uint64_t)n_channels * (t + 1) / nthreads);
    jobs[t].iq = iq;
    jobs[t].bytes_per_channel = bytes_per_channel;
    jobs[t].block_bytes = block_bytes;
    jobs[t].pcm = pcm;
    jobs[t].variant = variant;
    if (t > 0) pthread_create(&th[t], NULL, bank_worker, &jobs[t]);
  }
  bank_worker(&jobs[0]);
  for (uint32_t t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(jobs);
  free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
