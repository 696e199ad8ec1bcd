/* TEST INFRASTRUCTURE -- CPU restatement of the reference IQ->PCM path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this library. The product path
 * (rtlsdrdiags_b200/csrc, include/sdr_b200.h) never links or calls it.
 *
 * Parity status: PINNED. The restatement is checked bit-for-bit against the
 * reference itself, compiled unmodified from /root/reference into
 * oracle/_ref/ (oracle/Makefile), and against golden vectors that the
 * compiled reference produced (tests/golden/, tests/golden/make_golden.py).
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* IqDataProcessor::demodulatorType, radioDiags/hdr_diags/IqDataProcessor.h:20 */
enum { SDRO_MODE_NONE = 0, SDRO_MODE_AM = 1, SDRO_MODE_FM = 2, SDRO_MODE_WBFM = 3,
       SDRO_MODE_LSB = 4, SDRO_MODE_USB = 5 };
/* which demodulator object a gain / reset addresses */
enum { SDRO_KIND_AM = 1, SDRO_KIND_FM = 2, SDRO_KIND_WBFM = 3, SDRO_KIND_SSB = 4 };
/* FM/WBFM PCM scaling: radioDiags/ tree vs demodulatorResearch/ tree */
enum { SDRO_VARIANT_RADIODIAGS = 0, SDRO_VARIANT_RESEARCH = 1 };

/* ---- filter building blocks ---- */
typedef struct sdro_dec16 sdro_dec16;   /* Decimator_int16 */
sdro_dec16 *sdro_dec16_new(int L, const float *h, int M);
void sdro_dec16_free(sdro_dec16 *d);
void sdro_dec16_reset(sdro_dec16 *d);
uint32_t sdro_dec16_run(sdro_dec16 *d, const int16_t *in, uint32_t n, int16_t *out);
void sdro_dec16_taps(const sdro_dec16 *d, int16_t *q);

typedef struct sdro_dec16 sdro_fir16;   /* FirFilter_int16 == decimator with M=1 */
sdro_fir16 *sdro_fir16_new(int L, const float *h);
void sdro_fir16_free(sdro_fir16 *f);
void sdro_fir16_reset(sdro_fir16 *f);
void sdro_fir16_run(sdro_fir16 *f, const int16_t *in, uint32_t n, int16_t *out);

typedef struct sdro_fir sdro_fir;       /* FirFilter (float) */
sdro_fir *sdro_fir_new(int L, const float *h);
void sdro_fir_free(sdro_fir *f);
void sdro_fir_run(sdro_fir *f, const float *in, uint32_t n, float *out);

typedef struct sdro_iir sdro_iir;       /* IirFilter (float, DF-I) */
sdro_iir *sdro_iir_new(int nb, const float *b, int na, const float *a);
void sdro_iir_free(sdro_iir *f);
void sdro_iir_run(sdro_iir *f, const float *in, uint32_t n, float *out);

/* ---- the generic multirate classes the path does not instantiate:
 * kind 1 Decimator (float), 2 Interpolator (float), 3 Decimator_int16, 4 Interpolator_int16;
 * any tap count. Samples are float for kinds 1-2 and int16_t for kinds 3-4. ---- */
typedef struct sdro_mr sdro_mr;
sdro_mr *sdro_mr_new(int kind, int N, const float *h, int factor);
void sdro_mr_free(sdro_mr *m);
void sdro_mr_reset(sdro_mr *m);
/* feeds n samples one by one; returns the outputs written (n*L, or one per M inputs) */
uint64_t sdro_mr_run(sdro_mr *m, const void *in, uint64_t n, void *out);

/* ---- one channel: IqDataProcessor + Am/Fm/WbFm/Ssb demodulators ---- */
typedef struct sdro_chain sdro_chain;
sdro_chain *sdro_chain_new(int variant);
void sdro_chain_free(sdro_chain *c);
void sdro_chain_set_mode(sdro_chain *c, int mode);
void sdro_chain_set_gain(sdro_chain *c, int kind, float gain);
void sdro_chain_reset(sdro_chain *c, int kind);
/* IqDataProcessor::acceptIqData: u8 offset-binary in, converted + rotated in
 * place, dispatched on the current mode. Returns PCM samples produced. */
uint32_t sdro_chain_accept_u8(sdro_chain *c, uint8_t *buf, uint32_t nbytes,
                              int16_t *pcm, uint32_t cap);
/* XDemodulator::acceptIqData: signed, already rotated IQ, straight into the
 * demodulator that `mode` selects (LSB/USB also set the sideband). */
uint32_t sdro_chain_accept_s8(sdro_chain *c, int mode, int8_t *buf, uint32_t nbytes,
                              int16_t *pcm, uint32_t cap);

/* squelch: threshold in dBFS (default -200 = always open), tuner gain in dB, and the
 * state / mean magnitude the reference hands to its signal callbacks after a block */
void sdro_chain_set_threshold(sdro_chain *c, int32_t threshold);
void sdro_chain_set_rx_gain(sdro_chain *c, uint32_t gain_db);
void sdro_chain_signal(const sdro_chain *c, int *allowed, uint32_t *magnitude);

/* IQ dump: the in-place conversion + Fs/4 rotation of acceptIqData, i.e. the bytes the
 * reference hands to UdpClient::sendData, and the datagram sizes sendData cuts them into */
void sdro_front_end(uint8_t *buf, uint32_t nbytes);
uint32_t sdro_dump_datagrams(uint32_t nbytes, uint32_t *sizes, uint32_t cap);

/* quantised taps of every Q15 filter on the path, for table cross-checks.
 * id: 0 am1 1 am2 2 am3 3 fm_tuner 4 fm_post 5 audio40 6 wb_pre 7 wb_dec1
 *     8 ssb_delay 9 ssb_hilbert. Returns the tap count. */
int sdro_q15_taps(int id, int16_t *q);
/* (float)atan2((double)q,(double)i): the value both discriminators use */
float sdro_atan2f(int q, int i);

/* host-thread channel bank (bench.py cpu_baseline kind "port"); same contract
 * as ref_bank_run in oracle/ref_shim.cc. Returns elapsed seconds. */
double sdro_bank_run(const uint8_t *modes, uint32_t n_channels, uint8_t *iq,
                     uint64_t bytes_per_channel, uint32_t block_bytes, int16_t *pcm,
                     uint32_t nthreads, int variant);

#ifdef __cplusplus
}
#endif
#endif
