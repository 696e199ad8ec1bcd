// Shared machinery of the drop-in demodulator classes: a one-channel engine
// behind the reference's per-object surface.
//
// The reference's AmDemodulator / FmDemodulator / WbFmDemodulator / SsbDemodulator
// (radioDiags/*Demodulator/*.h) each own their filters and hand PCM to a callback
// from inside acceptIqData. Here each object owns a 1-channel sdr_engine fixed to
// its mode; acceptIqData ships the block to the GPU, waits for the PCM and fires
// the callback on the calling thread, once per call, exactly as the reference does
// (FmDemodulator.cc:334-352, 585). A bank of radios should use the C ABI directly
// (include/sdr_b200.h); this facade exists so existing callers compile unchanged.
#ifndef B200_DEMODULATOR_H
#define B200_DEMODULATOR_H

#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/sdr_b200.h"

class IqDataProcessor;

class B200Demodulator
{
  public:
  typedef void (*PcmCallback)(int16_t *bufferPtr, uint32_t bufferLength);

  void resetDemodulator(void);
  void setDemodulatorGain(float gain);
  // bufferLength is a byte count (I,Q interleaved, signed, already Fs/4-rotated)
  void acceptIqData(int8_t *bufferPtr, uint32_t bufferLength);
  // FM / WBFM only: scale like the demodulatorResearch/ tree (gain used directly,
  // WBFM nominal gain 64000/2pi) instead of radioDiags/ (gain/deviation*32767)
  void setResearchScaling(void);
  // status of the last engine call (SDR_OK or SDR_E_*); the reference's methods are void
  int lastStatus(void) const { return status; }

  protected:
  B200Demodulator(int kind, int mode, const char *name, PcmCallback pcmCallbackPtr);
  ~B200Demodulator(void);
  void setMode(int mode);
  void display(const char *extra);

  private:
  friend class IqDataProcessor;
  // IqDataProcessor hands over the raw u8 block; offset removal and the Fs/4
  // rotation then happen on the GPU (IqDataProcessor.cc:735-749)
  void acceptRawIqData(const uint8_t *bufferPtr, uint32_t bufferLength);
  void process(const uint8_t *bufferPtr, uint32_t bufferLength, uint32_t format);
  void discardPending(void);

  sdr_engine *engine;
  int kind, mode, status;
  const char *name;
  float demodulatorGain;
  PcmCallback pcmCallbackPtr;
  // bytes that do not yet make a whole PCM sample (64 bytes = 32 complex samples):
  // the reference holds the same samples in its decimators' input buffers
  // (Decimator_int16.cc:320-325)
  std::vector<uint8_t> pending;
  uint32_t pendingFormat;
  std::vector<uint8_t> staging;
  std::vector<int16_t> pcmData;
};

extern void nprintf(FILE *s, const char *formatPtr, ...);

#endif // B200_DEMODULATOR_H
