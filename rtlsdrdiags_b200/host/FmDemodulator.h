// Drop-in for radioDiags/FmDemodulator/FmDemodulator.h: same class name, constructor
// and public methods; the signal processing runs on the B200 engine.
#ifndef __FMDEMODULATOR__
#define __FMDEMODULATOR__

#include "B200Demodulator.h"

class FmDemodulator : public B200Demodulator
{
  public:
  FmDemodulator(void (*pcmCallbackPtr)(int16_t *bufferPtr, uint32_t bufferLength))
      : B200Demodulator(SDR_KIND_FM, SDR_MODE_FM, "FM Demodulator", pcmCallbackPtr) {}
  ~FmDemodulator(void) {}

  // inherited: resetDemodulator, setDemodulatorGain, acceptIqData
  void displayInternalInformation(void) { display(NULL); }
};

#endif // __FMDEMODULATOR__
