// Drop-in for radioDiags/AmDemodulator/AmDemodulator.h: same class name, constructor
// and public methods; the signal processing runs on the B200 engine.
#ifndef __AMDEMODULATOR__
#define __AMDEMODULATOR__

#include "B200Demodulator.h"

class AmDemodulator : public B200Demodulator
{
  public:
  AmDemodulator(void (*pcmCallbackPtr)(int16_t *bufferPtr, uint32_t bufferLength))
      : B200Demodulator(SDR_KIND_AM, SDR_MODE_AM, "AM Demodulator", pcmCallbackPtr) {}
  ~AmDemodulator(void) {}

  // inherited: resetDemodulator, setDemodulatorGain, acceptIqData
  void displayInternalInformation(void) { display(NULL); }
};

#endif // __AMDEMODULATOR__
