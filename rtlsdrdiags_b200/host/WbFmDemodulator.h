// Drop-in for radioDiags/WbFmDemodulator/WbFmDemodulator.h: same class name, constructor
// and public methods; the signal processing runs on the B200 engine.
#ifndef __WBFMDEMODULATOR__
#define __WBFMDEMODULATOR__

#include "B200Demodulator.h"

class WbFmDemodulator : public B200Demodulator
{
  public:
  WbFmDemodulator(void (*pcmCallbackPtr)(int16_t *bufferPtr, uint32_t bufferLength))
      : B200Demodulator(SDR_KIND_WBFM, SDR_MODE_WBFM, "Wideband FM Demodulator", pcmCallbackPtr) {}
  ~WbFmDemodulator(void) {}

  // inherited: resetDemodulator, setDemodulatorGain, acceptIqData
  void displayInternalInformation(void) { display(NULL); }
};

#endif // __WBFMDEMODULATOR__
