// Drop-in for radioDiags/SsbDemodulator/SsbDemodulator.h.
#ifndef __SSBDEMODULATOR__
#define __SSBDEMODULATOR__

#include "B200Demodulator.h"

class SsbDemodulator : public B200Demodulator
{
  public:
  SsbDemodulator(void (*pcmCallbackPtr)(int16_t *bufferPtr, uint32_t bufferLength))
      : B200Demodulator(SDR_KIND_SSB, SDR_MODE_LSB, "SSB Demodulator", pcmCallbackPtr),
        lsbDemodulationMode(true) {}
  ~SsbDemodulator(void) {}

  // SsbDemodulator.cc:333-369: the sideband is a property of this object
  void setLsbDemodulationMode(void) { lsbDemodulationMode = true; setMode(SDR_MODE_LSB); }
  void setUsbDemodulationMode(void) { lsbDemodulationMode = false; setMode(SDR_MODE_USB); }

  void displayInternalInformation(void)
  {
    display(lsbDemodulationMode ? "Demodulation Mode        : LSB\n" : "Demodulation Mode        : USB\n");
  }

  private:
  bool lsbDemodulationMode;
};

#endif // __SSBDEMODULATOR__
