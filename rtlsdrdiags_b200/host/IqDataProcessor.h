// Drop-in for radioDiags/hdr_diags/IqDataProcessor.h, reduced to the demodulation
// path: mode switch, demodulator registration and acceptIqData. The u8 -> s8
// conversion and the Fs/4 rotation (IqDataProcessor.cc:735-749) run on the GPU as
// the first phase of the demodulation kernel. Squelch, IQ dump and the signal
// callbacks are not rebuilt; the squelch default (-200 dBFS,
// IqDataProcessor.cc:41) is always open, which is what this class implements.
#ifndef _IQDATAPROCESSOR_H_
#define _IQDATAPROCESSOR_H_

#include <stdint.h>

#include "AmDemodulator.h"
#include "FmDemodulator.h"
#include "SsbDemodulator.h"
#include "WbFmDemodulator.h"

class IqDataProcessor
{
  public:
  enum demodulatorType {None = 0, Am = 1, Fm = 2, WbFm = 3, Lsb = 4, Usb = 5};

  IqDataProcessor(char *hostIpAddress, int hostPort);
  ~IqDataProcessor(void);

  void setDemodulatorMode(demodulatorType mode);
  void setAmDemodulator(AmDemodulator *demodulatorPtr);
  void setFmDemodulator(FmDemodulator *demodulatorPtr);
  void setWbFmDemodulator(WbFmDemodulator *demodulatorPtr);
  void setSsbDemodulator(SsbDemodulator *demodulatorPtr);

  void acceptIqData(unsigned long timeStamp, unsigned char *bufferPtr, unsigned long byteCount);

  void displayInternalInformation(void);

  private:
  demodulatorType demodulatorMode;
  AmDemodulator *amDemodulatorPtr;
  FmDemodulator *fmDemodulatorPtr;
  WbFmDemodulator *wbFmDemodulatorPtr;
  SsbDemodulator *ssbDemodulatorPtr;
};

#endif // _IQDATAPROCESSOR_H_
