// Mode switch in front of the four drop-in demodulators.
#include "IqDataProcessor.h"

#include <stdio.h>

IqDataProcessor::IqDataProcessor(char *hostIpAddress, int hostPort)
{
  (void)hostIpAddress; // the IQ-dump UDP client is not rebuilt
  (void)hostPort;

  // Default to no demodulation of the signal (IqDataProcessor.cc:38).
  demodulatorMode = None;
  amDemodulatorPtr = NULL;
  fmDemodulatorPtr = NULL;
  wbFmDemodulatorPtr = NULL;
  ssbDemodulatorPtr = NULL;
} // IqDataProcessor

IqDataProcessor::~IqDataProcessor(void)
{
} // ~IqDataProcessor

void IqDataProcessor::setAmDemodulator(AmDemodulator *demodulatorPtr) { amDemodulatorPtr = demodulatorPtr; }
void IqDataProcessor::setFmDemodulator(FmDemodulator *demodulatorPtr) { fmDemodulatorPtr = demodulatorPtr; }
void IqDataProcessor::setWbFmDemodulator(WbFmDemodulator *demodulatorPtr) { wbFmDemodulatorPtr = demodulatorPtr; }
void IqDataProcessor::setSsbDemodulator(SsbDemodulator *demodulatorPtr) { ssbDemodulatorPtr = demodulatorPtr; }

void IqDataProcessor::setDemodulatorMode(demodulatorType mode)
{
  demodulatorMode = mode;

  // IqDataProcessor.cc:247-258: the sideband lives in the SSB demodulator
  if (ssbDemodulatorPtr != NULL)
  {
    if (mode == Lsb) ssbDemodulatorPtr->setLsbDemodulationMode();
    if (mode == Usb) ssbDemodulatorPtr->setUsbDemodulationMode();
  } // if
} // setDemodulatorMode

void IqDataProcessor::acceptIqData(unsigned long timeStamp, unsigned char *bufferPtr, unsigned long byteCount)
{
  (void)timeStamp;

  // The rotation restarts at the head of every call and works on groups of four
  // complex samples (IqDataProcessor.cc:584-609); a block that is not a whole
  // number of groups has no defined continuation.
  if ((byteCount % 8) != 0)
  {
    fprintf(stderr, "IqDataProcessor: byteCount %lu is not a multiple of 8, block dropped\n", byteCount);
    return;
  } // if

  switch (demodulatorMode)
  {
    case Am:
      if (amDemodulatorPtr != NULL) amDemodulatorPtr->acceptRawIqData(bufferPtr, (uint32_t)byteCount);
      break;
    case Fm:
      if (fmDemodulatorPtr != NULL) fmDemodulatorPtr->acceptRawIqData(bufferPtr, (uint32_t)byteCount);
      break;
    case WbFm:
      if (wbFmDemodulatorPtr != NULL) wbFmDemodulatorPtr->acceptRawIqData(bufferPtr, (uint32_t)byteCount);
      break;
    case Lsb:
    case Usb:
      if (ssbDemodulatorPtr != NULL) ssbDemodulatorPtr->acceptRawIqData(bufferPtr, (uint32_t)byteCount);
      break;
    default:
      break;
  } // switch
} // acceptIqData

void IqDataProcessor::displayInternalInformation(void)
{
  static const char *names[] = {"None", "AM", "FM", "WBFM", "LSB", "USB"};
  nprintf(stderr, "\n--------------------------------------------\n");
  nprintf(stderr, "IQ Data Processor Internal Information\n");
  nprintf(stderr, "--------------------------------------------\n");
  nprintf(stderr, "Demodulator Mode         : %s\n", names[demodulatorMode]);
} // displayInternalInformation
