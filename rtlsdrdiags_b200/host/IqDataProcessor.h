// Drop-in for radioDiags/hdr_diags/IqDataProcessor.h, reduced to the demodulation
// path: mode switch, demodulator registration and acceptIqData. The u8 -> s8
// conversion and the Fs/4 rotation (IqDataProcessor.cc:735-749) run on the GPU as
// the first step of the demodulation kernels. The squelch gate and the signal-state /
// signal-magnitude callbacks (IqDataProcessor.cc:764-790) run on a one-channel engine of
// their own, because the reference keeps one Squelch per IqDataProcessor, shared by all
// modes. The same engine emits the IQ dump (enableIqDump, IqDataProcessor.cc:756-760), which
// leaves for hostIpAddress:hostPort in datagrams of at most 2048 bytes like UdpClient's.
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

  void setSignalDetectThreshold(int32_t threshold);

  // Frequency translation by -Fs/4 / +Fs/4 of signed interleaved IQ, in place on the host
  // (hdr_diags/IqDataProcessor.h:32-33). acceptIqData does the +Fs/4 translation on the GPU;
  // these serve callers that use the reference's public helpers directly.
  void downconvertByFsOver4(int8_t *bufferPtr, uint32_t byteCount);
  void upconvertByFsOver4(int8_t *bufferPtr, uint32_t byteCount);

  void enableIqDump(void);
  void disableIqDump(void);
  bool isIqDumpEnabled(void);

  void acceptIqData(unsigned long timeStamp, unsigned char *bufferPtr, unsigned long byteCount);

  void enableSignalNotification(void);
  void disableSignalNotification(void);
  void registerSignalStateCallback(void (*signalCallbackPtr)(bool signalPresent, void *contextPtr), void *contextPtr);
  void enableSignalMagnitudeNotification(void);
  void disableSignalMagnitudeNotification(void);
  void registerSignalMagnitudeCallback(void (*callbackPtr)(uint32_t signalMagnitude, void *contextPtr),
                                       void *contextPtr);

  void displayInternalInformation(void);

  private:
  bool runSquelch(unsigned char *bufferPtr, unsigned long byteCount);
  void sendIqDump(unsigned long byteCount);

  bool iqDumpEnabled;
  int dumpSocket;           // 0 = not open, as UdpClient keeps it (UdpClient.cc:77-96)
  unsigned char dumpPeer[16]; // struct sockaddr_in of the link partner

  demodulatorType demodulatorMode;
  int32_t signalDetectThreshold;
  sdr_engine *gateEngine;  // mode None: runs only the squelch kernel
  // what the gate engine was last told, so the setters run only on a change
  int32_t gateThreshold;
  int32_t gateReceiveGain;
  int gateDump;
  unsigned long blocksSeen; // blocks that passed while no gate engine existed yet
  bool signalNotificationEnabled;
  void *signalCallbackContextPtr;
  void (*signalCallbackPtr)(bool signalPresent, void *contextPtr);
  bool signalMagnitudeNotificationEnabled;
  void *signalMagnitudeCallbackContextPtr;
  void (*signalMagnitudeCallbackPtr)(uint32_t signalMagnitude, void *contextPtr);
  AmDemodulator *amDemodulatorPtr;
  FmDemodulator *fmDemodulatorPtr;
  WbFmDemodulator *wbFmDemodulatorPtr;
  SsbDemodulator *ssbDemodulatorPtr;
};

#endif // _IQDATAPROCESSOR_H_
