// Mode switch in front of the four drop-in demodulators.
#include "IqDataProcessor.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <vector>

// Owned by the application, as in the reference (IqDataProcessor.cc:8, Radio.cc).
extern int32_t radio_adjustableReceiveGainInDb;

IqDataProcessor::IqDataProcessor(char *hostIpAddress, int hostPort)
{
  // The dump's link partner (UdpClient.cc:53-96): one UDP socket, 32768-byte send buffer.
  iqDumpEnabled = false; // IqDataProcessor.cc:62
  dumpSocket = socket(PF_INET, SOCK_DGRAM, 0);
  if (dumpSocket != -1)
  {
    struct sockaddr_in peer;
    int bufferLength = 32768;
    static_assert(sizeof(peer) <= sizeof(dumpPeer), "peer address storage");
    memset(&peer, 0, sizeof(peer));
    peer.sin_family = AF_INET;
    peer.sin_addr.s_addr = inet_addr(hostIpAddress);
    peer.sin_port = htons(hostPort);
    memcpy(dumpPeer, &peer, sizeof(peer));
    if (setsockopt(dumpSocket, SOL_SOCKET, SO_SNDBUF, &bufferLength, sizeof(int)) == -1)
    {
      close(dumpSocket);
      dumpSocket = 0;
    } // if
  } // if
  else
  {
    dumpSocket = 0;
  } // else

  // Default to no demodulation of the signal (IqDataProcessor.cc:38).
  demodulatorMode = None;

  // Let all signal exceed threshold (IqDataProcessor.cc:41).
  signalDetectThreshold = -200;
  gateEngine = NULL;
  gateThreshold = -200;
  gateReceiveGain = 0;
  gateDump = 0;
  blocksSeen = 0;
  signalNotificationEnabled = false;
  signalCallbackPtr = NULL;
  signalCallbackContextPtr = NULL;
  signalMagnitudeNotificationEnabled = false;
  signalMagnitudeCallbackPtr = NULL;
  signalMagnitudeCallbackContextPtr = NULL;
  amDemodulatorPtr = NULL;
  fmDemodulatorPtr = NULL;
  wbFmDemodulatorPtr = NULL;
  ssbDemodulatorPtr = NULL;
} // IqDataProcessor

IqDataProcessor::~IqDataProcessor(void)
{
  if (gateEngine != NULL)
  {
    sdr_engine_destroy(gateEngine);
  } // if
  if (dumpSocket != 0)
  {
    close(dumpSocket);
  } // if
} // ~IqDataProcessor

void IqDataProcessor::enableIqDump(void) { iqDumpEnabled = true; }
void IqDataProcessor::disableIqDump(void) { iqDumpEnabled = false; }
bool IqDataProcessor::isIqDumpEnabled(void) { return (iqDumpEnabled); }

// networkInterfacePtr->sendData(signedBufferPtr, byteCount) (IqDataProcessor.cc:759): the
// engine's dump of this block, in UdpClient::sendData's slices (UdpClient.cc:199-231).
void IqDataProcessor::sendIqDump(unsigned long byteCount)
{
  static thread_local std::vector<int8_t> dump;
  uint64_t n = 0;
  dump.resize(byteCount);
  if (sdr_get_iq_dump(gateEngine, 0, dump.data(), byteCount, &n) != SDR_OK)
  {
    fprintf(stderr, "IqDataProcessor: IQ dump failed: %s\n", sdr_last_error(gateEngine));
    return;
  } // if
  if (dumpSocket == 0)
  {
    return; // UdpClient only sends to open sockets
  } // if
  const uint64_t maxPayloadLength = 2048;
  for (uint64_t off = 0; off < n; off += maxPayloadLength)
  {
    uint64_t len = n - off < maxPayloadLength ? n - off : maxPayloadLength;
    sendto(dumpSocket, dump.data() + off, len, 0, (struct sockaddr *)dumpPeer, sizeof(struct sockaddr));
  } // for
} // sendIqDump

void IqDataProcessor::setSignalDetectThreshold(int32_t threshold)
{
  signalDetectThreshold = threshold;
} // setSignalDetectThreshold

void IqDataProcessor::enableSignalNotification(void) { signalNotificationEnabled = true; }
void IqDataProcessor::disableSignalNotification(void) { signalNotificationEnabled = false; }
void IqDataProcessor::registerSignalStateCallback(void (*callbackPtr)(bool signalPresent, void *contextPtr),
                                                  void *contextPtr)
{
  signalCallbackContextPtr = contextPtr;
  signalCallbackPtr = callbackPtr;
} // registerSignalStateCallback
void IqDataProcessor::enableSignalMagnitudeNotification(void) { signalMagnitudeNotificationEnabled = true; }
void IqDataProcessor::disableSignalMagnitudeNotification(void) { signalMagnitudeNotificationEnabled = false; }
void IqDataProcessor::registerSignalMagnitudeCallback(void (*callbackPtr)(uint32_t signalMagnitude, void *contextPtr),
                                                      void *contextPtr)
{
  signalMagnitudeCallbackContextPtr = contextPtr;
  signalMagnitudeCallbackPtr = callbackPtr;
} // registerSignalMagnitudeCallback

// Squelch::run on the block (IqDataProcessor.cc:764-790): returns signalAllowed and fires the
// signal-state and signal-magnitude callbacks, in the reference's order, before any
// demodulation.
bool IqDataProcessor::runSquelch(unsigned char *bufferPtr, unsigned long byteCount)
{
  const bool wanted = iqDumpEnabled || signalDetectThreshold != -200 || radio_adjustableReceiveGainInDb != 0 ||
                      (signalNotificationEnabled && signalCallbackPtr != NULL) ||
                      (signalMagnitudeNotificationEnabled && signalMagnitudeCallbackPtr != NULL);
  if (!wanted && gateEngine == NULL)
  {
    blocksSeen++;
    return (true); // the default threshold never closes and nobody listens
  } // if
  if ((byteCount % 64) != 0 || byteCount == 0)
  {
    blocksSeen++;
    return (true); // the engine works on whole PCM samples; such a block passes ungated
  } // if
  if (gateEngine == NULL)
  {
    if (sdr_engine_create(1, 0, 1u << 20, &gateEngine) != SDR_OK)
    {
      fprintf(stderr, "IqDataProcessor: cannot create the B200 engine: %s\n", sdr_last_error(NULL));
      gateEngine = NULL;
      return (true);
    } // if
    sdr_enable_signal_reports(gateEngine, 1);
    if (blocksSeen != 0)
    {
      // Every block so far was above the -200 dBFS default, so the reference's tracker is
      // in its Tracking state and grants this block the one-block tail: replay that.
      static const uint8_t primer[64 + 16] = {0};
      const uint8_t *p16 = (const uint8_t *)(((uintptr_t)primer + 15) & ~(uintptr_t)15);
      sdr_accept_iq(gateEngine, p16, 64, 64, SDR_IQ_HOST | SDR_IQ_U8_OFFSET);
      sdr_sync(gateEngine);
    } // if
  } // if
  if (byteCount > (1u << 20))
  {
    return (true);
  } // if

  uint8_t allowed = 1;
  uint32_t magnitude = 0;
  // each setter makes the engine re-upload its squelch tables: only on a change
  if (signalDetectThreshold != gateThreshold)
  {
    gateThreshold = signalDetectThreshold;
    sdr_set_squelch_threshold(gateEngine, 0, gateThreshold);
  } // if
  if (radio_adjustableReceiveGainInDb != gateReceiveGain)
  {
    gateReceiveGain = radio_adjustableReceiveGainInDb;
    sdr_set_receive_gain_db(gateEngine, 0, (uint32_t)gateReceiveGain);
  } // if
  if ((iqDumpEnabled ? 1 : 0) != gateDump)
  {
    gateDump = iqDumpEnabled ? 1 : 0;
    sdr_set_iq_dump(gateEngine, 0, gateDump);
  } // if
  // the block may sit at any address: stage it 16-byte aligned
  static thread_local std::vector<uint8_t> staging;
  staging.resize(byteCount + 16);
  uint8_t *aligned = (uint8_t *)(((uintptr_t)staging.data() + 15) & ~(uintptr_t)15);
  for (unsigned long i = 0; i < byteCount; i++) aligned[i] = bufferPtr[i];
  const bool dumping = iqDumpEnabled;
  int rc = sdr_accept_iq(gateEngine, aligned, byteCount, byteCount, SDR_IQ_HOST | SDR_IQ_U8_OFFSET);
  if (rc == SDR_OK && dumping)
  {
    sendIqDump(byteCount); // before the squelch speaks, so a display stays live (IqDataProcessor.cc:753)
  } // if
  if (rc != SDR_OK || sdr_get_signal(gateEngine, &allowed, &magnitude) != SDR_OK)
  {
    fprintf(stderr, "IqDataProcessor: squelch failed: %s\n", sdr_last_error(gateEngine));
    return (true);
  } // if

  if ((signalNotificationEnabled) && (signalCallbackPtr != NULL))
  {
    signalCallbackPtr(allowed != 0, signalCallbackContextPtr);
  } // if
  if ((signalMagnitudeNotificationEnabled) && (signalMagnitudeCallbackPtr != NULL))
  {
    signalMagnitudeCallbackPtr(magnitude, signalMagnitudeCallbackContextPtr);
  } // if

  return (allowed != 0);
} // runSquelch

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

// Multiplication of z(n) = x(n) + j y(n) by exp(-j n pi / 2) = {1, -j, -1, j, ...}: within every
// group of four complex samples, sample 1 becomes (y, -x), sample 2 (-x, -y), sample 3 (-y, x)
// (Lyons, Understanding DSP, 13.1.2; IqDataProcessor.cc:500-545). int8 negation wraps: -(-128)
// stays -128. The phase restarts at the head of every call.
void IqDataProcessor::downconvertByFsOver4(int8_t *bufferPtr, uint32_t byteCount)
{
  for (uint32_t g = 0; g + 8 <= byteCount; g += 8)
  {
    int8_t *z = bufferPtr + g;
    const int8_t x1 = z[2], y1 = z[3], x2 = z[4], y2 = z[5], x3 = z[6], y3 = z[7];
    z[2] = y1;
    z[3] = (int8_t)(-x1);
    z[4] = (int8_t)(-x2);
    z[5] = (int8_t)(-y2);
    z[6] = (int8_t)(-y3);
    z[7] = x3;
  } // for
} // downconvertByFsOver4

// The same with exp(+j n pi / 2) = {1, j, -1, -j, ...}: sample 1 becomes (-y, x), sample 2
// (-x, -y), sample 3 (y, -x) (IqDataProcessor.cc:567-611).
void IqDataProcessor::upconvertByFsOver4(int8_t *bufferPtr, uint32_t byteCount)
{
  for (uint32_t g = 0; g + 8 <= byteCount; g += 8)
  {
    int8_t *z = bufferPtr + g;
    const int8_t x1 = z[2], y1 = z[3], x2 = z[4], y2 = z[5], x3 = z[6], y3 = z[7];
    z[2] = (int8_t)(-y1);
    z[3] = x1;
    z[4] = (int8_t)(-x2);
    z[5] = (int8_t)(-y2);
    z[6] = y3;
    z[7] = (int8_t)(-x3);
  } // for
} // upconvertByFsOver4

// The reference leaves the caller's buffer converted in place -- signed and Fs/4-translated
// (IqDataProcessor.cc:735-749) -- and a caller that looks at it afterwards (or reuses it) sees
// that. The GPU works on its own copy, so the host buffer is converted here, after the engine
// took the raw bytes. (WbFmDemodulator additionally overwrites it with its pre-filter output,
// WbFmDemodulator.cc:389-398; that is not reproduced.)
static void convertCallerBuffer(IqDataProcessor *p, unsigned char *bufferPtr, unsigned long byteCount)
{
  int8_t *signedBufferPtr = (int8_t *)bufferPtr;
  for (unsigned long i = 0; i < byteCount; i++)
  {
    signedBufferPtr[i] = (int8_t)(bufferPtr[i] - 128);
  } // for
  p->upconvertByFsOver4(signedBufferPtr, (uint32_t)byteCount);
} // convertCallerBuffer

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

  if (!runSquelch(bufferPtr, byteCount))
  {
    convertCallerBuffer(this, bufferPtr, byteCount);
    return; // squelched: no demodulator is called (IqDataProcessor.cc:793)
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
  convertCallerBuffer(this, bufferPtr, byteCount);
} // acceptIqData

void IqDataProcessor::displayInternalInformation(void)
{
  static const char *names[] = {"None", "AM", "FM", "WBFM", "LSB", "USB"};
  nprintf(stderr, "\n--------------------------------------------\n");
  nprintf(stderr, "IQ Data Processor Internal Information\n");
  nprintf(stderr, "--------------------------------------------\n");
  nprintf(stderr, "Demodulator Mode         : %s\n", names[demodulatorMode]);
} // displayInternalInformation
