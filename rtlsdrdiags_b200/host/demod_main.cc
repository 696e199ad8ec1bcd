// b200_demod -- offline file demodulator on the B200 engine, the counterpart of
// demodulatorResearch/demodulators/demod.cc (stdin IQ -> stdout S16_LE PCM at 8 kS/s).
//
//   b200_demod -d [1|2|3|4|5] [-u] [-r] [-l] [-b bytes] < in.iq > out.raw
//
//   -d   1 AM, 2 FM, 3 wideband FM, 4 LSB, 5 USB. As in demod.cc:233-241 the switch
//        falls through, so -d 4 selects USB as well; -l asks for the real LSB.
//   -u   input is raw u8 offset-binary from the dongle: route it through the
//        IqDataProcessor drop-in (offset removal + Fs/4 rotation) like radioDiags does.
//        Without -u the input is signed and already rotated (demod.cc:8-11).
//   -r   research-tree FM/WBFM scaling (what demod.cc itself links against).
//   -b   bytes per read (default 16384 like demod.cc:250; 32768 with -u).
//   -p   with -u: enableIqDump towards 127.0.0.1:<port> (IqDataProcessor::enableIqDump).
//   -s   squelch threshold in dBFS (with -u; IqDataProcessor::setSignalDetectThreshold),
//        and print every block's signal state and magnitude to stderr.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <unistd.h>

#include "IqDataProcessor.h"

int32_t radio_adjustableReceiveGainInDb = 0;  // Radio.cc owns this in the reference

static void onSignalState(bool signalPresent, void *contextPtr)
{
  fprintf(stderr, "signal %d", signalPresent ? 1 : 0);
  (void)contextPtr;
} // onSignalState

static void onSignalMagnitude(uint32_t signalMagnitude, void *contextPtr)
{
  fprintf(stderr, " magnitude %u\n", signalMagnitude);
  (void)contextPtr;
} // onSignalMagnitude

void nprintf(FILE *s, const char *formatPtr, ...)
{
  va_list ap;
  va_start(ap, formatPtr);
  vfprintf(s, formatPtr, ap);
  va_end(ap);
} // nprintf

static void processPcmData(int16_t *bufferPtr, uint32_t bufferLength)
{
  fwrite(bufferPtr, 2, bufferLength, stdout);
} // processPcmData

int main(int argc, char **argv)
{
  int demodulatorType = 2;
  bool rawInput = false, research = false, realLsb = false, squelch = false;
  long blockBytes = 0, threshold = -200, dumpPort = 0;
  int opt;

  while ((opt = getopt(argc, argv, "d:urlb:s:p:h")) != -1)
  {
    switch (opt)
    {
      case 'd': demodulatorType = atoi(optarg); break;
      case 'u': rawInput = true; break;
      case 'r': research = true; break;
      case 'l': realLsb = true; break;
      case 'b': blockBytes = atol(optarg); break;
      case 'p': dumpPort = atol(optarg); break;
      case 's': squelch = true; threshold = atol(optarg); break;
      default:
        fprintf(stderr, "./b200_demod -d [1 - AM | 2 - FM | 3 - WBFM | 4 - LSB | 5 - USB] [-u] [-r] [-l] [-b bytes]"
                        " < inputFile > outputFile\n");
        return 0;
    } // switch
  } // while
  if (demodulatorType < 1 || demodulatorType > 5)
  {
    fprintf(stderr, "invalid demodulator type\n");
    return 1;
  } // if
  if (blockBytes <= 0) blockBytes = rawInput ? 32768 : 16384;

  AmDemodulator *amDemodPtr = new AmDemodulator(processPcmData);
  FmDemodulator *fmDemodPtr = new FmDemodulator(processPcmData);
  WbFmDemodulator *wbFmDemodPtr = new WbFmDemodulator(processPcmData);
  SsbDemodulator *ssbDemodPtr = new SsbDemodulator(processPcmData);
  if (amDemodPtr->lastStatus() != SDR_OK) return 2;

  if (research)
  {
    // the research tree scales by the gain alone and starts WBFM at 64000/2pi
    fmDemodPtr->setResearchScaling();
    wbFmDemodPtr->setResearchScaling();
  } // if

  static char host[] = "127.0.0.1";
  IqDataProcessor *processorPtr = new IqDataProcessor(host, dumpPort ? (int)dumpPort : 8001);
  if (dumpPort)
  {
    processorPtr->enableIqDump();
  } // if
  processorPtr->setAmDemodulator(amDemodPtr);
  processorPtr->setFmDemodulator(fmDemodPtr);
  processorPtr->setWbFmDemodulator(wbFmDemodPtr);
  processorPtr->setSsbDemodulator(ssbDemodPtr);
  if (squelch)
  {
    processorPtr->setSignalDetectThreshold((int32_t)threshold);
    processorPtr->registerSignalStateCallback(onSignalState, NULL);
    processorPtr->enableSignalNotification();
    processorPtr->registerSignalMagnitudeCallback(onSignalMagnitude, NULL);
    processorPtr->enableSignalMagnitudeNotification();
  } // if

  switch (demodulatorType)
  {
    case 4:
      ssbDemodPtr->setLsbDemodulationMode();
      if (realLsb) break;
      // fall through, as demod.cc does
    case 5:
      ssbDemodPtr->setUsbDemodulationMode();
      break;
  } // switch
  IqDataProcessor::demodulatorType mode = (IqDataProcessor::demodulatorType)demodulatorType;
  if (demodulatorType == 4 && !realLsb) mode = IqDataProcessor::Usb;
  if (rawInput) processorPtr->setDemodulatorMode(mode);

  uint8_t *inputBuffer = (uint8_t *)malloc(blockBytes);
  for (;;)
  {
    size_t count = fread(inputBuffer, 1, blockBytes, stdin);
    if (count == 0) break;
    if (rawInput)
    {
      processorPtr->acceptIqData(0, inputBuffer, count);
      continue;
    } // if
    switch (demodulatorType)
    {
      case 1: amDemodPtr->acceptIqData((int8_t *)inputBuffer, count); break;
      case 2: fmDemodPtr->acceptIqData((int8_t *)inputBuffer, count); break;
      case 3: wbFmDemodPtr->acceptIqData((int8_t *)inputBuffer, count); break;
      default: ssbDemodPtr->acceptIqData((int8_t *)inputBuffer, count); break;
    } // switch
  } // for
  fflush(stdout);

  free(inputBuffer);
  delete processorPtr;
  delete amDemodPtr;
  delete fmDemodPtr;
  delete wbFmDemodPtr;
  delete ssbDemodPtr;
  return 0;
} // main
