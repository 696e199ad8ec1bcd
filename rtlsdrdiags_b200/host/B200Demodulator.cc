// One-channel engine facade shared by the four drop-in demodulator classes.
#include "B200Demodulator.h"

#include <math.h>
#include <string.h>

namespace
{
// the reference's member arrays cap a call at 32768 bytes (FmDemodulator.h:50);
// longer buffers are walked in pieces of this size
const uint32_t kMaxBlockBytes = 32768;
const uint32_t kBytesPerPcmSample = 64;
} // namespace

B200Demodulator::B200Demodulator(int kind, int mode, const char *name, PcmCallback pcmCallbackPtr)
{
  this->kind = kind;
  this->mode = mode;
  this->name = name;
  this->pcmCallbackPtr = pcmCallbackPtr;
  this->pendingFormat = SDR_IQ_S8_ROTATED;
  engine = NULL;
  status = sdr_engine_create(1, 0, kMaxBlockBytes, &engine);
  if (status != SDR_OK)
  {
    // There is no CPU path to fall back to: say so loudly, stay inert.
    fprintf(stderr, "%s: cannot create the B200 engine: %s\n", name, sdr_last_error(NULL));
    engine = NULL;
    demodulatorGain = 0;
    return;
  } // if

  sdr_set_mode(engine, 0, mode);

  // the same nominal gains the reference constructors set
  switch (kind)
  {
    case SDR_KIND_FM: demodulatorGain = 64000 / (2 * M_PI); break;
    case SDR_KIND_WBFM: demodulatorGain = 256000 / (2 * M_PI); break;
    default: demodulatorGain = 300; break;
  } // switch

  pcmData.resize(kMaxBlockBytes / kBytesPerPcmSample);
  staging.reserve(kMaxBlockBytes);
} // B200Demodulator

B200Demodulator::~B200Demodulator(void)
{
  if (engine != NULL)
  {
    sdr_engine_destroy(engine);
  } // if
} // ~B200Demodulator

void B200Demodulator::setMode(int mode)
{
  this->mode = mode;
  if (engine != NULL)
  {
    status = sdr_set_mode(engine, 0, mode);
  } // if
} // setMode

void B200Demodulator::discardPending(void)
{
  pending.clear();
} // discardPending

void B200Demodulator::resetDemodulator(void)
{
  // resetFilterState() also empties the decimators' input buffers
  discardPending();
  if (engine != NULL)
  {
    status = sdr_reset(engine, 0, kind);
  } // if
} // resetDemodulator

void B200Demodulator::setDemodulatorGain(float gain)
{
  demodulatorGain = gain;
  if (engine != NULL)
  {
    status = sdr_set_gain(engine, 0, kind, gain);
  } // if
} // setDemodulatorGain

void B200Demodulator::setResearchScaling(void)
{
  if (engine != NULL)
  {
    status = sdr_set_scaling(engine, SDR_SCALING_RESEARCH);
    if (kind == SDR_KIND_WBFM)
    {
      demodulatorGain = 64000 / (2 * M_PI);
    } // if
  } // if
} // setResearchScaling

void B200Demodulator::acceptIqData(int8_t *bufferPtr, uint32_t bufferLength)
{
  process((const uint8_t *)bufferPtr, bufferLength, SDR_IQ_S8_ROTATED);
} // acceptIqData

void B200Demodulator::acceptRawIqData(const uint8_t *bufferPtr, uint32_t bufferLength)
{
  process(bufferPtr, bufferLength, SDR_IQ_U8_OFFSET);
} // acceptRawIqData

void B200Demodulator::process(const uint8_t *bufferPtr, uint32_t bufferLength, uint32_t format)
{
  uint32_t produced = 0;

  if (engine == NULL)
  {
    status = SDR_E_CUDA;
    return;
  } // if

  if (!pending.empty() && pendingFormat != format)
  {
    // a caller that switches entry points mid-stream abandons the partial sample
    discardPending();
  } // if
  pendingFormat = format;

  if (pcmData.size() < (pending.size() + (size_t)bufferLength) / kBytesPerPcmSample + 1)
  {
    pcmData.resize((pending.size() + (size_t)bufferLength) / kBytesPerPcmSample + 1);
  } // if

  uint32_t offset = 0;
  while (offset < bufferLength || pending.size() >= kBytesPerPcmSample)
  {
    // top the carried bytes up to at most one reference block
    uint32_t room = kMaxBlockBytes - (uint32_t)pending.size();
    uint32_t take = bufferLength - offset < room ? bufferLength - offset : room;
    pending.insert(pending.end(), bufferPtr + offset, bufferPtr + offset + take);
    offset += take;

    uint32_t whole = (uint32_t)(pending.size() / kBytesPerPcmSample) * kBytesPerPcmSample;
    if (whole == 0)
    {
      break;
    } // if

    staging.assign(pending.begin(), pending.begin() + whole);
    staging.resize((whole + 15u) & ~15u);
    status = sdr_accept_iq(engine, staging.data(), whole, (whole + 15u) & ~15u, SDR_IQ_HOST | format);
    if (status == SDR_OK)
    {
      status = sdr_get_pcm(engine, &pcmData[produced], NULL);
    } // if
    if (status != SDR_OK)
    {
      fprintf(stderr, "%s: engine error %d: %s\n", name, status, sdr_last_error(engine));
      discardPending();
      return;
    } // if
    produced += whole / kBytesPerPcmSample;
    pending.erase(pending.begin(), pending.begin() + whole);
  } // while

  // Send the PCM data to the information sink: once per call, even when the
  // block was too short to complete a sample (the reference does the same).
  pcmCallbackPtr(pcmData.data(), produced);
} // process

void B200Demodulator::display(const char *extra)
{
  nprintf(stderr, "\n--------------------------------------------\n");
  nprintf(stderr, "%s Internal Information\n", name);
  nprintf(stderr, "--------------------------------------------\n");
  if (extra != NULL)
  {
    nprintf(stderr, "%s", extra);
  } // if
  nprintf(stderr, "Demodulator Gain         : %f\n", demodulatorGain);
  nprintf(stderr, "Engine                   : %s (status %d)\n", sdr_version(), status);
} // display
