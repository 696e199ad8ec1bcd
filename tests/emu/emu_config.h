// Shapes of the phase-structured test kernels: S = complex samples per sub-block.
#pragma once
#include "sdr_phase.cuh"

#ifndef SDR_S_AM
#define SDR_S_AM 1024
#endif
#ifndef SDR_S_SSB
#define SDR_S_SSB 1024
#endif
#ifndef SDR_S_FM
#define SDR_S_FM 1024
#endif
#ifndef SDR_S_WBFM
#define SDR_S_WBFM 512
#endif

namespace sdr {
using AmPipe = AmSsb<SDR_S_AM, false>;
using SsbPipe = AmSsb<SDR_S_SSB, true>;
using FmPipe = Fm<SDR_S_FM>;
using WbFmPipe = WbFm<SDR_S_WBFM>;
enum { KIND_AM = 1, KIND_FM = 2, KIND_WBFM = 3, KIND_SSB = 4 };
}  // namespace sdr
