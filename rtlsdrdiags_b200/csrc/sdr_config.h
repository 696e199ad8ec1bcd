// Compile-time shape of the pipelines: S = complex samples per sub-block.
// Larger S amortises barriers and history moves, smaller S lets more channels
// (or more CTAs) share an SM's shared memory. Overridable with -D for sweeps.
#pragma once
#include "sdr_device.cuh"
#include "sdr_tile.cuh"

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
