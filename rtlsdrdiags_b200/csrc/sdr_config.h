// What the engine is built from.
#pragma once
#include "sdr_device.cuh"
#include "sdr_tile.cuh"
#include "sdr_wbfm4.cuh"

namespace sdr {
enum { KIND_AM = 1, KIND_FM = 2, KIND_WBFM = 3, KIND_SSB = 4 };
}  // namespace sdr
