// Fast path of the harmonic kernel (hop % 64 == 0).  Placeholder until built.
#pragma once
#include "harmonic.cuh"
namespace ddsp {
inline bool harmonic_fast_supported(const HarmonicParams&) { return false; }
inline int launch_harmonic_fast(const HarmonicParams&, cudaStream_t) { return 1; }
}  // namespace ddsp
