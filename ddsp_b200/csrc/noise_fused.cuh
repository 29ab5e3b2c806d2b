// Fused filtered-noise kernel (IR synthesis + Philox + FIR + Add in one pass).
// Placeholder until built.
#pragma once
#include "noise.cuh"
namespace ddsp {
inline bool noise_fused_supported(int, int, int, int) { return false; }
inline int launch_noise_fused(const float*, const float*, uint64_t, uint64_t,
                              float*, int, int, int, int, int, int,
                              cudaStream_t) {
  return DDSP_B200_E_UNSUPPORTED;
}
}  // namespace ddsp
