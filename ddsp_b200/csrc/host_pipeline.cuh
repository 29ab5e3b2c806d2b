// Host-buffer decoder pipeline: the ae.gin decoder for callers whose network
// outputs and audio live in HOST memory (the reference's Processor call takes
// numpy arrays, processors_test.py:35-42).  A batch is cut into chunks of items;
// chunk c's host->device copies (copy stream), its two decoder kernels (the
// caller's stream) and its device->host audio copy (second copy stream) overlap
// with the neighbouring chunks', so PCIe runs in both directions while the SMs
// work: the step costs about max(H2D, compute, D2H) instead of their sum.
//
// The handle owns what such a pipeline needs beyond the caller's buffers: one
// device staging allocation, three copy streams and the events.  It is the one place
// where the library allocates; everything is released by *_destroy.
#pragma once
#include <vector>

#include "common.cuh"

namespace ddsp {

struct HostPipeline {
  int device = -1;
  int max_B = 0, F = 0, K = 0, nb = 0, N = 0, max_chunks = 0;
  float* d_base = nullptr;
  float *d_amps = nullptr, *d_hd = nullptr, *d_f0 = nullptr, *d_mags = nullptr,
        *d_audio = nullptr;
  cudaStream_t s_h2d = nullptr, s_h2d2 = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_start = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_h2d, ev_h2d2, ev_comp;
  bool used = false;
};

inline void host_pipeline_free(HostPipeline* hp) {
  if (!hp) return;
  if (hp->device >= 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(hp->device);
    if (hp->s_h2d) cudaStreamSynchronize(hp->s_h2d);
    if (hp->s_h2d2) cudaStreamSynchronize(hp->s_h2d2);
    if (hp->s_d2h) cudaStreamSynchronize(hp->s_d2h);
    for (auto e : hp->ev_h2d) cudaEventDestroy(e);
    for (auto e : hp->ev_h2d2) cudaEventDestroy(e);
    for (auto e : hp->ev_comp) cudaEventDestroy(e);
    if (hp->ev_start) cudaEventDestroy(hp->ev_start);
    if (hp->ev_done) cudaEventDestroy(hp->ev_done);
    if (hp->s_h2d) cudaStreamDestroy(hp->s_h2d);
    if (hp->s_h2d2) cudaStreamDestroy(hp->s_h2d2);
    if (hp->s_d2h) cudaStreamDestroy(hp->s_d2h);
    if (hp->d_base) cudaFree(hp->d_base);
    cudaSetDevice(cur);
  }
  delete hp;
}

}  // namespace ddsp
