// profile.cuh — optional per-KERNEL CUDA-event timing (bench.py's roofline): every launch site of the hot
// path sits in a KScope, which records an event pair on the launch stream when profiling is switched on
// (bliss_profile_enable) and costs one predictable branch when it is off.  Never enabled while a stream
// is being captured; timing with events between kernels also removes the programmatic-dependent-launch
// overlap, so per-kernel times are the kernels' own durations, not their share of a chained step.
#pragma once
#include <cuda_runtime.h>

namespace bliss {
extern bool g_profile_on;
void profile_begin(const char* name, cudaStream_t st);
void profile_end(cudaStream_t st);
struct KScope {
  cudaStream_t st;
  bool on;
  __host__ KScope(const char* name, cudaStream_t s) : st(s), on(g_profile_on) {
    if (on) profile_begin(name, st);
  }
  __host__ ~KScope() {
    if (on) profile_end(st);
  }
};
}  // namespace bliss
#define BLISS_KSCOPE(name, st) bliss::KScope kscope__(name, (cudaStream_t)(st))
// plain launch inside a profiling scope (needs BLISS_CHECK_LAUNCH of common.cuh)
#define BLISS_LAUNCH(kernel, grid, block, smem, st, ...)                             \
  do {                                                                               \
    BLISS_KSCOPE(#kernel, st);                                                       \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);                                  \
    BLISS_CHECK_LAUNCH();                                                            \
  } while (0)
