// profile.cu — registry behind profile.cuh: event pairs per launch, aggregated per kernel name on read.
#include <string.h>
#include <string>
#include <vector>
#include "../../include/bliss_b200.h"
#include "profile.cuh"

namespace bliss {
bool g_profile_on = false;
namespace {
struct Rec {
  const char* name;
  cudaEvent_t a, b;
};
std::vector<Rec> g_recs;          // launches recorded since the last enable / read
std::vector<cudaEvent_t> g_pool;  // recycled events
size_t g_open = (size_t)-1;
const size_t kMaxRecs = 1 << 18;

cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void profile_begin(const char* name, cudaStream_t st) {
  if (g_recs.size() >= kMaxRecs) {
    g_open = (size_t)-1;
    return;
  }
  Rec r{name, take_event(), take_event()};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  g_open = g_recs.size() - 1;
}
void profile_end(cudaStream_t st) {
  if (g_open == (size_t)-1) return;
  cudaEventRecord(g_recs[g_open].b, st);
  g_open = (size_t)-1;
}
}  // namespace bliss

extern "C" {

int bliss_profile_enable(int32_t on) {
  for (auto& r : bliss::g_recs) {
    bliss::g_pool.push_back(r.a);
    bliss::g_pool.push_back(r.b);
  }
  bliss::g_recs.clear();
  bliss::g_open = (size_t)-1;
  bliss::g_profile_on = on != 0;
  return 0;
}

// Synchronises the recorded events and writes one entry per distinct kernel name:
// names = '\n'-separated list, ms[i] = total milliseconds, calls[i] = launches.  Returns the number of
// entries (<= cap), or <0 when a buffer is too small.
int bliss_profile_read(char* names, int32_t names_cap, float* ms, int32_t* calls, int32_t cap) {
  if (!names || !ms || !calls || cap <= 0 || names_cap <= 0) return -1;
  std::vector<const char*> keys;
  std::vector<double> tot;
  std::vector<int> cnt;
  for (auto& r : bliss::g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    size_t k = 0;
    for (; k < keys.size(); ++k)
      if (keys[k] == r.name || strcmp(keys[k], r.name) == 0) break;
    if (k == keys.size()) {
      keys.push_back(r.name);
      tot.push_back(0.0);
      cnt.push_back(0);
    }
    tot[k] += t;
    cnt[k] += 1;
  }
  (void)cudaGetLastError();
  if ((int)keys.size() > cap) return -2;
  std::string all;
  for (size_t k = 0; k < keys.size(); ++k) {
    if (k) all += '\n';
    all += keys[k];
    ms[k] = (float)tot[k];
    calls[k] = cnt[k];
  }
  if ((int)all.size() + 1 > names_cap) return -3;
  memcpy(names, all.c_str(), all.size() + 1);
  return (int)keys.size();
}

}  // extern "C"
