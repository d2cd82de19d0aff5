// Library-wide state: last error string, launch counter, cached device properties.
#include "common.cuh"
#include "../../include/b200_distill.h"

#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace b200 {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      sms = 148;
    }
  }
  return sms;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// ---- optional per-launch timing of the dense kernels (bench.py's roofline leg): CUDA events on the launch stream
struct ProfRec { cudaEvent_t a, b; double flops; int cat; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_prof_mu;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

int prof_begin(cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> g(g_prof_mu);
  ProfRec r;
  r.a = take_event(); r.b = take_event(); r.flops = 0; r.cat = 0;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  return (int)g_recs.size() - 1;
}

void prof_end(int idx, cudaStream_t st, double flops, int cat) {
  if (idx < 0) return;
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (idx >= (int)g_recs.size()) return;
  g_recs[idx].flops = flops;
  g_recs[idx].cat = cat;
  cudaEventRecord(g_recs[idx].b, st);
}

}  // namespace b200

extern "C" void b200_profile_enable(int on) { b200::g_prof_on = on != 0; }

extern "C" int b200_profile_read(int n_cat, double* ms, double* flops, long long* launches) {
  using namespace b200;
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (int i = 0; i < n_cat; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) { set_error("profile: event sync failed"); return -2; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.cat >= 0 && r.cat < n_cat) { ms[r.cat] += t; flops[r.cat] += r.flops; launches[r.cat] += 1; }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}

extern "C" const char* b200_last_error(void) { return b200::g_last_error.c_str(); }
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" long long b200_launch_count(void) { return b200::g_launches.load(); }
extern "C" void b200_reset_launch_count(void) { b200::g_launches.store(0); }
