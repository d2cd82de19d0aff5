// Library-wide state: last error string, launch counter, cached device properties.
#include "common.cuh"
#include "../../include/b200_distill.h"

#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <algorithm>

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

// ---- dispatch options: a table of ints the launchers read (no getenv on any launch path). Defaults = the production
// path; the environment is consulted ONCE, when the library is loaded, so that A/B runs of a tool need no rebuild, and
// tests flip entries through b200_set_option().
struct OptDef { const char* name; const char* env; int value; };
static OptDef g_opts[OPT_COUNT] = {
    {"pdl", "B200_PDL", 1},                          // programmatic dependent launch attribute on every launch
    {"attn_tc_fwd", "B200_ATTN_TC", 1},              // tcgen05 attention forward (head_dim 64)
    {"attn_tc_fwd_long", "B200_ATTN_TC_LONG", 1},    // ... also for key ranges > 272 (online softmax)
    {"attn_tc_bwd", "B200_ATTN_TC_BWD", 1},          // tcgen05 attention backward (Nq, Nk <= 256)
    {"attn_bwd_fused", "B200_ATTN_BWD_FUSED", 1},    // mma.sync fallback: one-kernel backward when Nk <= 256
    {"attn_tc_bwd_long", "B200_ATTN_TC_BWD_LONG", 1},   // tcgen05 backward also for Nq / Nk > 256 (fp32 dQ accumulator)
    {"gemm_v2", "B200_GEMM_V2", 1},                  // bulk-store GEMM kernel (0: first-generation kernel everywhere)
    {"gemm_bn", "B200_GEMM_BN", 0},                  // force the tile width (128 / 192 / 256; 0 = cost model)
    {"gemm_2cta", "B200_GEMM_2CTA", -1},             // cta_group::2: -1 by K, 0 never, 1 whenever possible
    {"gemm_inplace_red", "B200_GEMM_INPLACE_RED", 1},
    {"gemm_ew", "B200_GEMM_EW", 16},                 // epilogue warps where the epilogue stages no operand tile
    {"gemm_dbg", "B200_GEMM_DBG", 0},                // probe mask (only in -DB200_GEMM_PROBES builds)
    {"attn_probe_skip", "B200_ATTN_PROBE_SKIP", 0},  // work-skipping mask (only in -DB200_ATTN_PROBES builds)
    {"gemm_ln", "B200_GEMM_LN", 1},                  // LayerNorm-prologue GEMM (K <= 384) instead of layernorm_fwd + GEMM
    {"stat_attn_tc_bwd", "", 0},                     // counters (read with b200_get_option, reset with b200_set_option):
    {"stat_attn_mma_bwd", "", 0},                    //   attention backward launches per path
    {"stat_attn_tc_fwd", "", 0},
    {"stat_attn_mma_fwd", "", 0},
    {"attn_pp_fwd", "B200_ATTN_PP", 1},              // two-tile tcgen05 attention forward (Nk <= 256, head dims 16-64)
    {"stat_attn_pp_fwd", "", 0},
    {"gemm_aux_deep", "B200_GEMM_AUX_DEEP", 1},      // aux-epilogue GEMMs: aux tiles two units ahead, across tiles
};
static const bool g_opts_init = [] {
  for (auto& o : g_opts) {
    const char* e = o.env[0] ? getenv(o.env) : nullptr;
    if (e && ((e[0] >= '0' && e[0] <= '9') || e[0] == '-')) o.value = atoi(e);
  }
  return true;
}();
int option(int id) { return g_opts[id].value; }
void bump_stat(int id) { ++g_opts[id].value; }

bool pdl_enabled() { return g_opts[OPT_PDL].value != 0; }

// ---- optional per-launch timing of the dense kernels (bench.py's roofline leg): CUDA events on the launch stream
struct ProfRec { cudaEvent_t a, b; double flops; int cat; double bytes; };
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
  r.a = take_event(); r.b = take_event(); r.flops = 0; r.cat = 0; r.bytes = 0;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  return (int)g_recs.size() - 1;
}

void prof_end(int idx, cudaStream_t st, double flops, int cat, double bytes) {
  if (idx < 0) return;
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (idx >= (int)g_recs.size()) return;
  g_recs[idx].flops = flops;
  g_recs[idx].cat = cat;
  g_recs[idx].bytes = bytes;
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

// What bracketing ONE kernel with an event pair adds to its measured duration (no programmatic overlap with the
// predecessor across the event, launch latency after the first event, timestamp granularity): a null kernel is timed
// (a) bracketed launch by launch exactly like prof_begin / prof_end do, (b) back to back between one event pair;
// *bracketed_us and *back_to_back_us are per-launch medians / means, their difference is the overhead.
namespace b200 {
__global__ void profile_null_kernel() {
  pdl_trigger();
  pdl_wait();
}
}  // namespace b200
extern "C" int b200_profile_event_overhead(int reps, float* bracketed_us, float* back_to_back_us, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(reps >= 8 && reps <= 4096 && bracketed_us && back_to_back_us, "bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev(2 * reps + 2);
  for (auto& e : ev) B200_CUDA_OK(cudaEventCreate(&e));
  for (int i = 0; i < 8; ++i) B200_CUDA_OK(launch_pdl(profile_null_kernel, dim3(148), dim3(128), 0, st));
  for (int i = 0; i < reps; ++i) {
    B200_CUDA_OK(cudaEventRecord(ev[2 * i], st));
    B200_CUDA_OK(launch_pdl(profile_null_kernel, dim3(148), dim3(128), 0, st));
    B200_CUDA_OK(cudaEventRecord(ev[2 * i + 1], st));
  }
  B200_CUDA_OK(cudaEventRecord(ev[2 * reps], st));
  for (int i = 0; i < reps; ++i) B200_CUDA_OK(launch_pdl(profile_null_kernel, dim3(148), dim3(128), 0, st));
  B200_CUDA_OK(cudaEventRecord(ev[2 * reps + 1], st));
  B200_CUDA_OK(cudaEventSynchronize(ev[2 * reps + 1]));
  std::vector<float> t(reps);
  for (int i = 0; i < reps; ++i) B200_CUDA_OK(cudaEventElapsedTime(&t[i], ev[2 * i], ev[2 * i + 1]));
  std::sort(t.begin(), t.end());
  float all = 0.f;
  B200_CUDA_OK(cudaEventElapsedTime(&all, ev[2 * reps], ev[2 * reps + 1]));
  *bracketed_us = t[reps / 2] * 1e3f;
  *back_to_back_us = all * 1e3f / reps;
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}

// per-launch records in launch order (ms, algorithmic flops, category); clears them like b200_profile_read
extern "C" int b200_profile_read_records(int cap, float* ms, double* flops, int* cat, double* bytes) {
  using namespace b200;
  std::lock_guard<std::mutex> g(g_prof_mu);
  int n = 0;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) { set_error("profile: event sync failed"); return -2; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (n < cap) { ms[n] = t; flops[n] = r.flops; cat[n] = r.cat; if (bytes) bytes[n] = r.bytes; ++n; }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return n;
}

extern "C" int b200_set_option(const char* name, int value) {
  using namespace b200;
  for (auto& o : g_opts)
    if (name && strcmp(name, o.name) == 0) { o.value = value; return 0; }
  set_error(std::string("b200_set_option: unknown option '") + (name ? name : "(null)") + "'");
  return -1;
}
extern "C" int b200_get_option(const char* name) {
  using namespace b200;
  for (auto& o : g_opts)
    if (name && strcmp(name, o.name) == 0) return o.value;
  return -1;
}

extern "C" const char* b200_last_error(void) { return b200::g_last_error.c_str(); }
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" long long b200_launch_count(void) { return b200::g_launches.load(); }
extern "C" void b200_reset_launch_count(void) { b200::g_launches.store(0); }
