// Library-wide state: last error string, launch counter, cached device properties.
#include "common.cuh"
#include "../../include/b200_distill.h"

#include <atomic>
#include <mutex>

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

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_last_error.c_str(); }
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" long long b200_launch_count(void) { return b200::g_launches.load(); }
extern "C" void b200_reset_launch_count(void) { b200::g_launches.store(0); }
