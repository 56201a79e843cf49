// Host-side plumbing of libaura_hippo: thread-local error string, cached device attributes,
// version.  (C ABI: include/aura_hippo.h.)
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>

#include "aura_common.cuh"

namespace aura {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return AURA_ERR_CUDA;
}

static int cached_attr(cudaDeviceAttr attr, int* cache /* per device, up to 16 */) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = 0;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess || v <= 0) v = -1;
    cache[dev] = v;
  }
  return cache[dev];
}

int sm_count() {
  static int cache[16] = {0};
  const int v = cached_attr(cudaDevAttrMultiProcessorCount, cache);
  return v > 0 ? v : 148;  // B200; only hit when no device is visible (workspace sizing on a CPU box)
}

int max_smem_optin() {
  static int cache[16] = {0};
  const int v = cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache);
  return v > 0 ? v : 232448;
}

}  // namespace aura

extern "C" int aura_version(void) { return AURA_HIPPO_VERSION; }
extern "C" const char* aura_last_error_string(void) { return aura::g_err; }
extern "C" uint64_t aura_kernel_launches(void) { return aura::launches(); }
