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

// Where a watchdog trap fired.  The mbarrier / flag waits of the tensor-core kernels trap instead of hanging the GPU when a
// protocol bug (or a lost co-residency assumption) keeps them waiting; a trap kills the context, so the waiting thread
// first writes {tag, block, thread, parity} into MAPPED PINNED HOST memory, which the process can still read afterwards.
static unsigned* g_trace_host = nullptr;
static unsigned* g_trace_dev = nullptr;
unsigned* trap_trace_device() {
  if (g_trace_dev == nullptr) {
    void* h = nullptr;
    void* d = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
      for (int i = 0; i < 16; ++i) reinterpret_cast<unsigned*>(h)[i] = 0u;
      g_trace_host = reinterpret_cast<unsigned*>(h);
      g_trace_dev = reinterpret_cast<unsigned*>(d);
    } else {
      (void)cudaGetLastError();
    }
  }
  return g_trace_dev;
}

}  // namespace aura

extern "C" int aura_debug_last_trap(uint32_t out[4]) {
  if (out == nullptr) return AURA_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) out[i] = aura::g_trace_host ? aura::g_trace_host[i] : 0u;
  return AURA_OK;
}
extern "C" int aura_version(void) { return AURA_HIPPO_VERSION; }
extern "C" const char* aura_last_error_string(void) { return aura::g_err; }
extern "C" uint64_t aura_kernel_launches(void) { return aura::launches(); }
