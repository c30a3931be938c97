#include "driver_api.h"

#include <cstdarg>
#include <atomic>
#include <mutex>

namespace adapcc {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (log_level() >= 1) fprintf(stderr, "[adapcc][error] %s\n", g_err);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int log_level() {
  static int lvl = [] {
    const char* e = getenv("ADAPCC_LOG");
    return e ? atoi(e) : 0;
  }();
  return lvl;
}

template <typename F>
static bool resolve(const char* name, F& out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
    (void)cudaGetLastError();
    out = nullptr;
    return false;
  }
  out = reinterpret_cast<F>(fn);
  return true;
}

const DriverApi& driver() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
    ok &= resolve("cuGetErrorString", api.GetErrorString);
    ok &= resolve("cuDeviceGet", api.DeviceGet);
    ok &= resolve("cuDeviceGetAttribute", api.DeviceGetAttribute);
    ok &= resolve("cuMemGetAllocationGranularity", api.MemGetAllocationGranularity);
    ok &= resolve("cuMemCreate", api.MemCreate);
    ok &= resolve("cuMemRelease", api.MemRelease);
    ok &= resolve("cuMemExportToShareableHandle", api.MemExportToShareableHandle);
    ok &= resolve("cuMemImportFromShareableHandle", api.MemImportFromShareableHandle);
    ok &= resolve("cuMemAddressReserve", api.MemAddressReserve);
    ok &= resolve("cuMemAddressFree", api.MemAddressFree);
    ok &= resolve("cuMemMap", api.MemMap);
    ok &= resolve("cuMemUnmap", api.MemUnmap);
    ok &= resolve("cuMemSetAccess", api.MemSetAccess);
    api.ok = ok;
    bool mc = true;
    mc &= resolve("cuMulticastCreate", api.MulticastCreate);
    mc &= resolve("cuMulticastAddDevice", api.MulticastAddDevice);
    mc &= resolve("cuMulticastBindMem", api.MulticastBindMem);
    mc &= resolve("cuMulticastUnbind", api.MulticastUnbind);
    mc &= resolve("cuMulticastGetGranularity", api.MulticastGetGranularity);
    api.has_multicast = ok && mc;
  });
  return api;
}

const char* cu_error_string(CUresult r) {
  const DriverApi& d = driver();
  const char* s = nullptr;
  if (d.GetErrorString && d.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  static thread_local char buf[32];
  snprintf(buf, sizeof(buf), "CUresult %d", (int)r);
  return buf;
}

}  // namespace adapcc
