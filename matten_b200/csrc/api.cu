// Error handling, version and device gate of the C ABI (include/matten_b200.h).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace mt {

uint64_t launch_count();

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return set_error(MT_ECUDA, "cudaGetDevice failed: %s (no CUDA device? there is no CPU fallback)",
                     cudaGetErrorString(e));
  // cache per device
  static int cached[64] = {0};  // 0 unknown, 1 ok, 2 bad
  if (dev >= 0 && dev < 64 && cached[dev] == 1) return MT_OK;
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return set_error(MT_ECUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  if (major != 10) {
    if (dev >= 0 && dev < 64) cached[dev] = 2;
    return set_error(MT_EARCH, "device %d is sm_%d%d; matten_b200 is built for sm_100a only", dev, major, minor);
  }
  if (dev >= 0 && dev < 64) cached[dev] = 1;
  return MT_OK;
}

}  // namespace mt

extern "C" {

int mt_abi_version(void) { return MT_ABI_VERSION; }

const char* mt_last_error(void) { return mt::last_error_buf(); }

uint64_t mt_launch_count(void) { return mt::launch_count(); }

int mt_device_supported(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return mt::set_error(MT_ECUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  if (major != 10) return mt::set_error(MT_EARCH, "device %d is not compute capability 10.x", device);
  return MT_OK;
}

}  // extern "C"
