// Library-level entry points: version, error text, device capability.
#include "dp_common.cuh"
#include <string.h>

namespace dp {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = 0;         // measured: slower on the whole step (18.91 vs 18.24 ms), kept as an option
int g_pdl_small = 0;   // programmatic dependent launch of the small finalize / split-reduce kernels only
long long* g_dbg = nullptr;
size_t g_dbg_slots = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DP_MAX_DEVICES) return 0;
  return dev;
}

int num_sms() {   // per-device cache: one process may drive several GPUs
  static int cached[DP_MAX_DEVICES] = {};
  const int dev = current_device();
  if (cached[dev]) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[dev] = n;
  return n;
}

}  // namespace dp

DP_API int dp_version(void) { return 100; }

DP_API const char* dp_last_error(void) { return dp::g_err; }

DP_API int dp_num_sms(void) { return dp::num_sms(); }

DP_API unsigned long long dp_launch_count(void) { return dp::g_launches; }

DP_API int dp_set_debug_buffer(void* ptr, size_t bytes) {
  dp::g_dbg = static_cast<long long*>(ptr);
  dp::g_dbg_slots = ptr ? bytes / sizeof(long long) : 0;
  return DP_OK;
}

DP_API int dp_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    dp::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return DP_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    dp::set_error("device %d is sm_%d%d; this library holds sm_100a code only", dev, major, minor);
    return DP_ERR_ARCH;
  }
  return DP_OK;
}
