// api.cu -- library-level entry points: version, error string, arch check, launch accounting.
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace sdvar {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_arch() {
  static thread_local int cached_dev = -1, cached_rc = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: libsdvar_b200 has no CPU fallback");
    return SDVAR_ERR_CUDA;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cached_dev = dev;
  cached_rc = (major == 10) ? SDVAR_OK : SDVAR_ERR_ARCH;
  if (cached_rc) set_error("device %d is sm_%d0: libsdvar_b200 is built for sm_100a only (no fallback)", dev, major);
  return cached_rc;
}

}  // namespace sdvar

extern "C" int sdvar_abi_version(void) { return SDVAR_ABI_VERSION; }
extern "C" const char* sdvar_last_error(void) { return sdvar::g_err; }
extern "C" long long sdvar_launch_count(void) { return sdvar::g_launches.load(); }
extern "C" int sdvar_arch_check(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
    sdvar::set_error("cannot query device %d", device);
    return SDVAR_ERR_CUDA;
  }
  if (major != 10) {
    sdvar::set_error("device %d is sm_%d0, need sm_100", device, major);
    return SDVAR_ERR_ARCH;
  }
  return SDVAR_OK;
}
extern "C" int sdvar_num_sms(int device) {
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return SDVAR_ERR_CUDA;
  return sms;
}
