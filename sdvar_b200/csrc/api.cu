// api.cu -- library-level entry points: version, error string, arch check, launch accounting.
#include <atomic>
#include <mutex>
#include <vector>

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

int sm_count() {
  static thread_local int dev_cached = -1, sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
  }
  return sms;
}

// ---- profiling ----------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int family; double work; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

ProfileScope::ProfileScope(cudaStream_t s, int family, double work) : st(s), slot(-1) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r;
  r.family = family;
  r.work = work;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfileScope::~ProfileScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace sdvar

extern "C" int sdvar_profile_begin(void) {
  std::lock_guard<std::mutex> lk(sdvar::g_prof_mu);
  for (auto& r : sdvar::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  sdvar::g_prof.clear();
  sdvar::g_prof_on.store(1);
  return SDVAR_OK;
}
// ms[f], work[f], launches[f] for f < SDVAR_PROFILE_FAMILIES; synchronises the device
extern "C" int sdvar_profile_end(double* ms, double* work, long long* launches) {
  sdvar::g_prof_on.store(0);
  if (cudaDeviceSynchronize() != cudaSuccess) { sdvar::set_error("cudaDeviceSynchronize failed in sdvar_profile_end"); return SDVAR_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(sdvar::g_prof_mu);
  for (int f = 0; f < sdvar::FAM_COUNT; ++f) { ms[f] = 0; work[f] = 0; launches[f] = 0; }
  for (auto& r : sdvar::g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.family] += t; work[r.family] += r.work; launches[r.family] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  sdvar::g_prof.clear();
  return SDVAR_OK;
}

extern "C" int sdvar_abi_version(void) { return SDVAR_ABI_VERSION; }
extern "C" const char* sdvar_last_error(void) { return sdvar::g_err; }
extern "C" long long sdvar_launch_count(void) { return sdvar::g_launches.load(); }
// a replayed CUDA graph launches kernels the library never sees: the host adds the graph's kernel count per replay
extern "C" void sdvar_count_launches(long long n) { sdvar::g_launches.fetch_add(n, std::memory_order_relaxed); }
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
