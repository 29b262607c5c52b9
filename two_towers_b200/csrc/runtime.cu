// runtime.cu -- library plumbing: error strings, device gate, launch counter.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace tt {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TT_OK;
  set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return TT_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("TT_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// cached per-device verdict: 0 unknown, 1 ok, -1 not sm_100
static std::atomic<int> g_dev_ok[64];

static int check_device(int dev) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    (void)cudaGetLastError();
    set_error("libtt_b200: no CUDA device available (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return TT_ERR_ARCH;
  }
  if (dev < 0 || dev >= ndev) {
    set_error("libtt_b200: device %d out of range (%d devices)", dev, ndev);
    return TT_ERR_INVALID;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("libtt_b200: device %d is sm_%d%d; kernels are built for sm_100a (B200) only", dev,
              major, minor);
    return TT_ERR_ARCH;
  }
  return TT_OK;
}

int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("libtt_b200: no CUDA device available (%s); this library has no CPU fallback",
              cudaGetErrorString(e));
    return TT_ERR_ARCH;
  }
  if (dev >= 0 && dev < 64) {
    int s = g_dev_ok[dev].load(std::memory_order_relaxed);
    if (s == 1) return TT_OK;
  }
  int rc = check_device(dev);
  if (rc == TT_OK && dev >= 0 && dev < 64) g_dev_ok[dev].store(1, std::memory_order_relaxed);
  return rc;
}

}  // namespace tt

extern "C" {

int tt_abi_version(void) { return TT_ABI_VERSION; }
const char* tt_last_error(void) { return tt::g_err; }
int tt_require_sm100(int device) { return tt::check_device(device); }
int64_t tt_launch_count(void) { return tt::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
