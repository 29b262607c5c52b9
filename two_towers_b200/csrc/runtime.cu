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

// one mapped, portable host word per process: kernels store (id * 2 + 1) of an out-of-range token id into it
static long long* g_bad_host = nullptr;
static long long* g_bad_dev = nullptr;
static std::atomic<int> g_bad_state{0};                   // 0 not tried, 1 ready, -1 unavailable
long long* bad_id_word() {
  int st = g_bad_state.load(std::memory_order_acquire);
  if (st == 0) {
    long long* h = nullptr;
    long long* d = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess && h &&
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) == cudaSuccess && d) {
      *h = 0;
      g_bad_host = h; g_bad_dev = d;
      g_bad_state.store(1, std::memory_order_release);
      st = 1;
    } else {
      (void)cudaGetLastError();
      g_bad_state.store(-1, std::memory_order_release);
      st = -1;
    }
  }
  return st == 1 ? g_bad_dev : nullptr;
}

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

int tt_bad_token_id(int64_t* id, int clear) {
  if (tt::g_bad_state.load(std::memory_order_acquire) != 1) return 0;
  const long long w = *reinterpret_cast<volatile long long*>(tt::g_bad_host);
  if (w == 0) return 0;
  if (id) *id = (int64_t)(w >> 1);
  if (clear) *reinterpret_cast<volatile long long*>(tt::g_bad_host) = 0;
  return 1;
}

}  // extern "C"
