// C-ABI plumbing shared by every entry point: thread-local error string, launch counter, device probe.
#include "common.cuh"
#include <atomic>
#include <string.h>

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void qeb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void qeb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

QEB_API const char* qeb_last_error(void) { return g_err; }

QEB_API int qeb_abi_version(void) { return 1; }

QEB_API long long qeb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

QEB_API void qeb_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

// 0 when the current device is an sm_100 part this library was compiled for
QEB_API int qeb_check_device(void) {
  int dev = 0;
  QEB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  QEB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    qeb_set_error("qeb kernels are built for sm_100a only; device %d is sm_%d%d (%s)", dev, p.major, p.minor, p.name);
    return QEB_ERR_UNSUPPORTED;
  }
  return QEB_OK;
}
