// Error plumbing and device checks shared by every C-ABI entry point.
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "../../include/gpvb200.h"
#include "host_util.h"

namespace gpv {

static char g_err[1024] = "";
static std::mutex g_err_mu;

void set_last_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> g(g_err_mu);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  return GPV_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPVB200_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

int ensure_arch() {
  static int cached = 1;  // 1 = unknown
  if (cached != 1) return cached;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_last_error("no CUDA device: %s", cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_last_error("gpvb200 kernels are built for sm_100a only; device has compute capability %d.x (no fallback)", major);
    cached = GPV_ERR_ARCH;
    return cached;
  }
  cached = GPV_OK;
  return cached;
}

}  // namespace gpv

extern "C" int gpvb200_version(void) { return 100; }

extern "C" int gpvb200_last_error(char* buf, size_t n) {
  std::lock_guard<std::mutex> g(gpv::g_err_mu);
  if (buf && n) {
    strncpy(buf, gpv::g_err, n - 1);
    buf[n - 1] = 0;
  }
  return (int)strlen(gpv::g_err);
}
