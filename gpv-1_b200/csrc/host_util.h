// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

namespace gpv {

// Error codes returned through the C ABI (0 = ok).
enum {
  GPV_OK = 0,
  GPV_ERR_ARG = -1,      // bad shape / alignment / unsupported combination
  GPV_ERR_CUDA = -2,     // a CUDA runtime / driver call failed
  GPV_ERR_ARCH = -3,     // device is not compute capability 10.x
  GPV_ERR_WORKSPACE = -4 // caller-provided workspace too small
};

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> error code
int ensure_arch();                   // GPV_OK only on sm_100

#define GPV_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      gpv::set_last_error(__VA_ARGS__);   \
      return gpv::GPV_ERR_ARG;            \
    }                                     \
  } while (0)

}  // namespace gpv
