// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
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

// Kernel launch with programmatic dependent launch allowed (GPVB200_PDL=0 turns it off): the kernel may start while the
// previous kernel of the stream drains; every kernel launched this way begins with pdl_sync() (common.cuh), so global
// memory is only touched once the predecessor has completed and flushed.  Inside a stream capture this becomes a
// programmatic edge of the CUDA graph.
bool pdl_enabled();
template <typename K, typename... A>
inline void launch_k(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);   // errors surface through check_launch() (cudaGetLastError)
}

// 4-D bf16 tensor map (SWIZZLE_128B, zero out-of-bounds fill), cached by (pointer, shape); gemm_umma.cu.  dims / box /
// element strides innermost first; strides_el = element strides of dims 1..3.
int make_map(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_el[3], const uint32_t box[4],
             const uint32_t estr[4]);

#define GPV_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      gpv::set_last_error(__VA_ARGS__);   \
      return gpv::GPV_ERR_ARG;            \
    }                                     \
  } while (0)

}  // namespace gpv
