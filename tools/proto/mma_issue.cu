// Micro-benchmark (developer tool): what one stage of the contraction kernel's MMA-issue loop costs the issuing warp, as the kernel
// issues it: a full warp, `elect_one()`, four tcgen05.mma (128 x N x 16, both operands in shared memory, SWIZZLE_128B K-major) and
// one tcgen05.commit per 64-deep stage.  Variants: N, commit per stage or only at the end, the number of CTAs.  Usage: mma_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gpv-1_b200/csrc/common.cuh"
using namespace gpv;

// variant: 0 = the kernel's loop (elect per stage, descriptors rebuilt per MMA); 1 = one elect around the whole loop, descriptors advanced
// by adding constants to a base; 2 = as 1 with four stages (16 MMAs) unrolled per iteration; 3 = as 1 plus tcgen05.fence::after_thread_sync
// per stage; 4 = as 1 plus an mbarrier try_wait (on a completed phase) per stage; 5 = as 1 plus both (the kernel's per-stage sequence)
__global__ void __launch_bounds__(128, 1) issue_kernel(long long* out, int N, int stages, int commit_each, int nslots, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8], done, ready;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    mbar_init(&done, 1);
    mbar_init(&ready, 1);
    fence_barrier_init();
    mbar_arrive(&ready);    // phase 0 of `ready` is complete: a wait on parity 0 passes at once
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t stage_bytes = 16384u + (uint32_t)N * 128u;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (variant >= 1) {
        if (elect_one()) {
          const uint64_t a0 = make_sdesc_sw128(smem_u32(smem), 0, 1024), b0 = make_sdesc_sw128(smem_u32(smem) + 16384u, 0, 1024);
          const uint64_t sstep = stage_bytes >> 4;
          if (variant == 1 || variant >= 3) {
#pragma unroll 1
            for (int st = 0; st < stages; ++st) {
              const int s = st % nslots;
              if (variant == 4 || variant == 5) mbar_wait(&ready, 0u);
              if (variant == 3 || variant == 5) tc_fence_after();
              const uint64_t a = a0 + s * sstep, b = b0 + s * sstep;
              umma_f16(tm, a, b, idesc, st > 0 ? 1u : 0u);
              umma_f16(tm, a + 2, b + 2, idesc, 1u);
              umma_f16(tm, a + 4, b + 4, idesc, 1u);
              umma_f16(tm, a + 6, b + 6, idesc, 1u);
              if (commit_each) umma_commit(&bar[s]);
            }
          } else {
#pragma unroll 1
            for (int st = 0; st < stages; st += 4) {
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint64_t a = a0 + s * sstep, b = b0 + s * sstep;
                umma_f16(tm, a, b, idesc, (st > 0 || s > 0) ? 1u : 0u);
                umma_f16(tm, a + 2, b + 2, idesc, 1u);
                umma_f16(tm, a + 4, b + 4, idesc, 1u);
                umma_f16(tm, a + 6, b + 6, idesc, 1u);
                if (commit_each) umma_commit(&bar[s]);
              }
            }
          }
        }
        __syncwarp();
      } else
      for (int st = 0; st < stages; ++st) {
        const int s = st % nslots;
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem) + s * stage_bytes, sb = sa + 16384u;
          for (int k = 0; k < 4; ++k)
            umma_f16(tm, make_sdesc_sw128(sa + k * 32, 0, 1024), make_sdesc_sw128(sb + k * 32, 0, 1024), idesc, (st > 0 || k > 0) ? 1u : 0u);
          if (commit_each) umma_commit(&bar[s]);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&done);
      __syncwarp();
      mbar_wait(&done, rep & 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        out[rep * 2] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int variant : {1, 3, 4, 5})
  for (int grid : {148})
    for (int commit_each : {1})
      for (int N : {32, 64, 128, 256}) {
        const int stages = 64, nslots = 4;
        issue_kernel<<<grid, 128, 200 * 1024>>>(d, N, stages, commit_each, nslots, variant);
        long long h[6];
        cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("variant %d grid=%3d N=%3d commit %s: issue %.0f clk per stage (4 MMAs), complete %.0f clk per stage  (tensor floor %d)\n", variant, grid, N,
               commit_each ? "per stage" : "at the end", (double)h[4] / stages, (double)h[5] / stages, 4 * N / 2);
      }
  return 0;
}
