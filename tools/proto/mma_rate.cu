// Micro-benchmark (developer tool): cycles per tcgen05.mma as a function of N, operand source (SS: A from shared memory,
// TS: A from tensor memory) and accumulate chains.  One CTA, one issuing thread, operands are whatever the shared memory
// holds (timing only).  Usage: mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gpv-1_b200/csrc/common.cuh"
using namespace gpv;

__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int N, int ts, int reps, int kblocks, int nacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    uint32_t ph = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int it = 0; it < reps; ++it) {
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = make_sdesc_sw128(sb + kb * (N * 128) + k * 32, 0, 1024);
            const uint32_t dd = tm + (uint32_t)(((kb * 4 + k) % nacc) * N);   // nacc independent accumulators, round-robin
            if (ts) umma_f16_ts(dd, tm + 384 + (kb * 4 + k) * 8, bd, idesc, 1);
            else umma_f16(dd, make_sdesc_sw128(sa + kb * 16384 + k * 32, 0, 1024), bd, idesc, 1);
          }
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// TMA pull rate of ONE CTA (or `grid` CTAs): `slots` boxes of [64 cols x rows] bf16 (SWIZZLE_128B) in flight, L2-resident source.
// rank: 2 = 2-D tensor map + .2d instruction, 4 = 4-D map (two unit dims) + .4d instruction.  nprod producer threads (one per warp)
// take alternate slots.
GPV_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, long long* out, int rows, int slots,
                                                     int iters, int src_rows, int rank, int nprod, int lanes, int spin, const __grid_constant__ CUtensorMap tm2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  const int w = lanes ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
  if ((lanes ? threadIdx.x < 32 : (threadIdx.x & 31) == 0) && w < nprod) {
    const uint32_t bytes = rows * 128;
    long long issue = 0;
    const long long t0 = clock64();
    for (int it = w; it < iters + slots; it += nprod) {
      const int s = it % slots;
      if (it >= slots) {
        if (spin) {
          uint32_t ok = 0;
          while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"((uint32_t)(((it / slots) - 1) & 1)) : "memory");
        } else {
          mbar_wait(&full[s], ((it / slots) - 1) & 1);
        }
      }
      if (it < iters) {
        const int r0 = ((blockIdx.x * 64) + it * rows) % (src_rows - rows);
        const long long a = clock64();
        mbar_expect_tx(&full[s], bytes);
        if (rank == 2) tma_load_2d(smem + s * bytes, &tm, &full[s], 0, r0);
        else tma_load_4d(smem + s * bytes, (spin & 2) && (it & 1) ? &tm2 : &tm, &full[s], 0, r0, 0, 0);
        issue += clock64() - a;
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = issue; }
  }
}

#include <cuda.h>
static CUtensorMap make_map2d(void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows, int rank) {
  CUtensorMap m;
  cuuint64_t gdim[4] = {cols, rows, 1, 1}, gstr[3] = {cols * 2, cols * 2 * rows, cols * 2 * rows};
  cuuint32_t box[4] = {64, box_rows, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return m;
}

static void tma_bench(long long* d) {
  void* src;
  const int src_rows = 4096, cols = 256;      // 2 MB, L2-resident after the first pass
  cudaMalloc(&src, (size_t)src_rows * cols * 2);
  cudaMemset(src, 0, (size_t)src_rows * cols * 2);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rank : {4})
    for (int spin : {0, 1, 2, 3})
    for (int lanes : {0})
    for (int nprod : {1})
      for (int grid : {1})
        for (int rows : {64, 256})
          for (int slots : {1, 4}) {
            if (rows * 128 * slots > 196 * 1024 || slots < nprod) continue;
            CUtensorMap m = make_map2d(src, cols, src_rows, rows, rank);
            const int iters = 64;
            for (int w = 0; w < 2; ++w) tma_kernel<<<grid, 128, 200 * 1024>>>(m, d, rows, slots, iters, src_rows, rank, nprod, lanes, spin, make_map2d((char*)src + 1024 * 512, cols, src_rows - 1024, rows, rank));
            long long h[2];
            cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
            printf("TMA %s rank=%d producers=%d grid=%3d box 64x%3d (%2d KB) slots=%d: %.1f B/clk/SM  (%.0f cyc per box, issue %.0f cyc per op)\n", spin == 0 ? "try_wait" : spin == 1 ? "test_wait spin" : spin == 2 ? "try_wait, two maps" : "spin, two maps", rank,
                   nprod, grid, rows, rows * 128 / 1024, slots, (double)iters * rows * 128 / h[0], (double)h[0] / iters,
                   (double)h[1] / (iters / nprod));
          }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int nacc : {1, 2, 4})
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 256}) {
      const int reps = 16, kblocks = 4;
      if (nacc * N > (ts ? 384 : 512)) continue;
      rate_kernel<<<1, 128, 200 * 1024>>>(d, N, ts, reps, kblocks, nacc);
      long long h[6];
      cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      const int n = reps * kblocks * 4;
      printf("acc=%d %s N=%3d: %d MMAs  issue %.1f cyc/MMA  complete %.1f cyc/MMA  (floor 128*N/256 = %d)\n", nacc, ts ? "TS" : "SS", N, n,
             (double)h[4] / n, (double)h[5] / n, N / 2);
    }
  return 0;
}
