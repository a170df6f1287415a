// Developer prototype (not part of libgpvb200.so): the next main-loop generation of the contraction kernel.
//
// the production kernel (csrc/gemm_umma.cu) is one CTA per SM with a 128 x BN tile, and its main loop is
// bound by the rate at which ONE SM's TMA stream is served from L2 (~70 GB/s per SM): a 128x256x64 stage is 48 KB.
// This prototype pairs the two SMs of a TPC (`cta_group::2`): the pair owns a 256 x 256 tile, each CTA stages only ITS
// half of A (128 rows) and ITS half of B (128 rows) -- 32 KB per stage for the same 128x256x64 of math per SM -- and
// the leader CTA issues one `tcgen05.mma.cta_group::2` (M = 256) that reads both halves of B through the pair's shared
// memory and writes each CTA's 128 accumulator rows into that CTA's own TMEM.
//
//   both CTAs   warp 0      TMA producer: own halves; complete_tx is sent to the LEADER's full barrier
//   leader      warp 1      MMA issuer; tcgen05.commit multicast frees the stage in both CTAs / publishes the accumulator
//   both CTAs   warps 2-9   epilogue from their own TMEM; the accumulator-empty barrier lives in the leader
//
// Standalone: builds D = A[M,K] * B[N,K]^T (bf16 in, fp32 accumulate, bf16 out), checks it against a CUDA-core kernel,
// and times it beside the production kernel (gpvb200_gemm through the C-ABI) on the same operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/proto/gemm_2cta.cu -o tools/proto/_build/gemm_2cta \
//        -Lgpv-1_b200/lib -lgpvb200 -lcuda -Xlinker -rpath -Xlinker '$ORIGIN/../../../gpv-1_b200/lib'
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/gpvb200.h"
#include "../../gpv-1_b200/csrc/common.cuh"

using namespace gpv;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

// ---------------------------------------------------------------- pair primitives (the rest live in csrc/common.cuh)
GPV_DEVINL uint32_t map_to_cta(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
// This CTA's half of a pair operand into its own shared memory; the bytes are counted on the barrier `bar_cluster`
// (a shared::cluster address: the leader's full barrier).
GPV_DEVINL void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}

GPV_DEVINL uint32_t leader_addr(uint32_t cta_addr, int mode) { return mode == 0 ? map_to_cta(cta_addr, 0) : (cta_addr & 0xFEFFFFFFu); }

constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int BK = 64;
constexpr int BN = 256;            // pair tile: 256 rows x 256 columns; per CTA 128 rows of A and 128 rows of B
constexpr int kStageBytes = 128 * BK * 2 * 2;   // A half + B half = 32 KB

struct PParams {
  int M, N, K, m_pairs, n_tiles, total, k_iters, nstages;
  int kblk;       // 64-deep k-blocks per pipeline stage (1..3): stage = kblk x 32 KB per CTA
  int flags;      // dissection: 1 = no MMA (stages released at once), 2 = no TMA (stages declared full at once), 4 = no stores
  int bar_mode;   // how the leader's barriers are addressed from the peer: 0 = mapa to rank 0, 1 = clear the peer bit (CUTLASS Sm100MmaPeerBitMask)
  bf16* D;
  long long ldd;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int S = p.nstages;
  const uint32_t stage_bytes = (uint32_t)p.kblk * kStageBytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)S * stage_bytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
  constexpr uint32_t kTmemCols = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);   // multicast commit from the leader's MMA warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);                 // multicast commit
      mbar_init(&acc_empty[b], 2 * kEpiWarps);    // leader's: epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                // barriers of both CTAs initialised, both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int gi = 0;
      for (int w = pair; w < p.total; w += npairs) {
        const int nt = w % p.n_tiles, mp = w / p.n_tiles;
        const int m0 = mp * 256 + (int)rank * 128, n0 = nt * BN + (int)rank * 128;
        for (int it = 0; it < p.k_iters; ++it, ++gi) {
          const int s = gi % S;
          const uint32_t ph = (uint32_t)(gi / S) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + (uint32_t)p.kblk * 16384u;
          if (p.flags & 2) {
            if (leader) mbar_arrive(&full_bar[s]);
            continue;
          }
          if (leader) mbar_expect_tx(&full_bar[s], 2u * stage_bytes);
          const uint32_t bar = leader_addr(smem_u32(&full_bar[s]), p.bar_mode);
          for (int kb = 0; kb < p.kblk; ++kb) {
            tma_load_2d_pair(sa + kb * 16384u, &tmA, bar, (it * p.kblk + kb) * BK, m0);
            tma_load_2d_pair(sb + kb * 16384u, &tmB, bar, (it * p.kblk + kb) * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader only)
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      int gi = 0, j = 0;
      for (int w = pair; w < p.total; w += npairs, ++j) {
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
        for (int it = 0; it < p.k_iters; ++it, ++gi) {
          const int s = gi % S;
          const uint32_t ph = (uint32_t)(gi / S) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + (uint32_t)p.kblk * 16384u;
            if (!(p.flags & 1)) {
              for (int kb = 0; kb < p.kblk; ++kb) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t ad = make_sdesc_sw128(sa + kb * 16384u + k * 32, 0u, 1024u);
                  const uint64_t bd = make_sdesc_sw128(sb + kb * 16384u + k * 32, 0u, 1024u);
                  umma_f16_pair(tacc, ad, bd, idesc, (it > 0 || kb > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            umma_commit_pair(&empty_bar[s]);
            if (it == p.k_iters - 1) umma_commit_pair(&acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================== epilogue (both CTAs, own TMEM rows)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int cbase = half * (BN / 2);
    int j = 0;
    for (int w = pair; w < p.total; w += npairs, ++j) {
      const int nt = w % p.n_tiles, mp = w / p.n_tiles;
      const int buf = j & 1;
      const long long row = (long long)mp * 256 + rank * 128 + r;
      mbar_wait(&acc_full[buf], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN / 2 / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cbase + c * 32), acc);
        tmem_ld_wait();
        const int col = nt * BN + cbase + c * 32;
        if (row < p.M && col + 32 <= p.N && !(p.flags & 4)) {
          bf16* dp = p.D + row * p.ldd + col;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(acc[8 * i + 0]), __uint_as_float(acc[8 * i + 1]));
            o.y = pack_bf16x2(__uint_as_float(acc[8 * i + 2]), __uint_as_float(acc[8 * i + 3]));
            o.z = pack_bf16x2(__uint_as_float(acc[8 * i + 4]), __uint_as_float(acc[8 * i + 5]));
            o.w = pack_bf16x2(__uint_as_float(acc[8 * i + 6]), __uint_as_float(acc[8 * i + 7]));
            *reinterpret_cast<uint4*>(dp + 8 * i) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_addr(smem_u32(&acc_empty[buf]), p.bar_mode));
    }
  }

  tc_fence_before();
  cluster_sync_all();                // nobody leaves while its pair may still touch its shared memory / TMEM
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------- reference + harness
__global__ void ref_gemm_kernel(const bf16* A, const bf16* B, float* D, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(B[(size_t)n * K + k]);
  D[(size_t)m * N + n] = acc;
}
__global__ void fill_kernel(bf16* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 0x9E3779B1u + seed;
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    p[i] = __float2bfloat16(((int)(x & 0xFFFF) - 32768) / 32768.0f);
  }
}
__global__ void cmp_kernel(const bf16* D, const float* R, size_t n, unsigned long long* bad, float* maxerr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = __bfloat162float(D[i]), r = R[i];
    const float e = fabsf(d - r);
    if (!(e <= 0.02f * fabsf(r) + 0.08f)) atomicAdd(bad, 1ull);
    atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(e));
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map_2d(PFN_encodeTiled enc, const void* ptr, uint64_t rows, uint64_t K, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {K, rows}, str[1] = {K * 2};
  cuuint32_t box[2] = {BK, box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(2);
  }
  return m;
}

int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 38400, N = argc > 2 ? atoi(argv[2]) : 256, K = argc > 3 ? atoi(argv[3]) : 1024;
  int stages = argc > 4 ? atoi(argv[4]) : 6, reps = argc > 5 ? atoi(argv[5]) : 20, bar_mode = argc > 6 ? atoi(argv[6]) : 0;
  int flags = argc > 7 ? atoi(argv[7]) : 0, kblk = argc > 8 ? atoi(argv[8]) : 1, max_pairs = argc > 9 ? atoi(argv[9]) : 1000;
  if (N % 256 || K % (64 * kblk)) {
    fprintf(stderr, "prototype needs N %% 256 == 0 and K %% (64 * kblk) == 0\n");
    return 2;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  bf16 *A, *B, *D, *D1;
  float* R;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&B, (size_t)N * K * 2));
  CK(cudaMalloc(&D, (size_t)M * N * 2));
  CK(cudaMalloc(&D1, (size_t)M * N * 2));
  CK(cudaMalloc(&R, (size_t)M * N * 4));
  fill_kernel<<<1024, 256>>>(A, (size_t)M * K, 1u);
  fill_kernel<<<256, 256>>>(B, (size_t)N * K, 2u);
  CK(cudaMemset(D, 0, (size_t)M * N * 2));
  ref_gemm_kernel<<<dim3((N + 127) / 128, M), 128>>>(A, B, R, M, N, K);
  CK(cudaDeviceSynchronize());

  PParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_pairs = (M + 255) / 256;
  p.n_tiles = N / BN;
  p.total = p.m_pairs * p.n_tiles;
  p.k_iters = K / (BK * kblk);
  p.kblk = kblk;
  p.flags = flags;
  p.nstages = stages;
  p.bar_mode = bar_mode;
  p.D = D;
  p.ldd = N;
  CUtensorMap ma = make_map_2d(enc, A, M, K, 128), mb = make_map_2d(enc, B, N, K, 128);
  const size_t smem = (size_t)stages * kblk * kStageBytes + 1024 + 256;
  CK(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int npairs = p.total < sms / 2 ? p.total : sms / 2;
  if (npairs > max_pairs) npairs = max_pairs;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  gemm_pair_kernel<<<2 * npairs, kThreads, smem>>>(ma, mb, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  unsigned long long* bad;
  float* maxerr;
  CK(cudaMalloc(&bad, 8));
  CK(cudaMalloc(&maxerr, 4));
  CK(cudaMemset(bad, 0, 8));
  CK(cudaMemset(maxerr, 0, 4));
  cmp_kernel<<<1024, 256>>>(D, R, (size_t)M * N, bad, maxerr);
  unsigned long long hbad;
  float hmax;
  CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hmax, maxerr, 4, cudaMemcpyDeviceToHost));
  if (flags) hbad = 0;   // dissection modes do not compute the product
  printf("pair kernel  M=%d N=%d K=%d stages=%d x %d KB pairs=%d flags=%d: mismatches %llu / %zu, max abs err %.4f\n", M, N, K, stages,
         32 * kblk, npairs, flags, hbad, (size_t)M * N, hmax);
  const double flop = 2.0 * M * N * K;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) gemm_pair_kernel<<<2 * npairs, kThreads, smem>>>(ma, mb, p);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("pair kernel  %.2f us / launch, %.1f TFLOP/s\n", 1e3 * ms / reps, flop * reps / (ms * 1e-3) / 1e12);

  if (flags || max_pairs < 1000) return 0;
  // production kernel (one CTA per SM, 128 x BN tiles) on the same operands, through the C-ABI
  gpvb200_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.mode = 0; d.M = M; d.N = N; d.K = K; d.batch = 1; d.alpha = 1.0f; d.splits = 1;
  d.A = A; d.B = B; d.D = D1; d.lda = K; d.ldb = K; d.ldd = N;
  int rc = gpvb200_gemm(&d, nullptr);
  CK(cudaDeviceSynchronize());
  if (rc != 0) {
    char msg[512];
    gpvb200_last_error(msg, sizeof(msg));
    printf("production kernel failed: %d %s\n", rc, msg);
    return hbad ? 1 : 0;
  }
  CK(cudaMemset(bad, 0, 8));
  CK(cudaMemset(maxerr, 0, 4));
  cmp_kernel<<<1024, 256>>>(D1, R, (size_t)M * N, bad, maxerr);
  unsigned long long hbad1;
  CK(cudaMemcpy(&hbad1, bad, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hmax, maxerr, 4, cudaMemcpyDeviceToHost));
  printf("production   mismatches %llu, max abs err %.4f\n", hbad1, hmax);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) gpvb200_gemm(&d, nullptr);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("production   %.2f us / launch, %.1f TFLOP/s\n", 1e3 * ms / reps, flop * reps / (ms * 1e-3) / 1e12);
  return hbad ? 1 : 0;
}
