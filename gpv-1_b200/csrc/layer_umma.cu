// Row-tile-resident transformer sub-layer kernels for sm_100a (tcgen05 + TMA): one CTA owns a 128-token tile of the
// d_model = 256 DETR encoder / decoder stream and keeps it in shared memory across the whole sub-layer, so the hidden
// activation of the feed-forward network never has to be a GEMM operand read back from HBM and bias / ReLU / dropout /
// residual / LayerNorm ride in the epilogues.
//
//   mlp_block_fwd:   y = LN(x + drop(W2 drop_h(relu(W1 x + b1)) + b2))        transformer.py:157-160, 228-231
//
//   attn_block_fwd:  y = LN(x + drop(concat_h softmax(Q_h K_h^T) V_h  Wo^T + bo))   transformer.py:153-157   (second half of this file)
//
// Roles (384 threads): warps 0-2 = TMA producers (one elected lane each), warp 3 = tcgen05.mma issuer (one elected lane), warps
// 4-11 = epilogue / softmax (TMEM lane quarter = warp & 3, column half = (warp - 4) >> 2).
//
// mlp_block_fwd pipeline, per 64-wide chunk j of the hidden dimension (d_ff / 64 chunks):
//   FFN1(j):  acc1[j&1] (TMEM, 64 cols)  = X[128x256] (smem, resident) . W1[64j..64j+63, :]^T        16 MMAs 128x64x16
//   epi(j):   h = drop_h(relu(acc1 + b1)) -> bf16 -> H[j&1] (smem, SWIZZLE_128B K-major: an A operand) (+ global copy
//             for the backward pass)
//   FFN2(j):  acc2 (TMEM, 256 cols)     += H[j&1][128x64] . W2[:, 64j..64j+63]^T                       4 MMAs 128x256x16
// issued as FFN1(0) FFN1(1) FFN2(0) FFN1(2) FFN2(1) ..., so the tensor pipe runs FFN1(j+1) while the epilogue warps turn
// chunk j around.  W1 / W2 chunks (32 KB each) stream through a 4-slot TMA ring in exactly that order.  The final epilogue
// adds b2, the dropout mask and the residual (read back from the resident X tile), computes the LayerNorm statistics of
// the row across the two column halves (shared-memory exchange) and writes y, the pre-norm sum and (mean, rstd).
//
// Algorithmic work per 128-row tile: 4 * 128 * 256 * d_ff FLOP (268 MFLOP at d_ff = 2048) against 2 * 256 * d_ff * 2 B of
// weights from L2 (2 MB) and 128 * (256 * 3 + d_ff) * 2 B of HBM traffic (x in; y, pre, h out).
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

constexpr int kLtProducers = 3;            // TMA producer warps (see gemm_umma.cu: bulk loads of one thread do not overlap)
constexpr int kLtMmaWarp = kLtProducers;
constexpr int kLtEpiWarp0 = kLtProducers + 1;
constexpr int kLtThreads = 32 * (kLtProducers + 1 + 8);
constexpr int kLtEpiThreads = 256;
constexpr int kRingSlots = 4;
constexpr uint32_t kSlotBytes = 32768u;   // one W1 chunk [64 x 256] (4 k-blocks of 8 KB) or one W2 chunk [256 x 64]
constexpr uint32_t kTileBytes = 65536u;   // [128 x 256] bf16 as 4 k-blocks of [128 x 64] (16 KB each), SWIZZLE_128B
constexpr uint32_t kKblkBytes = 16384u;
constexpr uint32_t kHBytes = 16384u;      // one hidden chunk [128 x 64]

struct MlpParams {
  int M, S, tps;        // rows; rows per sequence and 128-row tiles per sequence (flat tiling: S = M)
  int nchunk, dff;
  float eps;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  bf16* y;
  bf16* h;              // [M, dff] hidden activation (post ReLU / dropout) for the backward pass, or null
  bf16* pre;            // [M, 256] pre-LayerNorm sum for the backward pass, or null
  float* stats;         // [M, 2] (mean, rstd) or null
  long long ldy, ldh, ldpre;
  DropArgs drop_h, drop_o;
  long long* trace;     // developer timeline (tools/trace_layer.py): [3 roles][nchunk + 1][8] clock64 stamps of CTA 0, or null
  // backward variant (mlp_block_bwd): the saved hidden activation gates the hidden gradient, the residual-branch gradient is added
  const bf16* hmask;    // [M, dff] h of the forward pass (post ReLU / dropout): dh = alpha * acc1 where h > 0
  const bf16* res;      // [M, 256] gradient of the residual branch, added to the output
  long long ldhm, ldres;
  float alpha;
};

#define LT_STAMP(role, j, slot) do { if (p.trace != nullptr && blockIdx.x == 0) p.trace[((role) * (p.nchunk + 1) + (j)) * 8 + (slot)] = clock64(); } while (0)

GPV_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 64] bf16 SWIZZLE_128B K-major block whose base is 1024-aligned
GPV_DEVINL uint32_t sw128_off(int r, int c) { return (uint32_t)r * 128u + ((uint32_t)(c ^ (r & 7)) << 4); }

template <bool DROP, bool BWD>
__global__ void __launch_bounds__(kLtThreads, 1)
mlp_block_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                       // TMA landing zone of the X tile; once X sits in TMEM: H buffers + output slabs
  uint8_t* sH = sX;                         // 2 x 16 KB
  uint8_t* sSlab = sX + 32768;              // 8 warps x 2 x 2 KB (final epilogue)
  uint8_t* sRing = sX + kTileBytes;
  uint64_t* bars = (uint64_t*)(sRing + kRingSlots * kSlotBytes);
  uint64_t* ring_full = bars;          // [kRingSlots]
  uint64_t* ring_empty = bars + 6;     // [kRingSlots]
  uint64_t* acc1_full = bars + 12;     // [2]
  uint64_t* acc1_empty = bars + 14;    // [2]
  uint64_t* h_full = bars + 16;        // [2]
  uint64_t* h_empty = bars + 18;       // [2]
  uint64_t* x_full = bars + 20;
  uint64_t* xt_full = bars + 21;       // X copied into tensor memory (8 epilogue warps)
  uint64_t* acc2_full = bars + 22;
  uint32_t* tmem_slot = (uint32_t*)(bars + 23);
  float* sVec = reinterpret_cast<float*>(bars + 32);   // b2 | gamma | beta (256 each) | b1 (dff): staged once, read by every epilogue step
  float* sB2 = sVec;
  float* sGamma = sVec + 256;
  float* sBeta = sVec + 512;
  float* sB1 = sVec + 768;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.nchunk;

  const int tile = blockIdx.x;
  const int seq = tile / p.tps, tt = tile % p.tps;
  const int m0 = seq * p.S + tt * 128;
  int rows_valid = p.S - tt * 128;
  rows_valid = rows_valid > 128 ? 128 : rows_valid;
  if (m0 + rows_valid > p.M) rows_valid = p.M - m0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < kRingSlots; ++s) {
      mbar_init(&ring_full[s], 1);
      mbar_init(&ring_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 4);
      mbar_init(&h_full[b], 4);
      mbar_init(&h_empty[b], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(xt_full, 8);
    mbar_init(acc2_full, 1);
    fence_barrier_init();
  }
  if (warp == kLtMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tensor memory: [0,256) FFN2 accumulator, [256,384) two FFN1 accumulators, [384,512) the X tile as packed bf16 pairs
  const uint32_t t_acc2 = tmem_base, t_acc1 = tmem_base + 256u, t_x = tmem_base + 384u;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < kLtProducers) {
    // ================================================================== TMA producers.  Ring item g = W1 chunk / W2 chunk in the
    // order the MMA warp consumes them (W1(0) W1(1) W2(0) W1(2) W2(1) ...); item g belongs to warp g % kLtProducers: the bulk loads
    // issued by one thread complete one after the other (~650 clocks each), loads of different warps overlap.  Every item is ONE
    // 32 KB box: W1 rows [64j, 64j+64) as [4 k-blocks][64 rows][64] through a 3-D map, W2 columns [64j, 64j+64) as [256 rows][64].
    if (elect_one()) {
      if (warp == 0) {
        mbar_expect_tx(x_full, kTileBytes);
        tma_load_4d(sX, &tmX, x_full, 0, m0, 0, 0);           // [4 k-blocks][128 rows][64]: one 64 KB box
      }
      int g = 0;
      for (int j = 0; j <= n; ++j) {
        if (j < n) {
          if (g % kLtProducers == warp) {
            const int s = g % kRingSlots;
            mbar_wait(&ring_empty[s], (((uint32_t)(g / kRingSlots)) & 1u) ^ 1u);
            LT_STAMP(0, j, 0);
            mbar_expect_tx(&ring_full[s], kSlotBytes);
            tma_load_4d(sRing + (size_t)s * kSlotBytes, &tmW1, &ring_full[s], 0, j * 64, 0, 0);
          }
          ++g;
        }
        if (j >= 1) {
          if (g % kLtProducers == warp) {
            const int s = g % kRingSlots;
            mbar_wait(&ring_empty[s], (((uint32_t)(g / kRingSlots)) & 1u) ^ 1u);
            LT_STAMP(0, j - 1, 1);
            mbar_expect_tx(&ring_full[s], kSlotBytes);
            tma_load_4d(sRing + (size_t)s * kSlotBytes, &tmW2, &ring_full[s], (j - 1) * 64, 0, 0, 0);
          }
          ++g;
        }
      }
    }
    __syncwarp();
  } else if (warp == kLtMmaWarp) {
    // ================================================================== MMA issuer
    const uint32_t idesc1 = make_idesc_bf16(128, 64, 0, 0), idesc2 = make_idesc_bf16(128, 256, 0, 0);
    mbar_wait(xt_full, 0);
    tc_fence_after();
    int g = 0;
    for (int j = 0; j <= n; ++j) {
      if (j < n) {
        const int s = g % kRingSlots, b = j & 1;
        mbar_wait(&ring_full[s], ((uint32_t)(g / kRingSlots)) & 1u);
        if (lane == 0) LT_STAMP(1, j, 0);
        mbar_wait(&acc1_empty[b], (((uint32_t)j >> 1) & 1u) ^ 1u);
        if (lane == 0) LT_STAMP(1, j, 1);
        tc_fence_after();
        if (elect_one()) {
          // A = X from tensor memory (8 columns per 16-deep k-step): no shared-memory read for the 4 KB A slice of every step
          const uint64_t bd0 = make_sdesc_sw128(smem_u32(sRing + (size_t)s * kSlotBytes), 0u, 1024u);
          const uint32_t d = t_acc1 + (uint32_t)(b * 64);
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_f16_ts(d, t_x + (uint32_t)(k * 8), bd0 + (uint64_t)(((k >> 2) * 8192 + (k & 3) * 32) >> 4), idesc1, k > 0 ? 1u : 0u);
          umma_commit(&ring_empty[s]);
          umma_commit(&acc1_full[b]);
          LT_STAMP(1, j, 2);
        }
        __syncwarp();
        ++g;
      }
      if (j >= 1) {
        const int jj = j - 1, s = g % kRingSlots, b = jj & 1;
        mbar_wait(&ring_full[s], ((uint32_t)(g / kRingSlots)) & 1u);
        if (lane == 0) LT_STAMP(1, jj, 3);
        mbar_wait(&h_full[b], ((uint32_t)jj >> 1) & 1u);
        if (lane == 0) LT_STAMP(1, jj, 4);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad0 = make_sdesc_sw128(smem_u32(sH + (size_t)b * kHBytes), 0u, 1024u);
          const uint64_t bd0 = make_sdesc_sw128(smem_u32(sRing + (size_t)s * kSlotBytes), 0u, 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(t_acc2, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc2, (jj > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ring_empty[s]);
          umma_commit(&h_empty[b]);
          if (jj == n - 1) umma_commit(acc2_full);
          LT_STAMP(1, jj, 5);
        }
        __syncwarp();
        ++g;
      }
    }
  } else {
    // ================================================================== epilogue warps
    // Main loop: two groups of four warps take alternate hidden chunks (group = chunk parity = FFN1 accumulator = H buffer), so the
    // latency chain of one chunk (accumulator ready -> tcgen05.ld -> bias / ReLU / dropout -> H tile -> proxy fence -> barrier)
    // overlaps the other group's.  A thread owns one tile row (its TMEM lane) and all 64 hidden columns of the chunk.
    // Final epilogue: (lane quarter, column half) as in the GEMM kernel.
    const int e = warp - kLtEpiWarp0;
    const int q = warp & 3, half = e >> 2, grp = e >> 2;
    const int r = q * 32 + lane;                       // tile row of this thread (= its TMEM lane)
    const long long grow = (long long)m0 + r;          // global row
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const uint32_t dkey_h = DROP && p.drop_h.seed ? drop_key(*p.drop_h.seed, p.drop_h.site) : 0u;
    const uint32_t dkey_o = DROP && p.drop_o.seed ? drop_key(*p.drop_o.seed, p.drop_o.site) : 0u;
    // rows this lane stores in the coalesced arrangement: 8jj + (lane >> 2) of the warp's 32-row group, chunk lane & 3
    uint32_t ok = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (q * 32 + 8 * jj + (lane >> 2) < rows_valid) ok |= 1u << jj;
    const long long crow0 = (long long)m0 + q * 32 + (lane >> 2);

    // ---- bias / LayerNorm vectors -> shared memory (one cold L2 round trip here instead of one per epilogue step)
    {
      const int t = threadIdx.x - kLtEpiWarp0 * 32;     // 0..255
      if constexpr (!BWD) {
        if (t < 64) reinterpret_cast<float4*>(sB2)[t] = __ldg(reinterpret_cast<const float4*>(p.b2) + t);
        else if (t < 128) reinterpret_cast<float4*>(sGamma)[t - 64] = __ldg(reinterpret_cast<const float4*>(p.gamma) + t - 64);
        else if (t < 192) reinterpret_cast<float4*>(sBeta)[t - 128] = __ldg(reinterpret_cast<const float4*>(p.beta) + t - 128);
        for (int i = t; i < p.dff / 4; i += kLtEpiThreads) reinterpret_cast<float4*>(sB1)[i] = __ldg(reinterpret_cast<const float4*>(p.b1) + i);
      }
    }
    // ---- X tile: shared memory (TMA, SWIZZLE_128B) -> tensor memory, packed bf16 pairs; this warp copies columns
    // [128 half, 128 half + 128) of its 32 rows = TMEM columns [64 half, 64 half + 64)
    mbar_wait(x_full, 0);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int kb = half * 2 + kk;
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = lds128(smem_u32(sX) + (uint32_t)kb * kKblkBytes + sw128_off(r, i));
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
      }
      tmem_st_32x32b_x32(t_x + t_lane + (uint32_t)(kb * 32), w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(xt_full);     // also: this warp no longer reads sX (the H buffers and slabs reuse it)
    named_bar_sync(1, kLtEpiThreads);        // the staged vectors are visible to every epilogue warp

    for (int j = grp; j < n; j += 2) {
      const int b = grp;
      const uint32_t use = (uint32_t)(j >> 1);           // how many times this group's buffers were used before
      const bool tr = e == 0 && lane == 0;
      if (tr) LT_STAMP(2, j, 0);
      uint4 hm[8];                                       // backward: this row's 64 saved hidden activations of chunk j (in flight during the wait)
      if constexpr (BWD) {
        const bf16* hrow = p.hmask + grow * p.ldhm + j * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) hm[i] = r < rows_valid ? ldg_u4(hrow + 8 * i) : make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&acc1_full[b], use & 1u);
      if (tr) LT_STAMP(2, j, 1);
      tc_fence_after();
      uint32_t acc[2][32];
      tmem_ld_32x32b_x32(t_acc1 + t_lane + (uint32_t)(b * 64), acc[0]);
      tmem_ld_32x32b_x32(t_acc1 + t_lane + (uint32_t)(b * 64 + 32), acc[1]);
      tmem_ld_wait();
      if (tr) LT_STAMP(2, j, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[b]);
      const int nb = j * 64;                             // first hidden column of the chunk
      uint4 pk[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        if constexpr (BWD) {
          // dh = alpha * (dy W2)  where the forward's h is positive (relu' and, in train mode, the hidden dropout mask in one test)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 hq = hm[hh * 4 + i];
            const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float2 hf = unpack_bf16x2(hw[w]);
              v[8 * i + 2 * w] = hf.x > 0.f ? __uint_as_float(acc[hh][8 * i + 2 * w]) * p.alpha : 0.f;
              v[8 * i + 2 * w + 1] = hf.y > 0.f ? __uint_as_float(acc[hh][8 * i + 2 * w + 1]) * p.alpha : 0.f;
            }
          }
        } else {
        const float4* b4 = reinterpret_cast<const float4*>(sB1 + nb + hh * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = b4[i];
          v[4 * i] = fmaxf(__uint_as_float(acc[hh][4 * i]) + bb.x, 0.f);
          v[4 * i + 1] = fmaxf(__uint_as_float(acc[hh][4 * i + 1]) + bb.y, 0.f);
          v[4 * i + 2] = fmaxf(__uint_as_float(acc[hh][4 * i + 2]) + bb.z, 0.f);
          v[4 * i + 3] = fmaxf(__uint_as_float(acc[hh][4 * i + 3]) + bb.w, 0.f);
        }
        }
        if (DROP && p.drop_h.seed) {
          const uint32_t base = (uint32_t)grow * (uint32_t)((p.dff + 1) >> 1) + (uint32_t)((nb + hh * 32) >> 1);
#pragma unroll
          for (int i = 0; i < 16; ++i) drop_pair(v[2 * i], v[2 * i + 1], dkey_h, base + i, p.drop_h.thresh16, p.drop_h.scale);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[hh * 4 + i] = pack8(v + 8 * i);
      }
      if (tr) LT_STAMP(2, j, 3);
      mbar_wait(&h_empty[b], (use & 1u) ^ 1u);           // FFN2(j-2) has finished reading this buffer
      if (tr) LT_STAMP(2, j, 4);
      const uint32_t hb = smem_u32(sH + (size_t)b * kHBytes);
#pragma unroll
      for (int i = 0; i < 8; ++i) sts128(hb + sw128_off(r, i), pk[i]);
      if (tr) LT_STAMP(2, j, 5);
      fence_proxy_async();            // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[b]);
      if (tr) LT_STAMP(2, j, 6);
      if (p.h != nullptr) {           // saved for the backward pass: re-read this warp's 32 rows x 128 bytes coalesced (4 rows per instruction)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int rr = q * 32 + 4 * jj + (lane >> 3);
          const uint4 o = lds128(hb + sw128_off(rr, lane & 7));
          if (rr < rows_valid) *reinterpret_cast<uint4*>(p.h + ((long long)m0 + rr) * p.ldh + nb + (lane & 7) * 8) = o;
        }
      }
    }

    // ---- final epilogue: pre = acc2 + b2 (dropout) + x;  y = LayerNorm(pre)
    // pass 1: pre in fp32, written back over the accumulator (tcgen05.st) and out as bf16; row statistics.
    // pass 2: re-read, normalise, store y.  The TMEM load of the next 32-column slice is in flight while this one is processed.
    if (e == 0 && lane == 0) LT_STAMP(2, n, 0);
    mbar_wait(acc2_full, 0);
    if (e == 0 && lane == 0) LT_STAMP(2, n, 1);
    tc_fence_after();
    const uint32_t slab = smem_u32(sSlab) + (uint32_t)e * 4096u;
    const uint32_t co = 16u * slab_slot(lane >> 2, lane & 3);
    uint32_t own[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) own[i] = 16u * slab_slot(lane, i);
    if constexpr (BWD) {
      // dx = acc2 + gradient of the residual branch (read from global memory, row-per-thread), one pass, no LayerNorm
      const bf16* rrow = p.res + grow * p.ldres + half * 128;
      const bool rv = r < rows_valid;
      uint32_t acc[2][32];
      uint4 xr[2][4];
      tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(half * 128), acc[0]);
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[0][i] = rv ? ldg_u4(rrow + 8 * i) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col0 = half * 128 + c * 32;
        tmem_ld_wait();
        if (c < 3) {
          tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(col0 + 32), acc[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) xr[(c + 1) & 1][i] = rv ? ldg_u4(rrow + (c + 1) * 32 + 8 * i) : make_uint4(0, 0, 0, 0);
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[c & 1][i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) unpack8(xr[c & 1][i], v + 8 * i, true);
        const uint32_t sl = slab + 2048u * (uint32_t)(c & 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(sl + own[i], pack8(v + 8 * i));
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if ((ok >> jj) & 1u)
            *reinterpret_cast<uint4*>(p.y + (crow0 + 8 * jj) * p.ldy + col0 + (lane & 3) * 8) = lds128(sl + co + 512u * jj);
      }
    } else {
    float sum = 0.f, sq = 0.f;
    {
      uint32_t acc[2][32], xr[2][16];
      tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(half * 128), acc[0]);
      tmem_ld_32x32b_x16(t_x + t_lane + (uint32_t)(half * 64), xr[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col0 = half * 128 + c * 32;
        tmem_ld_wait();
        if (c < 3) {
          tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(col0 + 32), acc[(c + 1) & 1]);
          tmem_ld_32x32b_x16(t_x + t_lane + (uint32_t)((col0 + 32) >> 1), xr[(c + 1) & 1]);
        }
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(sB2 + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = b4[i];
          v[4 * i] = __uint_as_float(acc[c & 1][4 * i]) + bb.x;
          v[4 * i + 1] = __uint_as_float(acc[c & 1][4 * i + 1]) + bb.y;
          v[4 * i + 2] = __uint_as_float(acc[c & 1][4 * i + 2]) + bb.z;
          v[4 * i + 3] = __uint_as_float(acc[c & 1][4 * i + 3]) + bb.w;
        }
        if (DROP && p.drop_o.seed) {
          const uint32_t base = (uint32_t)grow * 128u + (uint32_t)(col0 >> 1);
#pragma unroll
          for (int i = 0; i < 16; ++i) drop_pair(v[2 * i], v[2 * i + 1], dkey_o, base + i, p.drop_o.thresh16, p.drop_o.scale);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 f = unpack_bf16x2(xr[c & 1][i]);
          v[2 * i] += f.x;
          v[2 * i + 1] += f.y;
        }
        uint32_t wv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          sum += v[i];
          sq += v[i] * v[i];
          wv[i] = __float_as_uint(v[i]);
        }
        tmem_st_32x32b_x32(t_acc2 + t_lane + (uint32_t)col0, wv);
        if (p.pre != nullptr) {
          const uint32_t sl = slab + 2048u * (uint32_t)(c & 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) sts128(sl + own[i], pack8(v + 8 * i));
          __syncwarp();
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if ((ok >> jj) & 1u)
              *reinterpret_cast<uint4*>(p.pre + (crow0 + 8 * jj) * p.ldpre + col0 + (lane & 3) * 8) = lds128(sl + co + 512u * jj);
        }
      }
    }
    tmem_st_wait();
    float2* part = reinterpret_cast<float2*>(sH);     // [2][128]: the H buffers are idle once acc2 is complete
    part[half * 128 + r] = make_float2(sum, sq);
    named_bar_sync(1, kLtEpiThreads);
    const float2 other = part[(half ^ 1) * 128 + r];
    const float mean = (sum + other.x) * (1.0f / 256.0f);
    const float var = fmaxf((sq + other.y) * (1.0f / 256.0f) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    if (half == 0 && r < rows_valid && p.stats != nullptr) {
      p.stats[grow * 2] = mean;
      p.stats[grow * 2 + 1] = rstd;
    }
    {
      uint32_t acc[2][32];
      tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(half * 128), acc[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col0 = half * 128 + c * 32;
        tmem_ld_wait();
        if (c < 3) tmem_ld_32x32b_x32(t_acc2 + t_lane + (uint32_t)(col0 + 32), acc[(c + 1) & 1]);
        float v[32];
        const float4* g4 = reinterpret_cast<const float4*>(sGamma + col0);
        const float4* e4 = reinterpret_cast<const float4*>(sBeta + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 gg = g4[i], ee = e4[i];
          v[4 * i] = (__uint_as_float(acc[c & 1][4 * i]) - mean) * rstd * gg.x + ee.x;
          v[4 * i + 1] = (__uint_as_float(acc[c & 1][4 * i + 1]) - mean) * rstd * gg.y + ee.y;
          v[4 * i + 2] = (__uint_as_float(acc[c & 1][4 * i + 2]) - mean) * rstd * gg.z + ee.z;
          v[4 * i + 3] = (__uint_as_float(acc[c & 1][4 * i + 3]) - mean) * rstd * gg.w + ee.w;
        }
        const uint32_t sl = slab + 2048u * (uint32_t)(c & 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(sl + own[i], pack8(v + 8 * i));
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if ((ok >> jj) & 1u)
            *reinterpret_cast<uint4*>(p.y + (crow0 + 8 * jj) * p.ldy + col0 + (lane & 3) * 8) = lds128(sl + co + 512u * jj);
      }
    }
    }   // !BWD
    if (e == 0 && lane == 0) LT_STAMP(2, n, 2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLtMmaWarp) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// attn_block_fwd: multi-head attention core on tcgen05 for d_model = 256, 8 heads of 32 (DETR encoder / decoder,
// transformer.py:153-156, 216-227), optionally fused with the output projection, the residual and the LayerNorm:
//     o = concat_h softmax(scale Q_h K_h^T + key mask) V_h          y = LN(x + drop(o Wo^T + bo))
// One CTA owns 128 query rows of one image and walks the 8 heads:
//   S_h  (TMEM [0,320), fp32)  = Q_h K_h^T        SS MMAs, K = 32: A = Q tile (shared memory, resident), B = K tile of the head pair
//   P_h  (TMEM [320,480), packed bf16 pairs)       softmax by the 8 epilogue warps: row = TMEM lane = thread, the two warps of a
//                                                  lane quarter split the keys; row maxima / sums cross through shared memory
//   O_h  (TMEM [480,512))      = P_h V_h           TS MMAs (A = P from tensor memory), B = V tile read MN-major, N = 32
//   O tile (shared memory, over the Q tile: head h's 32 columns replace Q_h, which is dead by then) = O_h / rowsum, bf16
// K / V of a head PAIR ([Sk x 64] each, one 128-byte swizzle row per key) stream through a 2-slot ring; scores never leave the
// chip.  Fused tail: acc (TMEM [0,256)) = O tile . Wo^T (Wo's four k-blocks land in the K/V ring), epilogue as mlp_block_fwd.
// Dropout on the probabilities uses the same counter-based mask as attention.cu (row = (b H + h) Sq + q, col = key), so the
// backward kernel regenerates it.
// =====================================================================================================================
constexpr uint32_t kKvTileBytes = 38912u;     // [304 keys x 64] bf16
constexpr int kPolyNum = 2, kPolyDen = 5;      // of every kPolyDen pairs of probabilities, kPolyNum take exp2 on the FMA pipe (exp2_poly2)
constexpr uint32_t kKvSlotBytes = 2 * kKvTileBytes;

struct AttnBlkParams {
  int B, H, Sq, Sk, Skp, split, tps;     // Skp = Sk rounded up to 16; split = first key of column-half 1 (multiple of 32)
  int nbox, box_rows;                    // a K / V pair tile is loaded as nbox boxes of box_rows keys
  int fuse;                              // 1: + out-proj + residual + LayerNorm
  float sl2, eps;
  const uint8_t* kmask;                  // [B, Sk], 1 = masked key, or null
  const float* bo;
  const float* gamma;
  const float* beta;
  const bf16* xres;
  bf16* o;
  float* lse;
  bf16* pre;
  bf16* y;
  float* stats;
  long long ldx, ldo, ldpre, ldy;
  DropArgs drop_p, drop_o;
  long long* trace;
  int nchunk;                            // (trace layout only: rows of 8 stamps per role = nchunk + 1)
};

template <bool DROP>
__global__ void __launch_bounds__(kLtThreads, 1)
attn_block_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmWo,
                      const __grid_constant__ AttnBlkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                         // Q tile, then O tile (A operand of the output projection)
  uint8_t* sKV = smem + kTileBytes;           // 2 slots x (K pair tile | V pair tile); later Wo (4 x 32 KB), then output slabs
  uint64_t* bars = (uint64_t*)(sKV + 2 * kKvSlotBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;       // [2]
  uint64_t* v_full = bars + 3;       // [2]
  uint64_t* kv_empty = bars + 5;     // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint64_t* o_empty = bars + 10;
  uint64_t* otile_full = bars + 11;
  uint64_t* wo_full = bars + 12;
  uint64_t* acc_full = bars + 13;
  uint64_t* kv_done = bars + 14;     // every MMA that reads the K / V ring has completed (Wo may land there)
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);
  uint64_t* s_empty = bars + 16;     // the softmax warps hold S_h in registers: S_{h+1} may overwrite it
  float* sVec = reinterpret_cast<float*>(bars + 32);   // bo | gamma | beta
  float* sBo = sVec;
  float* sGamma = sVec + 256;
  float* sBeta = sVec + 512;
  float* sMax = sVec + 768;          // [2 head parity][2 halves][128]
  float* sSum = sMax + 512;          // [2][2][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int img = tile / p.tps, tt = tile % p.tps;
  const int m0 = img * p.Sq + tt * 128;                 // first query row (global)
  int rows_valid = p.Sq - tt * 128;
  rows_valid = rows_valid > 128 ? 128 : rows_valid;
  const int k0 = img * p.Sk;                            // first key row (global)
  const int nks = p.Skp >> 4;                           // 16-key steps of P V
  const int n1 = p.Skp > 256 ? 256 : p.Skp, n2 = p.Skp - n1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (p.fuse) tma_prefetch_desc(&tmWo);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], p.nbox);
      mbar_init(&v_full[i], p.nbox);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    mbar_init(otile_full, 8);
    mbar_init(wo_full, 4);
    mbar_init(acc_full, 1);
    mbar_init(kv_done, 1);
    mbar_init(s_empty, 8);
    fence_barrier_init();
  }
  if (warp == kLtMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_p = tmem_base + 320u, t_o = tmem_base + 480u;
  // warp group 0 (three producers + the MMA issuer) needs few registers; the two softmax warp groups keep a whole half row of
  // scores (160 fp32) in registers: 128 x 64 + 256 x 208 = 61440 <= 65536
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
  if (warp < kLtProducers) {
    // ================================================================== TMA producers: operations dealt round-robin to the warps
    if (elect_one()) {
      int op = 0;
      if (op++ % kLtProducers == warp) {
        mbar_expect_tx(q_full, kTileBytes);
        tma_load_4d(sQ, &tmQ, q_full, 0, m0, 0, 0);
      }
      const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;
      for (int g = 0; g < 4; ++g) {
        const int slot = g & 1;
        uint8_t* dk = sKV + (size_t)slot * kKvSlotBytes;
        uint8_t* dv = dk + kKvTileBytes;
        for (int kv = 0; kv < 2; ++kv) {
          for (int i = 0; i < p.nbox; ++i) {
            if (op++ % kLtProducers != warp) continue;
            mbar_wait(&kv_empty[slot], (((uint32_t)g >> 1) & 1u) ^ 1u);
            uint64_t* bar = kv ? &v_full[slot] : &k_full[slot];
            mbar_expect_tx(bar, box_bytes);
            tma_load_4d((kv ? dv : dk) + (size_t)i * box_bytes, kv ? &tmV : &tmK, bar, g * 64, k0 + i * p.box_rows, 0, 0);
          }
        }
      }
      if (p.fuse) {
        for (int kb = 0; kb < 4; ++kb) {
          if (op++ % kLtProducers != warp) continue;
          mbar_wait(kv_done, 0);
          mbar_expect_tx(wo_full, 32768u);
          tma_load_4d(sKV + (size_t)kb * 32768u, &tmWo, wo_full, kb * 64, 0, 0, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == kLtMmaWarp) {
    // ================================================================== MMA issuer
    const uint32_t idesc_s1 = make_idesc_bf16(128, n1, 0, 0), idesc_s2 = make_idesc_bf16(128, n2 > 0 ? n2 : 16, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 32, 0, 1), idesc_out = make_idesc_bf16(128, 256, 0, 0);
    const uint32_t qa = smem_u32(sQ);
    auto issue_s = [&](int h) {       // S = Q_h K_h^T  (two 16-deep steps over d_h = 32)
      const uint32_t kb = smem_u32(sKV + (size_t)((h >> 1) & 1) * kKvSlotBytes) + (uint32_t)(h & 1) * 64u;
      const uint32_t ab = qa + (uint32_t)(h >> 1) * kKblkBytes + (uint32_t)(h & 1) * 64u;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t ad = make_sdesc_sw128(ab + ks * 32u, 0u, 1024u);
        umma_f16(t_s, ad, make_sdesc_sw128(kb + ks * 32u, 0u, 1024u), idesc_s1, ks > 0 ? 1u : 0u);
        if (n2 > 0) umma_f16(t_s + 256u, ad, make_sdesc_sw128(kb + 32768u + ks * 32u, 0u, 1024u), idesc_s2, ks > 0 ? 1u : 0u);
      }
      umma_commit(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
    for (int h = 0; h < 8; ++h) {
      const int slot = (h >> 1) & 1;
      if (lane == 0) LT_STAMP(1, h, 0);
      if (h < 7) {                                               // S_{h+1}: K of its pair has landed, the softmax warps hold S_h in registers
        if (((h + 1) & 1) == 0) mbar_wait(&k_full[((h + 1) >> 1) & 1], ((uint32_t)(h + 1) >> 2) & 1u);
        mbar_wait(s_empty, (uint32_t)h & 1u);
        tc_fence_after();
        if (elect_one()) issue_s(h + 1);
        __syncwarp();
      }
      mbar_wait(p_full, (uint32_t)h & 1u);                       // P_h is in tensor memory
      if (lane == 0) LT_STAMP(1, h, 1);
      if ((h & 1) == 0) mbar_wait(&v_full[slot], ((uint32_t)h >> 2) & 1u);
      mbar_wait(o_empty, ((uint32_t)h & 1u) ^ 1u);               // O_{h-1} has been read out
      if (lane == 0) LT_STAMP(1, h, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vb = smem_u32(sKV + (size_t)slot * kKvSlotBytes + kKvTileBytes) + (uint32_t)(h & 1) * 64u;
        const uint64_t vd = make_sdesc_sw128(vb, 1024u, 1024u);
#pragma unroll
        for (int ks = 0; ks < 19; ++ks)                          // Skp <= 304: at most 19 sixteen-key steps
          if (ks < nks) umma_f16_ts(t_o, t_p + (uint32_t)(ks * 8), vd + (uint64_t)(ks * 128), idesc_pv, ks > 0 ? 1u : 0u);
        umma_commit(o_full);
        LT_STAMP(1, h, 3);
        if (h & 1) umma_commit(&kv_empty[slot]);
        if (h == 7) umma_commit(kv_done);
      }
      __syncwarp();
    }
    if (p.fuse) {
      mbar_wait(otile_full, 0);
      mbar_wait(wo_full, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad0 = make_sdesc_sw128(qa, 0u, 1024u);
        const uint64_t bd0 = make_sdesc_sw128(smem_u32(sKV), 0u, 1024u);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_f16(t_s, ad0 + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4), bd0 + (uint64_t)(((k >> 2) * 32768 + (k & 3) * 32) >> 4), idesc_out,
                   k > 0 ? 1u : 0u);
        umma_commit(acc_full);
      }
      __syncwarp();
    }
  }
  } else {
    // ================================================================== softmax / epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
    const int e = warp - kLtEpiWarp0;
    const int q = warp & 3, half = e >> 2;
    const int r = q * 32 + lane;
    const long long grow = (long long)m0 + r;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const int base_key = half ? p.split : 0;
    const int nkeys = half ? p.Skp - p.split : p.split;
    const int nch = (nkeys + 31) >> 5;                 // <= 5
    const uint32_t dkey_p = DROP && p.drop_p.seed ? drop_key(*p.drop_p.seed, p.drop_p.site) : 0u;
    const uint32_t dkey_o = DROP && p.drop_o.seed ? drop_key(*p.drop_o.seed, p.drop_o.site) : 0u;
    uint32_t ok = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (q * 32 + 8 * jj + (lane >> 2) < rows_valid) ok |= 1u << jj;
    const long long crow0 = (long long)m0 + q * 32 + (lane >> 2);
    if (p.fuse) {
      const int t = threadIdx.x - kLtEpiWarp0 * 32;
      if (t < 64) reinterpret_cast<float4*>(sBo)[t] = __ldg(reinterpret_cast<const float4*>(p.bo) + t);
      else if (t < 128) reinterpret_cast<float4*>(sGamma)[t - 64] = __ldg(reinterpret_cast<const float4*>(p.gamma) + t - 64);
      else if (t < 192) reinterpret_cast<float4*>(sBeta)[t - 128] = __ldg(reinterpret_cast<const float4*>(p.beta) + t - 128);
    }
    // keys this thread's chunks may use: inside [0, Sk) and not masked (same for every head)
    uint32_t vmask[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      uint32_t m = 0;
      if (c < nch) {
        for (int i = 0; i < 32; ++i) {
          const int key = base_key + c * 32 + i;
          bool v = key < p.Sk && (c * 32 + i) < nkeys;
          if (v && p.kmask != nullptr) v = p.kmask[(long long)img * p.Sk + key] == 0;
          m |= (uint32_t)v << i;
        }
      }
      vmask[c] = m;
    }
    const uint32_t hs = (uint32_t)((p.Sk + 1) >> 1);
    float m_prev = 0.f;

    // O_h (32 fp32 columns; this warp takes 16) / rowsum -> bf16 -> O tile in shared memory (where Q_h was); lse
    auto o_epilogue = [&](int hh, float m_h) {
      mbar_wait(o_full, (uint32_t)hh & 1u);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld_32x32b_x16(t_o + t_lane + (uint32_t)(half * 16), acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const float l = sSum[((hh & 1) * 2 + 0) * 128 + r] + sSum[((hh & 1) * 2 + 1) * 128 + r];
      const float inv = l > 0.f ? (DROP && p.drop_p.seed ? p.drop_p.scale : 1.0f) / l : 0.f;
      uint4 w0, w1;
      w0.x = pack_bf16x2(__uint_as_float(acc[0]) * inv, __uint_as_float(acc[1]) * inv);
      w0.y = pack_bf16x2(__uint_as_float(acc[2]) * inv, __uint_as_float(acc[3]) * inv);
      w0.z = pack_bf16x2(__uint_as_float(acc[4]) * inv, __uint_as_float(acc[5]) * inv);
      w0.w = pack_bf16x2(__uint_as_float(acc[6]) * inv, __uint_as_float(acc[7]) * inv);
      w1.x = pack_bf16x2(__uint_as_float(acc[8]) * inv, __uint_as_float(acc[9]) * inv);
      w1.y = pack_bf16x2(__uint_as_float(acc[10]) * inv, __uint_as_float(acc[11]) * inv);
      w1.z = pack_bf16x2(__uint_as_float(acc[12]) * inv, __uint_as_float(acc[13]) * inv);
      w1.w = pack_bf16x2(__uint_as_float(acc[14]) * inv, __uint_as_float(acc[15]) * inv);
      const uint32_t ob = smem_u32(sQ) + (uint32_t)(hh >> 1) * kKblkBytes;
      sts128(ob + sw128_off(r, (hh & 1) * 4 + half * 2), w0);
      sts128(ob + sw128_off(r, (hh & 1) * 4 + half * 2 + 1), w1);
      if (half == 0 && r < rows_valid && p.lse != nullptr)
        p.lse[((long long)img * p.H + hh) * p.Sq + tt * 128 + r] = m_h * p.sl2 + log2f(l);
    };

    bool full_chunk[5];                                  // every key of the chunk is live: no select in the inner loops (warp-uniform)
#pragma unroll
    for (int c = 0; c < 5; ++c) full_chunk[c] = vmask[c] == 0xffffffffu;
    for (int h = 0; h < 8; ++h) {
      const bool tr = e == 0 && lane == 0;
      if (tr) LT_STAMP(2, h, 0);
      mbar_wait(s_full, (uint32_t)h & 1u);
      if (tr) LT_STAMP(2, h, 1);
      tc_fence_after();
      // ---- this warp's half row of S_h into registers (one round trip), then S may be overwritten by head h + 1
      uint32_t sc[5][32];
#pragma unroll
      for (int c = 0; c < 5; ++c)
        if (c < nch) tmem_ld_32x32b_x32(t_s + t_lane + (uint32_t)(base_key + c * 32), sc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      // ---- dead keys (past Sk, or masked) become -inf here, so that both passes below are single straight-line paths: the whole
      //      per-head body must stay small enough for the instruction cache (the first version, with a masked and an unmasked
      //      copy of each pass, was 120 KB of code and spent 30 % of its issue slots waiting for instruction fetch)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch && !full_chunk[c]) {
          const uint32_t vm = vmask[c];
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[c][i] = ((vm >> i) & 1u) ? sc[c][i] : 0xff800000u;
        }
      }
      // ---- pass 1: row maximum over this warp's keys
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sc[c][i]));
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (tr) LT_STAMP(2, h, 2);
      sMax[((h & 1) * 2 + half) * 128 + r] = mx;
      named_bar_sync(1, kLtEpiThreads);
      if (tr) LT_STAMP(2, h, 3);
      const float m = fmaxf(mx, sMax[((h & 1) * 2 + (half ^ 1)) * 128 + r]);
      const float ms = (m == -INFINITY) ? 0.f : m;
      if (tr) LT_STAMP(2, h, 4);
      // ---- pass 2: p = exp2((s - m) * scale * log2 e), row sum over every key, dropout on what goes into P V
      uint64_t sum2[2] = {0ull, 0ull};
      const uint32_t prow = ((uint32_t)(img * p.H + h) * (uint32_t)p.Sq + (uint32_t)(tt * 128 + r)) * hs;
      const float nms = -ms * p.sl2;
      const uint64_t sl2x2 = pk2(p.sl2, p.sl2), nmsx2 = pk2(nms, nms);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t0, t1, p0, p1;
            upk2(ffma2(pk2(__uint_as_float(sc[c][2 * j]), __uint_as_float(sc[c][2 * j + 1])), sl2x2, nmsx2), t0, t1);
            if ((j % kPolyDen) < kPolyNum) {
              exp2_poly2(t0, t1, p0, p1);               // a dead key gives 2^-125 here instead of 0: below fp32 resolution of the row sum
            } else {
              p0 = ex2_approx(t0);
              p1 = ex2_approx(t1);
            }
            sc[c][2 * j] = __float_as_uint(p0);
            sc[c][2 * j + 1] = __float_as_uint(p1);
            sum2[j & 1] = fadd2(sum2[j & 1], pk2(p0, p1));
          }
        }
      }
      // P V of head h-1 has had passes 1 and 2 of this head to finish: O_{h-1} out, after which P may be overwritten
      if (h > 0) o_epilogue(h - 1, m_prev);
      m_prev = ms;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch) {
          if (DROP && p.drop_p.seed) {
            const uint32_t cb = prow + (uint32_t)((base_key + c * 32) >> 1);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a = __uint_as_float(sc[c][2 * i]), b = __uint_as_float(sc[c][2 * i + 1]);
              drop_pair(a, b, dkey_p, cb + i, p.drop_p.thresh16, 1.0f);       // 1 / (1 - p) rides on the normalisation of O
              sc[c][2 * i] = __float_as_uint(a);
              sc[c][2 * i + 1] = __float_as_uint(b);
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(sc[c][2 * i]), __uint_as_float(sc[c][2 * i + 1]));
          tmem_st_32x32b_x16(t_p + t_lane + (uint32_t)((base_key + c * 32) >> 1), pk);
        }
      }
      if (tr) LT_STAMP(2, h, 5);
      {
        float a0, a1, b0, b1;
        upk2(sum2[0], a0, a1);
        upk2(sum2[1], b0, b1);
        // a row without any live key (m = -inf) sums to exactly 0, as in the mma.sync kernel: its output row is 0
        sSum[((h & 1) * 2 + half) * 128 + r] = (m == -INFINITY) ? 0.f : (a0 + a1) + (b0 + b1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr) LT_STAMP(2, h, 6);
    }
    named_bar_sync(1, kLtEpiThreads);                // the row sums of head 7 are visible
    o_epilogue(7, m_prev);
    fence_proxy_async();                             // O tile (generic-proxy stores) -> visible to the output projection's MMAs
    named_bar_sync(1, kLtEpiThreads);
    if (lane == 0) mbar_arrive(otile_full);
    // ---- attention output for the backward pass: this warp's 32 rows x 2 k-blocks, 4 rows x 128 bytes per instruction
    if (p.o != nullptr) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int kb = half * 2 + kk;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int rr = q * 32 + 4 * jj + (lane >> 3);
          const uint4 v = lds128(smem_u32(sQ) + (uint32_t)kb * kKblkBytes + sw128_off(rr, lane & 7));
          if (rr < rows_valid) *reinterpret_cast<uint4*>(p.o + ((long long)m0 + rr) * p.ldo + kb * 64 + (lane & 7) * 8) = v;
        }
      }
    }
    if (p.fuse) {
      // ---- pre = acc + bo (dropout) + x;  y = LayerNorm(pre): as mlp_block_fwd's final epilogue, residual read from global memory
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t slab = smem_u32(sKV) + (uint32_t)e * 4096u;      // Wo is dead once acc is complete
      const uint32_t co = 16u * slab_slot(lane >> 2, lane & 3);
      uint32_t own[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) own[i] = 16u * slab_slot(lane, i);
      const bf16* xrow = p.xres + grow * p.ldx + half * 128;
      const bool rv = r < rows_valid;
      float sum = 0.f, sq = 0.f;
      {
        uint32_t acc[2][32];
        uint4 xr[2][4];
        tmem_ld_32x32b_x32(t_s + t_lane + (uint32_t)(half * 128), acc[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) xr[0][i] = rv ? ldg_u4(xrow + 8 * i) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = half * 128 + c * 32;
          tmem_ld_wait();
          if (c < 3) {
            tmem_ld_32x32b_x32(t_s + t_lane + (uint32_t)(col0 + 32), acc[(c + 1) & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) xr[(c + 1) & 1][i] = rv ? ldg_u4(xrow + (c + 1) * 32 + 8 * i) : make_uint4(0, 0, 0, 0);
          }
          float v[32];
          const float4* b4 = reinterpret_cast<const float4*>(sBo + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            v[4 * i] = __uint_as_float(acc[c & 1][4 * i]) + bb.x;
            v[4 * i + 1] = __uint_as_float(acc[c & 1][4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(acc[c & 1][4 * i + 2]) + bb.z;
            v[4 * i + 3] = __uint_as_float(acc[c & 1][4 * i + 3]) + bb.w;
          }
          if (DROP && p.drop_o.seed) {
            const uint32_t base = (uint32_t)grow * 128u + (uint32_t)(col0 >> 1);
#pragma unroll
            for (int i = 0; i < 16; ++i) drop_pair(v[2 * i], v[2 * i + 1], dkey_o, base + i, p.drop_o.thresh16, p.drop_o.scale);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) unpack8(xr[c & 1][i], v + 8 * i, true);
          uint32_t wv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sum += v[i];
            sq += v[i] * v[i];
            wv[i] = __float_as_uint(v[i]);
          }
          tmem_st_32x32b_x32(t_s + t_lane + (uint32_t)col0, wv);
          if (p.pre != nullptr) {
            const uint32_t sl = slab + 2048u * (uint32_t)(c & 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) sts128(sl + own[i], pack8(v + 8 * i));
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if ((ok >> jj) & 1u)
                *reinterpret_cast<uint4*>(p.pre + (crow0 + 8 * jj) * p.ldpre + col0 + (lane & 3) * 8) = lds128(sl + co + 512u * jj);
          }
        }
      }
      tmem_st_wait();
      float2* part = reinterpret_cast<float2*>(sMax);    // [2][128] float2 = 2 KB: the softmax exchange arrays are idle
      part[half * 128 + r] = make_float2(sum, sq);
      named_bar_sync(1, kLtEpiThreads);
      const float2 other = part[(half ^ 1) * 128 + r];
      const float mean = (sum + other.x) * (1.0f / 256.0f);
      const float var = fmaxf((sq + other.y) * (1.0f / 256.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      if (half == 0 && rv && p.stats != nullptr) {
        p.stats[grow * 2] = mean;
        p.stats[grow * 2 + 1] = rstd;
      }
      {
        uint32_t acc[2][32];
        tmem_ld_32x32b_x32(t_s + t_lane + (uint32_t)(half * 128), acc[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = half * 128 + c * 32;
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32b_x32(t_s + t_lane + (uint32_t)(col0 + 32), acc[(c + 1) & 1]);
          float v[32];
          const float4* g4 = reinterpret_cast<const float4*>(sGamma + col0);
          const float4* e4 = reinterpret_cast<const float4*>(sBeta + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 gg = g4[i], ee = e4[i];
            v[4 * i] = (__uint_as_float(acc[c & 1][4 * i]) - mean) * rstd * gg.x + ee.x;
            v[4 * i + 1] = (__uint_as_float(acc[c & 1][4 * i + 1]) - mean) * rstd * gg.y + ee.y;
            v[4 * i + 2] = (__uint_as_float(acc[c & 1][4 * i + 2]) - mean) * rstd * gg.z + ee.z;
            v[4 * i + 3] = (__uint_as_float(acc[c & 1][4 * i + 3]) - mean) * rstd * gg.w + ee.w;
          }
          const uint32_t sl = slab + 2048u * (uint32_t)(c & 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) sts128(sl + own[i], pack8(v + 8 * i));
          __syncwarp();
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if ((ok >> jj) & 1u)
              *reinterpret_cast<uint4*>(p.y + (crow0 + 8 * jj) * p.ldy + col0 + (lane & 3) * 8) = lds128(sl + co + 512u * jj);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLtMmaWarp) tmem_dealloc(tmem_base, 512);
}

// Launch with programmatic dependent launch allowed (GPVB200_PDL=0 turns it off), opting in to the full shared memory.
template <typename K, typename... A>
static int launch_layer(K kern, const char* what, int grid, size_t smem, cudaStream_t st, A... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("%s: cudaFuncSetAttribute(%zu bytes of shared memory) failed: %s", what, smem, cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kLtThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) {
    set_last_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  return check_launch(what);
}

static int map2d(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  const uint64_t dims[4] = {cols, rows, 1, 1};
  const uint64_t str[3] = {ld, ld * rows, ld * rows};
  const uint32_t box[4] = {box_cols, box_rows, 1, 1};
  const uint32_t one4[4] = {1, 1, 1, 1};
  return make_map(m, ptr, dims, str, box, one4);
}

static void fill_drop(DropArgs* d, const void* seed, uint32_t site, float p) {
  d->seed = nullptr;
  d->site = 0;
  d->thresh16 = 0;
  d->scale = 1.f;
  if (seed != nullptr && p > 0.f) {
    d->seed = (const unsigned long long*)seed;
    d->site = site;
    d->thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
    d->scale = 1.0f / (1.0f - p);
  }
}

}  // namespace gpv

using namespace gpv;

static long long* g_layer_trace = nullptr;
/* developer hook (tools/trace_layer.py): device buffer that CTA 0 of the next layer kernels fills with clock64 stamps; NULL = off */
extern "C" int gpvb200_layer_trace(void* buf) {
  g_layer_trace = (long long*)buf;
  return GPV_OK;
}

extern "C" int gpvb200_mlp_block_fwd(const void* x, int64_t ldx, const void* w1, int64_t ldw1, const float* b1, const void* w2,
                                     int64_t ldw2, const float* b2, const float* gamma, const float* beta, float eps, void* y,
                                     int64_t ldy, void* h, int64_t ldh, void* pre, int64_t ldpre, float* stats, int64_t M,
                                     int32_t d_model, int32_t d_ff, int32_t seq_len, const void* drop_seed, uint32_t site_h,
                                     float p_h, uint32_t site_o, float p_o, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && w1 && b1 && w2 && b2 && gamma && beta && y, "mlp_block_fwd: null operand");
  GPV_REQUIRE(d_model == 256, "mlp_block_fwd: d_model must be 256 (got %d)", d_model);
  GPV_REQUIRE(d_ff >= 64 && d_ff % 64 == 0, "mlp_block_fwd: d_ff must be a multiple of 64 (got %d)", d_ff);
  GPV_REQUIRE(M > 0 && M < (1ll << 31), "mlp_block_fwd: bad M");
  GPV_REQUIRE((ldy & 7) == 0 && (ldh & 7) == 0 && (ldpre & 7) == 0, "mlp_block_fwd: output row strides must be multiples of 8");
  GPV_REQUIRE((((uintptr_t)y | (uintptr_t)h | (uintptr_t)pre | (uintptr_t)b1 | (uintptr_t)b2 | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
              "mlp_block_fwd: outputs / vectors must be 16-byte aligned");
  GPV_REQUIRE(p_h < 1.f && p_o < 1.f, "mlp_block_fwd: dropout p must be < 1");
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M;
  p.S = seq_len > 0 ? seq_len : (int)M;
  GPV_REQUIRE(p.M % p.S == 0, "mlp_block_fwd: M (%d) is not a multiple of seq_len (%d)", p.M, p.S);
  p.tps = (p.S + 127) / 128;
  p.dff = d_ff;
  p.nchunk = d_ff / 64;
  p.eps = eps;
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta;
  p.y = (bf16*)y; p.h = (bf16*)h; p.pre = (bf16*)pre; p.stats = stats;
  p.ldy = ldy; p.ldh = ldh; p.ldpre = ldpre;
  fill_drop(&p.drop_h, drop_seed, site_h, p_h);
  fill_drop(&p.drop_o, drop_seed, site_o, p_o);
  p.trace = g_layer_trace;
  GPV_REQUIRE(d_ff <= 4096, "mlp_block_fwd: d_ff > 4096 does not fit the shared-memory bias stage");
  CUtensorMap mx, m1, m2;
  {
    // X tile and W1 chunk as ONE box each: [4 k-blocks][rows][64] (k-block stride 64 elements), i.e. exactly the SWIZZLE_128B
    // K-major operand layout of four 64-deep k-blocks
    const uint32_t one4[4] = {1, 1, 1, 1};
    const uint64_t dx[4] = {64, (uint64_t)M, 4, 1}, sx[3] = {(uint64_t)ldx, 64, 256};
    const uint32_t bx[4] = {64, 128, 4, 1};
    if ((rc = make_map(&mx, x, dx, sx, bx, one4))) return rc;
    const uint64_t d1[4] = {64, (uint64_t)d_ff, 4, 1}, s1[3] = {(uint64_t)ldw1, 64, 256};
    const uint32_t b1x[4] = {64, 64, 4, 1};
    if ((rc = make_map(&m1, w1, d1, s1, b1x, one4))) return rc;
  }
  if ((rc = map2d(&m2, w2, (uint64_t)d_ff, 256, (uint64_t)ldw2, 64, 256))) return rc;
  const int grid = (p.M / p.S) * p.tps;
  const size_t smem = 1024 + kTileBytes + kRingSlots * kSlotBytes + 256 + (768 + (size_t)d_ff) * 4;
  const bool drop = p.drop_h.seed != nullptr || p.drop_o.seed != nullptr;
  if (drop) return launch_layer(mlp_block_fwd_kernel<true, false>, "mlp_block_fwd", grid, smem, (cudaStream_t)stream, mx, m1, m2, p);
  return launch_layer(mlp_block_fwd_kernel<false, false>, "mlp_block_fwd", grid, smem, (cudaStream_t)stream, mx, m1, m2, p);
}

// Data gradients of the same sub-layer in one launch (the mirror of mlp_block_fwd with the transposed weights as K-major operands):
//     dh = alpha * (dy W2) (*) [h > 0]          dx = dh W1 + dres
// dy [M, 256] = gradient of the sub-layer output before the residual add (LayerNorm backward's masked output), w2t = W2^T [d_ff, 256],
// w1t = W1^T [256, d_ff], h = the forward's hidden activation, dres = gradient of the residual branch.  dh is written for the two
// weight-gradient GEMMs (dW2 = dy^T h, dW1 = dh^T x), which stay separate launches beside the chain.
extern "C" int gpvb200_mlp_block_bwd(const void* dy, int64_t lddy, const void* w2t, int64_t ldw2t, const void* w1t, int64_t ldw1t,
                                     const void* h, int64_t ldh, float alpha, const void* dres, int64_t lddres, void* dh, int64_t lddh,
                                     void* dx, int64_t lddx, int64_t M, int32_t d_model, int32_t d_ff, int32_t seq_len, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(dy && w2t && w1t && h && dres && dh && dx, "mlp_block_bwd: null operand");
  GPV_REQUIRE(d_model == 256, "mlp_block_bwd: d_model must be 256 (got %d)", d_model);
  GPV_REQUIRE(d_ff >= 64 && d_ff % 64 == 0 && d_ff <= 4096, "mlp_block_bwd: d_ff must be a multiple of 64, <= 4096 (got %d)", d_ff);
  GPV_REQUIRE(M > 0 && M < (1ll << 31), "mlp_block_bwd: bad M");
  GPV_REQUIRE((lddh & 7) == 0 && (lddx & 7) == 0 && (ldh & 7) == 0 && (lddres & 7) == 0, "mlp_block_bwd: row strides must be multiples of 8");
  GPV_REQUIRE((((uintptr_t)dh | (uintptr_t)dx | (uintptr_t)h | (uintptr_t)dres) & 15) == 0, "mlp_block_bwd: buffers must be 16-byte aligned");
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M;
  p.S = seq_len > 0 ? seq_len : (int)M;
  GPV_REQUIRE(p.M % p.S == 0, "mlp_block_bwd: M (%d) is not a multiple of seq_len (%d)", p.M, p.S);
  p.tps = (p.S + 127) / 128;
  p.dff = d_ff;
  p.nchunk = d_ff / 64;
  p.y = (bf16*)dx; p.h = (bf16*)dh;
  p.ldy = lddx; p.ldh = lddh;
  p.hmask = (const bf16*)h; p.ldhm = ldh;
  p.res = (const bf16*)dres; p.ldres = lddres;
  p.alpha = alpha;
  p.trace = g_layer_trace;
  CUtensorMap mx, m1, m2;
  {
    const uint32_t one4[4] = {1, 1, 1, 1};
    const uint64_t dx4[4] = {64, (uint64_t)M, 4, 1}, sx[3] = {(uint64_t)lddy, 64, 256};
    const uint32_t bx[4] = {64, 128, 4, 1};
    if ((rc = make_map(&mx, dy, dx4, sx, bx, one4))) return rc;
    const uint64_t d1[4] = {64, (uint64_t)d_ff, 4, 1}, s1[3] = {(uint64_t)ldw2t, 64, 256};     // W2^T plays W1's role: [d_ff rows, 256]
    const uint32_t b1x[4] = {64, 64, 4, 1};
    if ((rc = make_map(&m1, w2t, d1, s1, b1x, one4))) return rc;
  }
  if ((rc = map2d(&m2, w1t, (uint64_t)d_ff, 256, (uint64_t)ldw1t, 64, 256))) return rc;          // W1^T plays W2's role: [256 rows, d_ff]
  const int grid = (p.M / p.S) * p.tps;
  const size_t smem = 1024 + kTileBytes + kRingSlots * kSlotBytes + 256 + (768 + (size_t)d_ff) * 4;
  return launch_layer(mlp_block_fwd_kernel<false, true>, "mlp_block_bwd", grid, smem, (cudaStream_t)stream, mx, m1, m2, p);
}

extern "C" int gpvb200_attn_block_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      const uint8_t* key_mask, int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh, float scale,
                                      void* o, int64_t ldo, float* lse, const void* wo, int64_t ldwo, const float* bo, const void* x,
                                      int64_t ldx, const float* gamma, const float* beta, float eps, void* pre, int64_t ldpre, void* y,
                                      int64_t ldy, float* stats, const void* drop_seed, uint32_t site_p, float p_p, uint32_t site_o,
                                      float p_o, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(q && k && v, "attn_block_fwd: null operand");
  GPV_REQUIRE(H == 8 && dh == 32, "attn_block_fwd: built for 8 heads of 32 (got %d x %d)", H, dh);
  GPV_REQUIRE(B > 0 && Sq > 0 && Sk > 0 && Sk <= 304, "attn_block_fwd: needs 1 <= Sk <= 304 (got Sq %d, Sk %d)", Sq, Sk);
  GPV_REQUIRE(o != nullptr || wo != nullptr, "attn_block_fwd: nothing to write");
  GPV_REQUIRE((ldo & 7) == 0 && (ldpre & 7) == 0 && (ldy & 7) == 0 && (ldx & 7) == 0, "attn_block_fwd: row strides must be multiples of 8");
  GPV_REQUIRE((((uintptr_t)o | (uintptr_t)pre | (uintptr_t)y | (uintptr_t)x | (uintptr_t)bo | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
              "attn_block_fwd: outputs / vectors must be 16-byte aligned");
  GPV_REQUIRE(p_p < 1.f && p_o < 1.f, "attn_block_fwd: dropout p must be < 1");
  AttnBlkParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk;
  p.Skp = (Sk + 15) & ~15;
  p.split = ((p.Skp / 2) + 31) & ~31;
  if (p.split > p.Skp) p.split = p.Skp;
  p.tps = (Sq + 127) / 128;
  if (p.Skp <= 256) { p.nbox = 1; p.box_rows = p.Skp; } else { p.nbox = 2; p.box_rows = p.Skp / 2; }
  p.fuse = wo != nullptr;
  if (p.fuse) GPV_REQUIRE(bo && x && gamma && beta && y, "attn_block_fwd: the fused tail needs bo, x, gamma, beta and y");
  p.sl2 = scale * 1.4426950408889634f;
  p.eps = eps;
  p.kmask = key_mask;
  p.bo = bo; p.gamma = gamma; p.beta = beta;
  p.xres = (const bf16*)x; p.o = (bf16*)o; p.lse = lse; p.pre = (bf16*)pre; p.y = (bf16*)y; p.stats = stats;
  p.ldx = ldx; p.ldo = ldo; p.ldpre = ldpre; p.ldy = ldy;
  fill_drop(&p.drop_p, drop_seed, site_p, p_p);
  fill_drop(&p.drop_o, drop_seed, site_o, p_o);
  p.trace = g_layer_trace;
  p.nchunk = 8;
  CUtensorMap mq, mk, mv, mw;
  {
    const uint32_t one4[4] = {1, 1, 1, 1};
    const uint64_t dq[4] = {64, (uint64_t)B * Sq, 4, 1}, sq[3] = {(uint64_t)ldq, 64, 256};
    const uint32_t bq[4] = {64, 128, 4, 1};
    if ((rc = make_map(&mq, q, dq, sq, bq, one4))) return rc;
  }
  if ((rc = map2d(&mk, k, 256, (uint64_t)B * Sk, (uint64_t)ldk, 64, (uint32_t)p.box_rows))) return rc;
  if ((rc = map2d(&mv, v, 256, (uint64_t)B * Sk, (uint64_t)ldv, 64, (uint32_t)p.box_rows))) return rc;
  if (p.fuse) {
    if ((rc = map2d(&mw, wo, 256, 256, (uint64_t)ldwo, 64, 256))) return rc;
  } else {
    mw = mk;
  }
  const int grid = B * p.tps;
  const size_t smem = 1024 + kTileBytes + 2 * kKvSlotBytes + 256 + (768 + 512 + 512) * 4;
  const bool drop = p.drop_p.seed != nullptr || p.drop_o.seed != nullptr;
  if (drop) return launch_layer(attn_block_fwd_kernel<true>, "attn_block_fwd", grid, smem, (cudaStream_t)stream, mq, mk, mv, mw, p);
  return launch_layer(attn_block_fwd_kernel<false>, "attn_block_fwd", grid, smem, (cudaStream_t)stream, mq, mk, mv, mw, p);
}
