// tcgen05 tensor-core contraction kernel for sm_100a: plain / batched GEMM, NHWC implicit-GEMM convolution
// (forward and data-gradient) and convolution weight-gradient, all through one warp-specialised pipeline:
//
//   warp 0 (one lane)  TMA producer: 4-D tiled cp.async.bulk.tensor loads, SWIZZLE_128B, zero-filled halo
//   warp 1 (one lane)  tcgen05.mma issuer (128 x BN x 16 bf16 -> fp32 in TMEM), tcgen05.commit -> mbarriers
//   warps 2-9          epilogue: tcgen05.ld TMEM -> registers -> bias/residual/activation/mask -> global
// The kernel is persistent (one CTA per SM walks the tile list) with two TMEM accumulator buffers, so the epilogue
// of one tile (HBM-bound for the K = 64..256 1x1 convolutions) overlaps the TMA/MMA main loop of the next; epilogue
// warps prefetch the residual / mask rows of the next 32-column slice while they process the current one.
//
// Pair variant (umma_gemm_kernel<BN, F, true>, `cta_group::2`): the two SMs of a TPC form a cluster of two CTAs that own
// TWO 128-row tiles with the same columns.  Each CTA stages its own A tile but only ITS half of the B tile (32 KB per
// 64-deep k-block instead of 48 KB: the main loop of the single-CTA kernel is bound by the chip's L2 -> SM throughput);
// the leader issues one M = 256 MMA per k-step that reads both halves of B through the pair's shared memory and writes
// each CTA's 128 accumulator rows into that CTA's TMEM; stages are 128 deep.  Measured on the prototype
// (tools/proto/gemm_2cta.cu, profiles/r1H_summary.md): 1.12-1.33x on the deep-K shapes of the trunk.
//
// Replaces the cuDNN / cuBLAS calls behind every nn.Conv2d / nn.Linear of the reference hot path
// (backbone.py:72, detr_roi_head.py:79-84, transformer.py:153-160,218-231, vilbert.py:748-761,847-898,
//  gpv.py:140,145,162, answer_head.py:31-33) and their autograd backward.
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>
#include <string>

#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

// Division by a run-time constant as multiply-high + shift (dividend < 2^31).  Every role of the persistent kernel decodes every work
// item; with plain `/` and `%` that was ~600-800 clocks of dependent arithmetic per item and role (clock64 timeline,
// profiles/r4a_trace_gemm.txt) -- more than the whole main loop of a 128 x 64 x 64 item.
struct FastDiv {
  uint32_t d, mul, shr;
};
GPV_DEVINL int fdiv(int n, const FastDiv& f) { return f.d == 1u ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr); }
static FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)(d < 1 ? 1 : d);
  f.mul = 0;
  f.shr = 0;
  if (f.d > 1) {
    int l = 0;
    while ((1ull << l) < f.d) ++l;                       // smallest l with 2^l >= d
    const int p = 31 + l;
    f.mul = (uint32_t)(((1ull << p) + f.d - 1) / f.d);   // ceil(2^p / d) fits 32 bits because 2^(l-1) < d
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}

struct KParams {
  int mode, M, N, a_mn, b_mn, bk, k_iters, splits, nstages, b_batched;
  int m_tiles, n_tiles, gy, total_work, k_per_split;
  int m_tiles_cta;  // 128-row tiles of the problem; m_tiles counts work-item rows (pairs of tiles in the pair variant)
  int kblk, kb_total;  // 64-deep k-blocks per pipeline stage of a K-major operand (1; 2 in the pair variant) and in the whole contraction
  int kc_per_tap;
  FastDiv fd_nt, fd_mt, fd_gy, fd_tpi, fd_tws, fd_tw, fd_kc;   // n_tiles, m_tiles, gy, tiles_h * tiles_w, tiles_w, tw, kc_per_tap
  int Ho, Wo, th, tw, tiles_h, tiles_w, stride;
  int ntaps;
  int tap_dh[9], tap_dw[9], tap_w[9];
  int OH, OW, os, ooh, oow;
  int act, aux_mode, d_fp32, d_atomic, vec_ok, res_fp32, coal, pf_mode, max_ctas;
  uint32_t epi_warp_bytes, epi_aux_off, epi_slot_stride;   // per-warp epilogue slabs (coalesced path)
  int b_res;      // 1: the whole K-major B operand of the (single) column tile stays in shared memory, stages carry A only
  int nprod;      // producer warps in use (1..kProducers; GPVB200_PRODUCERS, default kProducers)
  int drop_mode;  // 0 none, 1 before the residual add, 2 after the activation (DropArgs below)
  DropArgs drop;
  int epi_full;   // 1: every slab of a work item is requested up front (one slot per chunk); 0: two-slot ring, one slab ahead
  float alpha;
  void* D;
  bf16* D2;
  const float* bias;
  const float* rowscale;
  const bf16* residual;
  const bf16* aux;
  long long ldd, ldr, ldaux, d_batch_stride;
};

constexpr int kEpiWarps = 8;
// TMA producer warps.  Measured on B200 (tools/proto/mma_rate.cu, profiles/r2_tma_issue.md): the bulk-tensor loads issued by ONE
// thread complete one after the other, ~650 clocks each whatever the box size (8-32 KB) and however many stages are free, while
// loads issued from different warps overlap (2 warps: 330 clocks per box, 4 warps: 175).  A single producer thread therefore caps a
// 128 x 256 x 64 stage (two boxes) at ~1300 clocks against 512 clocks of MMA time.  Stages are dealt round-robin to kProducers
// warps (one lane each); 12 warps cost the same registers as 10 (allocation is per 128 threads).
constexpr int kProducers = 3;
constexpr int kMmaWarp = kProducers;
constexpr int kEpiWarp0 = kProducers + 1;
constexpr int kThreads = 32 * (kProducers + 1 + kEpiWarps);
// Narrow tiles (BN = 64) move 24 KB per stage in two small boxes and are bound by the issue side, not by bandwidth: three producer
// warps deliver one stage per ~850 clocks (28 B/clk per SM; profiles/r4a_trace_gemm.txt) against 128 clocks of MMA.  Their kernels
// are launched with four more producer warps behind the epilogue warps (warps 12-15; 512 threads still leave 128 registers each).
constexpr int kExtraProducers = 4;
template <int BN>
constexpr int kProdWarps = BN == 64 ? kProducers + kExtraProducers : kProducers;
template <int BN>
constexpr int kThreadsBN = kThreads + (BN == 64 ? 32 * kExtraProducers : 0);
constexpr int BM = 128;
constexpr int kChunk = 32;  // accumulator columns per epilogue step

struct Work {
  int nt, mt, bz, it0, it1;
};

// Developer timeline (tools/trace_gemm.py builds a second library with -DGPV_GEMM_TRACE; the shipped library carries none of this):
// clock64 stamps of CTA 0, [6 roles][kTraceN uses][8 slots] -- roles 0-2 producer warps (index = stage use), 3 MMA issuer, 4 / 5 the
// first / last epilogue warp (index = work item).
#ifdef GPV_GEMM_TRACE
constexpr int kTraceN = 96;
__device__ long long* g_gemm_trace = nullptr;
#define GT_STAMP(role, idx, slot)                                                                      \
  do {                                                                                                 \
    if (gt_buf != nullptr && (idx) < kTraceN) gt_buf[((role) * kTraceN + (idx)) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define GT_STAMP(role, idx, slot) do { } while (0)
#endif

// Position in the stage ring, advanced once per stage use (no `%` / `/` by the run-time stage count in the loops).
struct Ring {
  int s;         // stage slot, gi % nstages
  uint32_t ph;   // (gi / nstages) & 1
  int turn;      // gi % nprod: the producer warp this stage use belongs to
};
GPV_DEVINL void ring_next(Ring& r, int S, int nprod) {
  if (++r.s == S) {
    r.s = 0;
    r.ph ^= 1u;
  }
  if (++r.turn == nprod) r.turn = 0;
}

GPV_DEVINL Work decode_work(const KParams& p, int w) {
  Work k;
  const int q1 = fdiv(w, p.fd_nt);
  k.nt = w - q1 * p.n_tiles;
  const int q2 = fdiv(q1, p.fd_mt);
  k.mt = q1 - q2 * p.m_tiles;
  const int sp = fdiv(q2, p.fd_gy);
  k.bz = q2 - sp * p.gy;
  k.it0 = sp * p.k_per_split;
  k.it1 = min(k.it0 + p.k_per_split, p.k_iters);
  return k;
}

// ---- coalesced epilogue -----------------------------------------------------------------------------------
// Each epilogue warp owns a 32-row x 32-column slab of the tile per step.  A thread's natural view is one ROW
// (tcgen05.ld 32x32b: lane = row), but one row-slice is only 64 bytes of global memory, so row-per-thread accesses
// touch 32 cache lines per instruction.  Instead every global access uses the COALESCED arrangement -- access j of
// lane l is 16-byte chunk (l & 3) of row 8j + (l >> 2): 8 rows x 64 contiguous bytes per warp instruction -- and a
// swizzled 2 KB shared-memory slab converts between the two views: chunk c of row r lives at slot
// r*4 + (c ^ ((r >> 1) & 3)) (conflict-free both ways).
// Residual / aux slices are brought in with cp.async (global -> shared, no registers, no scoreboard) one slab ahead,
// into a two-slot ring per warp; the TMA producer warp L2-prefetches the tile's residual / aux rows several tiles
// earlier, so the cp.async traffic hits L2.  (ncu on the first version, which prefetched into registers with LDG:
// every epilogue warp sat on long-scoreboard stalls because later loads share scoreboards with the prefetch.)
// (slab_slot, sts128 / lds128, cp.async and bf16 pack helpers: common.cuh)

// Global row (pixel) index of tile row r of work item wk, and whether it exists.
GPV_DEVINL bool tile_row(const KParams& p, const Work& wk, int r, long long* pix) {
  if (p.mode == 1) {
    const int img = fdiv(wk.mt, p.fd_tpi);
    const int rr_ = wk.mt - img * (p.tiles_h * p.tiles_w);
    const int trow = fdiv(rr_, p.fd_tws), prow = fdiv(r, p.fd_tw);
    const int ho = trow * p.th + prow, wo = (rr_ - trow * p.tiles_w) * p.tw + (r - prow * p.tw);
    *pix = ((long long)img * p.OH + (ho * p.os + p.ooh)) * p.OW + (wo * p.os + p.oow);
    return (r < p.th * p.tw) && ho < p.Ho && wo < p.Wo;
  }
  *pix = (long long)wk.mt * BM + r;
  return *pix < p.M;
}

// Epilogue variant F: -1 = every flag read from KParams at run time (any combination, any output type);
// F >= 0 = compile-time flags of the bf16 coalesced path (bit 0 bias, bit 1 bf16 residual, bits 2-3 activation,
// bits 4-5 aux mode, bit 6 second output D2, bits 7-8 dropout mode), so the hot variants carry no flag tests.
template <int F>
struct EpiFlags {
  GPV_DEVINL static bool bias(const KParams& p) { if constexpr (F < 0) return p.bias != nullptr; else return (F & 1) != 0; }
  GPV_DEVINL static bool res(const KParams& p) { if constexpr (F < 0) return p.residual != nullptr; else return (F & 2) != 0; }
  GPV_DEVINL static int act(const KParams& p) { if constexpr (F < 0) return p.act; else return (F >> 2) & 3; }
  GPV_DEVINL static int aux(const KParams& p) { if constexpr (F < 0) return p.aux_mode; else return (F >> 4) & 3; }
  GPV_DEVINL static bool d2(const KParams& p) { if constexpr (F < 0) return p.D2 != nullptr; else return (F & 64) != 0; }
  GPV_DEVINL static bool res_fp32(const KParams& p) { if constexpr (F < 0) return p.res_fp32 != 0; else return false; }
  GPV_DEVINL static bool d_fp32(const KParams& p) { if constexpr (F < 0) return p.d_fp32 != 0; else return false; }
  GPV_DEVINL static int drop(const KParams& p) { if constexpr (F < 0) return p.drop_mode; else return (F >> 7) & 3; }
};

template <int F>
GPV_DEVINL void epi_activation(const KParams& p, float (&v)[kChunk]) {
  const int act = EpiFlags<F>::act(p);
  if (act == GPVB200_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (act == GPVB200_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = gelu_erf(v[j]);
  } else if (act == GPVB200_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = 1.0f / (1.0f + __expf(-v[j]));
  }
}
template <int F>
GPV_DEVINL void epi_aux(const KParams& p, float (&v)[kChunk], const float (&a)[kChunk]) {
  if (EpiFlags<F>::aux(p) == GPVB200_AUX_RELU_MASK) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = a[j] > 0.0f ? v[j] : 0.0f;
  } else {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] *= gelu_erf_grad(a[j]);
  }
}

// One 32-column slice of an accumulator row, the thread's own row addressed directly:
// v = alpha*acc*rowscale + bias + residual; D2 = v; v = act(v); v *= mask(aux) / gelu'(aux); store.
//   vec = false  scalar (ragged N edge, unaligned operands);  vec = true  16-byte vectors (fp32 / atomic outputs)
// Dropout on the 32-column slice [nb, nb+32) of row `pix` (N even; see common.cuh for the mask definition).
GPV_DEVINL void epi_dropout(const KParams& p, float (&v)[kChunk], uint32_t dkey, long long pix, int nb) {
  const uint32_t base = (uint32_t)pix * (uint32_t)((p.N + 1) >> 1) + (uint32_t)(nb >> 1);
#pragma unroll
  for (int j = 0; j < kChunk / 2; ++j) drop_pair(v[2 * j], v[2 * j + 1], dkey, base + j, p.drop.thresh16, p.drop.scale);
}

template <int F>
GPV_DEVINL void epi_chunk(const KParams& p, const uint32_t (&acc)[kChunk], float rs, long long row_off, long long pix, int nb,
                          int nvalid, bool vec, uint32_t dkey) {
  typedef EpiFlags<F> E;
  const long long off_d = row_off + pix * p.ldd + nb, off_r = row_off + pix * p.ldr + nb, off_a = row_off + pix * p.ldaux + nb;
  float v[kChunk];
#pragma unroll
  for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(acc[j]) * rs;
  if (E::bias(p)) {
    if (vec) {
#pragma unroll
      for (int j = 0; j < kChunk; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) v[j] += __ldg(p.bias + nb + j);
    }
  }
  if (E::drop(p) == 1) epi_dropout(p, v, dkey, pix, nb);
  if (E::res(p)) {
    if (E::res_fp32(p)) {
      const float* rp = reinterpret_cast<const float*>(p.residual) + off_r;
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) v[j] += rp[j];
    } else if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack8(ldg_u4(p.residual + off_r + 8 * i), v + 8 * i, true);
    } else {
      const bf16* rp = p.residual + off_r;
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) v[j] += __bfloat162float(rp[j]);
    }
  }
  if (E::d2(p)) {
    bf16* dp = p.D2 + off_d;
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dp + 8 * i) = pack8(v + 8 * i);
    } else {
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) dp[j] = __float2bfloat16(v[j]);
    }
  }
  epi_activation<F>(p, v);
  if (E::drop(p) == 2) epi_dropout(p, v, dkey, pix, nb);
  if (E::aux(p) != GPVB200_AUX_NONE) {
    float a[kChunk];
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack8(ldg_u4(p.aux + off_a + 8 * i), a + 8 * i, false);
    } else {
      const bf16* ap = p.aux + off_a;
#pragma unroll
      for (int j = 0; j < kChunk; ++j) a[j] = (j < nvalid) ? __bfloat162float(ap[j]) : 0.0f;
    }
    epi_aux<F>(p, v, a);
  }
  if (E::d_fp32(p)) {
    float* dp = reinterpret_cast<float*>(p.D) + off_d;
    if (p.d_atomic) {
      if (vec) {  // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic requests
#pragma unroll
        for (int j = 0; j < kChunk; j += 4)
          atomicAdd(reinterpret_cast<float4*>(dp + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
          if (j < nvalid) atomicAdd(dp + j, v[j]);
      }
    } else if (vec) {
#pragma unroll
      for (int j = 0; j < kChunk; j += 4)
        *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) dp[j] = v[j];
    }
  } else {
    bf16* dp = reinterpret_cast<bf16*>(p.D) + off_d;
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dp + 8 * i) = pack8(v + 8 * i);
    } else {
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (j < nvalid) dp[j] = __float2bfloat16(v[j]);
    }
  }
}

// ---- coalesced bf16 path.  Per work item and lane: the global addresses of this lane's four coalesced 16-byte chunks
// (slab row 8j + (lane >> 2), chunk lane & 3, first column of the warp's half) of the residual, aux, D and D2.
struct CoalRows {
  const bf16* r[4];
  const bf16* a[4];
  bf16* d[4];
  bf16* d2[4];
  uint32_t ok;   // bit j: slab row 8j + (lane >> 2) exists
};
// Byte offsets inside a 2 KB slab: co = this lane's coalesced chunk of row-group 0 (row-group j is 512 B further),
// own[i] = chunk i of this lane's own row.
struct SlabOffs {
  uint32_t co, own[4];
};

// Own-row bf16 slice (4 x 16 bytes) -> swizzled slab -> coalesced global store at column offset `col` (elements).
GPV_DEVINL void store_slab(bf16* const (&dst)[4], uint32_t ok, int col, const float* v, uint32_t slab, const SlabOffs& so) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts128(slab + so.own[i], pack8(v + 8 * i));
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 o = lds128(slab + so.co + 512u * j);
    if ((ok >> j) & 1u) *reinterpret_cast<uint4*>(dst[j] + col) = o;
  }
  __syncwarp();
}

// Same arithmetic as epi_chunk for one full 32-column slab of a warp; residual / aux slices wait in the slabs res_s /
// aux_s (cp.async), outputs leave through out_s (the slab of an input already consumed, or a slab of its own).
template <int F>
GPV_DEVINL void epi_chunk_coal(const KParams& p, const uint32_t (&acc)[kChunk], float rs, const float4 (&b4)[kChunk / 4], int col,
                               uint32_t res_s, uint32_t aux_s, uint32_t out_s, const CoalRows& cr, const SlabOffs& so,
                               uint32_t dkey, long long pix, int nb) {
  typedef EpiFlags<F> E;
  float v[kChunk];
  if (rs != 1.0f) {  // rs = alpha * rowscale; 1 for most layers
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(acc[j]) * rs;
  } else {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (E::bias(p)) {
#pragma unroll
    for (int j = 0; j < kChunk; j += 4) {
      v[j] += b4[j / 4].x; v[j + 1] += b4[j / 4].y; v[j + 2] += b4[j / 4].z; v[j + 3] += b4[j / 4].w;
    }
  }
  if (E::drop(p) == 1) epi_dropout(p, v, dkey, pix, nb);   // compile-time in the F >= 0 variants: only the two dropout variants carry it
  if (E::res(p)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack8(lds128(res_s + so.own[i]), v + 8 * i, true);
  }
  if (E::d2(p)) store_slab(cr.d2, cr.ok, col, v, out_s, so);
  epi_activation<F>(p, v);
  if (E::drop(p) == 2) epi_dropout(p, v, dkey, pix, nb);
  if (E::aux(p) != GPVB200_AUX_NONE) {
    float a[kChunk];
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack8(lds128(aux_s + so.own[i]), a + 8 * i, false);
    epi_aux<F>(p, v, a);
  }
  store_slab(cr.d, cr.ok, col, v, out_s, so);
}

// Persistent kernel: each CTA walks work items w = blockIdx.x, blockIdx.x + gridDim.x, ... (an item = one
// 128 x BN output tile of one batch/tap and one K split).  Two TMEM accumulator buffers let the epilogue of item j
// overlap the TMA/MMA main loop of item j+1.
template <int BN, int F, bool PAIR = false>
__global__ void __launch_bounds__(kThreadsBN<BN>, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Pair variant: rank 0 (the leader) issues the MMAs for both CTAs; work items are walked per pair.
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int first_work = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int BNC = PAIR ? BN / 2 : BN;   // B columns this CTA stages
  constexpr bool kAltEpi = BN == 64 && !PAIR;   // narrow tiles: the two epilogue warp groups take alternate work items

  // ---- shared memory carve-up ----------------------------------------------------------------------
  const int rowsA = (p.mode == 1) ? p.th * p.tw : BM;
  const uint32_t a_blk = 128u * 128u;                  // one 64-deep k-block of a K-major A tile (reserved; conv tiles may write less)
  const uint32_t b_blk = (uint32_t)BNC * 128u;         // one 64-deep k-block of a K-major B tile
  const uint32_t a_bytes = p.a_mn ? 2u * p.bk * 128u : (uint32_t)p.kblk * a_blk;  // reserved per stage
  // Resident B (narrow forward convolutions: one column tile, <= 96 KB of weights): loaded once per CTA in front of the stage ring;
  // a 24 KB stage becomes 16 KB, and these launches are bound by the chip-wide L2 -> SM rate (~43 B/clk per SM with every SM pulling).
  const uint32_t b_res_bytes = p.b_res ? (uint32_t)p.kb_total * b_blk : 0u;
  const uint32_t b_bytes = p.b_res ? 0u : (p.b_mn ? (uint32_t)(BNC / 64) * p.bk * 128u : (uint32_t)p.kblk * b_blk);
  const uint32_t a_tx = p.a_mn ? a_bytes : (uint32_t)p.kblk * rowsA * 128u;    // bytes TMA actually writes
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int S = p.nstages;
  uint8_t* const bres = smem;
  smem += b_res_bytes;                                  // (a multiple of 8 KB: the stage ring stays 1024-byte aligned)
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)S * stage_bytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* bres_bar = acc_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(bres_bar + 1);
  // per epilogue warp: residual ring 2 x 2 KB (also the output transposition slab), aux ring 2 x 2 KB, bias slice 512 B
  const uint32_t stg_base = (smem_u32(tmem_slot + 4) + 127u) & ~127u;
  constexpr uint32_t kTmemCols = 2 * BN;  // 128 / 256 / 512: powers of two

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      // pair: the leader's barrier collects both CTAs' epilogue warps; BN = 64: one group of four warps per accumulator (see the epilogue)
      mbar_init(&acc_empty[b], PAIR ? 2 * kEpiWarps : (kAltEpi ? kEpiWarps / 2 : kEpiWarps));
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run while
  // the previous kernel of the stream is still draining; global memory is touched only after this point.  Dependents
  // are released at once -- their own prologue then overlaps this kernel's work.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef GPV_GEMM_TRACE
  long long* const gt_buf = blockIdx.x == 0 ? g_gemm_trace : nullptr;   // read once: a stamp is one clock read and one store
#endif

  if (warp < kProducers || warp >= kEpiWarp0 + kEpiWarps) {
    // ================================================================== TMA producers (stage gi belongs to producer gi % nprod)
    const int pw = warp < kProducers ? warp : warp - (kEpiWarp0 + kEpiWarps) + kProducers;   // producer index (BN = 64: 0..6)
    // Lane 0 issues the TMA loads; before that, all 32 lanes L2-prefetch the residual / aux rows the epilogue of this
    // work item will read (the producer runs 2+ tiles ahead of the epilogue, so they are L2 hits by then).
    const bool pf_r = p.pf_mode && p.coal && p.residual != nullptr, pf_a = p.pf_mode && p.coal && p.aux_mode != GPVB200_AUX_NONE;
    int gi = 0;  // stage-use counter, runs across work items
    Ring rg = {0, 0u, 0};
    if (p.b_res && pw == 0 && first_work < p.total_work) {
      if (elect_one()) {
        mbar_expect_tx(bres_bar, b_res_bytes);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          if (p.mode == 0) {
            tma_load_4d(bres + (size_t)kb * b_blk, &tmB, bres_bar, kb * 64, 0, 0, 0);
          } else {
            const int tap = fdiv(kb, p.fd_kc), kc = kb - tap * p.kc_per_tap;
            tma_load_4d(bres + (size_t)kb * b_blk, &tmB, bres_bar, kc * 64, 0, p.tap_w[tap], 0);
          }
        }
      }
      __syncwarp();
    }
    for (int w = first_work; w < p.total_work; w += work_step) {
      Work wk = decode_work(p, w);
      if constexpr (PAIR) wk.mt = 2 * wk.mt + rank;   // may be one past the last tile: TMA zero-fills, the epilogue skips it
      const int n0 = wk.nt * BN + rank * BNC, m0 = wk.mt * BM, bz = wk.bz;
      if ((pf_r || pf_a) && pw == 0) {
        const uint32_t bytes = (uint32_t)(min(BN, p.N - n0) * 2) & ~15u;
        const long long row_off = (long long)bz * p.d_batch_stride + n0;
        if (bytes > 0) {
#pragma unroll
          for (int r = lane; r < BM; r += 32) {
            long long pix;
            if (tile_row(p, wk, r, &pix)) {
              if (pf_r) prefetch_l2_row(p.residual + row_off + pix * p.ldr, bytes, p.pf_mode);
              if (pf_a) prefetch_l2_row(p.aux + row_off + pix * p.ldaux, bytes, p.pf_mode);
            }
          }
        }
      }
      const bool elected = elect_one();
      if (elected) {
        int img = 0, ho0 = 0, wo0 = 0;
        if (p.mode == 1) {
          img = fdiv(wk.mt, p.fd_tpi);
          const int r = wk.mt - img * (p.tiles_h * p.tiles_w);
          const int trow = fdiv(r, p.fd_tws);
          ho0 = trow * p.th;
          wo0 = (r - trow * p.tiles_w) * p.tw;
        }
        for (int it = wk.it0; it < wk.it1; ++it, ++gi, ring_next(rg, S, p.nprod)) {
          if (rg.turn != pw) continue;
          const int s = rg.s;
          const uint32_t ph = rg.ph;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          GT_STAMP(0, gi, 0);
          GT_STAMP(0, gi, 2 + (pw & 3));
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if constexpr (PAIR) {
            // both CTAs' bytes land on the leader's barrier; 128-deep stages (modes 0 and 1)
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * (a_tx + b_bytes));
            const uint32_t fb = leader_smem_addr(smem_u32(&full_bar[s]));
            const int bzB = p.b_batched ? bz : 0;
            for (int kb = 0; kb < p.kblk; ++kb) {
              const int kbi = it * p.kblk + kb;         // 64-deep k-block of the contraction; past the end: zero-filled
              if (p.mode == 0) {
                if (!p.a_mn) {
                  tma_load_4d_pair(sa + kb * a_blk, &tmA, fb, kbi * 64, m0, bz, 0);
                } else {   // MN-major A (weight gradients): two 64-column slabs of [bk rows of K][64], one 64-row box per k-block
                  tma_load_4d_pair(sa + kb * 64 * 128, &tmA, fb, m0, kbi * 64, bz, 0);
                  tma_load_4d_pair(sa + p.bk * 128 + kb * 64 * 128, &tmA, fb, m0 + 64, kbi * 64, bz, 0);
                }
                if (!p.b_mn) tma_load_4d_pair(sb + kb * b_blk, &tmB, fb, kbi * 64, n0, bzB, 0);
              } else {
                const bool live = kbi < p.kb_total;
                const int tap = live ? kbi / p.kc_per_tap : 0, kc = live ? kbi % p.kc_per_tap : p.kc_per_tap;
                tma_load_4d_pair(sa + kb * a_blk, &tmA, fb, kc * 64, wo0 * p.stride + p.tap_dw[tap], ho0 * p.stride + p.tap_dh[tap], img);
                if (!p.b_mn) tma_load_4d_pair(sb + kb * b_blk, &tmB, fb, kc * 64, n0, p.tap_w[tap], 0);
              }
            }
            if (p.b_mn) {   // MN-major B: per 64 columns a [bk rows of K][64] slab (LBO = bk * 128); one 64-row box per k-block
              for (int kb = 0; kb < p.kblk; ++kb) {
                const int kbi = it * p.kblk + kb;
                const bool live = p.mode == 0 || kbi < p.kb_total;
                const int tap = (p.mode == 1 && live) ? kbi / p.kc_per_tap : 0;
                const int krow = p.mode == 0 ? kbi * 64 : (live ? (kbi % p.kc_per_tap) * 64 : p.kc_per_tap * 64);
                const int c2 = p.mode == 0 ? bzB : p.tap_w[tap];
#pragma unroll
                for (int j = 0; j < BNC / 64; ++j)
                  tma_load_4d_pair(sb + j * p.bk * 128 + kb * 64 * 128, &tmB, fb, n0 + 64 * j, krow, c2, 0);
              }
            }
            continue;
          }
          mbar_expect_tx(&full_bar[s], a_tx + b_bytes);
          if (p.mode == 0) {
            const int k0 = it * p.bk;
            const int bzB = p.b_batched ? bz : 0;
            if (!p.a_mn) {
              tma_load_4d(sa, &tmA, &full_bar[s], k0, m0, bz, 0);
            } else {
              tma_load_4d(sa, &tmA, &full_bar[s], m0, k0, bz, 0);
              tma_load_4d(sa + p.bk * 128, &tmA, &full_bar[s], m0 + 64, k0, bz, 0);
            }
            if (p.b_res) {
            } else if (!p.b_mn) {
              tma_load_4d(sb, &tmB, &full_bar[s], k0, n0, bzB, 0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d(sb + j * p.bk * 128, &tmB, &full_bar[s], n0 + 64 * j, k0, bzB, 0);
            }
          } else if (p.mode == 1) {
            const int tap = fdiv(it, p.fd_kc), kc = it - tap * p.kc_per_tap;
            tma_load_4d(sa, &tmA, &full_bar[s], kc * 64, wo0 * p.stride + p.tap_dw[tap], ho0 * p.stride + p.tap_dh[tap], img);
            if (p.b_res) {
            } else if (!p.b_mn) {
              tma_load_4d(sb, &tmB, &full_bar[s], kc * 64, n0, p.tap_w[tap], 0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d(sb + j * p.bk * 128, &tmB, &full_bar[s], n0 + 64 * j, kc * 64, p.tap_w[tap], 0);
            }
          } else {
            const int im = fdiv(it, p.fd_tpi), r = it - im * (p.tiles_h * p.tiles_w);
            const int trow = fdiv(r, p.fd_tws);
            const int h0 = trow * p.th, w0 = (r - trow * p.tiles_w) * p.tw;
            tma_load_4d(sa, &tmA, &full_bar[s], m0, w0, h0, im);
            tma_load_4d(sa + p.bk * 128, &tmA, &full_bar[s], m0 + 64, w0, h0, im);
            const int wi = w0 * p.stride + p.tap_dw[bz], hi = h0 * p.stride + p.tap_dh[bz];
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * p.bk * 128, &tmB, &full_bar[s], n0 + 64 * j, wi, hi, im);
          }
          GT_STAMP(0, gi, 1);
        }
      } else {
        gi += wk.it1 - wk.it0;
      }
      __syncwarp();
      {   // the ring position travels from the lane that walked the stages to the others (any lane may be elected next time)
        const int src = __ffs(__ballot_sync(0xffffffffu, elected)) - 1;
        const uint32_t packed = __shfl_sync(0xffffffffu, (uint32_t)rg.s | (rg.ph << 8) | ((uint32_t)rg.turn << 16), src);
        rg.s = (int)(packed & 0xffu);
        rg.ph = (packed >> 8) & 1u;
        rg.turn = (int)(packed >> 16);
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issuer
    const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN, p.a_mn, p.b_mn);
    const uint32_t a_lbo = p.a_mn ? (uint32_t)p.bk * 128u : 0u;
    const uint32_t b_lbo = p.b_mn ? (uint32_t)p.bk * 128u : 0u;
    const uint32_t a_kstep = p.a_mn ? 2048u : 32u;  // bytes per UMMA_K = 16 step
    const uint32_t b_kstep = p.b_mn ? 2048u : 32u;
    const int ksteps = p.bk / 16;
    int gi = 0, j = 0;
    Ring rg = {0, 0u, 0};
    if constexpr (!PAIR) {
      // ONE elected thread runs the whole role loop (waits, MMAs, commits), and the operand descriptors advance by adding constants to
      // a base descriptor.  Measured (tools/proto/mma_issue.cu, profiles/r4g_mma_issue.txt): electing per stage and rebuilding both
      // descriptors per MMA costs the issuing warp 376 / 392 / 513 clocks per 64-deep stage at N = 64 / 128 / 256 whatever the tensor
      // time (128 / 256 / 512); this form costs 219 / 260 / 511.  With the waits and fences on top, the per-stage chain of this warp
      // (650-750 clocks for every tile width: clock64 timeline, profiles/r4f_trace_gemm.txt) was what bounded every main loop.
      if (elect_one()) {
        if (p.b_res) {
          mbar_wait(bres_bar, 0u);
          tc_fence_after();
        }
        const uint32_t stage0 = smem_u32(smem);
        const uint64_t a_base = make_sdesc_sw128(stage0, a_lbo, 1024u), b_base = make_sdesc_sw128(stage0 + a_bytes, b_lbo, 1024u);
        const uint64_t bres_base = make_sdesc_sw128(smem_u32(bres), 0u, 1024u);
        const uint64_t sstep = stage_bytes >> 4, bres_step = b_blk >> 4;       // descriptor address units (16 bytes)
        const uint64_t a_kd = a_kstep >> 4, b_kd = b_kstep >> 4;
        for (int w = first_work; w < p.total_work; w += work_step, ++j) {
          const Work wk = decode_work(p, w);
          const int buf = j & 1;
          GT_STAMP(3, j, 0);
          mbar_wait(&acc_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          GT_STAMP(3, j, 1);
          const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
          uint32_t accum = 0u;                                         // the first MMA of an item overwrites the accumulator
          for (int it = wk.it0; it < wk.it1; ++it, ++gi, ring_next(rg, S, 1)) {
            const int s = rg.s;
            // (measured and not kept, profiles/r4k_mma_loop_switches.txt: dropping this per-stage fence, or testing the next stage's
            // barrier before this stage's MMAs are issued, changes nothing / costs 1-5 % on the narrow shapes)
            mbar_wait(&full_bar[s], rg.ph);
            tc_fence_after();
            GT_STAMP(1, gi, 0);
            if (it == wk.it0) GT_STAMP(3, j, 2);
            if (it == wk.it1 - 1) GT_STAMP(3, j, 3);
            uint64_t ad = a_base + (uint64_t)s * sstep;
            uint64_t bd = p.b_res ? bres_base + (uint64_t)it * bres_step : b_base + (uint64_t)s * sstep;
            if (ksteps == 4) {
              umma_f16(tacc, ad, bd, idesc, accum);
              umma_f16(tacc, ad + a_kd, bd + b_kd, idesc, 1u);
              umma_f16(tacc, ad + 2 * a_kd, bd + 2 * b_kd, idesc, 1u);
              umma_f16(tacc, ad + 3 * a_kd, bd + 3 * b_kd, idesc, 1u);
            } else {
              umma_f16(tacc, ad, bd, idesc, accum);
              for (int k = 1; k < ksteps; ++k) {
                ad += a_kd;
                bd += b_kd;
                umma_f16(tacc, ad, bd, idesc, 1u);
              }
            }
            accum = 1u;
            umma_commit(&empty_bar[s]);                         // frees the smem stage once these MMAs retire
            if (it == wk.it1 - 1) umma_commit(&acc_full[buf]);  // accumulator complete
            GT_STAMP(1, gi, 1);
            if (it == wk.it1 - 1) GT_STAMP(3, j, 4);
          }
        }
      }
      __syncwarp();
    } else {
    for (int w = first_work; w < p.total_work && rank == 0; w += work_step, ++j) {
      const Work wk = decode_work(p, w);
      const int buf = j & 1;
      mbar_wait(&acc_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
      for (int it = wk.it0; it < wk.it1; ++it, ++gi, ring_next(rg, S, 1)) {
        const int s = rg.s;
        const uint32_t ph = rg.ph;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          for (int k = 0; k < ksteps; ++k) {
            // K-major operands: 64-deep k-blocks (4 steps of 32 bytes inside the 128-byte swizzle row), one after the other
            const uint32_t ao = p.a_mn ? k * a_kstep : (uint32_t)(k >> 2) * a_blk + (uint32_t)(k & 3) * 32u;
            const uint32_t bo = p.b_mn ? k * b_kstep : (uint32_t)(k >> 2) * b_blk + (uint32_t)(k & 3) * 32u;
            const uint64_t ad = make_sdesc_sw128(sa + ao, a_lbo, 1024u);
            const uint64_t bd = make_sdesc_sw128(sb + bo, b_lbo, 1024u);
            umma_f16_pair(tacc, ad, bd, idesc, (it > wk.it0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s]);                          // frees the stage in both CTAs
          if (it == wk.it1 - 1) umma_commit_pair(&acc_full[buf]);   // both epilogues
        }
        __syncwarp();
      }
    }
    }
  } else {
    // ================================================================== epilogue (warps kEpiWarp0 .. kEpiWarp0 + 7)
    typedef EpiFlags<F> E;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    // BN >= 128: the two groups of four warps split the columns of every item.  BN = 64 (kAltEpi): a group takes every other work item
    // whole -- its accumulator buffer is its own -- so two epilogues are in flight and each warp stores full 128-byte rows; the per-item
    // epilogue chain of ~1700-2600 clocks (addressing, TMEM read, transposition, stores) bounded the 1- to 4-stage items of the
    // frozen trunk prefix (clock64 timeline, profiles/r4c_trace_gemm.txt).
    const int grp = (warp - kEpiWarp0) >> 2;
    const int half = kAltEpi ? 0 : grp;             // which half of the BN columns
    const int r = q * 32 + lane;
    constexpr int kChunksPerHalf = kAltEpi ? BN / kChunk : BN / 2 / kChunk;
    const int cbase = half * (BN / 2);
    const bool coal = F >= 0 || p.coal != 0;
    const uint32_t res_ring = stg_base + (uint32_t)(warp - kEpiWarp0) * p.epi_warp_bytes, aux_ring = res_ring + p.epi_aux_off;
    const uint32_t slot_stride = p.epi_slot_stride;
    const bool full_pf = p.epi_full != 0;
    const bool pre_r = E::res(p) && !E::res_fp32(p), pre_a = E::aux(p) != GPVB200_AUX_NONE;
    const uint32_t out_ring = (pre_r || !pre_a) ? res_ring : aux_ring;
    SlabOffs so;
    so.co = 16u * slab_slot(lane >> 2, lane & 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) so.own[i] = 16u * slab_slot(lane, i);
    const long long ldd = p.ldd, ldr = p.ldr, lda = p.ldaux;
    const uint32_t dkey = E::drop(p) ? drop_key(*p.drop.seed, p.drop.site) : 0u;
    const int jstep = kAltEpi ? 2 : 1;
    int j = kAltEpi ? grp : 0;
    for (int w = first_work + j * work_step; w < p.total_work; w += jstep * work_step, j += jstep) {
#ifdef GPV_GEMM_TRACE
      const int trole = warp == kEpiWarp0 ? 4 : (warp == kEpiWarp0 + kEpiWarps - 1 ? 5 : -1);
#define GT_EPI(slot) do { if (trole >= 0 && lane == 0) GT_STAMP(trole, j, slot); } while (0)
#else
#define GT_EPI(slot) do { } while (0)
#endif
      GT_EPI(0);
      Work wk = decode_work(p, w);
      if constexpr (PAIR) wk.mt = 2 * wk.mt + rank;
      const int n0 = wk.nt * BN, bz = wk.bz;
      const int buf = j & 1;
      long long pix;
      const bool row_ok = tile_row(p, wk, r, &pix) && wk.mt < p.m_tiles_cta;
      const long long row_off = (long long)bz * p.d_batch_stride;  // mode 0: batch, mode 2: tap, mode 1: bz == 0
      const float rs = p.alpha * ((p.rowscale != nullptr && row_ok && p.mode != 1) ? p.rowscale[wk.mt * BM + r] : 1.0f);
      CoalRows cr;
      cr.ok = 0;
      if (coal) {
        const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
        const long long cofs = row_off + n0 + cbase + (lane & 3) * 8;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int row = 8 * jj + (lane >> 2);
          const long long pj = __shfl_sync(0xffffffffu, pix, row);
          cr.ok |= ((okmask >> row) & 1u) << jj;
          cr.d[jj] = reinterpret_cast<bf16*>(p.D) + pj * ldd + cofs;
          if (E::d2(p)) cr.d2[jj] = p.D2 + pj * ldd + cofs;
          if (pre_r) cr.r[jj] = p.residual + pj * ldr + cofs;
          if (pre_a) cr.a[jj] = p.aux + pj * lda + cofs;
        }
      }
      const float* bias_t = p.bias + n0 + cbase;

      // cp.async the residual / aux slices of chunk c into its slot (coalesced arrangement); one commit per call
      auto issue = [&](int c) {
        if (n0 + cbase + (c + 1) * kChunk <= p.N) {
          const uint32_t so_c = (uint32_t)(full_pf ? c : (c & 1)) * slot_stride + so.co;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            if ((cr.ok >> jj) & 1u) {
              if (pre_r) cp_async16(res_ring + so_c + 512u * jj, cr.r[jj] + c * kChunk);
              if (pre_a) cp_async16(aux_ring + so_c + 512u * jj, cr.a[jj] + c * kChunk);
            }
          }
        }
        cp_async_commit();
      };
      if (coal) {  // independent of the accumulator: overlaps the wait below
        issue(0);
        if (full_pf) {
#pragma unroll
          for (int c = 1; c < kChunksPerHalf; ++c) issue(c);
        }
      }
      GT_EPI(1);
      mbar_wait(&acc_full[buf], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      GT_EPI(2);
#pragma unroll
      for (int c = 0; c < kChunksPerHalf; ++c) {
        const int nb = n0 + cbase + c * kChunk;
        if (nb < p.N) {  // warp-uniform
          uint32_t acc[kChunk];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cbase + c * kChunk), acc);
          float4 b4[kChunk / 4];
          if (coal && E::bias(p) && nb + kChunk <= p.N) {   // bias slice of this slab: in flight across the waits below
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias_t + c * kChunk) + i);
          }
          if (coal) {
            if (full_pf) {
              cp_async_wait_n(kChunksPerHalf - 1 - c);   // slabs 0..c have landed (this thread's copies) ...
            } else {
              if (c + 1 < kChunksPerHalf) issue(c + 1);
              else cp_async_commit();
              cp_async_wait<1>();   // everything but the slab just requested has landed
            }
            __syncwarp();           // ... and every other lane's
          }
          tmem_ld_wait();
          if (c == kChunksPerHalf - 1) GT_EPI(3);
          const int nvalid = min(kChunk, p.N - nb);
          const bool full = nvalid == kChunk;
          if (coal && full) {                     // whole warp takes part (shared-memory slabs)
            const uint32_t sc = (uint32_t)(full_pf ? c : (c & 1)) * slot_stride;
            epi_chunk_coal<F>(p, acc, rs, b4, c * kChunk, res_ring + sc, aux_ring + sc, out_ring + sc, cr, so, dkey, pix, nb);
          } else if (row_ok) {
            epi_chunk<F>(p, acc, rs, row_off, pix, nb, nvalid, p.vec_ok && full, dkey);
          }
        }
      }
      GT_EPI(4);
      if (coal) {
        cp_async_wait<0>();
        __syncwarp();           // the slabs are free for the next work item
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(leader_smem_addr(smem_u32(&acc_empty[buf])));
        else mbar_arrive(&acc_empty[buf]);
      }
      GT_EPI(5);
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();   // neither CTA leaves while its peer may still read its shared memory or write its TMEM
    if (warp == kMmaWarp) tmem_dealloc_pair(tmem_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================
// Host side
// =====================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

struct MapKey {
  uint64_t v[13];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 13; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};

// 4-D bf16 tensor map, SWIZZLE_128B, zero OOB fill. dims/strides innermost first; strides in elements for dims 1..3.
// (declared in host_util.h: layer_umma.cu builds its maps through the same cache)
int make_map(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_el[3],
                    const uint32_t box[4], const uint32_t estr[4]) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key;
  key.v[0] = (uint64_t)(uintptr_t)ptr;
  for (int i = 0; i < 4; ++i) key.v[1 + i] = dims[i];
  for (int i = 0; i < 3; ++i) key.v[5 + i] = strides_el[i];
  for (int i = 0; i < 4; ++i) key.v[8 + i] = ((uint64_t)box[i] << 32) | estr[i];
  key.v[12] = 0;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return GPV_OK;
    }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return GPV_ERR_CUDA;
  }
  if (((uintptr_t)ptr & 15) != 0) {
    set_last_error("tensor map base pointer %p not 16-byte aligned", ptr);
    return GPV_ERR_ARG;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = estr[i];
  }
  for (int i = 0; i < 3; ++i) {
    gstr[i] = strides_el[i] * 2;
    if (gstr[i] % 16 != 0) {
      set_last_error("tensor map stride %llu bytes (dim %d) not a multiple of 16", (unsigned long long)gstr[i], i + 1);
      return GPV_ERR_ARG;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                   (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    return GPV_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return GPV_OK;
}

// Choose a th x tw pixel tile with th*tw <= limit (or == exact when `exact`), maximising useful rows.
static void pick_tile(int H, int W, int limit, bool exact_mult16, int* th_out, int* tw_out) {
  double best = -1.0;
  int bth = 1, btw = 1;
  // weight-gradient tiles may hang over the image edge (TMA zero-fills), which lets tiny maps reach a multiple of 16
  const int Wm = exact_mult16 ? (W < 8 ? 8 : W) : W, Hm = exact_mult16 ? (H < 8 ? 8 : H) : H;
  for (int tw = 1; tw <= Wm && tw <= limit; ++tw) {
    for (int th = 1; th <= Hm && th * tw <= limit; ++th) {
      const int rows = th * tw;
      if (exact_mult16 && (rows % 16 != 0)) continue;
      const long long tiles = (long long)((H + th - 1) / th) * ((W + tw - 1) / tw);
      // cost of a tile is a full MMA pass (limit rows) for mode 1, `rows` of contraction for mode 2
      const double work = exact_mult16 ? (double)tiles * rows + tiles * 8.0 : (double)tiles * limit;
      const double eff = (double)H * W / work;
      if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > btw)) {
        best = eff;
        bth = th;
        btw = tw;
      }
    }
  }
  *th_out = bth;
  *tw_out = btw;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int F, bool PAIR = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<BN, F, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return GPV_ERR_CUDA;
    }
    configured = true;
  }
  // one persistent CTA per SM; the pair variant launches clusters of two CTAs (the two SMs of a TPC), one pair per work item
  int slots = PAIR ? num_sms() / 2 : num_sms();
  {
    // Launches that run on a lane beside a dependent chain of kernels (weight gradients beside the data-gradient chain) ask for a
    // capped grid (gpvb200_gemm_desc.max_ctas): as persistent kernels they would otherwise occupy every SM and the chain's next
    // kernel would wait for their CTAs to retire (measured: 16.51 -> 16.05 ms per step at 72 of 148, profiles/r2t_wgrad_cta_cap.txt).
    const int wg_cap = kp.max_ctas;
    if (wg_cap > 0 && !PAIR && slots > wg_cap) slots = wg_cap;
  }
  const int grid = (kp.total_work < slots ? kp.total_work : slots) * (PAIR ? 2 : 1);
  static int pdl = -1;   // GPVB200_PDL=0 disables programmatic dependent launch
  if (pdl < 0) {
    const char* e = getenv("GPVB200_PDL");
    pdl = e ? atoi(e) : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsBN<BN>);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, umma_gemm_kernel<BN, F, PAIR>, ma, mb, kp);
  if (e != cudaSuccess) {
    set_last_error("umma_gemm_kernel launch failed: %s", cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  return check_launch("umma_gemm_kernel");
}

// Epilogue variants compiled with their flags fixed (the bf16 coalesced path of the hot layers); anything else runs
// the run-time-flag kernel (F = -1).  bit 0 bias, bit 1 residual, bits 2-3 act, bits 4-5 aux, bit 6 D2.
#define GPV_EPI_VARIANTS(X) X(0) X(1) X(2) X(3) X(5) X(7) X(16) X(18) X(32) X(73) X(131) X(261)   // 131 / 261: + dropout before the residual / after ReLU

template <int BN, bool PAIR = false>
static int launch_bn(int f, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, size_t smem, cudaStream_t st) {
  switch (f) {
#define GPV_CASE(V) case V: return launch<BN, V, PAIR>(ma, mb, kp, smem, st);
    GPV_EPI_VARIANTS(GPV_CASE)
#undef GPV_CASE
    default: return launch<BN, -1, PAIR>(ma, mb, kp, smem, st);
  }
}

// Pair variant (cta_group::2) for contractions at least this many 64-deep k-blocks long; 0 = never (the default).
// GPVB200_PAIR overrides.  Measured (profiles/r1H_summary.md): as a stand-alone kernel the pair main loop is 1.12-1.33x
// faster on the deep-K shapes of the trunk, but with the threshold at 16 the whole step got 4.5 % SLOWER (19.81 vs 18.92
// ms): a pair launch costs ~2 us more (cluster scheduling, two cluster barriers), needs both SMs of a TPC free while the
// weight-gradient lane occupies single SMs, halves the work-item count (worse tail waves), and the two-ring epilogue
// variants keep only two 64 KB stages.  Opt-in until the scheduling side is solved.
static int pair_min_kblocks() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPVB200_PAIR");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace gpv

using namespace gpv;

extern "C" size_t gpvb200_gemm_desc_size(void) { return sizeof(gpvb200_gemm_desc); }

#ifdef GPV_GEMM_TRACE
/* developer hook of the trace build only (tools/trace_gemm.py): device buffer [6][96][8] int64 that CTA 0 fills; NULL = off */
extern "C" int gpvb200_gemm_trace(void* buf) {
  long long* b = (long long*)buf;
  return cudaMemcpyToSymbol(gpv::g_gemm_trace, &b, sizeof(b)) == cudaSuccess ? GPV_OK : GPV_ERR_CUDA;
}
#endif

static long long g_pair_launches = 0;   // statistics only (tests assert that the pair variant really ran)
extern "C" int64_t gpvb200_gemm_pair_launches(void) { return (int64_t)g_pair_launches; }

extern "C" int gpvb200_gemm(const gpvb200_gemm_desc* d, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(d != nullptr, "gemm: null descriptor");
  GPV_REQUIRE(d->mode >= 0 && d->mode <= 2, "gemm: bad mode %d", d->mode);
  GPV_REQUIRE(d->A && d->B && d->D, "gemm: null operand");
  GPV_REQUIRE(d->N > 0 && d->K >= 0, "gemm: bad N/K");
  const int splits_in = d->splits > 1 ? d->splits : (d->splits == 1 ? 1 : 0);
  const bool auto_split = d->splits <= 0 && d->d_atomic && d->d_fp32 && d->mode != 2;   // splits 0: choose here
  GPV_REQUIRE(splits_in <= 1 || (d->d_atomic && d->d_fp32), "gemm: split-K needs fp32 atomic output");
  GPV_REQUIRE(!d->d_atomic || d->d_fp32, "gemm: atomic output must be fp32");
  if (d->aux_mode != GPVB200_AUX_NONE) GPV_REQUIRE(d->aux != nullptr, "gemm: aux_mode set without aux");

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.mode = d->mode;
  kp.M = d->M;
  kp.N = d->N;
  kp.a_mn = d->a_mn ? 1 : 0;
  kp.b_mn = d->b_mn ? 1 : 0;
  kp.splits = 1;
  kp.act = d->act;
  kp.aux_mode = d->aux_mode;
  kp.d_fp32 = d->d_fp32;
  kp.d_atomic = d->d_atomic;
  kp.max_ctas = d->max_ctas > 0 ? d->max_ctas : 0;
  kp.alpha = d->alpha;
  kp.res_fp32 = d->res_fp32;
  kp.D = d->D;
  kp.D2 = (bf16*)d->D2;
  kp.bias = d->bias;
  kp.rowscale = d->rowscale;
  kp.residual = (const bf16*)d->residual;
  kp.aux = (const bf16*)d->aux;
  kp.ldd = d->ldd;
  kp.ldr = d->residual ? d->ldr : 8;
  kp.ldaux = d->aux ? d->ldaux : 8;
  kp.d_batch_stride = d->d_batch_stride;
  {
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    kp.vec_ok = ((kp.ldd & 7) == 0) && ((kp.ldr & 7) == 0) && ((kp.ldaux & 7) == 0) && ((kp.d_batch_stride & 7) == 0) &&
                al16(d->D) && al16(d->D2) && al16(d->residual) && al16(d->aux) && al16(d->bias);
  }
  // coalesced epilogue: bf16 outputs (and bf16 residual / aux) move as 8 rows x 64 bytes per warp instruction
  {
    static int pf = -1;
    if (pf < 0) {
      const char* e = getenv("GPVB200_PF");
      pf = e ? atoi(e) : 0;
    }
    kp.pf_mode = pf;
  }
  {
    static int nprod = -1;   // GPVB200_PRODUCERS=1 reproduces the single-producer kernel (A/B measurements)
    if (nprod < 0) {
      const char* e = getenv("GPVB200_PRODUCERS");
      nprod = e ? atoi(e) : kProducers + kExtraProducers;
      if (nprod < 1) nprod = 1;
      if (nprod > kProducers + kExtraProducers) nprod = kProducers + kExtraProducers;
    }
    kp.nprod = nprod;   // clamped to the producer warps of the chosen tile width below
  }
  kp.drop_mode = 0;
  if (d->drop_mode != 0 && d->drop_p > 0.f) {
    GPV_REQUIRE(d->drop_mode == 1 || d->drop_mode == 2, "gemm: bad drop_mode %d", d->drop_mode);
    GPV_REQUIRE(d->drop_seed != nullptr && d->drop_p < 1.f, "gemm: dropout needs a seed and p < 1");
    GPV_REQUIRE(d->mode != 2 && (d->mode == 1 || d->batch <= 1) && (d->N % 2 == 0) && splits_in <= 1 && !d->d_atomic,
                "gemm: dropout needs an unbatched, unsplit problem with even N");
    kp.drop_mode = d->drop_mode;
    kp.drop.seed = (const unsigned long long*)d->drop_seed;
    kp.drop.site = d->drop_site;
    kp.drop.thresh16 = (uint32_t)(d->drop_p * 65536.0f + 0.5f);
    kp.drop.scale = 1.0f / (1.0f - d->drop_p);
  }
  kp.coal = kp.vec_ok && !kp.d_fp32 && !kp.res_fp32 && !(d->D2 && d->aux_mode != GPVB200_AUX_NONE);
  kp.stride = d->stride > 0 ? d->stride : 1;
  kp.os = d->out_stride > 0 ? d->out_stride : 1;
  kp.ooh = d->out_off_h;
  kp.oow = d->out_off_w;
  kp.OH = d->OH;
  kp.OW = d->OW;
  kp.Ho = d->Ho;
  kp.Wo = d->Wo;
  kp.ntaps = d->ntaps;
  for (int i = 0; i < 9; ++i) {
    kp.tap_dh[i] = d->tap_dh[i];
    kp.tap_dw[i] = d->tap_dw[i];
    kp.tap_w[i] = d->tap_w[i];
  }

  // ---- tile width: the widest BN that still yields about one work item per SM -------------------------------
  int m_tiles_pre, gy_pre;
  if (d->mode == 1) {
    GPV_REQUIRE(d->Ho > 0 && d->Wo > 0 && d->n_img > 0, "gemm: bad conv geometry");
    int th, tw;
    pick_tile(d->Ho, d->Wo, 128, false, &th, &tw);
    m_tiles_pre = d->n_img * ((d->Ho + th - 1) / th) * ((d->Wo + tw - 1) / tw);
    gy_pre = 1;
  } else {
    m_tiles_pre = (d->M + BM - 1) / BM;
    gy_pre = d->mode == 0 ? (d->batch > 0 ? d->batch : 1) : d->ntaps;
  }
  // Cost model (us): a k-iteration is L2->smem bound (A 16 KB + B BN/8 KB at ~68 GB/s per SM), the epilogue moves
  // 128 x BN outputs (plus residual) at the SM's share of HBM; split-K pays one vector atomic per output per split.
  const int bn_cap = d->mode == 2 ? 128 : 256;
  const int k_iters_pre = d->mode == 0 ? (d->K + 63) / 64 : (d->mode == 1 ? d->ntaps * ((d->K + 63) / 64) : 0);
  int BN = 64, splits = splits_in > 0 ? splits_in : 1;
  static int min_kper = -1;   // GPVB200_MIN_KPER=<n>: automatic split-K keeps at least n 64-deep k-iterations per split (default 8: measured 17.44 -> 17.25 ms per step against 1, profiles/r2p)
  if (min_kper < 0) {
    const char* e = getenv("GPVB200_MIN_KPER");
    min_kper = e ? atoi(e) : 8;
    if (min_kper < 1) min_kper = 1;
  }
  static int force_bn = -1;   // developer override (tools/sweep_bn.py): GPVB200_FORCE_BN=64|128|256
  if (force_bn < 0) {
    const char* e = getenv("GPVB200_FORCE_BN");
    force_bn = e ? atoi(e) : 0;
  }
  // per-k-iteration cost a + b * BN (us); GPVB200_TITER="a,b" overrides (developer sweeps)
  static double titer_a = -1.0, titer_b = 0.0;
  if (titer_a < 0.0) {
    titer_a = 0.24;
    titer_b = 0.0018;
    const char* e = getenv("GPVB200_TITER");
    if (e) {
      double a_ = 0, b_ = 0;
      if (sscanf(e, "%lf,%lf", &a_, &b_) == 2 && a_ >= 0.0) {
        titer_a = a_;
        titer_b = b_;
      }
    }
  }
  {
    double best = 1e30;
    for (int bn = 64; bn <= bn_cap; bn *= 2) {
      if (bn > 64 && d->N <= bn / 2) break;
      if (force_bn && bn != force_bn && !(force_bn > bn_cap && bn == bn_cap)) continue;
      const long long tiles = (long long)m_tiles_pre * ((d->N + bn - 1) / bn) * gy_pre;
      int sp = splits_in > 0 ? splits_in : 1;
      if (auto_split && k_iters_pre > 0) {
        sp = (int)((num_sms() + tiles - 1) / tiles);
        if (sp > k_iters_pre / min_kper) sp = k_iters_pre / min_kper;   // every split keeps at least min_kper k-iterations
        if (sp < 1) sp = 1;
      }
      const double t_iter = titer_a + titer_b * bn;
      const double t_epi = (d->d_atomic ? 4.0 : 3.0) * bn / 256.0;
      const int kper = k_iters_pre > 0 ? (k_iters_pre + sp - 1) / sp : 8;
      const long long waves = (tiles * sp + num_sms() - 1) / num_sms();
      const double t = waves * (kper * t_iter + t_epi + 1.5);
      if (t < best - 1e-9) {
        best = t;
        BN = bn;
        splits = sp;
      }
    }
  }

  if (BN == 256 && d->residual && d->aux_mode != GPVB200_AUX_NONE && k_iters_pre > 0 && k_iters_pre <= 4 && splits == 1)
    BN = 128;   // both epilogue input streams on a shallow-K item: full slab prefetch fits only with two chunks per warp

  // ---- pair variant: K-major A (plain GEMM / implicit-GEMM convolution), bf16 or fp32 stores, deep contraction, 256 columns
  const int kb_total_pre = d->mode == 0 ? (d->K + 63) / 64 : k_iters_pre;
  static int pair_fwd_only = -1;   // GPVB200_PAIR_FWD=1: pairs only for forward-direction contractions (K-major B), which run without
  if (pair_fwd_only < 0) {         // the single-CTA weight-gradient lane beside them (a pair needs both SMs of a TPC free at once)
    const char* e = getenv("GPVB200_PAIR_FWD");
    pair_fwd_only = e ? atoi(e) : 0;
  }
  const bool pair_shape_ok = !kp.a_mn && splits == 1 && !d->d_atomic && !(pair_fwd_only && d->b_mn);
  // GPVB200_PAIR_BN=128 also pairs the 128-column tiles (per CTA: A 16 KB + B 8 KB per k-block -- the L2 traffic of the single-CTA
  // 256-column tile, so only the halved item count remains; default 256)
  static int pair_min_bn = -1;
  if (pair_min_bn < 0) {
    const char* e = getenv("GPVB200_PAIR_BN");
    pair_min_bn = e ? atoi(e) : 256;
    if (pair_min_bn < 128) pair_min_bn = 128;
  }
  const bool pair = pair_min_kblocks() > 0 && d->mode != 2 && pair_shape_ok && BN >= pair_min_bn &&
                    kb_total_pre >= pair_min_kblocks() && m_tiles_pre >= 2;
  kp.kblk = pair ? 2 : 1;
  const uint32_t BNC = pair ? BN / 2 : BN;   // B columns one CTA stages

  CUtensorMap ma, mb;
  const uint32_t one4[4] = {1, 1, 1, 1};

  if (d->mode == 0) {
    GPV_REQUIRE(d->M > 0 && d->K > 0 && d->batch > 0, "gemm: bad plain shape");
    kp.bk = 64 * kp.kblk;
    kp.kb_total = (d->K + 63) / 64;
    kp.k_iters = (kp.kb_total + kp.kblk - 1) / kp.kblk;
    kp.b_batched = d->b_batch_stride != 0;
    {
      uint64_t dims[4], str[3];
      uint32_t box[4];
      if (!kp.a_mn) {
        dims[0] = d->K; dims[1] = d->M; box[0] = 64; box[1] = 128;
      } else {
        dims[0] = d->M; dims[1] = d->K; box[0] = 64; box[1] = 64;
      }
      dims[2] = d->batch; dims[3] = 1; box[2] = 1; box[3] = 1;
      str[0] = d->lda;
      str[1] = d->batch > 1 ? (uint64_t)d->a_batch_stride : (uint64_t)d->lda * dims[1];
      str[2] = str[1] * dims[2];
      rc = make_map(&ma, d->A, dims, str, box, one4);
      if (rc) return rc;
    }
    {
      uint64_t dims[4], str[3];
      uint32_t box[4];
      if (!kp.b_mn) {
        dims[0] = d->K; dims[1] = d->N; box[0] = 64; box[1] = BNC;
      } else {
        dims[0] = d->N; dims[1] = d->K; box[0] = 64; box[1] = 64;
      }
      const int bb = kp.b_batched ? d->batch : 1;
      dims[2] = bb; dims[3] = 1; box[2] = 1; box[3] = 1;
      str[0] = d->ldb;
      str[1] = bb > 1 ? (uint64_t)d->b_batch_stride : (uint64_t)d->ldb * dims[1];
      str[2] = str[1] * dims[2];
      rc = make_map(&mb, d->B, dims, str, box, one4);
      if (rc) return rc;
    }
    kp.m_tiles_cta = (d->M + BM - 1) / BM;
    kp.m_tiles = pair ? (kp.m_tiles_cta + 1) / 2 : kp.m_tiles_cta;
    kp.gy = d->batch;
  } else if (d->mode == 1) {
    GPV_REQUIRE(d->n_img > 0 && d->Hi > 0 && d->Wi > 0 && d->Ho > 0 && d->Wo > 0, "gemm: bad conv geometry");
    GPV_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, "gemm: ntaps must be 1..9");
    GPV_REQUIRE(!kp.a_mn, "gemm: conv A must be channel-contiguous");
    GPV_REQUIRE(kp.stride <= 2, "gemm: conv stride > 2 unsupported");
    kp.bk = 64 * kp.kblk;
    kp.kc_per_tap = (d->K + 63) / 64;
    kp.kb_total = d->ntaps * kp.kc_per_tap;
    kp.k_iters = (kp.kb_total + kp.kblk - 1) / kp.kblk;
    pick_tile(d->Ho, d->Wo, 128 / 1, false, &kp.th, &kp.tw);
    if (kp.tw * kp.stride > 256 || kp.th * kp.stride > 256) {
      set_last_error("gemm: conv tile exceeds TMA box limit");
      return GPV_ERR_ARG;
    }
    kp.tiles_h = (d->Ho + kp.th - 1) / kp.th;
    kp.tiles_w = (d->Wo + kp.tw - 1) / kp.tw;
    {
      uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->Wi, (uint64_t)d->Hi, (uint64_t)d->n_img};
      uint64_t str[3] = {(uint64_t)d->lda, (uint64_t)d->lda * d->Wi, (uint64_t)d->lda * d->Wi * d->Hi};
      uint32_t box[4] = {64, (uint32_t)(kp.tw * kp.stride), (uint32_t)(kp.th * kp.stride), 1};
      uint32_t es[4] = {1, (uint32_t)kp.stride, (uint32_t)kp.stride, 1};
      rc = make_map(&ma, d->A, dims, str, box, es);
      if (rc) return rc;
    }
    {
      int ntw = 0;
      for (int i = 0; i < d->ntaps; ++i) ntw = d->tap_w[i] + 1 > ntw ? d->tap_w[i] + 1 : ntw;
      uint64_t dims[4], str[3];
      uint32_t box[4];
      if (!kp.b_mn) {
        dims[0] = d->K; dims[1] = d->N; box[0] = 64; box[1] = BNC;
      } else {
        dims[0] = d->N; dims[1] = d->K; box[0] = 64; box[1] = 64;
      }
      dims[2] = ntw; dims[3] = 1; box[2] = 1; box[3] = 1;
      str[0] = d->ldb;
      str[1] = d->b_batch_stride ? (uint64_t)d->b_batch_stride : (uint64_t)d->ldb * dims[1];
      str[2] = str[1] * dims[2];
      rc = make_map(&mb, d->B, dims, str, box, one4);
      if (rc) return rc;
    }
    if (kp.OH == 0) { kp.OH = d->Ho; kp.OW = d->Wo; }
    kp.m_tiles_cta = d->n_img * kp.tiles_h * kp.tiles_w;
    kp.m_tiles = pair ? (kp.m_tiles_cta + 1) / 2 : kp.m_tiles_cta;
    kp.gy = 1;
  } else {
    GPV_REQUIRE(d->n_img > 0 && d->Hi > 0 && d->Wi > 0 && d->Ho > 0 && d->Wo > 0, "gemm: bad wgrad geometry");
    GPV_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, "gemm: ntaps must be 1..9");
    GPV_REQUIRE(d->d_atomic && d->d_fp32, "gemm: wgrad output must be fp32 atomic");
    GPV_REQUIRE(d->M > 0, "gemm: bad wgrad M");
    kp.a_mn = 1;
    kp.b_mn = 1;
    pick_tile(d->Ho, d->Wo, 96, true, &kp.th, &kp.tw);
    kp.bk = kp.th * kp.tw;
    GPV_REQUIRE(kp.bk % 16 == 0 && kp.bk >= 16, "gemm: no wgrad pixel tile for %dx%d", d->Ho, d->Wo);
    kp.tiles_h = (d->Ho + kp.th - 1) / kp.th;
    kp.tiles_w = (d->Wo + kp.tw - 1) / kp.tw;
    kp.k_iters = d->n_img * kp.tiles_h * kp.tiles_w;
    {
      uint64_t dims[4] = {(uint64_t)d->M, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->n_img};
      uint64_t str[3] = {(uint64_t)d->lda, (uint64_t)d->lda * d->Wo, (uint64_t)d->lda * d->Wo * d->Ho};
      uint32_t box[4] = {64, (uint32_t)kp.tw, (uint32_t)kp.th, 1};
      rc = make_map(&ma, d->A, dims, str, box, one4);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->Wi, (uint64_t)d->Hi, (uint64_t)d->n_img};
      uint64_t str[3] = {(uint64_t)d->ldb, (uint64_t)d->ldb * d->Wi, (uint64_t)d->ldb * d->Wi * d->Hi};
      uint32_t box[4] = {64, (uint32_t)(kp.tw * kp.stride), (uint32_t)(kp.th * kp.stride), 1};
      uint32_t es[4] = {1, (uint32_t)kp.stride, (uint32_t)kp.stride, 1};
      rc = make_map(&mb, d->B, dims, str, box, es);
      if (rc) return rc;
    }
    kp.m_tiles = kp.m_tiles_cta = (d->M + BM - 1) / BM;
    kp.kb_total = kp.k_iters;
    kp.gy = d->ntaps;
  }
  // K splits: equal chunks of k_per_split iterations, none empty
  {
    int sp = splits > kp.k_iters ? kp.k_iters : splits;
    kp.k_per_split = (kp.k_iters + sp - 1) / sp;
    kp.splits = (kp.k_iters + kp.k_per_split - 1) / kp.k_per_split;
  }
  kp.n_tiles = (d->N + BN - 1) / BN;
  {
    const long long tw_ = (long long)kp.m_tiles * kp.n_tiles * kp.gy * kp.splits;
    GPV_REQUIRE(tw_ > 0 && tw_ < (1ll << 31), "gemm: work list size %lld out of range", tw_);
    kp.total_work = (int)tw_;
  }

  // ---- epilogue slabs (coalesced path) -----------------------------------------------------------------
  // One 2 KB slot per warp, streamed input (residual, aux) and slab in flight.  Shallow-K work items (the 1x1
  // convolutions with K = 64..256) are epilogue-bound and need every slab of the item requested up front to cover the
  // HBM latency (one slot per chunk); deep-K items hide it under the main loop and keep a two-slot ring, which leaves
  // room for a fourth pipeline stage.  The output is transposed through the slot of an input already consumed, or
  // through a slot of its own when there is none.
  {
    const int rings = (kp.coal && d->residual ? 1 : 0) + (kp.coal && d->aux_mode != GPVB200_AUX_NONE ? 1 : 0);
    const int chunks = (BN == 64 && !pair) ? BN / kChunk : BN / 2 / kChunk;   // slabs per epilogue warp and work item
    kp.epi_full = rings > 0 && kp.k_per_split <= 4 && rings * chunks <= 4 && !pair;   // pair items are deep-K by construction
    const int depth = kp.epi_full ? chunks : 2;
    kp.epi_warp_bytes = rings ? 2048u * depth * rings : 2048u;
    kp.epi_aux_off = (d->residual && rings == 2) ? 2048u * depth : 0u;
    kp.epi_slot_stride = rings ? 2048u : 0u;
  }

  // ---- pipeline depth ---------------------------------------------------------------------------------
  const uint32_t a_bytes = kp.a_mn ? 2u * kp.bk * 128u : (uint32_t)kp.kblk * 128u * 128u;
  uint32_t b_bytes = kp.b_mn ? (BNC / 64) * kp.bk * 128u : (uint32_t)kp.kblk * BNC * 128u;
  // resident B: one column tile, K-major un-batched B, one split (every CTA walks whole contractions), at most 96 KB of weights
  static int bres_on = -1;   // GPVB200_BRES=0 switches it off (A/B measurements)
  if (bres_on < 0) {
    const char* e = getenv("GPVB200_BRES");
    bres_on = e ? atoi(e) : 1;
  }
  uint32_t b_res_bytes = 0;
  kp.b_res = 0;
  if (bres_on && !pair && kp.mode != 2 && !kp.a_mn && !kp.b_mn && !kp.b_batched && kp.n_tiles == 1 && kp.gy == 1 && kp.splits == 1 &&
      kp.kblk == 1 && (uint32_t)kp.kb_total * BNC * 128u <= 96u * 1024u) {
    kp.b_res = 1;
    b_res_bytes = (uint32_t)kp.kb_total * BNC * 128u;
    b_bytes = 0;
  }
  const uint32_t stage = a_bytes + b_bytes;
  const uint32_t epi_smem = kEpiWarps * kp.epi_warp_bytes + 128;     // 16 / 32 / 64 KB of epilogue slabs
  int nst = (int)((227u * 1024u - 1024u - 256u - epi_smem - b_res_bytes) / stage);
  if (nst > (BN == 64 ? 8 : 6)) nst = BN == 64 ? 8 : 6;
  GPV_REQUIRE(nst >= 2, "gemm: stage of %u bytes does not fit twice in shared memory", stage);
  kp.nstages = nst;
  const size_t smem = (size_t)nst * stage + b_res_bytes + 1024 /*alignment slack*/ + (2 * nst + 5) * 8 + 32 + epi_smem;

  if (kp.nprod > (BN == 64 ? kProducers + kExtraProducers : kProducers)) kp.nprod = BN == 64 ? kProducers + kExtraProducers : kProducers;
  // A producer waits on the PARITY of a slot's empty barrier, which tells the current phase from the previous one only: the warp that
  // fills stage use g has passed the wait of use g - nprod, i.e. uses up to g - nprod - nst are consumed, and needs use g - 2 nst
  // consumed for its parity test to be unambiguous -> nprod <= nst.
  if (kp.nprod > nst) kp.nprod = nst;
  kp.fd_nt = make_fastdiv(kp.n_tiles);
  kp.fd_mt = make_fastdiv(kp.m_tiles);
  kp.fd_gy = make_fastdiv(kp.gy);
  kp.fd_tpi = make_fastdiv(kp.mode == 0 ? 1 : kp.tiles_h * kp.tiles_w);
  kp.fd_tws = make_fastdiv(kp.mode == 0 ? 1 : kp.tiles_w);
  kp.fd_tw = make_fastdiv(kp.mode == 0 ? 1 : kp.tw);
  kp.fd_kc = make_fastdiv(kp.mode == 1 ? kp.kc_per_tap : 1);

  cudaStream_t st = (cudaStream_t)stream;
  int f = -1;
  if (kp.coal)
    f = (d->bias ? 1 : 0) | (d->residual ? 2 : 0) | ((d->act & 3) << 2) | ((d->aux_mode & 3) << 4) | (d->D2 ? 64 : 0) | (kp.drop_mode << 7);
  if (pair) {
    kp.nprod = 1;   // the pair hand-shake (both CTAs' bytes on the leader's barrier) was validated with one producer warp only
    ++g_pair_launches;
    return BN == 256 ? launch_bn<256, true>(f, ma, mb, kp, smem, st) : launch_bn<128, true>(f, ma, mb, kp, smem, st);
  }
  if (BN == 256) return launch_bn<256>(f, ma, mb, kp, smem, st);
  if (BN == 128) return launch_bn<128>(f, ma, mb, kp, smem, st);
  return launch_bn<64>(f, ma, mb, kp, smem, st);
}
