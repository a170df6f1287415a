// Decode-loop bookkeeping on the device (exp/gpv/models/gpv.py:178-196 greedy, 256-362 beam search): the arg-max over
// the vocabulary with the additive vocab mask, one beam-search update (log-softmax, per-beam top-K, candidate merge
// with the reference's order and tie rule, sequence / score / parent update) and the permutation of the KV caches by
// the parent index, and the single-query attention of the KV-cached decode step.  HBM/L2-bound row scans: one CTA per row (arg-max,
// beam candidates), per image (beam merge) or per (K/V batch, head) (decode attention); warp-shuffle reductions.
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

constexpr int kDecThreads = 512;
constexpr int kMaxBeams = 8;

struct Best {
  float v;
  int i;
};
// order of the reference's arg-max / top-k / stable descending sort: larger value first, then the smaller index
GPV_DEVINL bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

GPV_DEVINL Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (better(ov, oi, b.v, b.i)) {
      b.v = ov;
      b.i = oi;
    }
  }
  return b;
}

// block-wide best of (value, index) pairs; every thread gets the result.  red: [2 * 16] words of shared memory.
GPV_DEVINL Best block_best(Best b, float* red_v, int* red_i) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  b = warp_best(b);
  __syncthreads();
  if (lane == 0) {
    red_v[warp] = b.v;
    red_i[warp] = b.i;
  }
  __syncthreads();
  Best r;
  r.v = lane < (kDecThreads >> 5) ? red_v[lane] : -INFINITY;
  r.i = lane < (kDecThreads >> 5) ? red_i[lane] : 0x7fffffff;
  return warp_best(r);
}

GPV_DEVINL float block_sum(float s, float* red_v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red_v[warp] = s;
  __syncthreads();
  float r = lane < (kDecThreads >> 5) ? red_v[lane] : 0.f;
  return warp_sum(r);
}

// ------------------------------------------------------------------------------------------------
// Greedy step (gpv.py:186-190): out[row, :] = logits[row, :V] + vocab_mask; ids[row] = first arg-max of it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads) argmax_kernel(const float* __restrict__ logits, long long ld, int V,
                                                             const float* __restrict__ vmask, float* __restrict__ out,
                                                             long long ldo, int64_t* __restrict__ ids) {
  pdl_sync();
  __shared__ float red_v[16];
  __shared__ int red_i[16];
  const long long row = blockIdx.x;
  const float* lr = logits + row * ld;
  float* orow = out != nullptr ? out + row * ldo : nullptr;
  Best b;
  b.v = -INFINITY;
  b.i = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += kDecThreads) {
    float v = lr[i];
    if (vmask != nullptr) v += vmask[i];
    if (orow != nullptr) orow[i] = v;
    if (better(v, i, b.v, b.i)) {
      b.v = v;
      b.i = i;
    }
  }
  b = block_best(b, red_v, red_i);
  if (threadIdx.x == 0 && ids != nullptr) ids[row] = b.i == 0x7fffffff ? 0 : b.i;     // all-NaN row: torch returns some index; 0 here
}

// ------------------------------------------------------------------------------------------------
// One beam-search step in two launches.
// beam_rows_kernel (one CTA per hypothesis row b K + k1): the log-softmax of the row's next-token logits and its K best tokens
// (value descending, lower token id first on ties); candidates cand[k1][k2] = score[k1] + logp go to the workspace (at t = 0 only
// hypothesis 0 of an image is live: the others hold the same prefix, gpv.py:283-286, and get -1e9).
// beam_merge_kernel (one CTA per image): the K best of the K*K candidates in stable descending order (first occurrence in (k1 major, k2
// minor) order wins ties, as torch.sort(stable=True) in the reference's restatement); then ids_out[b, k, :t+1] = ids_in[b, k1, :t+1],
// ids_out[b, k, t+1] = token, score_out[b, k] = candidate, parent[b K + k] = b K + k1, tok[b K + k] = token.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads) beam_rows_kernel(const float* __restrict__ logits, long long ld, int K, int V, int t,
                                                                const float* __restrict__ score_in, float* __restrict__ cand_v,
                                                                int* __restrict__ cand_tok) {
  pdl_sync();
  __shared__ float red_v[16];
  __shared__ int red_i[16];
  const int row = blockIdx.x, k1 = row % K;
  if (t == 0 && k1 > 0) {
    if (threadIdx.x < K) {
      cand_v[(long long)row * K + threadIdx.x] = -1e9f;
      cand_tok[(long long)row * K + threadIdx.x] = 0;
    }
    return;
  }
  const float* lr = logits + (long long)row * ld;
  Best m;
  m.v = -INFINITY;
  m.i = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += kDecThreads) {
    const float v = lr[i];
    if (better(v, i, m.v, m.i)) {
      m.v = v;
      m.i = i;
    }
  }
  m = block_best(m, red_v, red_i);
  const float mx = m.v;
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += kDecThreads) s += expf(lr[i] - mx);
  s = block_sum(s, red_v);
  const float logsum = logf(s);
  const float sc = score_in[row];
  Best prev = m;                                   // the best token is the row maximum found above
  for (int k2 = 0; k2 < K; ++k2) {
    if (k2 > 0) {
      Best c;
      c.v = -INFINITY;
      c.i = 0x7fffffff;
      for (int i = threadIdx.x; i < V; i += kDecThreads) {
        const float v = lr[i];
        const bool after = v < prev.v || (v == prev.v && i > prev.i);      // strictly after the previous pick in the order
        if (after && better(v, i, c.v, c.i)) {
          c.v = v;
          c.i = i;
        }
      }
      prev = block_best(c, red_v, red_i);
    }
    if (threadIdx.x == 0) {
      cand_v[(long long)row * K + k2] = sc + ((prev.v - mx) - logsum);          // log_softmax = (x - max) - log(sum exp(x - max))
      cand_tok[(long long)row * K + k2] = prev.i == 0x7fffffff ? 0 : prev.i;
    }
  }
}

__global__ void __launch_bounds__(128) beam_merge_kernel(int K, int t, int L, const float* __restrict__ cand_v_g, const int* __restrict__ cand_tok_g,
                                                         const int64_t* __restrict__ ids_in, float* __restrict__ score_out,
                                                         int64_t* __restrict__ ids_out, int64_t* __restrict__ parent, int64_t* __restrict__ tok) {
  pdl_sync();
  __shared__ float cand_v[kMaxBeams * kMaxBeams];
  __shared__ int cand_tok[kMaxBeams * kMaxBeams];
  __shared__ int sel_k1[kMaxBeams];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    cand_v[i] = cand_v_g[(long long)b * K * K + i];
    cand_tok[i] = cand_tok_g[(long long)b * K * K + i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // K rounds of selection over K*K candidates in (value descending, flat index ascending) order
    float pv = INFINITY;
    int pi = -1;
    for (int k = 0; k < K; ++k) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = 0; c < K * K; ++c) {
        const float v = cand_v[c];
        const bool after = v < pv || (v == pv && c > pi);
        if (after && better(v, c, bv, bi)) {
          bv = v;
          bi = c;
        }
      }
      if (bi == 0x7fffffff) bi = k;                   // NaN scores: keep the layout defined
      pv = bv;
      pi = bi;
      const int k1 = bi / K;
      sel_k1[k] = k1;
      score_out[b * K + k] = cand_v[bi];
      parent[b * K + k] = (int64_t)b * K + k1;
      tok[b * K + k] = cand_tok[bi];
      ids_out[((long long)b * K + k) * L + t + 1] = cand_tok[bi];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * (t + 1); i += blockDim.x) {
    const int k = i / (t + 1), j = i % (t + 1);
    ids_out[((long long)b * K + k) * L + j] = ids_in[((long long)b * K + sel_k1[k]) * L + j];
  }
}

// ------------------------------------------------------------------------------------------------
// Attention of ONE query row per hypothesis against cached keys / values (the KV-cached decode step, gpv.py:178-196 / 318-326).  One CTA
// per (K/V batch, head): K and V of the head are staged in shared memory ONCE and serve every hypothesis that shares them (`rep`
// hypotheses per K/V batch: the beams of an image attend to the same encoder memory; rep = 1 for the self-attention caches).  K is
// staged transposed as bf16 pairs (a lane owns keys lane, lane + 32, ...: conflict-free), V row-major (a lane owns output dimensions).
// HBM/L2-bound: K and V of a step are read once per (batch, head) instead of once per hypothesis by a 16-row MMA tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) decode_attn_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ kk, long long ldk,
                                                          long long bsk, const bf16* __restrict__ vv, long long ldv, long long bsv,
                                                          bf16* __restrict__ o, long long ldo, int H, int Sk, int dh, int rep, float sl2) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_dec[];
  const int kvb = blockIdx.x / H, h = blockIdx.x % H;
  const int Skp = Sk | 1;                                  // odd word stride of the transposed K rows
  const int dh2 = dh >> 1;
  uint32_t* Kt = reinterpret_cast<uint32_t*>(smem_dec);   // [dh / 2][Skp] bf16 pairs (d, d + 1) of key s
  bf16* Vs = reinterpret_cast<bf16*>(Kt + (size_t)dh2 * Skp);   // [Sk][dh]
  float* prob = reinterpret_cast<float*>(Vs + (size_t)Sk * dh);   // [4 warps][Sk]
  float* qs = prob + 4 * Sk;                                  // [4 warps][dh]
  const int CH = dh >> 3;
  const bf16* kb = kk + (long long)kvb * bsk + h * dh;
  const bf16* vb = vv + (long long)kvb * bsv + h * dh;
  for (int i = threadIdx.x; i < Sk * CH; i += blockDim.x) {
    const int s = i / CH, c = i % CH;
    const uint4 kq = *reinterpret_cast<const uint4*>(kb + (long long)s * ldk + c * 8);
    Kt[(c * 4 + 0) * Skp + s] = kq.x;
    Kt[(c * 4 + 1) * Skp + s] = kq.y;
    Kt[(c * 4 + 2) * Skp + s] = kq.z;
    Kt[(c * 4 + 3) * Skp + s] = kq.w;
    *reinterpret_cast<uint4*>(Vs + (size_t)s * dh + c * 8) = *reinterpret_cast<const uint4*>(vb + (long long)s * ldv + c * 8);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pw = prob + warp * Sk;
  float* qw = qs + warp * dh;
  for (int r = warp; r < rep; r += 4) {
    const long long row = (long long)kvb * rep + r;
    const bf16* qr = q + row * ldq + h * dh;
    for (int d = lane; d < dh; d += 32) qw[d] = __bfloat162float(qr[d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int s = lane; s < Sk; s += 32) {
      float acc = 0.f;
      for (int d2 = 0; d2 < dh2; ++d2) {
        const float2 kf = unpack_bf16x2(Kt[d2 * Skp + s]);
        acc = fmaf(qw[2 * d2], kf.x, acc);
        acc = fmaf(qw[2 * d2 + 1], kf.y, acc);
      }
      acc *= sl2;                                           // log2 domain
      pw[s] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < Sk; s += 32) {
      const float e = ex2_approx(pw[s] - mx);
      pw[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int s = 0; s < Sk; ++s) acc = fmaf(pw[s], __bfloat162float(Vs[(size_t)s * dh + d]), acc);
      o[row * ldo + h * dh + d] = __float2bfloat16(acc * inv);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// KV-cache permutation: dst[r, :n] = src[parent[r], :n] for rows of `row_elems` bf16 (16-byte aligned), only the first
// n elements of each row (the positions decoded so far) are moved.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reorder_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                           const int64_t* __restrict__ parent, long long row_vec, int n_vec) {
  pdl_sync();
  const long long r = blockIdx.y;
  const uint4* s = src + parent[r] * row_vec;
  uint4* d = dst + r * row_vec;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n_vec; i += gridDim.x * 256) d[i] = s[i];
}

}  // namespace gpv

using namespace gpv;

extern "C" int gpvb200_argmax(const float* logits, int64_t ld, int32_t rows, int32_t V, const float* vocab_mask, float* out, int64_t ldo,
                              int64_t* ids, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(logits && rows >= 0 && V > 0 && (out || ids), "argmax: bad arguments");
  if (rows == 0) return GPV_OK;
  launch_k(argmax_kernel, dim3(rows), dim3(kDecThreads), 0, (cudaStream_t)stream, logits, (long long)ld, V, vocab_mask, out, (long long)ldo, ids);
  return check_launch("argmax_kernel");
}

extern "C" int gpvb200_beam_update(const float* logits, int64_t ld, int32_t B, int32_t K, int32_t V, int32_t t, int32_t L,
                                   const float* score_in, const int64_t* ids_in, float* score_out, int64_t* ids_out, int64_t* parent,
                                   int64_t* tok, void* workspace, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(logits && score_in && ids_in && score_out && ids_out && parent && tok && workspace, "beam_update: null pointer");
  GPV_REQUIRE(B >= 0 && K >= 1 && K <= kMaxBeams && V >= K && t >= 0 && t + 1 < L, "beam_update: needs 1 <= K <= %d <= V and t + 1 < L (K %d, V %d, t %d, L %d)",
              kMaxBeams, K, V, t, L);
  GPV_REQUIRE(ids_in != ids_out && score_in != score_out, "beam_update: in-place update is not supported (double-buffer ids and scores)");
  GPV_REQUIRE(((uintptr_t)workspace & 7) == 0, "beam_update: workspace must be 8-byte aligned");
  if (B == 0) return GPV_OK;
  float* cand_v = (float*)workspace;                       // [B, K, K] candidate scores, then [B, K, K] candidate tokens
  int* cand_tok = (int*)(cand_v + (size_t)B * K * K);
  launch_k(beam_rows_kernel, dim3(B * K), dim3(kDecThreads), 0, (cudaStream_t)stream, logits, (long long)ld, K, V, t, score_in, cand_v, cand_tok);
  rc = check_launch("beam_rows_kernel");
  if (rc != GPV_OK) return rc;
  launch_k(beam_merge_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, K, t, L, (const float*)cand_v, (const int*)cand_tok, ids_in, score_out,
           ids_out, parent, tok);
  return check_launch("beam_merge_kernel");
}

extern "C" int gpvb200_decode_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, int64_t bsk, const void* v, int64_t ldv,
                                        int64_t bsv, void* o, int64_t ldo, int32_t Bq, int32_t rep, int32_t H, int32_t Sk, int32_t dh,
                                        float scale, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(q && k && v && o, "decode_attention: null operand");
  GPV_REQUIRE(Bq >= 0 && rep >= 1 && Bq % rep == 0 && H >= 1 && Sk >= 1, "decode_attention: needs Bq %% rep == 0 (Bq %d, rep %d), H, Sk >= 1", Bq, rep);
  GPV_REQUIRE(dh >= 8 && dh <= 256 && dh % 8 == 0, "decode_attention: head dim %d must be a multiple of 8, <= 256", dh);
  GPV_REQUIRE((ldk & 7) == 0 && (ldv & 7) == 0 && (bsk & 7) == 0 && (bsv & 7) == 0 && (((uintptr_t)k | (uintptr_t)v) & 15) == 0,
              "decode_attention: K / V rows must be 16-byte aligned");
  if (Bq == 0) return GPV_OK;
  const size_t smem = (size_t)(dh / 2) * (Sk | 1) * 4 + (size_t)Sk * dh * 2 + (size_t)4 * Sk * 4 + (size_t)4 * dh * 4 + 16;
  GPV_REQUIRE(smem <= 200 * 1024, "decode_attention: Sk = %d, dh = %d needs %zu bytes of shared memory (> 200 KB)", Sk, dh, smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_last_error("decode_attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return GPV_ERR_CUDA;
    }
    configured = smem;
  }
  launch_k(decode_attn_kernel, dim3((Bq / rep) * H), dim3(128), smem, (cudaStream_t)stream, (const bf16*)q, (long long)ldq, (const bf16*)k,
           (long long)ldk, (long long)bsk, (const bf16*)v, (long long)ldv, (long long)bsv, (bf16*)o, (long long)ldo, H, Sk, dh, rep,
           scale * 1.4426950408889634f);
  return check_launch("decode_attn_kernel");
}

extern "C" int gpvb200_reorder_rows(const void* src, void* dst, const int64_t* parent, int32_t rows, int64_t row_elems, int64_t n_elems,
                                    void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(src && dst && parent && src != dst, "reorder_rows: bad pointers (in-place permutation is not supported)");
  GPV_REQUIRE(rows >= 0 && rows <= 65535 && row_elems > 0 && n_elems >= 0 && n_elems <= row_elems && (row_elems & 7) == 0 && (n_elems & 7) == 0,
              "reorder_rows: rows of bf16 must be multiples of 8 elements");
  GPV_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "reorder_rows: buffers must be 16-byte aligned");
  if (rows == 0 || n_elems == 0) return GPV_OK;
  const int n_vec = (int)(n_elems / 8);
  int gx = (n_vec + 255) / 256;
  gx = gx > 8 ? 8 : gx;
  launch_k(reorder_rows_kernel, dim3(gx, rows), dim3(256), 0, (cudaStream_t)stream, (const uint4*)src, (uint4*)dst, parent, (long long)(row_elems / 8), n_vec);
  return check_launch("reorder_rows_kernel");
}
