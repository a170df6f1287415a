// Decode-loop bookkeeping on the device (exp/gpv/models/gpv.py:178-196 greedy, 256-362 beam search): the arg-max over
// the vocabulary with the additive vocab mask, one beam-search update (log-softmax, per-beam top-K, candidate merge
// with the reference's order and tie rule, sequence / score / parent update) and the permutation of the KV caches by
// the parent index.  HBM/L2-bound row scans: one CTA per row (arg-max) or per image (beam update), float4 loads,
// warp-shuffle reductions.
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
// One beam-search step for image b (one CTA): for each live beam k1 the log-softmax of its next-token logits and its K
// best tokens (value descending, lower token id first on ties); candidates cand[k1][k2] = score[k1] + logp (at t = 0
// only beam 0 is live: the others hold the same prefix, gpv.py:283-286); the K best of the K*K candidates in stable
// descending order (first occurrence in (k1 major, k2 minor) order wins ties, as torch.sort(stable=True) in the
// reference's restatement); then ids_out[b, k, :t+1] = ids_in[b, k1, :t+1], ids_out[b, k, t+1] = token,
// score_out[b, k] = candidate, parent[b K + k] = b K + k1, tok[b K + k] = token.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads) beam_update_kernel(const float* __restrict__ logits, long long ld, int K, int V, int t,
                                                                  int L, const float* __restrict__ score_in,
                                                                  const int64_t* __restrict__ ids_in, float* __restrict__ score_out,
                                                                  int64_t* __restrict__ ids_out, int64_t* __restrict__ parent,
                                                                  int64_t* __restrict__ tok) {
  pdl_sync();
  __shared__ float red_v[16];
  __shared__ int red_i[16];
  __shared__ float cand_v[kMaxBeams * kMaxBeams];
  __shared__ int cand_tok[kMaxBeams * kMaxBeams];
  __shared__ int sel_k1[kMaxBeams];
  const int b = blockIdx.x;
  const int live = t == 0 ? 1 : K;
  for (int k1 = 0; k1 < K; ++k1) {
    if (k1 >= live) {
      if (threadIdx.x < K) {
        cand_v[k1 * K + threadIdx.x] = -1e9f;
        cand_tok[k1 * K + threadIdx.x] = 0;
      }
      continue;
    }
    const float* lr = logits + ((long long)b * K + k1) * ld;
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
    const float sc = score_in[b * K + k1];
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
        cand_v[k1 * K + k2] = sc + ((prev.v - mx) - logsum);          // log_softmax = (x - max) - log(sum exp(x - max))
        cand_tok[k1 * K + k2] = prev.i == 0x7fffffff ? 0 : prev.i;
      }
    }
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
  for (int i = threadIdx.x; i < K * (t + 1); i += kDecThreads) {
    const int k = i / (t + 1), j = i % (t + 1);
    ids_out[((long long)b * K + k) * L + j] = ids_in[((long long)b * K + sel_k1[k]) * L + j];
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
                                   int64_t* tok, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(logits && score_in && ids_in && score_out && ids_out && parent && tok, "beam_update: null pointer");
  GPV_REQUIRE(B >= 0 && K >= 1 && K <= kMaxBeams && V >= K && t >= 0 && t + 1 < L, "beam_update: needs 1 <= K <= %d <= V and t + 1 < L (K %d, V %d, t %d, L %d)",
              kMaxBeams, K, V, t, L);
  GPV_REQUIRE(ids_in != ids_out && score_in != score_out, "beam_update: in-place update is not supported (double-buffer ids and scores)");
  if (B == 0) return GPV_OK;
  launch_k(beam_update_kernel, dim3(B), dim3(kDecThreads), 0, (cudaStream_t)stream, logits, (long long)ld, K, V, t, L, score_in, ids_in, score_out,
           ids_out, parent, tok);
  return check_launch("beam_update_kernel");
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
