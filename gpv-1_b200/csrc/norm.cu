// LayerNorm forward / backward over the last dimension, one warp per row, shuffle reductions, 16-byte vector access.
// Reference: nn.LayerNorm eps=1e-5 (transformer.py:139-140,199-201; nn.TransformerDecoderLayer gpv.py:38-43),
// BertLayerNorm eps=1e-12 (vilbert.py:296-316, same biased-variance formula), F.layer_norm without affine
// (detr_roi_head.py:91).  The residual add that precedes every LayerNorm is fused into the producing GEMM's
// epilogue, so the input here is already x + sublayer(x).
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

template <int MAXC, bool DROP>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x, long long ldx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, bf16* __restrict__ y, long long ldy,
                                                            float* __restrict__ stats, int M, int D, const DropArgs dr) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = D >> 3;
  const uint32_t dkey = DROP ? drop_key(*dr.seed, dr.site) : 0u;   // y = dropout(LN(x)) (BERT embeddings, vilbert.py:364)
  for (long long row = (long long)blockIdx.x * 8 + warp; row < M; row += (long long)gridDim.x * 8) {
    const bf16* xr = x + row * ldx;
    float v[MAXC][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nch) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + ch * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[c][2 * j] = f.x;
          v[c][2 * j + 1] = f.y;
          s += f.x + f.y;
        }
      }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (lane + c * 32 < nch) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[c][j] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (stats != nullptr && lane == 0) {
      stats[row * 2] = mean;
      stats[row * 2 + 1] = rstd;
    }
    bf16* yr = y + row * ldy;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nch) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = (v[c][j] - mean) * rstd;
          if (gamma != nullptr) t = t * __ldg(gamma + ch * 8 + j) + __ldg(beta + ch * 8 + j);
          o[j] = t;
        }
        if (DROP) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            drop_pair(o[2 * j], o[2 * j + 1], dkey, (uint32_t)row * (uint32_t)(D >> 1) + ch * 4 + j, dr.thresh16, dr.scale);
        }
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + ch * 8) = u;
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat;  dbeta += dy
template <int MAXC, bool AFFINE, bool DXM>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, long long lddy,
                                                            const bf16* __restrict__ x, long long ldx,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            bf16* __restrict__ dx, long long lddx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int D, bf16* __restrict__ dxm,
                                                            long long lddxm, const DropArgs dr) {
  pdl_sync();
  // dxm (optional) = dx (*) mask / (1 - p): the gradient of the dropped-out sub-layer output in y = LN(res + dropout(f)),
  // regenerated from the forward's (seed, site); dx itself is the gradient of the residual branch.
  const uint32_t dkey = DXM ? drop_key(*dr.seed, dr.site) : 0u;
  extern __shared__ float red[];  // [2][D] when affine
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = D >> 3;
  constexpr bool affine = AFFINE;
  if (affine) {
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
  }
  float ag[MAXC][8], ab[MAXC][8];
  if (affine) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) ag[c][j] = ab[c][j] = 0.f;
  }
  for (long long row = (long long)blockIdx.x * 8 + warp; row < M; row += (long long)gridDim.x * 8) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float g[MAXC][8], xh[MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nch) {
        const uint4 ux = *reinterpret_cast<const uint4*>(x + row * ldx + ch * 8);
        const uint4 ud = *reinterpret_cast<const uint4*>(dy + row * lddy + ch * 8);
        const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fx = unpack_bf16x2(wx[j]), fd = unpack_bf16x2(wd[j]);
          const float h0 = (fx.x - mean) * rstd, h1 = (fx.y - mean) * rstd;
          float g0 = fd.x, g1 = fd.y;
          if (affine) {
            ag[c][2 * j] += fd.x * h0;
            ag[c][2 * j + 1] += fd.y * h1;
            ab[c][2 * j] += fd.x;
            ab[c][2 * j + 1] += fd.y;
            g0 *= __ldg(gamma + ch * 8 + 2 * j);
            g1 *= __ldg(gamma + ch * 8 + 2 * j + 1);
          }
          g[c][2 * j] = g0; g[c][2 * j + 1] = g1;
          xh[c][2 * j] = h0; xh[c][2 * j + 1] = h1;
          s1 += g0 + g1;
          s2 += g0 * h0 + g1 * h1;
        }
      }
    }
    const float c1 = warp_sum(s1) / (float)D, c2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nch) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[c][j] - c1 - xh[c][j] * c2);
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + row * lddx + ch * 8) = u;
        if (DXM) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            drop_pair(o[2 * j], o[2 * j + 1], dkey, (uint32_t)row * (uint32_t)(D >> 1) + ch * 4 + j, dr.thresh16, dr.scale);
          u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dxm + row * lddxm + ch * 8) = u;
        }
      }
    }
  }
  if (affine) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nch) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&red[ch * 8 + j], ag[c][j]);
          atomicAdd(&red[D + ch * 8 + j], ab[c][j]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(dgamma + i, red[i]);
      atomicAdd(dbeta + i, red[D + i]);
    }
  }
}

// The same for D <= 256 (DETR encoder / decoder: one 16-byte chunk per lane), FOUR rows per warp iteration: all of a warp's loads
// (x, dy, statistics of four rows) are in flight before the first reduction, which is what bounds this kernel -- it sits on the
// critical path of the backward pass (two per transformer layer) and moves only 20 MB per launch.
template <bool AFFINE, bool DXM>
__global__ void __launch_bounds__(256) layernorm_bwd_d256_kernel(const bf16* __restrict__ dy, long long lddy,
                                                                 const bf16* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                 bf16* __restrict__ dx, long long lddx, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, int M, int D, bf16* __restrict__ dxm,
                                                                 long long lddxm, const DropArgs dr) {
  pdl_sync();
  constexpr int R = 4;
  const uint32_t dkey = DXM ? drop_key(*dr.seed, dr.site) : 0u;
  __shared__ float red[2 * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = lane < (D >> 3);
  float gam[8], ag[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gam[j] = (AFFINE && live) ? __ldg(gamma + lane * 8 + j) : 1.0f;
    ag[j] = ab[j] = 0.f;
  }
  if (AFFINE) {
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
  }
  const float invD = 1.0f / (float)D;
  for (long long row0 = ((long long)blockIdx.x * 8 + warp) * R; row0 < M; row0 += (long long)gridDim.x * 8 * R) {
    uint4 ux[R], ud[R];
    float2 st[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      const bool ok = row < M && live;
      ux[r] = ok ? *reinterpret_cast<const uint4*>(x + row * ldx + lane * 8) : make_uint4(0, 0, 0, 0);
      ud[r] = ok ? *reinterpret_cast<const uint4*>(dy + row * lddy + lane * 8) : make_uint4(0, 0, 0, 0);
      st[r] = row < M ? *reinterpret_cast<const float2*>(stats + row * 2) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      if (row >= M) break;                                       // warp-uniform
      const float mean = st[r].x, rstd = st[r].y;
      const uint32_t wx[4] = {ux[r].x, ux[r].y, ux[r].z, ux[r].w}, wd[4] = {ud[r].x, ud[r].y, ud[r].z, ud[r].w};
      float g[8], xh[8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fx = unpack_bf16x2(wx[j]), fd = unpack_bf16x2(wd[j]);
        const float h0 = live ? (fx.x - mean) * rstd : 0.f, h1 = live ? (fx.y - mean) * rstd : 0.f;
        if (AFFINE) {
          ag[2 * j] += fd.x * h0;
          ag[2 * j + 1] += fd.y * h1;
          ab[2 * j] += fd.x;
          ab[2 * j + 1] += fd.y;
        }
        g[2 * j] = fd.x * gam[2 * j];
        g[2 * j + 1] = fd.y * gam[2 * j + 1];
        xh[2 * j] = h0;
        xh[2 * j + 1] = h1;
        s1 += g[2 * j] + g[2 * j + 1];
        s2 += g[2 * j] * h0 + g[2 * j + 1] * h1;
      }
      const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
      if (live) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] - c1 - xh[j] * c2);
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + row * lddx + lane * 8) = u;
        if (DXM) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            drop_pair(o[2 * j], o[2 * j + 1], dkey, (uint32_t)row * (uint32_t)(D >> 1) + lane * 4 + j, dr.thresh16, dr.scale);
          u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dxm + row * lddxm + lane * 8) = u;
        }
      }
    }
  }
  if (AFFINE) {
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[lane * 8 + j], ag[j]);
        atomicAdd(&red[D + lane * 8 + j], ab[j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(dgamma + i, red[i]);
      atomicAdd(dbeta + i, red[D + i]);
    }
  }
}

}  // namespace gpv

using namespace gpv;

static DropArgs make_drop(const void* seed, uint32_t site, float p) {
  DropArgs d;
  d.seed = (p > 0.f) ? (const unsigned long long*)seed : nullptr;
  d.site = site;
  d.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

extern "C" int gpvb200_layernorm_fwd_drop(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                          int64_t ldy, float* stats, int32_t M, int32_t D, const void* drop_seed,
                                          uint32_t drop_site, float drop_p, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && y && M >= 0 && D > 0, "layernorm_fwd: bad arguments");
  GPV_REQUIRE(D % 8 == 0 && D <= 2304 && ldx % 8 == 0 && ldy % 8 == 0, "layernorm_fwd: D must be a multiple of 8, <= 2304");
  GPV_REQUIRE((gamma == nullptr) == (beta == nullptr), "layernorm_fwd: gamma and beta go together");
  GPV_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed), "layernorm_fwd: bad dropout arguments");
  if (M == 0) return GPV_OK;
  const DropArgs dr = make_drop(drop_seed, drop_site, drop_p);
  const int grid = (M + 7) / 8 < 148 * 8 ? (M + 7) / 8 : 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (D <= 768 && dr.seed != nullptr)
    launch_k(layernorm_fwd_kernel<3, true>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D, dr);
  else if (D <= 768)
    launch_k(layernorm_fwd_kernel<3, false>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D, dr);
  else if (dr.seed != nullptr)
    launch_k(layernorm_fwd_kernel<9, true>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D, dr);
  else
    launch_k(layernorm_fwd_kernel<9, false>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D, dr);
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int gpvb200_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                     int64_t ldy, float* stats, int32_t M, int32_t D, void* stream) {
  return gpvb200_layernorm_fwd_drop(x, ldx, gamma, beta, eps, y, ldy, stats, M, D, nullptr, 0, 0.f, stream);
}

extern "C" int gpvb200_layernorm_bwd_drop(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* stats,
                                          const float* gamma, void* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M,
                                          int32_t D, void* dx_masked, int64_t lddxm, const void* drop_seed, uint32_t drop_site,
                                          float drop_p, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(dy && x && stats && dx && M >= 0 && D > 0, "layernorm_bwd: bad arguments");
  GPV_REQUIRE(D % 8 == 0 && D <= 2304 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "layernorm_bwd: bad D / strides");
  GPV_REQUIRE(gamma == nullptr || (dgamma && dbeta), "layernorm_bwd: affine needs dgamma/dbeta");
  GPV_REQUIRE(gamma == nullptr || D <= 768, "layernorm_bwd: affine path supports D <= 768");
  GPV_REQUIRE(dx_masked == nullptr || (drop_seed && drop_p > 0.f && drop_p < 1.f && lddxm % 8 == 0), "layernorm_bwd: bad dropout arguments");
  if (M == 0) return GPV_OK;
  const DropArgs dr = make_drop(drop_seed, drop_site, dx_masked ? drop_p : 0.f);
  bf16* dxm = (bf16*)dx_masked;
  const int grid = (M + 7) / 8 < 148 * 2 ? (M + 7) / 8 : 148 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = gamma ? (size_t)2 * D * sizeof(float) : 0;
#define GPV_LN_BWD(MAXC, AFF, DXM)                                                                                          \
  launch_k(layernorm_bwd_kernel<MAXC, AFF, DXM>, dim3(grid), dim3(256), smem, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, stats, \
           gamma, (bf16*)dx, lddx, dgamma, dbeta, M, D, dxm, lddxm, dr)
  if (D <= 256) {
    const int g4 = (M + 31) / 32 < 148 * 4 ? (M + 31) / 32 : 148 * 4;
#define GPV_LN_BWD4(AFF, DXM)                                                                                                   \
  launch_k(layernorm_bwd_d256_kernel<AFF, DXM>, dim3(g4), dim3(256), 0, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, stats, gamma, \
           (bf16*)dx, lddx, dgamma, dbeta, M, D, dxm, lddxm, dr)
    if (gamma != nullptr) {
      if (dxm) GPV_LN_BWD4(true, true); else GPV_LN_BWD4(true, false);
    } else {
      if (dxm) GPV_LN_BWD4(false, true); else GPV_LN_BWD4(false, false);
    }
#undef GPV_LN_BWD4
  } else if (gamma != nullptr) {
    if (dxm) GPV_LN_BWD(3, true, true); else GPV_LN_BWD(3, true, false);
  } else if (D <= 768) {
    if (dxm) GPV_LN_BWD(3, false, true); else GPV_LN_BWD(3, false, false);
  } else {
    if (dxm) GPV_LN_BWD(9, false, true); else GPV_LN_BWD(9, false, false);
  }
#undef GPV_LN_BWD
  return check_launch("layernorm_bwd_kernel");
}

extern "C" int gpvb200_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* stats,
                                     const float* gamma, void* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M,
                                     int32_t D, void* stream) {
  return gpvb200_layernorm_bwd_drop(dy, lddy, x, ldx, stats, gamma, dx, lddx, dgamma, dbeta, M, D, nullptr, 0, nullptr, 0, 0.f, stream);
}

// Test / debug export of the dropout mask of a logical [rows, N] tensor: out[row*N + col] = 1 (kept) or 0.
namespace gpv {
__global__ void dropout_mask_kernel(uint8_t* __restrict__ out, long long rows, int N, const DropArgs dr) {
  const uint32_t key = drop_key(*dr.seed, dr.site);
  const long long total = rows * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / N;
    const int col = (int)(i % N);
    const uint32_t bits = drop_bits(key, (uint32_t)row * (uint32_t)((N + 1) >> 1) + (uint32_t)(col >> 1));
    out[i] = (((col & 1) ? (bits >> 16) : (bits & 0xFFFFu)) >= dr.thresh16) ? 1 : 0;
  }
}
}  // namespace gpv

extern "C" int gpvb200_dropout_mask(uint8_t* out, int64_t rows, int32_t N, const void* drop_seed, uint32_t drop_site, float drop_p,
                                    void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(out && drop_seed && rows >= 0 && N > 0 && drop_p > 0.f && drop_p < 1.f, "dropout_mask: bad arguments");
  if (rows == 0) return GPV_OK;
  long long blocks = (rows * N + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_mask_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, rows, N, make_drop(drop_seed, drop_site, drop_p));
  return check_launch("dropout_mask_kernel");
}
