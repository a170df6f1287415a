// LayerNorm forward / backward over the last dimension, one warp per row, shuffle reductions, 16-byte vector access.
// Reference: nn.LayerNorm eps=1e-5 (transformer.py:139-140,199-201; nn.TransformerDecoderLayer gpv.py:38-43),
// BertLayerNorm eps=1e-12 (vilbert.py:296-316, same biased-variance formula), F.layer_norm without affine
// (detr_roi_head.py:91).  The residual add that precedes every LayerNorm is fused into the producing GEMM's
// epilogue, so the input here is already x + sublayer(x).
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

template <int MAXC>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x, long long ldx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, bf16* __restrict__ y, long long ldy,
                                                            float* __restrict__ stats, int M, int D) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = D >> 3;
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
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + ch * 8) = u;
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat;  dbeta += dy
template <int MAXC, bool AFFINE>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, long long lddy,
                                                            const bf16* __restrict__ x, long long ldx,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            bf16* __restrict__ dx, long long lddx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int D) {
  pdl_sync();
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

}  // namespace gpv

using namespace gpv;

extern "C" int gpvb200_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y,
                                     int64_t ldy, float* stats, int32_t M, int32_t D, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && y && M >= 0 && D > 0, "layernorm_fwd: bad arguments");
  GPV_REQUIRE(D % 8 == 0 && D <= 2304 && ldx % 8 == 0 && ldy % 8 == 0, "layernorm_fwd: D must be a multiple of 8, <= 2304");
  GPV_REQUIRE((gamma == nullptr) == (beta == nullptr), "layernorm_fwd: gamma and beta go together");
  if (M == 0) return GPV_OK;
  const int grid = (M + 7) / 8 < 148 * 8 ? (M + 7) / 8 : 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (D <= 768)
    launch_k(layernorm_fwd_kernel<3>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D);
  else
    launch_k(layernorm_fwd_kernel<9>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, gamma, beta, eps, (bf16*)y, ldy, stats, M, D);
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int gpvb200_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* stats,
                                     const float* gamma, void* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M,
                                     int32_t D, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(dy && x && stats && dx && M >= 0 && D > 0, "layernorm_bwd: bad arguments");
  GPV_REQUIRE(D % 8 == 0 && D <= 2304 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "layernorm_bwd: bad D / strides");
  GPV_REQUIRE(gamma == nullptr || (dgamma && dbeta), "layernorm_bwd: affine needs dgamma/dbeta");
  GPV_REQUIRE(gamma == nullptr || D <= 768, "layernorm_bwd: affine path supports D <= 768");
  if (M == 0) return GPV_OK;
  const int grid = (M + 7) / 8 < 148 * 2 ? (M + 7) / 8 : 148 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = gamma ? (size_t)2 * D * sizeof(float) : 0;
  if (gamma != nullptr)
    launch_k(layernorm_bwd_kernel<3, true>, dim3(grid), dim3(256), smem, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, stats, gamma, (bf16*)dx,
                                                           lddx, dgamma, dbeta, M, D);
  else if (D <= 768)
    launch_k(layernorm_bwd_kernel<3, false>, dim3(grid), dim3(256), smem, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, stats, gamma, (bf16*)dx,
                                                            lddx, dgamma, dbeta, M, D);
  else
    launch_k(layernorm_bwd_kernel<9, false>, dim3(grid), dim3(256), smem, st, (const bf16*)dy, lddy, (const bf16*)x, ldx, stats, gamma, (bf16*)dx,
                                                            lddx, dgamma, dbeta, M, D);
  return check_launch("layernorm_bwd_kernel");
}
