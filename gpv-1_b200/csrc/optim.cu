// Fused gradient clipping + AdamW over the flat gradient arena (multi-tensor apply).
//
// Replaces, in the reference's training step (exp/gpv/train_distr.py:414-428):
//     torch.nn.utils.clip_grad_norm_(params['detr_backbone'] + params['detr_head'], cfg.training.clip_max_norm)
//     optimizer.step()          # torch.optim.AdamW, four parameter groups (train_distr.py:228-253)
// torch runs these as ~400 per-tensor norms + a stack/norm + ~400 multiplies + the foreach AdamW kernels; here the
// engine already owns every gradient in ONE fp32 arena (model/engine.py), so the step is two launches:
//   1. grad_sqnorm:  sum of squares of the clipped subset -> one device scalar (fp32 atomics per CTA)
//   2. clip_adamw:   g *= min(1, max_norm / (sqrt(total) + 1e-6)) on the clipped subset (written back, as
//                    clip_grad_norm_ does), then the decoupled-weight-decay Adam update of p, m, v
// HBM-bound: 28 bytes per parameter (read g, m, v, p; write m, v, p) + 4 for the written-back clipped gradients.
// Arithmetic follows torch.optim.AdamW (single-tensor form, amsgrad off, maximize off):
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),   bc1 = 1 - b1^t, bc2 = 1 - b2^t.
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

struct OptItem {
  float* p;         // parameter tensor (fp32 master, owned by PyTorch)
  long long goff;   // offset of its gradient in the arena (elements); m and v use the same offset in their arenas
  int n;            // elements
  int group;        // learning-rate group 0..3
  int clip;         // 1: member of the clipped subset
  int step0;        // optimizer step at which this tensor's state started: its Adam step is (global step - step0), as torch keeps
                    // state['step'] per parameter (parameters unfrozen for the second training phase start their bias correction at 1)
};
constexpr int kOptChunk = 4096;

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const OptItem* __restrict__ items, const int* __restrict__ blk_item,
                                                          const int* __restrict__ blk_chunk, const float* __restrict__ grads,
                                                          float* __restrict__ out_sq) {
  __shared__ float red[8];
  const OptItem it = items[blk_item[blockIdx.x]];
  const int base = blk_chunk[blockIdx.x] * kOptChunk;
  const int end = min(base + kOptChunk, it.n);
  const float* g = grads + it.goff;
  float acc = 0.f;
  if (((it.goff | base) & 3) == 0) {
    const int n4 = (end - base) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 x = g4[i];
      acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (int i = base + (n4 << 2) + threadIdx.x; i < end; i += 256) acc += g[i] * g[i];
  } else {
    for (int i = base + threadIdx.x; i < end; i += 256) acc += g[i] * g[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out_sq, t);
  }
}

struct AdamArgs {
  float lr[4];
  float beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm;   // bc1 / bc2_sqrt: for tensors with step0 = 0 (the common case)
  long long step;
};

GPV_DEVINL void adam_one(float& p, float& m, float& v, float g, float lr, const AdamArgs& a) {
  p *= 1.0f - lr * a.wd;
  m = a.beta1 * m + (1.0f - a.beta1) * g;
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (lr / a.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) clip_adamw_kernel(const OptItem* __restrict__ items, const int* __restrict__ blk_item,
                                                         const int* __restrict__ blk_chunk, float* __restrict__ grads,
                                                         float* __restrict__ ms, float* __restrict__ vs,
                                                         const float* __restrict__ total_sq, const AdamArgs a) {
  const OptItem it = items[blk_item[blockIdx.x]];
  const int base = blk_chunk[blockIdx.x] * kOptChunk;
  const int end = min(base + kOptChunk, it.n);
  float coef = 1.0f;
  if (it.clip && a.max_norm > 0.f) coef = fminf(1.0f, a.max_norm / (sqrtf(*total_sq) + 1e-6f));
  const float lr = a.lr[it.group];
  AdamArgs al = a;
  if (it.step0 != 0) {                                    // this tensor joined later: its own bias correction
    const double st = (double)(a.step - it.step0);
    al.bc1 = (float)(1.0 - pow((double)a.beta1, st));
    al.bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, st));
  }
  float* g = grads + it.goff;
  float* m = ms + it.goff;
  float* v = vs + it.goff;
  float* p = it.p;
  const bool vec = (((it.goff | base) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  int i0 = base + threadIdx.x;
  if (vec) {
    const int n4 = (end - base) >> 2;
    float4* g4 = reinterpret_cast<float4*>(g + base);
    float4* m4 = reinterpret_cast<float4*>(m + base);
    float4* v4 = reinterpret_cast<float4*>(v + base);
    float4* p4 = reinterpret_cast<float4*>(p + base);
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 gg = g4[i], mm = m4[i], vv = v4[i], pp = p4[i];
      if (coef != 1.0f) {
        gg.x *= coef; gg.y *= coef; gg.z *= coef; gg.w *= coef;
        g4[i] = gg;
      }
      adam_one(pp.x, mm.x, vv.x, gg.x, lr, al);
      adam_one(pp.y, mm.y, vv.y, gg.y, lr, al);
      adam_one(pp.z, mm.z, vv.z, gg.z, lr, al);
      adam_one(pp.w, mm.w, vv.w, gg.w, lr, al);
      m4[i] = mm;
      v4[i] = vv;
      p4[i] = pp;
    }
    i0 = base + (n4 << 2) + threadIdx.x;
  }
  for (int i = i0; i < end; i += 256) {
    float gg = g[i] * coef;
    if (coef != 1.0f) g[i] = gg;
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, mm, vv, gg, lr, al);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

}  // namespace gpv

using namespace gpv;

extern "C" size_t gpvb200_optim_item_size(void) { return sizeof(OptItem); }
extern "C" int gpvb200_optim_chunk(void) { return kOptChunk; }

extern "C" int gpvb200_grad_sqnorm(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks,
                                   const float* grads, float* out_sq, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(items && grads && out_sq && n_blocks >= 0 && (n_blocks == 0 || (blk_item && blk_chunk)), "grad_sqnorm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out_sq, 0, sizeof(float), st);
  if (e != cudaSuccess) {
    set_last_error("grad_sqnorm: memset failed: %s", cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  if (n_blocks == 0) return GPV_OK;
  grad_sqnorm_kernel<<<n_blocks, 256, 0, st>>>((const OptItem*)items, blk_item, blk_chunk, grads, out_sq);
  return check_launch("grad_sqnorm_kernel");
}

extern "C" int gpvb200_clip_adamw(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks,
                                  float* grads, float* m, float* v, const float* total_sq, float max_norm, float lr0, float lr1,
                                  float lr2, float lr3, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                  void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(items && blk_item && blk_chunk && grads && m && v && total_sq && n_blocks >= 0 && step >= 1, "clip_adamw: bad arguments");
  if (n_blocks == 0) return GPV_OK;
  AdamArgs a;
  a.lr[0] = lr0; a.lr[1] = lr1; a.lr[2] = lr2; a.lr[3] = lr3;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.max_norm = max_norm;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.step = step;
  clip_adamw_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>((const OptItem*)items, blk_item, blk_chunk, grads, m, v, total_sq, a);
  return check_launch("clip_adamw_kernel");
}
