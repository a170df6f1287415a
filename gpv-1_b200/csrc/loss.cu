// Criterion kernels: token cross-entropy over the vocabulary (losses.py:20-26) and the DETR set criterion on the
// Hungarian-matched pairs (utils/set_criterion.py:44-62, 78-97), forward and backward fused.
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

// ------------------------------------------------------------------------------------------------
// One CTA per (sample, position) row of fp32 logits [rows][V]:
//   loss += w[row] * (logsumexp(logits) - logits[target]);   dlogits = w[row] * (softmax - onehot)  (bf16)
// w[row] carries the reference's reduction: CE(reduction='none').mean(0).sum() per task times its loss weight
// (losses.py:26, 170-174) -> w = loss_wt(task) / (#samples of that task); rows of samples without an answer get 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_fwd_bwd_kernel(const float* __restrict__ logits, long long ldl,
                                                         const int64_t* __restrict__ targets, const float* __restrict__ w,
                                                         float* __restrict__ loss_sum, float* __restrict__ row_loss,
                                                         bf16* __restrict__ dlogits, long long ldd, int V) {
  pdl_sync();
  __shared__ float red[8];
  __shared__ float bcast;
  const long long row = blockIdx.x;
  const float* lr = logits + row * ldl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) mx = fmaxf(mx, lr[i]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) s += __expf(lr[i] - mx);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = t;
  }
  __syncthreads();
  s = bcast;
  const float wr = w[row];
  const int tgt = (int)targets[row];
  if (threadIdx.x == 0) {
    const float l = (logf(s) + mx) - lr[tgt];
    if (row_loss) row_loss[row] = l;
    if (wr != 0.f) atomicAdd(loss_sum, wr * l);
  }
  if (dlogits != nullptr) {
    const float inv = wr / s;
    bf16* dr = dlogits + row * ldd;
    for (int i = threadIdx.x; i < V; i += 256) {
      float g = __expf(lr[i] - mx) * inv;
      if (i == tgt) g -= wr;
      dr[i] = __float2bfloat16(g);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Set criterion on matched pairs.  One CTA per image.
//   loss_ce   = sum_{b in loc, q} wc * nll(logits[b,q], c) / sum wc,  c = 0 if q matched else 1, w = [1, eos_coef]
//   loss_bbox = sum_matched |box - tgt|_1 / num_boxes ;  loss_giou = sum_matched (1 - GIoU) / num_boxes
// out[0..2] accumulate the three (unweighted) losses; gradients are scaled by the loss weights wt_*:
//   dlogits [B,Q,ldl] fp32 (written), dbox_pre [B,Q,ldb] bf16 = d total / d(pre-sigmoid box) (written).
// ------------------------------------------------------------------------------------------------
struct GiouGrad {
  float giou;
  float d[4];  // d giou / d (x0, y0, x1, y1) of box a
};

GPV_DEVINL GiouGrad giou_with_grad(float ax0, float ay0, float ax1, float ay1, float bx0, float by0, float bx1, float by1) {
  GiouGrad r;
  const float aw = ax1 - ax0, ah = ay1 - ay0;
  const float area1 = aw * ah, area2 = (bx1 - bx0) * (by1 - by0);
  // d area1
  const float dA[4] = {-ah, -aw, ah, aw};
  const float ix0 = fmaxf(ax0, bx0), iy0 = fmaxf(ay0, by0), ix1 = fminf(ax1, bx1), iy1 = fminf(ay1, by1);
  const float iw = fmaxf(ix1 - ix0, 0.f), ih = fmaxf(iy1 - iy0, 0.f);
  const float inter = iw * ih;
  float dI[4] = {0.f, 0.f, 0.f, 0.f};
  if (ix1 - ix0 > 0.f && iy1 - iy0 > 0.f) {
    dI[0] = (ax0 > bx0) ? -ih : 0.f;
    dI[2] = (ax1 < bx1) ? ih : 0.f;
    dI[1] = (ay0 > by0) ? -iw : 0.f;
    dI[3] = (ay1 < by1) ? iw : 0.f;
  }
  const float uni = area1 + area2 - inter;
  const float ex0 = fminf(ax0, bx0), ey0 = fminf(ay0, by0), ex1 = fmaxf(ax1, bx1), ey1 = fmaxf(ay1, by1);
  const float ew = fmaxf(ex1 - ex0, 0.f), eh = fmaxf(ey1 - ey0, 0.f);
  const float earea = ew * eh;
  float dE[4];
  dE[0] = (ax0 < bx0) ? -eh : 0.f;
  dE[2] = (ax1 > bx1) ? eh : 0.f;
  dE[1] = (ay0 < by0) ? -ew : 0.f;
  dE[3] = (ay1 > by1) ? ew : 0.f;
  const float iou = inter / uni;
  r.giou = iou - (earea - uni) / earea;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dU = dA[k] - dI[k];
    const float diou = (dI[k] * uni - inter * dU) / (uni * uni);
    // d[-(E-U)/E] = d[U/E] = (dU*E - U*dE)/E^2
    r.d[k] = diou + (dU * earea - uni * dE[k]) / (earea * earea);
  }
  return r;
}

__global__ void __launch_bounds__(128) set_criterion_kernel(const float* __restrict__ logits, long long ldl,
                                                            const float* __restrict__ boxes, long long ldb,
                                                            const float* __restrict__ tboxes, const int32_t* __restrict__ toff,
                                                            const int64_t* __restrict__ idx_q, const int64_t* __restrict__ idx_t,
                                                            int Kmax, const uint8_t* __restrict__ loc_valid, int Q,
                                                            float eos_coef, float inv_wsum, float inv_num_boxes, float wt_ce,
                                                            float wt_bbox, float wt_giou, float* __restrict__ out,
                                                            float* __restrict__ dlogits, bf16* __restrict__ dbox_pre,
                                                            long long lddb) {
  pdl_sync();
  __shared__ int match_t[1024];
  __shared__ float red[3][4];
  __shared__ float s_norm[2];
  const int b = blockIdx.x;
  if (inv_wsum <= 0.0f) {
    // normalisers derived on the device from the ragged target offsets (set_criterion.py:163-168 num_boxes; the
    // weighted-mean denominator of F.cross_entropy with empty_weight = [1, eos_coef]) so that a captured CUDA graph
    // can be replayed with different targets
    if (threadIdx.x == 0) {
      int n_match = 0, n_loc = 0, sum_t = 0;
      for (int i = 0; i < (int)gridDim.x; ++i) {
        if (loc_valid[i]) {
          const int Ti = toff[i + 1] - toff[i];
          n_match += min(Q, Ti);
          sum_t += Ti;
          ++n_loc;
        }
      }
      const float ws = (float)n_match + eos_coef * (float)(n_loc * Q - n_match);
      s_norm[0] = 1.0f / fmaxf(ws, 1e-20f);
      s_norm[1] = 1.0f / (float)max(sum_t, 1);
    }
    __syncthreads();
    inv_wsum = s_norm[0];
    inv_num_boxes = s_norm[1];
  }
  const bool valid = loc_valid[b] != 0;
  const int t0 = toff[b], T = toff[b + 1] - t0;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) match_t[q] = -1;
  __syncthreads();
  if (valid) {
    const int K = min(Q, T);
    for (int k = threadIdx.x; k < K && k < Kmax; k += blockDim.x) {
      const int q = (int)idx_q[(long long)b * Kmax + k];
      if (q >= 0) match_t[q] = (int)idx_t[(long long)b * Kmax + k];
    }
  }
  __syncthreads();
  float l_ce = 0.f, l_l1 = 0.f, l_gi = 0.f;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    const long long r = (long long)b * Q + q;
    float dl0 = 0.f, dl1 = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float l0 = logits[r * ldl], l1 = logits[r * ldl + 1];
      const float mx = fmaxf(l0, l1);
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
      const float lse = logf(e0 + e1) + mx;
      const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
      const int mt = match_t[q];
      const int cls = mt >= 0 ? 0 : 1;
      const float wc = cls == 0 ? 1.0f : eos_coef;
      l_ce += wc * (lse - (cls == 0 ? l0 : l1));
      const float gs = wt_ce * wc * inv_wsum;
      dl0 = gs * (p0 - (cls == 0 ? 1.f : 0.f));
      dl1 = gs * (p1 - (cls == 1 ? 1.f : 0.f));
      if (mt >= 0) {
        const float* sb = boxes + r * ldb;
        const float* tb = tboxes + (long long)(t0 + mt) * 4;
        const float cx = sb[0], cy = sb[1], w = sb[2], h = sb[3];
        float gb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = sb[k] - tb[k];
          l_l1 += fabsf(d);
          gb[k] = wt_bbox * inv_num_boxes * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        const GiouGrad g = giou_with_grad(cx - 0.5f * w, cy - 0.5f * h, cx + 0.5f * w, cy + 0.5f * h, tb[0] - 0.5f * tb[2],
                                          tb[1] - 0.5f * tb[3], tb[0] + 0.5f * tb[2], tb[1] + 0.5f * tb[3]);
        l_gi += 1.0f - g.giou;
        const float sg = -wt_giou * inv_num_boxes;
        gb[0] += sg * (g.d[0] + g.d[2]);
        gb[1] += sg * (g.d[1] + g.d[3]);
        gb[2] += sg * 0.5f * (g.d[2] - g.d[0]);
        gb[3] += sg * 0.5f * (g.d[3] - g.d[1]);
#pragma unroll
        for (int k = 0; k < 4; ++k) db[k] = gb[k] * sb[k] * (1.0f - sb[k]);  // through the sigmoid
      }
    }
    dlogits[r * ldl] = dl0;
    dlogits[r * ldl + 1] = dl1;
#pragma unroll
    for (int k = 0; k < 4; ++k) dbox_pre[r * lddb + k] = __float2bfloat16(db[k]);
  }
  l_ce = warp_sum(l_ce);
  l_l1 = warp_sum(l_l1);
  l_gi = warp_sum(l_gi);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = l_ce;
    red[1][warp] = l_l1;
    red[2][warp] = l_gi;
  }
  __syncthreads();
  if (threadIdx.x == 0 && valid) {
    atomicAdd(out + 0, (red[0][0] + red[0][1] + red[0][2] + red[0][3]) * inv_wsum);
    atomicAdd(out + 1, (red[1][0] + red[1][1] + red[1][2] + red[1][3]) * inv_num_boxes);
    atomicAdd(out + 2, (red[2][0] + red[2][1] + red[2][2] + red[2][3]) * inv_num_boxes);
  }
}

}  // namespace gpv

using namespace gpv;

extern "C" int gpvb200_ce_fwd_bwd(const float* logits, int64_t ldl, const int64_t* targets, const float* row_weight,
                                  float* loss_sum, float* row_loss, void* dlogits, int64_t ldd, int32_t rows, int32_t V,
                                  void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(logits && targets && row_weight && loss_sum && rows >= 0 && V > 0, "ce: bad arguments");
  if (rows == 0) return GPV_OK;
  launch_k(ce_fwd_bwd_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, logits, ldl, targets, row_weight, loss_sum, row_loss, (bf16*)dlogits,
                                                            ldd, V);
  return check_launch("ce_fwd_bwd_kernel");
}

extern "C" int gpvb200_set_criterion(const float* logits, int64_t ldl, const float* boxes, int64_t ldb, const float* tgt_boxes,
                                     const int32_t* tgt_offsets, const int64_t* idx_q, const int64_t* idx_t, int32_t Kmax,
                                     const uint8_t* loc_valid, int32_t B, int32_t Q, float eos_coef, float weight_sum,
                                     float num_boxes, float wt_ce, float wt_bbox, float wt_giou, float* out3, float* dlogits,
                                     void* dbox_pre, int64_t lddb, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(logits && boxes && tgt_offsets && loc_valid && out3 && dlogits && dbox_pre, "set_criterion: null pointer");
  GPV_REQUIRE(B >= 0 && Q > 0 && Q <= 1024 && ((weight_sum > 0.f && num_boxes > 0.f) || (weight_sum <= 0.f && num_boxes <= 0.f)),
              "set_criterion: bad arguments");
  GPV_REQUIRE(Kmax == 0 || (idx_q && idx_t && tgt_boxes), "set_criterion: matches without indices");
  if (B == 0) return GPV_OK;
  launch_k(set_criterion_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, logits, ldl, boxes, ldb, tgt_boxes, tgt_offsets, idx_q, idx_t, Kmax,
                                                            loc_valid, Q, eos_coef, weight_sum > 0.f ? 1.0f / weight_sum : -1.0f,
                                                            weight_sum > 0.f ? 1.0f / num_boxes : -1.0f, wt_ce,
                                                            wt_bbox, wt_giou, out3, dlogits, (bf16*)dbox_pre, lddb);
  return check_launch("set_criterion_kernel");
}
