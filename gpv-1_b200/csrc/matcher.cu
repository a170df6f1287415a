// Hungarian matcher on the device: pairwise cost blocks and the rectangular assignment solver.
// Replaces utils/matcher.py:53-76 (softmax / cdist / generalized_box_iou on the flattened [B*Q, sum T] matrix,
// .cpu(), scipy.optimize.linear_sum_assignment per image) with two kernels and no host round trip.
#include <math.h>

#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

// ------------------------------------------------------------------------------------------------------
// Cost blocks.  One CTA per image; threads stride over (q, t).  fp32, explicit _rn intrinsics so the compiler
// cannot contract a*b+c into an FMA: the reference evaluates every step as a separate tensor op
// (matcher.py:56-72, box_ops.py:9-59), and near-tie assignments depend on those roundings.
// ------------------------------------------------------------------------------------------------------
__global__ void matcher_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                    const float* __restrict__ tboxes, const int64_t* __restrict__ tlabels,
                                    const int32_t* __restrict__ toff, int Q, int C, int Tmax, long long ldl, long long ldb,
                                    float w_class, float w_bbox, float w_giou, float* __restrict__ cost) {
  pdl_sync();
  const int b = blockIdx.x;
  const int t0 = toff[b], T = toff[b + 1] - t0;
  for (int idx = threadIdx.x; idx < Q * T; idx += blockDim.x) {
    const int q = idx / T, t = idx % T;
    // --- class term: -softmax(logits)[label]   (matcher.py:56, 63) ; x * (1/sum) like ATen's CPU softmax
    const float* lg = logits + ((size_t)b * Q + q) * ldl;
    float mx = lg[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, lg[c]);
    float sum = 0.0f;
    for (int c = 0; c < C; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(lg[c], mx)));
    const int lab = (int)tlabels[t0 + t];
    const float prob = __fmul_rn(expf(__fsub_rn(lg[lab], mx)), __fdiv_rn(1.0f, sum));
    const float cost_class = -prob;
    // --- L1 term: cdist(p=1)                   (matcher.py:66)
    const float4 ob = *reinterpret_cast<const float4*>(boxes + ((size_t)b * Q + q) * ldb);
    const float4 tb = *reinterpret_cast<const float4*>(tboxes + (size_t)(t0 + t) * 4);
    float l1 = fabsf(__fsub_rn(ob.x, tb.x));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(ob.y, tb.y)));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(ob.z, tb.z)));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(ob.w, tb.w)));
    // --- GIoU term                             (box_ops.py:9-13, 24-37, 40-59)
    const float ax0 = __fsub_rn(ob.x, __fmul_rn(0.5f, ob.z)), ay0 = __fsub_rn(ob.y, __fmul_rn(0.5f, ob.w));
    const float ax1 = __fadd_rn(ob.x, __fmul_rn(0.5f, ob.z)), ay1 = __fadd_rn(ob.y, __fmul_rn(0.5f, ob.w));
    const float bx0 = __fsub_rn(tb.x, __fmul_rn(0.5f, tb.z)), by0 = __fsub_rn(tb.y, __fmul_rn(0.5f, tb.w));
    const float bx1 = __fadd_rn(tb.x, __fmul_rn(0.5f, tb.z)), by1 = __fadd_rn(tb.y, __fmul_rn(0.5f, tb.w));
    const float area1 = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
    const float area2 = __fmul_rn(__fsub_rn(bx1, bx0), __fsub_rn(by1, by0));
    const float iw = fmaxf(__fsub_rn(fminf(ax1, bx1), fmaxf(ax0, bx0)), 0.0f);
    const float ih = fmaxf(__fsub_rn(fminf(ay1, by1), fmaxf(ay0, by0)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(area1, area2), inter);
    const float iou = __fdiv_rn(inter, uni);
    const float ew = fmaxf(__fsub_rn(fmaxf(ax1, bx1), fminf(ax0, bx0)), 0.0f);
    const float eh = fmaxf(__fsub_rn(fmaxf(ay1, by1), fminf(ay0, by0)), 0.0f);
    const float earea = __fmul_rn(ew, eh);
    const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(earea, uni), earea));
    const float cost_giou = -giou;
    // --- C = w_bbox*L1 + w_class*class + w_giou*giou, left to right   (matcher.py:72)
    float c = __fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, cost_class));
    c = __fadd_rn(c, __fmul_rn(w_giou, cost_giou));
    cost[((size_t)b * Q + q) * Tmax + t] = c;
  }
}

// ------------------------------------------------------------------------------------------------------
// Rectangular LSAP, one warp per image.  Shortest augmenting path with the column scan order, the
// "prefer an unassigned column among equal minima" rule and the dual updates of scipy's rectangular_lsap
// (Crouse 2016), in float64, so that ties resolve to the same assignment scipy returns.
// The scan over the remaining columns is split across the 32 lanes; the sequential scan's choice
//   index = last unassigned column among the minima if one exists, else the first minimum
// is reproduced with two warp reductions.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) lsap_kernel(const float* __restrict__ cost, const int32_t* __restrict__ toff,
                                                  int Q, int Tmax, int Kmax, int64_t* __restrict__ out_q,
                                                  int64_t* __restrict__ out_t) {
  pdl_sync();
  extern __shared__ double smem_d[];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int T = toff[b + 1] - toff[b];
  const bool transposed = T < Q;
  const int nr = transposed ? T : Q;  // rows of the solved problem
  const int nc = transposed ? Q : T;
  const int NMAX = Q > Tmax ? Q : Tmax;
  double* u = smem_d;
  double* v = u + NMAX;
  double* spc = v + NMAX;
  int* path = (int*)(spc + NMAX);
  int* col4row = path + NMAX;
  int* row4col = col4row + NMAX;
  int* remaining = row4col + NMAX;
  unsigned char* SR = (unsigned char*)(remaining + NMAX);
  unsigned char* SC = SR + NMAX;
  const float* cb = cost + (size_t)b * Q * Tmax;
  int64_t* oq = out_q + (size_t)b * Kmax;
  int64_t* ot = out_t + (size_t)b * Kmax;

  for (int k = lane; k < Kmax; k += 32) {
    oq[k] = -1;
    ot[k] = -1;
  }
  if (nr == 0) return;

  for (int i = lane; i < nr; i += 32) {
    u[i] = 0.0;
    col4row[i] = -1;
  }
  for (int j = lane; j < nc; j += 32) {
    v[j] = 0.0;
    row4col[j] = -1;
  }
  __syncwarp();
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  for (int cur = 0; cur < nr; ++cur) {
    // ---- augmenting_path(cur)
    double minVal = 0.0;
    int num_remaining = nc;
    for (int j = lane; j < nc; j += 32) {
      remaining[j] = nc - j - 1;
      SC[j] = 0;
      spc[j] = INF;
    }
    for (int i = lane; i < nr; i += 32) SR[i] = 0;
    __syncwarp();
    int sink = -1;
    int i = cur;
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      double lmin = INF;
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = remaining[it];
        const double cij = (double)(transposed ? cb[(size_t)j * Tmax + i] : cb[(size_t)i * Tmax + j]);
        const double r = ((minVal + cij) - ui) - v[j];
        double s = spc[j];
        if (r < s) {
          path[j] = i;
          spc[j] = r;
          s = r;
        }
        lmin = s < lmin ? s : lmin;
      }
      double m = lmin;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other < m ? other : m;
      }
      __syncwarp();
      int firstMin = 0x7fffffff, lastFree = -1;
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = remaining[it];
        if (spc[j] == m) {
          firstMin = it < firstMin ? it : firstMin;
          if (row4col[j] == -1) lastFree = it > lastFree ? it : lastFree;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        firstMin = min(firstMin, __shfl_xor_sync(0xffffffffu, firstMin, o));
        lastFree = max(lastFree, __shfl_xor_sync(0xffffffffu, lastFree, o));
      }
      minVal = m;
      if (m == INF) {  // infeasible (only with inf/NaN costs): leave -1 markers
        return;
      }
      const int index = lastFree >= 0 ? lastFree : firstMin;
      const int j = remaining[index];
      const int r4c = row4col[j];
      if (r4c == -1) sink = j; else i = r4c;
      __syncwarp();
      if (lane == 0) {
        SC[j] = 1;
        remaining[index] = remaining[num_remaining - 1];
      }
      --num_remaining;
      __syncwarp();
    }
    // ---- dual update
    if (lane == 0) u[cur] += minVal;
    for (int r = lane; r < nr; r += 32)
      if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
    for (int j = lane; j < nc; j += 32)
      if (SC[j]) v[j] -= minVal - spc[j];
    __syncwarp();
    // ---- augment along the path
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int tmp = col4row[r];
        col4row[r] = j;
        j = tmp;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }

  // ---- emit (query index, target index) sorted by query index, like scipy
  if (!transposed) {
    for (int r = lane; r < nr; r += 32) {
      oq[r] = r;
      ot[r] = col4row[r];
    }
  } else {
    for (int r = lane; r < nr; r += 32) {
      const int qv = col4row[r];
      int rank = 0;
      for (int k = 0; k < nr; ++k) rank += (col4row[k] < qv) ? 1 : 0;
      oq[rank] = qv;
      ot[rank] = r;
    }
  }
}

}  // namespace gpv

using namespace gpv;

extern "C" int gpvb200_matcher_cost(const float* logits, int64_t ldl, const float* boxes, int64_t ldb, const float* tgt_boxes,
                                    const int64_t* tgt_labels, const int32_t* tgt_offsets, int32_t B, int32_t Q,
                                    int32_t C, int32_t Tmax, float w_class, float w_bbox, float w_giou, float* cost,
                                    void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(B >= 0 && Q > 0 && C > 0 && Tmax >= 0, "matcher_cost: bad shape");
  if (B == 0 || Tmax == 0) return GPV_OK;
  GPV_REQUIRE(logits && boxes && tgt_boxes && tgt_labels && tgt_offsets && cost, "matcher_cost: null pointer");
  GPV_REQUIRE(ldl >= C && ldb >= 4 && ldb % 4 == 0 && ((uintptr_t)boxes & 15) == 0, "matcher_cost: bad row strides");
  launch_k(matcher_cost_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, logits, boxes, tgt_boxes, tgt_labels, tgt_offsets, Q, C, Tmax, ldl, ldb,
                                                           w_class, w_bbox, w_giou, cost);
  return check_launch("matcher_cost_kernel");
}

extern "C" int gpvb200_lsap(const float* cost, const int32_t* tgt_offsets, int32_t B, int32_t Q, int32_t Tmax,
                            int64_t* out_q, int64_t* out_t, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(B >= 0 && Q > 0 && Tmax >= 0, "lsap: bad shape");
  const int Kmax = Q < Tmax ? Q : Tmax;
  if (B == 0 || Kmax == 0) return GPV_OK;
  GPV_REQUIRE(cost && tgt_offsets && out_q && out_t, "lsap: null pointer");
  const int NMAX = Q > Tmax ? Q : Tmax;
  const size_t smem = (size_t)NMAX * (3 * sizeof(double) + 4 * sizeof(int) + 2);
  GPV_REQUIRE(smem <= 200 * 1024, "lsap: problem of size %d too large for shared memory", NMAX);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_last_error("lsap: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return GPV_ERR_CUDA;
    }
  }
  launch_k(lsap_kernel, dim3(B), dim3(32), smem, (cudaStream_t)stream, cost, tgt_offsets, Q, Tmax, Kmax, out_q, out_t);
  return check_launch("lsap_kernel");
}
