// HBM-bound helper kernels of the GPV-1 hot path: weight packing (fp32 master -> bf16, FrozenBN fold, conv layout),
// position-embedding adds, bias-gradient column sums, max-pool, stem im2col, ROI-align weight construction,
// relevance conditioning, embedding gathers and row-remapped copies.  All vectorised to 16-byte accesses where the
// layout allows, grids sized in multiples of the SM count.
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

constexpr int kSMs = 148;

// ------------------------------------------------------------------------------------------------
// out[m][:] = x[m][:] + p[m % P][:]     (q = k = src + pos: transformer.py:153, 218, 223-224)
// ------------------------------------------------------------------------------------------------
__global__ void add_rowbcast_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ pe, long long ldp,
                                    bf16* __restrict__ out, long long ldo, long long M, int D, int P) {
  pdl_sync();
  const int nch = D >> 3;
  const long long total = M * nch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nch;
    const int c = (int)(i % nch);
    const uint4 b = *reinterpret_cast<const uint4*>(pe + (m % P) * ldp + c * 8);
    uint4 o = b;
    if (x != nullptr) {
      const uint4 a = *reinterpret_cast<const uint4*>(x + m * ldx + c * 8);
      const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
      uint32_t wo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = unpack_bf16x2(wa[j]), fb = unpack_bf16x2(wb[j]);
        wo[j] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
      }
      o = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
    *reinterpret_cast<uint4*>(out + m * ldo + c * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// out[n] += sum_m dy[m][n]   (bias gradients).  Each CTA reduces a slab of rows for 64 columns.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, long long ld, float* __restrict__ out,
                                                     long long M, int N, int rows_per_cta) {
  pdl_sync();
  __shared__ float red[4][64];
  const int col = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rgrp = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float acc = 0.f;
  if (col < N)
    for (long long r = r0 + rgrp; r < r1; r += 4) acc += __bfloat162float(dy[r * ld + col]);
  red[rgrp][threadIdx.x & 63] = acc;
  __syncthreads();
  if (threadIdx.x < 64 && col < N) atomicAdd(out + col, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// Vector form: a warp reads 32 x 16 bytes = 256 contiguous columns of one row per instruction, 8 warps stride the
// rows, four rows in flight per thread; the 8 partial rows meet in shared memory and leave as one atomic per column.
__global__ void __launch_bounds__(256) colsum_vec_kernel(const bf16* __restrict__ dy, long long ld, float* __restrict__ out,
                                                         long long M, int N, int rows_per_cta) {
  pdl_sync();
  __shared__ float red[8][256 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < N) {
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(dy + (r + 8 * u) * ld + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(dy + r * ld + c0));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

// out[s][d] += sum_b x[b*S + s][d]   (gradient of a parameter broadcast over the batch, e.g. query_embed)
__global__ void batch_reduce_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ out, int B, int S, int D) {
  pdl_sync();
  const long long total = (long long)S * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / D), d = (int)(i % D);
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += __bfloat162float(x[((long long)b * S + s) * ld + d]);
    out[i] += acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-tensor weight packing: fp32 master [O][I][taps] -> bf16 [taps][O][I] (scaled per output channel by the
// folded FrozenBN scale, backbone.py:44-54).  mode 1: 7x7 stem -> [O][152] with k = tap*3 + c, zero padded.
// ------------------------------------------------------------------------------------------------
struct PackItem {
  const float* src;
  bf16* dst;
  const float* scale;
  int O, I, taps, mode;
};
constexpr int kPackChunk = 4096;

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackItem* __restrict__ items, const int* __restrict__ blk_item,
                                                           const int* __restrict__ blk_chunk) {
  const PackItem it = items[blk_item[blockIdx.x]];
  const long long base = (long long)blk_chunk[blockIdx.x] * kPackChunk;
  const long long n = it.mode == 1 ? (long long)it.O * 152 : (long long)it.taps * it.O * it.I;
  const long long OI = (long long)it.O * it.I;
  for (int j = threadIdx.x; j < kPackChunk; j += 256) {
    const long long e = base + j;
    if (e >= n) break;
    float v;
    if (it.mode == 1) {
      const int o = (int)(e / 152), k = (int)(e % 152);
      if (k < 147) {
        const int t = k / 3, c = k % 3;
        v = it.src[((long long)o * 3 + c) * 49 + t];
        if (it.scale) v *= it.scale[o];
      } else {
        v = 0.f;
      }
    } else {
      const int t = (int)(e / OI);
      const long long r = e % OI;
      const int o = (int)(r / it.I), i = (int)(r % it.I);
      v = it.src[((long long)o * it.I + i) * it.taps + t];
      if (it.scale) v *= it.scale[o];
    }
    it.dst[e] = __float2bfloat16(v);
  }
}

// FrozenBN fold: scale = w * rsqrt(rv + 1e-5); bias = b - rm * scale   (backbone.py:44-54)
__global__ void bn_fold_kernel(const float* w, const float* b, const float* rm, const float* rv, float* scale, float* bias, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = w[i] * rsqrtf(rv[i] + 1e-5f);
    scale[i] = s;
    bias[i] = b[i] - rm[i] * s;
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 max-pool on NHWC bf16 (torchvision resnet stem, backbone.py:72)
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  pdl_sync();
  const int nch = C >> 3;
  const long long total = (long long)B * Ho * Wo * nch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % nch);
    long long pth = i / nch;
    const int wo = (int)(pth % Wo);
    pth /= Wo;
    const int ho = (int)(pth % Ho);
    const int b = (int)(pth / Ho);
    // max of bf16 values is exact in bf16: packed __hmax2, no unpacking
    __nv_bfloat162 m[4];
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ninf;
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * ho - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * wo - 1 + s;
        if (w < 0 || w >= W) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + h) * W + w) * C + c * 8));
        const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], e[j]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    *reinterpret_cast<uint4*>(y + (((long long)b * Ho + ho) * Wo + wo) * C + c * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Stem im2col: NCHW fp32 image -> bf16 [B*Ho*Wo][152], k = (r*7+s)*3 + c, 7x7 stride 2 pad 3.
// ------------------------------------------------------------------------------------------------
__global__ void stem_im2col_kernel(const float* __restrict__ img, bf16* __restrict__ col, int B, int H, int W, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * 19;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 19);
    long long pth = i / 19;
    const int wo = (int)(pth % Wo);
    pth /= Wo;
    const int ho = (int)(pth % Ho);
    const int b = (int)(pth / Ho);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      float t = 0.f;
      if (k < 147) {
        const int tap = k / 3, c = k % 3, r = tap / 7, s = tap % 7;
        const int h = 2 * ho - 3 + r, w = 2 * wo - 3 + s;
        if (h >= 0 && h < H && w >= 0 && w < W) t = __ldg(img + (((long long)b * 3 + c) * H + h) * W + w);
      }
      v[j] = t;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(col + ((((long long)b * Ho + ho) * Wo + wo) * 152 + ch * 8)) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Stem space-to-depth: NCHW fp32 image -> zero-bordered bf16 [B][Ho+4][Wo+4][16] with
//   out[b][I][J][dy*6 + dx*3 + c] = img[b][c][2(I-2)+dy][2(J-2)+dx]   (channels 12..15 and out-of-image taps = 0).
// The 7x7 stride-2 pad-3 stem convolution (backbone.py:72, torchvision resnet conv1) is then a 4x4 stride-1
// convolution over this map: input row 2ho-3+r = 2(ho-2+a)+dy with r = 2a+dy-1.  Four horizontally adjacent pixels
// are 64 contiguous bf16, so the implicit-GEMM kernel reads it as an NHWC map with 64 "channels" whose pixel stride
// is 16 elements (overlapping TMA rows) and 4 vertical taps: no im2col buffer (which was 747 MB at batch 32).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int H, int W,
                                                       int Hp, int Wp) {
  pdl_sync();
  const long long total = (long long)B * Hp * Wp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int J = (int)(idx % Wp);
    const long long t = idx / Wp;
    const int I = (int)(t % Hp), b = (int)(t / Hp);
    const int h0 = 2 * (I - 2), w0 = 2 * (J - 2);
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int h = h0 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* row = img + (((long long)b * 3 + c) * H + h) * W;
        if (w0 >= 0 && w0 + 1 < W && (W & 1) == 0) {
          const float2 x2 = __ldg(reinterpret_cast<const float2*>(row + w0));
          v[dy * 6 + c] = x2.x;
          v[dy * 6 + 3 + c] = x2.y;
        } else {
          if (w0 >= 0 && w0 < W) v[dy * 6 + c] = __ldg(row + w0);
          if (w0 + 1 >= 0 && w0 + 1 < W) v[dy * 6 + 3 + c] = __ldg(row + w0 + 1);
        }
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);   o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(out + idx * 16);
    dst[0] = o0;
    dst[1] = o1;
  }
}

// Same map from the loader's raw format: uint8 NHWC [B][H][W][3] with the reference's normalisation
// (ToTensor + Normalize, datasets/coco_generic_dataset.py:31-32: x = (u8 / 255 - mean[c]) / std[c]) folded into the read.
// A quarter of the fp32 NCHW bytes over PCIe / HBM.
struct Norm3 {
  float scale[3], shift[3];   // x = u8 * scale[c] + shift[c]
};
__global__ void __launch_bounds__(256) stem_s2d_u8_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int B, int H, int W,
                                                          int Hp, int Wp, const Norm3 nm) {
  pdl_sync();
  const long long total = (long long)B * Hp * Wp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int J = (int)(idx % Wp);
    const long long t = idx / Wp;
    const int I = (int)(t % Hp), b = (int)(t / Hp);
    const int h0 = 2 * (I - 2), w0 = 2 * (J - 2);
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int h = h0 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int w = w0 + dx;
        if (w < 0 || w >= W) continue;
        const uint8_t* px = img + (((long long)b * H + h) * W + w) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[dy * 6 + dx * 3 + c] = (float)__ldg(px + c) * nm.scale[c] + nm.shift[c];
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);   o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(out + idx * 16);
    dst[0] = o0;
    dst[1] = o1;
  }
}

// ------------------------------------------------------------------------------------------------
// ROI-align-mean weights (detr_roi_head.py:44-56 + torchvision roi_align, output 7x7, aligned=True,
// sampling_ratio=-1, then mean over the 49 bins).  The mean is linear in the feature map with rank-1 separable
// weights w_y (x) w_x per box; this kernel writes Wroi[b][q][y*W + x] (bf16, row stride ldw, zero padded) so that
// the pooled feature is one batched GEMM  Wroi[b] (Q x HW) * C5[b] (HW x C).
// ------------------------------------------------------------------------------------------------
GPV_DEVINL void roi_axis_weights(float start, float size, int n, float* w /*[n]*/) {
  for (int i = 0; i < n; ++i) w[i] = 0.f;
  const float bin = size / 7.0f;
  const int grid = (int)ceilf(size / 7.0f);
  if (grid <= 0) return;
  const float norm = 1.0f / (7.0f * (float)grid);
  for (int pbin = 0; pbin < 7; ++pbin) {
    for (int i = 0; i < grid; ++i) {
      float y = start + pbin * bin + ((float)i + 0.5f) * bin / (float)grid;
      if (y < -1.0f || y > (float)n) continue;  // sample outside the map contributes zero
      if (y <= 0.f) y = 0.f;
      int lo = (int)y, hi;
      if (lo >= n - 1) {
        hi = lo = n - 1;
        y = (float)lo;
      } else {
        hi = lo + 1;
      }
      const float l = y - (float)lo, h = 1.0f - l;
      w[lo] += h * norm;
      w[hi] += l * norm;
    }
  }
}

__global__ void __launch_bounds__(64) roi_weights_kernel(const float* __restrict__ boxes, long long ldb, bf16* __restrict__ wroi,
                                                         long long ldw, int BQ, int H, int W) {
  pdl_sync();
  __shared__ float wy[64], wx[64];
  const int bq = blockIdx.x;
  if (bq >= BQ) return;
  const float* bx = boxes + (long long)bq * ldb;
  const float cx = bx[0], cy = bx[1], bw = bx[2], bh = bx[3];
  // scaled_boxes (detr_roi_head.py:48-52), then roi_align's aligned offset of 0.5
  const float x1 = (float)W * (cx - 0.5f * bw), y1 = (float)H * (cy - 0.5f * bh);
  const float x2 = (float)W * (cx + 0.5f * bw), y2 = (float)H * (cy + 0.5f * bh);
  if (threadIdx.x == 0) roi_axis_weights(y1 - 0.5f, (y2 - 0.5f) - (y1 - 0.5f), H, wy);
  if (threadIdx.x == 32) roi_axis_weights(x1 - 0.5f, (x2 - 0.5f) - (x1 - 0.5f), W, wx);
  __syncthreads();
  bf16* out = wroi + (long long)bq * ldw;
  for (int i = threadIdx.x; i < ldw; i += blockDim.x) {
    float v = 0.f;
    if (i < H * W) v = wy[i / W] * wx[i % W];
    out[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Relevance conditioning (gpv.py:364-375): out[m] = x[m] + softmax(logits[m])[0]*tok[0] + softmax(logits[m])[1]*tok[1]
// Row remap on the output lets it write straight into the [B, Q+Tl, D] decoder memory (gpv.py:175).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relevance_mix_fwd_kernel(const bf16* __restrict__ x, long long ldx,
                                                                const float* __restrict__ logits, long long ldl,
                                                                const float* __restrict__ tok, bf16* __restrict__ out,
                                                                long long ldo, int M, int D, int G, int out_gstride,
                                                                int out_off) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
    const float l0 = logits[m * ldl], l1 = logits[m * ldl + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx);
    const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
    const long long orow = (m / G) * out_gstride + out_off + (m % G);
    for (int c = lane; c < (D >> 3); c += 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + m * ldx + c * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        const int d = c * 8 + 2 * j;
        o[j] = pack_bf16x2(f.x + p0 * __ldg(tok + d) + p1 * __ldg(tok + D + d),
                           f.y + p0 * __ldg(tok + d + 1) + p1 * __ldg(tok + D + d + 1));
      }
      *reinterpret_cast<uint4*>(out + orow * ldo + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// backward: dy read with the same row remap; dlogits[m][c] += p_c * (a_c - sum_k p_k a_k), a_c = <dy, tok_c>;
// dtok[c] += sum_m p_c * dy[m];  (dx = dy, taken by the caller as a view)
__global__ void __launch_bounds__(256) relevance_mix_bwd_kernel(const bf16* __restrict__ dy, long long lddy,
                                                                const float* __restrict__ logits, long long ldl,
                                                                const float* __restrict__ tok, float* __restrict__ dlogits,
                                                                long long lddl, float* __restrict__ dtok, int M, int D, int G,
                                                                int gstride, int off) {
  pdl_sync();
  extern __shared__ float acc[];  // [2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  // a lane owns dimensions lane, lane + 32, ...: its share of dtok stays in registers over the warp's rows (D <= 32 * kJ) and meets the
  // other warps' once, at the end -- the first version issued two contended shared-memory atomics per element of every row (40 us at
  // M = 3200, D = 768, on the backward's critical path)
  constexpr int kJ = 24;
  float t0[kJ], t1[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) t0[j] = t1[j] = 0.f;
  const bool in_regs = D <= 32 * kJ;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
    const float l0 = logits[m * ldl], l1 = logits[m * ldl + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx);
    const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
    const long long row = (m / G) * gstride + off + (m % G);
    float a0 = 0.f, a1 = 0.f;
    if (in_regs) {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int d = lane + 32 * j;
        if (d < D) {
          const float g = __bfloat162float(dy[row * lddy + d]);
          a0 += g * __ldg(tok + d);
          a1 += g * __ldg(tok + D + d);
          t0[j] += p0 * g;
          t1[j] += p1 * g;
        }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float g = __bfloat162float(dy[row * lddy + d]);
        a0 += g * __ldg(tok + d);
        a1 += g * __ldg(tok + D + d);
        atomicAdd(&acc[d], p0 * g);
        atomicAdd(&acc[D + d], p1 * g);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      const float mean = p0 * a0 + p1 * a1;
      dlogits[m * lddl] += p0 * (a0 - mean);
      dlogits[m * lddl + 1] += p1 * (a1 - mean);
    }
  }
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int d = lane + 32 * j;
      if (d < D) {
        atomicAdd(&acc[d], t0[j]);
        atomicAdd(&acc[D + d], t1[j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) atomicAdd(dtok + i, acc[i]);
}

// ------------------------------------------------------------------------------------------------
// out[m][:] = table[ids[m]][:] (+ pos[m % T][:]) (+ cst[:])   fp32 tables -> bf16 rows
// (AnswerInputEmbedding gpv.py:53; BERT embeddings)
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids, const float* __restrict__ pos,
                                   const float* __restrict__ cst, bf16* __restrict__ out, long long ldo, long long M, int D, int T,
                                   uint8_t* __restrict__ pad_mask, long long pad_id) {
  pdl_sync();
  const int nch = D >> 2;
  const long long total = M * nch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nch;
    const int c = (int)(i % nch);
    const long long id = ids[m];
    if (pad_mask != nullptr && c == 0) pad_mask[m] = id == pad_id ? 1 : 0;   // key-padding mask of the token row (bert.py:12-15, padding=True)
    float4 v = *reinterpret_cast<const float4*>(table + id * (long long)D + c * 4);
    if (pos) {
      const float4 q = *reinterpret_cast<const float4*>(pos + (m % T) * (long long)D + c * 4);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (cst) {
      const float4 q = *reinterpret_cast<const float4*>(cst + c * 4);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + m * ldo + c * 4) = o;
  }
}

// dst[remap_d(m)][:] = src[remap_s(m)][:]   with remap(m) = (m / G) * gstride + off + m % G     (memory concat, gpv.py:175)
__global__ void copy_rows_kernel(const bf16* __restrict__ src, long long lds, int sG, int sgs, int soff, bf16* __restrict__ dst,
                                 long long ldd, int dG, int dgs, int doff, long long M, int D) {
  pdl_sync();
  const int nch = D >> 3;
  const long long total = M * nch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nch;
    const int c = (int)(i % nch);
    const long long sr = (m / sG) * sgs + soff + (m % sG), dr = (m / dG) * dgs + doff + (m % dG);
    *reinterpret_cast<uint4*>(dst + dr * ldd + c * 8) = *reinterpret_cast<const uint4*>(src + sr * lds + c * 8);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_sync();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i * 4) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16(src[(n4 << 2) + threadIdx.x]);
}

static inline int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)kSMs * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace gpv

using namespace gpv;

#define ST ((cudaStream_t)stream)

extern "C" int gpvb200_add_rowbcast(const void* x, int64_t ldx, const void* p, int64_t ldp, void* out, int64_t ldo, int64_t M,
                                    int32_t D, int32_t P, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(p && out && D % 8 == 0 && P > 0 && ldp % 8 == 0 && ldo % 8 == 0 && (x == nullptr || ldx % 8 == 0), "add_rowbcast: bad arguments");
  if (M == 0) return GPV_OK;
  launch_k(add_rowbcast_kernel, dim3(grid_for(M * (D / 8), 256)), dim3(256), 0, ST, (const bf16*)x, ldx, (const bf16*)p, ldp, (bf16*)out, ldo, M, D, P);
  return check_launch("add_rowbcast_kernel");
}

extern "C" int gpvb200_colsum(const void* dy, int64_t ld, float* out, int64_t M, int32_t N, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(dy && out && N > 0, "colsum: bad arguments");
  if (M == 0) return GPV_OK;
  if (N % 8 == 0 && ld % 8 == 0 && ((uintptr_t)dy & 15) == 0) {
    const int gx = (N + 255) / 256;
    int gy = (int)((2LL * kSMs + gx - 1) / gx);
    if ((long long)gy * 32 > M) gy = (int)((M + 31) / 32);
    if (gy < 1) gy = 1;
    const int rows = (int)((M + gy - 1) / gy);
    launch_k(colsum_vec_kernel, dim3(gx, gy), dim3(256), 0, ST, (const bf16*)dy, ld, out, M, N, rows);
    return check_launch("colsum_vec_kernel");
  }
  const int gx = (N + 63) / 64;
  int gy = (int)((2LL * kSMs + gx - 1) / gx);
  if ((long long)gy * 64 > M) gy = (int)((M + 63) / 64);
  if (gy < 1) gy = 1;
  const int rows = (int)((M + gy - 1) / gy);
  launch_k(colsum_kernel, dim3(gx, gy), dim3(256), 0, ST, (const bf16*)dy, ld, out, M, N, rows);
  return check_launch("colsum_kernel");
}

extern "C" int gpvb200_batch_reduce(const void* x, int64_t ld, float* out, int32_t B, int32_t S, int32_t D, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && out && B > 0 && S > 0 && D > 0, "batch_reduce: bad arguments");
  launch_k(batch_reduce_kernel, dim3(grid_for((long long)S * D, 256)), dim3(256), 0, ST, (const bf16*)x, ld, out, B, S, D);
  return check_launch("batch_reduce_kernel");
}

extern "C" size_t gpvb200_pack_item_size(void) { return sizeof(PackItem); }
extern "C" int gpvb200_pack_chunk(void) { return kPackChunk; }

extern "C" int gpvb200_pack_weights(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks,
                                    void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(items && blk_item && blk_chunk && n_blocks >= 0, "pack_weights: bad arguments");
  if (n_blocks == 0) return GPV_OK;
  pack_weights_kernel<<<n_blocks, 256, 0, ST>>>((const PackItem*)items, blk_item, blk_chunk);
  return check_launch("pack_weights_kernel");
}

extern "C" int gpvb200_bn_fold(const float* w, const float* b, const float* rm, const float* rv, float* scale, float* bias,
                               int32_t n, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(w && b && rm && rv && scale && bias && n > 0, "bn_fold: bad arguments");
  bn_fold_kernel<<<(n + 255) / 256, 256, 0, ST>>>(w, b, rm, rv, scale, bias, n);
  return check_launch("bn_fold_kernel");
}

extern "C" int gpvb200_maxpool3x3s2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C % 8 == 0, "maxpool: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  launch_k(maxpool_kernel, dim3(grid_for((long long)B * Ho * Wo * (C / 8), 256)), dim3(256), 0, ST, (const bf16*)x, (bf16*)y, B, H, W, C, Ho, Wo);
  return check_launch("maxpool_kernel");
}

extern "C" int gpvb200_stem_im2col(const float* img, void* col, int32_t B, int32_t H, int32_t W, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(img && col && B > 0 && H > 0 && W > 0, "stem_im2col: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  stem_im2col_kernel<<<grid_for((long long)B * Ho * Wo * 19, 256), 256, 0, ST>>>(img, (bf16*)col, B, H, W, Ho, Wo);
  return check_launch("stem_im2col_kernel");
}

extern "C" int gpvb200_stem_s2d(const float* img, void* out, int32_t B, int32_t H, int32_t W, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(img && out && B > 0 && H > 0 && W > 0, "stem_s2d: bad arguments");
  GPV_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)img & 7) == 0, "stem_s2d: unaligned buffers");
  const int Hp = (H - 1) / 2 + 1 + 4, Wp = (W - 1) / 2 + 1 + 4;
  launch_k(stem_s2d_kernel, dim3(grid_for((long long)B * Hp * Wp, 256)), dim3(256), 0, ST, img, (bf16*)out, B, H, W, Hp, Wp);
  return check_launch("stem_s2d_kernel");
}

extern "C" int gpvb200_stem_s2d_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, const float* mean3,
                                   const float* std3, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(img && out && mean3 && std3 && B > 0 && H > 0 && W > 0, "stem_s2d_u8: bad arguments");
  GPV_REQUIRE(((uintptr_t)out & 15) == 0, "stem_s2d_u8: unaligned output");
  Norm3 nm;
  for (int c = 0; c < 3; ++c) {   // mean3 / std3 are HOST arrays (three floats each)
    GPV_REQUIRE(std3[c] > 0.f, "stem_s2d_u8: std must be positive");
    nm.scale[c] = 1.0f / (255.0f * std3[c]);
    nm.shift[c] = -mean3[c] / std3[c];
  }
  const int Hp = (H - 1) / 2 + 1 + 4, Wp = (W - 1) / 2 + 1 + 4;
  launch_k(stem_s2d_u8_kernel, dim3(grid_for((long long)B * Hp * Wp, 256)), dim3(256), 0, ST, img, (bf16*)out, B, H, W, Hp, Wp, nm);
  return check_launch("stem_s2d_u8_kernel");
}

extern "C" int gpvb200_roi_weights(const float* boxes, int64_t ldb, void* wroi, int64_t ldw, int32_t BQ, int32_t H, int32_t W,
                                   void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(boxes && wroi && BQ >= 0 && H > 0 && W > 0 && H <= 64 && W <= 64 && ldw >= (int64_t)H * W, "roi_weights: bad arguments");
  if (BQ == 0) return GPV_OK;
  launch_k(roi_weights_kernel, dim3(BQ), dim3(64), 0, ST, boxes, ldb, (bf16*)wroi, ldw, BQ, H, W);
  return check_launch("roi_weights_kernel");
}

extern "C" int gpvb200_relevance_mix_fwd(const void* x, int64_t ldx, const float* logits, int64_t ldl, const float* tok, void* out,
                                         int64_t ldo, int32_t M, int32_t D, int32_t G, int32_t out_gstride, int32_t out_off,
                                         void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(x && logits && tok && out && D % 8 == 0 && G > 0, "relevance_mix_fwd: bad arguments");
  if (M == 0) return GPV_OK;
  launch_k(relevance_mix_fwd_kernel, dim3(grid_for(M, 8)), dim3(256), 0, ST, (const bf16*)x, ldx, logits, ldl, tok, (bf16*)out, ldo, M, D, G, out_gstride, out_off);
  return check_launch("relevance_mix_fwd_kernel");
}

extern "C" int gpvb200_relevance_mix_bwd(const void* dy, int64_t lddy, const float* logits, int64_t ldl, const float* tok,
                                         float* dlogits, int64_t lddl, float* dtok, int32_t M, int32_t D, int32_t G, int32_t gstride,
                                         int32_t off, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(dy && logits && tok && dlogits && dtok && G > 0, "relevance_mix_bwd: bad arguments");
  if (M == 0) return GPV_OK;
  int grid = (M + 7) / 8;
  if (grid > kSMs) grid = kSMs;
  launch_k(relevance_mix_bwd_kernel, dim3(grid), dim3(256), (size_t)2 * D * sizeof(float), ST, (const bf16*)dy, lddy, logits, ldl, tok, dlogits, lddl, dtok, M, D, G,
                                                                            gstride, off);
  return check_launch("relevance_mix_bwd_kernel");
}

extern "C" int gpvb200_gather_rows_mask(const float* table, const int64_t* ids, const float* pos, const float* cst, void* out, int64_t ldo,
                                        int64_t M, int32_t D, int32_t T, uint8_t* pad_mask, int64_t pad_id, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(table && ids && out && D % 4 == 0 && ldo % 4 == 0, "gather_rows: bad arguments");
  if (M == 0) return GPV_OK;
  launch_k(gather_rows_kernel, dim3(grid_for(M * (D / 4), 256)), dim3(256), 0, ST, table, ids, pos, cst, (bf16*)out, ldo, M, D, T > 0 ? T : 1,
           pad_mask, (long long)pad_id);
  return check_launch("gather_rows_kernel");
}

extern "C" int gpvb200_gather_rows(const float* table, const int64_t* ids, const float* pos, const float* cst, void* out, int64_t ldo,
                                   int64_t M, int32_t D, int32_t T, void* stream) {
  return gpvb200_gather_rows_mask(table, ids, pos, cst, out, ldo, M, D, T, nullptr, 0, stream);
}

extern "C" int gpvb200_copy_rows(const void* src, int64_t lds, int32_t sG, int32_t sgs, int32_t soff, void* dst, int64_t ldd, int32_t dG,
                                 int32_t dgs, int32_t doff, int64_t M, int32_t D, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(src && dst && D % 8 == 0 && sG > 0 && dG > 0 && lds % 8 == 0 && ldd % 8 == 0, "copy_rows: bad arguments");
  if (M == 0) return GPV_OK;
  launch_k(copy_rows_kernel, dim3(grid_for(M * (D / 8), 256)), dim3(256), 0, ST, (const bf16*)src, lds, sG, sgs, soff, (bf16*)dst, ldd, dG, dgs, doff, M, D);
  return check_launch("copy_rows_kernel");
}

extern "C" int gpvb200_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(src && dst && n >= 0, "cast: bad arguments");
  if (n == 0) return GPV_OK;
  launch_k(cast_f32_bf16_kernel, dim3(grid_for(n / 4 + 1, 256)), dim3(256), 0, ST, src, (bf16*)dst, n);
  return check_launch("cast_f32_bf16_kernel");
}

// ------------------------------------------------------------------------------------------------
// Conv weight-gradient unpack: packed fp32 [taps][O][I] (what the wgrad GEMM accumulates) -> master layout
// [O][I][taps] (torch Conv2d weight [O,I,kh,kw]); dst = src (accumulate == 0) or dst += src.
// ------------------------------------------------------------------------------------------------
namespace gpv {
__global__ void unpack_conv_grad_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int taps, int accumulate) {
  pdl_sync();
  const long long OI = (long long)O * I, n = OI * taps;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long oi = e / taps;
    const int t = (int)(e % taps);
    const float v = src[(long long)t * OI + oi];
    dst[e] = accumulate ? dst[e] + v : v;
  }
}

// out = a + b (bf16, row strides), vectorised 16 B
__global__ void add_bf16_kernel(const bf16* __restrict__ a, long long lda, const bf16* __restrict__ b, long long ldb,
                                bf16* __restrict__ out, long long ldo, long long M, int D) {
  pdl_sync();
  const int nch = D >> 3;
  const long long total = M * nch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nch;
    const int c = (int)(i % nch);
    const uint4 x = *reinterpret_cast<const uint4*>(a + m * lda + c * 8);
    const uint4 y = *reinterpret_cast<const uint4*>(b + m * ldb + c * 8);
    const uint32_t wa[4] = {x.x, x.y, x.z, x.w}, wb[4] = {y.x, y.y, y.z, y.w};
    uint32_t wo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(wa[j]), fb = unpack_bf16x2(wb[j]);
      wo[j] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    }
    *reinterpret_cast<uint4*>(out + m * ldo + c * 8) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
}
}  // namespace gpv

extern "C" int gpvb200_unpack_conv_grad(const float* src, float* dst, int32_t O, int32_t I, int32_t taps, int32_t accumulate,
                                        void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(src && dst && O > 0 && I > 0 && taps > 0, "unpack_conv_grad: bad arguments");
  launch_k(unpack_conv_grad_kernel, dim3(grid_for((long long)O * I * taps, 256)), dim3(256), 0, ST, src, dst, O, I, taps, accumulate);
  return check_launch("unpack_conv_grad_kernel");
}

extern "C" int gpvb200_add_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t M, int32_t D,
                                void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(a && b && out && D % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "add_bf16: bad arguments");
  if (M == 0) return GPV_OK;
  launch_k(add_bf16_kernel, dim3(grid_for(M * (D / 8), 256)), dim3(256), 0, ST, (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)out, ldo, M, D);
  return check_launch("add_bf16_kernel");
}
