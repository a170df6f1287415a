// Fused multi-head attention core (softmax(scale * Q K^T + mask) V) forward and backward for the short sequences
// of GPV-1: DETR encoder self-attention (S=300, d_h=32), decoder self/cross attention (100 / 100x300), ViLBERT
// co-attention (20<->100, d_h=48), text-decoder causal self-attention and cross-attention (d_h=96), BERT (d_h=64).
// Reference: torch.nn.MultiheadAttention via transformer.py:153-155,218-226 and nn.TransformerDecoderLayer
// (gpv.py:38-43); BertBiAttention.forward vilbert.py:737-824.  The attention-weight average the reference
// materialises and discards (need_weights) is never computed.
//
// One CTA (8-12 warps) owns one (batch, head): Q, K, V (and dO for backward) of that head live in shared memory, both
// row-major and transposed, so every product is a register-A x shared-B^T warp MMA (mma.sync m16n8k16 bf16,
// fp32 accumulate) with online softmax kept in registers.  Scores never touch HBM.
// NOTE: this is the legacy tensor path (HMMA); the sequences are too short to fill a 128-row tcgen05 tile per
// head -- see DESIGN.md for the packing plan.
#include "../../include/gpvb200.h"
#include "common.cuh"
#include "host_util.h"

namespace gpv {

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  const bf16* d_o; bf16* dq; bf16* dk; bf16* dv;
  float* lse;                 // [B,H,Sq], log2 domain
  const uint8_t* kmask;       // [B,Sk] 1 = masked key, or null
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;  // token-row strides (elements)
  long long bsq, bsk, bsv, bso;   // batch strides (elements) of q, k, v, o in the forward; S*ld unless a KV cache is read in place
  int B, H, Sq, Sk, causal;
  float scale;
  DropArgs drop;   // dropout on the attention probabilities (train mode); drop.seed == nullptr: none
};

GPV_DEVINL void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

GPV_DEVINL uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// B fragment (k16 x n8) of mma.m16n8k16 from a ROW-major [k][n] shared-memory matrix X (row stride LD elements, rows
// 16-byte aligned): ldmatrix.trans hands lane (g, t) the pair X[k0 + 2t .. 2t+1][n0 + g] (b0) and the same for k + 8 (b1),
// which is exactly the "col" operand layout -- no transposed copy of K / V / Q / dO in shared memory.
GPV_DEVINL void ldsm_bt(const bf16* X, int LD, int k0, int n0, int lane, uint32_t& b0, uint32_t& b1) {
  const uint32_t a = smem_u32(X + (k0 + (lane & 15)) * LD + n0);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(a));
}

// Stage rows [0,S) x DH of a token-major global matrix into smem row-major (stride LD) and optionally transposed
// (dst_t[d][s], stride LDT).  Rows >= S are zero-filled up to Sp.
template <int DH>
GPV_DEVINL void stage(const bf16* __restrict__ g, long long ld, int S, int Sp, bf16* dst, int LD, bf16* dst_t, int LDT) {
  constexpr int CH = DH / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < Sp * CH; i += blockDim.x) {
    const int r = i / CH, c = i % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < S) v = *reinterpret_cast<const uint4*>(g + (long long)r * ld + c * 8);
    if (dst) *reinterpret_cast<uint4*>(dst + r * LD + c * 8) = v;
    if (dst_t) {
      const bf16* e = reinterpret_cast<const bf16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst_t[(c * 8 + j) * LDT + r] = e[j];
    }
  }
}

template <int DH>
GPV_DEVINL void load_a_frags(const bf16* s, int LD, int r0, int g, int t, uint32_t (&a)[DH / 16][4]) {
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
    a[ks][0] = lds32(s + (r0 + g) * LD + ks * 16 + 2 * t);
    a[ks][1] = lds32(s + (r0 + g + 8) * LD + ks * 16 + 2 * t);
    a[ks][2] = lds32(s + (r0 + g) * LD + ks * 16 + 8 + 2 * t);
    a[ks][3] = lds32(s + (r0 + g + 8) * LD + ks * 16 + 8 + 2 * t);
  }
}

// ======================================================================================== forward
template <int DH, int NT, bool DROP>
__global__ void __launch_bounds__(NT) attn_fwd_kernel(const AttnParams p) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int Sq = p.Sq, Sk = p.Sk;
  const int Sqp = (Sq + 15) & ~15, Skp = (Sk + 63) & ~63;
  constexpr int LD = DH + 8;
  bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
  bf16* Ks = Qs + Sqp * LD;
  bf16* Vs = Ks + Skp * LD;
  uint8_t* msk = reinterpret_cast<uint8_t*>(Vs + Skp * LD);

  stage<DH>(p.q + b * p.bsq + h * DH, p.ldq, Sq, Sqp, Qs, LD, nullptr, 0);
  stage<DH>(p.k + b * p.bsk + h * DH, p.ldk, Sk, Skp, Ks, LD, nullptr, 0);
  stage<DH>(p.v + b * p.bsv + h * DH, p.ldv, Sk, Skp, Vs, LD, nullptr, 0);
  for (int j = threadIdx.x; j < Skp; j += blockDim.x)
    msk[j] = (j >= Sk) ? 1 : (p.kmask ? p.kmask[(long long)b * Sk + j] : 0);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float sl2 = p.scale * 1.4426950408889634f;
  const uint32_t dkey = DROP ? drop_key(*p.drop.seed, p.drop.site) : 0u;
  for (int r0 = warp * 16; r0 < Sqp; r0 += NT / 2) {
    uint32_t qa[DH / 16][4];
    load_a_frags<DH>(Qs, LD, r0, g, t, qa);
    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int row0 = r0 + g, row1 = r0 + g + 8;
    for (int kb = 0; kb < Skp; kb += 64) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const bf16* kr = Ks + (kb + nt * 8 + g) * LD + 2 * t;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) mma16816(s[nt], qa[ks], lds32(kr + ks * 16), lds32(kr + ks * 16 + 8));
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = kb + nt * 8 + 2 * t + e;
          const bool dead = msk[col] != 0;
          if (dead || (p.causal && col > row0)) s[nt][e] = -INFINITY;
          if (dead || (p.causal && col > row1)) s[nt][2 + e] = -INFINITY;
          mx0 = fmaxf(mx0, s[nt][e]);
          mx1 = fmaxf(mx1, s[nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float c0 = ex2_approx((m0 - ms0) * sl2), c1 = ex2_approx((m1 - ms1) * sl2);
      m0 = mn0;
      m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = ex2_approx((s[nt][0] - ms0) * sl2);
        s[nt][1] = ex2_approx((s[nt][1] - ms0) * sl2);
        s[nt][2] = ex2_approx((s[nt][2] - ms1) * sl2);
        s[nt][3] = ex2_approx((s[nt][3] - ms1) * sl2);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
      if (DROP) {   // O = dropout(softmax(S)) V: the normaliser keeps every key, the PV product the kept ones
        const uint32_t hs = (uint32_t)((Sk + 1) >> 1);
        const uint32_t pr0 = ((uint32_t)bh * (uint32_t)Sq + (uint32_t)row0) * hs, pr1 = pr0 + 8u * hs;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint32_t cp = (uint32_t)((kb + nt * 8) >> 1) + t;
          drop_pair(s[nt][0], s[nt][1], dkey, pr0 + cp, p.drop.thresh16, p.drop.scale);
          drop_pair(s[nt][2], s[nt][3], dkey, pr1 + cp, p.drop.thresh16, p.drop.scale);
        }
      }
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1;
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * k2][0], s[2 * k2][1]);
        pa[1] = pack_bf16x2(s[2 * k2][2], s[2 * k2][3]);
        pa[2] = pack_bf16x2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        pa[3] = pack_bf16x2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int nd = 0; nd < DH / 8; ++nd) {
          uint32_t b0, b1;
          ldsm_bt(Vs, LD, kb + k2 * 16, nd * 8, lane, b0, b1);
          mma16816(o[nd], pa, b0, b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    if (row0 < Sq) {
      bf16* op = p.o + b * p.bso + (long long)row0 * p.ldo + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) *reinterpret_cast<uint32_t*>(op + nd * 8) = pack_bf16x2(o[nd][0] * i0, o[nd][1] * i0);
      if (p.lse && t == 0) p.lse[((long long)b * p.H + h) * Sq + row0] = m0 * sl2 + log2f(l0);
    }
    if (row1 < Sq) {
      bf16* op = p.o + b * p.bso + (long long)row1 * p.ldo + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) *reinterpret_cast<uint32_t*>(op + nd * 8) = pack_bf16x2(o[nd][2] * i1, o[nd][3] * i1);
      if (p.lse && t == 0) p.lse[((long long)b * p.H + h) * Sq + row1] = m1 * sl2 + log2f(l1);
    }
  }
}

// ======================================================================================== backward
// Pass 1 (warp owns 16 queries): D = rowsum(dO*O), dQ = scale * [P o (dO V^T - D)] K
// Pass 2 (warp owns 16 keys):    dV = P^T dO,      dK = scale * [P o (dO V^T - D)]^T Q
template <int DH, int NT, bool DROP>
__global__ void __launch_bounds__(NT, (DH <= 32) ? 2 : 1) attn_bwd_kernel(const AttnParams p) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int Sq = p.Sq, Sk = p.Sk;
  const int Sqp = (Sq + 63) & ~63, Skp = (Sk + 63) & ~63;
  constexpr int LD = DH + 8;

  bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
  bf16* dOs = Qs + Sqp * LD;
  bf16* Ks = dOs + Sqp * LD;
  bf16* Vs = Ks + Skp * LD;
  float* lse_s = reinterpret_cast<float*>(Vs + Skp * LD);
  float* D_s = lse_s + Sqp;
  uint8_t* msk = reinterpret_cast<uint8_t*>(D_s + Sqp);

  stage<DH>(p.q + ((long long)b * Sq) * p.ldq + h * DH, p.ldq, Sq, Sqp, Qs, LD, nullptr, 0);
  stage<DH>(p.d_o + ((long long)b * Sq) * p.lddo + h * DH, p.lddo, Sq, Sqp, dOs, LD, nullptr, 0);
  stage<DH>(p.k + ((long long)b * Sk) * p.ldk + h * DH, p.ldk, Sk, Skp, Ks, LD, nullptr, 0);
  stage<DH>(p.v + ((long long)b * Sk) * p.ldv + h * DH, p.ldv, Sk, Skp, Vs, LD, nullptr, 0);
  for (int j = threadIdx.x; j < Skp; j += blockDim.x)
    msk[j] = (j >= Sk) ? 1 : (p.kmask ? p.kmask[(long long)b * Sk + j] : 0);
  for (int j = threadIdx.x; j < Sqp; j += blockDim.x)
    lse_s[j] = (j < Sq) ? p.lse[((long long)b * p.H + h) * Sq + j] : INFINITY;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float sl2 = p.scale * 1.4426950408889634f;
  constexpr bool drop = DROP;
  const uint32_t dkey = drop ? drop_key(*p.drop.seed, p.drop.site) : 0u;
  const uint32_t hs = (uint32_t)((Sk + 1) >> 1);   // mask pairs per (b, h, q) row

  // ---------------------------------------------------------------- pass 1: D and dQ
  for (int r0 = warp * 16; r0 < Sqp; r0 += NT / 2) {
    uint32_t qa[DH / 16][4], da[DH / 16][4];
    load_a_frags<DH>(Qs, LD, r0, g, t, qa);
    load_a_frags<DH>(dOs, LD, r0, g, t, da);
    const int row0 = r0 + g, row1 = r0 + g + 8;
    // D = rowsum(dO * O): O read from global with the A-fragment footprint
    float d0 = 0.f, d1 = 0.f;
    {
      const bf16* o0 = p.o + ((long long)b * Sq + row0) * p.ldo + h * DH;
      const bf16* o1 = p.o + ((long long)b * Sq + row1) * p.ldo + h * DH;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        if (row0 < Sq) {
          const float2 a = unpack_bf16x2(da[ks][0]), c = unpack_bf16x2(da[ks][2]);
          const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o0 + ks * 16 + 2 * t));
          const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o0 + ks * 16 + 8 + 2 * t));
          d0 += a.x * x.x + a.y * x.y + c.x * y.x + c.y * y.y;
        }
        if (row1 < Sq) {
          const float2 a = unpack_bf16x2(da[ks][1]), c = unpack_bf16x2(da[ks][3]);
          const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o1 + ks * 16 + 2 * t));
          const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o1 + ks * 16 + 8 + 2 * t));
          d1 += a.x * x.x + a.y * x.y + c.x * y.x + c.y * y.y;
        }
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      if (t == 0) {
        D_s[row0] = d0;
        D_s[row1] = d1;
      }
    }
    const float ls0 = lse_s[row0], ls1 = lse_s[row1];
    float dq[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kb = 0; kb < Skp; kb += 32) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        const bf16* kr = Ks + (kb + nt * 8 + g) * LD + 2 * t;
        const bf16* vr = Vs + (kb + nt * 8 + g) * LD + 2 * t;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          mma16816(s[nt], qa[ks], lds32(kr + ks * 16), lds32(kr + ks * 16 + 8));
          mma16816(dp[nt], da[ks], lds32(vr + ks * 16), lds32(vr + ks * 16 + 8));
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = kb + nt * 8 + 2 * t + e;
          const bool dead = msk[col] != 0;
          const float p0 = (dead || (p.causal && col > row0)) ? 0.f : ex2_approx(s[nt][e] * sl2 - ls0);
          const float p1 = (dead || (p.causal && col > row1)) ? 0.f : ex2_approx(s[nt][2 + e] * sl2 - ls1);
          s[nt][e] = p0;
          s[nt][2 + e] = p1;
        }
        if (drop) {   // dP reaches the probabilities through the forward's mask: dP <- dP (*) mask / (1 - p)
          const uint32_t cp = (uint32_t)((kb + nt * 8) >> 1) + t;
          const uint32_t pr0 = ((uint32_t)bh * (uint32_t)Sq + (uint32_t)row0) * hs;
          drop_pair(dp[nt][0], dp[nt][1], dkey, pr0 + cp, p.drop.thresh16, p.drop.scale);
          drop_pair(dp[nt][2], dp[nt][3], dkey, pr0 + 8u * hs + cp, p.drop.thresh16, p.drop.scale);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s[nt][e] = s[nt][e] * (dp[nt][e] - d0) * p.scale;
          s[nt][2 + e] = s[nt][2 + e] * (dp[nt][2 + e] - d1) * p.scale;
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * k2][0], s[2 * k2][1]);
        pa[1] = pack_bf16x2(s[2 * k2][2], s[2 * k2][3]);
        pa[2] = pack_bf16x2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        pa[3] = pack_bf16x2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int nd = 0; nd < DH / 8; ++nd) {
          uint32_t b0, b1;
          ldsm_bt(Ks, LD, kb + k2 * 16, nd * 8, lane, b0, b1);
          mma16816(dq[nd], pa, b0, b1);
        }
      }
    }
    if (row0 < Sq) {
      bf16* op = p.dq + ((long long)b * Sq + row0) * p.lddq + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) *reinterpret_cast<uint32_t*>(op + nd * 8) = pack_bf16x2(dq[nd][0], dq[nd][1]);
    }
    if (row1 < Sq) {
      bf16* op = p.dq + ((long long)b * Sq + row1) * p.lddq + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) *reinterpret_cast<uint32_t*>(op + nd * 8) = pack_bf16x2(dq[nd][2], dq[nd][3]);
    }
  }
  __syncthreads();  // D_s complete

  // ---------------------------------------------------------------- pass 2: dK and dV
  for (int r0 = warp * 16; r0 < Skp; r0 += NT / 2) {
    if (r0 >= ((Sk + 15) & ~15)) break;
    uint32_t ka[DH / 16][4], va[DH / 16][4];
    load_a_frags<DH>(Ks, LD, r0, g, t, ka);
    load_a_frags<DH>(Vs, LD, r0, g, t, va);
    const int key0 = r0 + g, key1 = r0 + g + 8;
    const bool dead0 = msk[key0] != 0, dead1 = msk[key1] != 0;
    float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qb = 0; qb < Sqp; qb += 32) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        const bf16* qr = Qs + (qb + nt * 8 + g) * LD + 2 * t;
        const bf16* dr = dOs + (qb + nt * 8 + g) * LD + 2 * t;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          mma16816(s[nt], ka[ks], lds32(qr + ks * 16), lds32(qr + ks * 16 + 8));
          mma16816(dp[nt], va[ks], lds32(dr + ks * 16), lds32(dr + ks * 16 + 8));
        }
      }
      float pt[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int qc = qb + nt * 8 + 2 * t + e;  // query index (column of S^T)
          const float ls = lse_s[qc], dd = D_s[qc];
          const float p0 = (dead0 || (p.causal && key0 > qc)) ? 0.f : ex2_approx(s[nt][e] * sl2 - ls);
          const float p1 = (dead1 || (p.causal && key1 > qc)) ? 0.f : ex2_approx(s[nt][2 + e] * sl2 - ls);
          float m0 = 1.0f, m1 = 1.0f;
          if (drop) {   // element (query qc, key): one half of the pair's 32 bits
            const uint32_t pr = ((uint32_t)bh * (uint32_t)Sq + (uint32_t)qc) * hs;
            const uint32_t b0 = drop_bits(dkey, pr + (uint32_t)(key0 >> 1)), b1 = drop_bits(dkey, pr + (uint32_t)(key1 >> 1));
            m0 = (((key0 & 1) ? (b0 >> 16) : (b0 & 0xFFFFu)) >= p.drop.thresh16) ? p.drop.scale : 0.0f;
            m1 = (((key1 & 1) ? (b1 >> 16) : (b1 & 0xFFFFu)) >= p.drop.thresh16) ? p.drop.scale : 0.0f;
          }
          pt[nt][e] = p0 * m0;                                   // dV = dropout(P)^T dO
          pt[nt][2 + e] = p1 * m1;
          s[nt][e] = p0 * (dp[nt][e] * m0 - dd) * p.scale;
          s[nt][2 + e] = p1 * (dp[nt][2 + e] * m1 - dd) * p.scale;
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        uint32_t pa[4], sa[4];
        pa[0] = pack_bf16x2(pt[2 * k2][0], pt[2 * k2][1]);
        pa[1] = pack_bf16x2(pt[2 * k2][2], pt[2 * k2][3]);
        pa[2] = pack_bf16x2(pt[2 * k2 + 1][0], pt[2 * k2 + 1][1]);
        pa[3] = pack_bf16x2(pt[2 * k2 + 1][2], pt[2 * k2 + 1][3]);
        sa[0] = pack_bf16x2(s[2 * k2][0], s[2 * k2][1]);
        sa[1] = pack_bf16x2(s[2 * k2][2], s[2 * k2][3]);
        sa[2] = pack_bf16x2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        sa[3] = pack_bf16x2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int nd = 0; nd < DH / 8; ++nd) {
          uint32_t b0, b1, c0, c1;
          ldsm_bt(dOs, LD, qb + k2 * 16, nd * 8, lane, b0, b1);
          ldsm_bt(Qs, LD, qb + k2 * 16, nd * 8, lane, c0, c1);
          mma16816(dv[nd], pa, b0, b1);
          mma16816(dk[nd], sa, c0, c1);
        }
      }
    }
    if (key0 < Sk) {
      bf16* kp = p.dk + ((long long)b * Sk + key0) * p.lddk + h * DH + 2 * t;
      bf16* vp = p.dv + ((long long)b * Sk + key0) * p.lddv + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) {
        *reinterpret_cast<uint32_t*>(kp + nd * 8) = pack_bf16x2(dk[nd][0], dk[nd][1]);
        *reinterpret_cast<uint32_t*>(vp + nd * 8) = pack_bf16x2(dv[nd][0], dv[nd][1]);
      }
    }
    if (key1 < Sk) {
      bf16* kp = p.dk + ((long long)b * Sk + key1) * p.lddk + h * DH + 2 * t;
      bf16* vp = p.dv + ((long long)b * Sk + key1) * p.lddv + h * DH + 2 * t;
#pragma unroll
      for (int nd = 0; nd < DH / 8; ++nd) {
        *reinterpret_cast<uint32_t*>(kp + nd * 8) = pack_bf16x2(dk[nd][2], dk[nd][3]);
        *reinterpret_cast<uint32_t*>(vp + nd * 8) = pack_bf16x2(dv[nd][2], dv[nd][3]);
      }
    }
  }
}

static size_t fwd_smem(int DH, int Sq, int Sk) {
  const int Sqp = (Sq + 15) & ~15, Skp = (Sk + 63) & ~63, LD = DH + 8;
  return (size_t)(Sqp * LD + 2 * Skp * LD) * 2 + Skp + 16;
}
static size_t bwd_smem(int DH, int Sq, int Sk) {
  const int Sqp = (Sq + 63) & ~63, Skp = (Sk + 63) & ~63, LD = DH + 8;
  return (size_t)(2 * Sqp * LD + 2 * Skp * LD) * 2 + (size_t)Sqp * 8 + Skp + 16;
}

template <int DH, int NT>
static int launch_attn_nt(const AttnParams& p, bool bwd, size_t smem, cudaStream_t st) {
  const bool drop = p.drop.seed != nullptr;
  auto kern = bwd ? (drop ? attn_bwd_kernel<DH, NT, true> : attn_bwd_kernel<DH, NT, false>)
                  : (drop ? attn_fwd_kernel<DH, NT, true> : attn_fwd_kernel<DH, NT, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return GPV_ERR_CUDA;
  }
  launch_k(kern, dim3(p.B * p.H), dim3(NT), smem, st, p);
  return check_launch(bwd ? "attn_bwd_kernel" : "attn_fwd_kernel");
}

template <int DH>
static int launch_attn(const AttnParams& p, bool bwd, cudaStream_t st) {
  const size_t smem = bwd ? bwd_smem(DH, p.Sq, p.Sk) : fwd_smem(DH, p.Sq, p.Sk);
  if (smem > 227 * 1024) {
    set_last_error("attention: Sq=%d Sk=%d dh=%d needs %zu bytes of shared memory (> 227 KB)", p.Sq, p.Sk, DH, smem);
    return GPV_ERR_ARG;
  }
  // 16 query (or key) rows per warp and pass.  Long sequences: 8 warps per CTA; with K / V / Q / dO kept row-major only
  // (ldmatrix.trans builds the transposed operand fragments) the encoder's S = 300, d_h = 32 backward needs 105 KB of
  // shared memory and <= 128 registers, so two CTAs share an SM and the 256 (batch, head) CTAs of a B = 32 step run in
  // one wave.  Short sequences (co-attention 100 x 20, text decoder 20 x 120): 4 warps already cover every row block,
  // and the smaller CTAs fit 2-3 per SM by registers (the d_h >= 48 backward uses 160-240 registers per thread).
  const int rows = (p.Sq > p.Sk ? p.Sq : p.Sk);
  if (rows <= 64 || (DH >= 48 && rows <= 128)) return launch_attn_nt<DH, 128>(p, bwd, smem, st);
  return launch_attn_nt<DH, 256>(p, bwd, smem, st);
}

static int dispatch(const AttnParams& p, int dh, bool bwd, cudaStream_t st) {
  switch (dh) {
    case 32: return launch_attn<32>(p, bwd, st);
    case 48: return launch_attn<48>(p, bwd, st);
    case 64: return launch_attn<64>(p, bwd, st);
    case 96: return launch_attn<96>(p, bwd, st);
    default:
      set_last_error("attention: head dim %d unsupported (32, 48, 64, 96)", dh);
      return GPV_ERR_ARG;
  }
}

}  // namespace gpv

using namespace gpv;

static DropArgs attn_drop(const void* seed, uint32_t site, float p) {
  DropArgs d;
  d.seed = (p > 0.f) ? (const unsigned long long*)seed : nullptr;
  d.site = site;
  d.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, const uint8_t* key_mask, int64_t ldq,
                         int64_t ldk, int64_t ldv, int64_t ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, int32_t B,
                         int32_t H, int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale, const DropArgs& drop,
                         void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(q && k && v && o, "attention_fwd: null pointer");
  GPV_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0, "attention_fwd: bad shape");
  GPV_REQUIRE((ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (ldo % 2 == 0), "attention_fwd: row strides must be multiples of 8");
  GPV_REQUIRE((bsq % 8 == 0) && (bsk % 8 == 0) && (bsv % 8 == 0) && (bso % 2 == 0), "attention_fwd: batch strides must be multiples of 8");
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)o;
  p.lse = lse; p.kmask = key_mask;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.bsq = bsq > 0 ? bsq : (long long)Sq * ldq;
  p.bsk = bsk > 0 ? bsk : (long long)Sk * ldk;
  p.bsv = bsv > 0 ? bsv : (long long)Sk * ldv;
  p.bso = bso > 0 ? bso : (long long)Sq * ldo;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  p.drop = drop;
  return dispatch(p, dh, false, (cudaStream_t)stream);
}

extern "C" int gpvb200_attention_fwd_bs(const void* q, const void* k, const void* v, void* o, float* lse,
                                        const uint8_t* key_mask, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                        int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, int32_t B, int32_t H, int32_t Sq,
                                        int32_t Sk, int32_t dh, int32_t causal, float scale, void* stream) {
  return attn_fwd_impl(q, k, v, o, lse, key_mask, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, B, H, Sq, Sk, dh, causal, scale,
                       attn_drop(nullptr, 0, 0.f), stream);
}

extern "C" int gpvb200_attention_fwd_drop(const void* q, const void* k, const void* v, void* o, float* lse,
                                          const uint8_t* key_mask, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                          int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale,
                                          const void* drop_seed, uint32_t drop_site, float drop_p, void* stream) {
  if (!(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed))) {
    set_last_error("attention_fwd_drop: bad dropout arguments");
    return GPV_ERR_ARG;
  }
  return attn_fwd_impl(q, k, v, o, lse, key_mask, ldq, ldk, ldv, ldo, 0, 0, 0, 0, B, H, Sq, Sk, dh, causal, scale,
                       attn_drop(drop_seed, drop_site, drop_p), stream);
}

extern "C" int gpvb200_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                                     const uint8_t* key_mask, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                     int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale,
                                     void* stream) {
  return gpvb200_attention_fwd_bs(q, k, v, o, lse, key_mask, ldq, ldk, ldv, ldo, 0, 0, 0, 0, B, H, Sq, Sk, dh, causal, scale, stream);
}

extern "C" int gpvb200_attention_bwd_drop(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                          const float* lse, const uint8_t* key_mask, void* dq, void* dk, void* dv,
                                          int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq,
                                          int64_t lddk, int64_t lddv, int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh,
                                          int32_t causal, float scale, const void* drop_seed, uint32_t drop_site, float drop_p,
                                          void* stream) {
  int rc = ensure_arch();
  if (rc != GPV_OK) return rc;
  GPV_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv, "attention_bwd: null pointer");
  GPV_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0, "attention_bwd: bad shape");
  GPV_REQUIRE((ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (lddo % 8 == 0), "attention_bwd: row strides must be multiples of 8");
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)const_cast<void*>(o);
  p.d_o = (const bf16*)d_o; p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.lse = const_cast<float*>(lse); p.kmask = key_mask;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  GPV_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed), "attention_bwd: bad dropout arguments");
  p.drop = attn_drop(drop_seed, drop_site, drop_p);
  return dispatch(p, dh, true, (cudaStream_t)stream);
}

extern "C" int gpvb200_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                     const float* lse, const uint8_t* key_mask, void* dq, void* dk, void* dv,
                                     int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq,
                                     int64_t lddk, int64_t lddv, int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh,
                                     int32_t causal, float scale, void* stream) {
  return gpvb200_attention_bwd_drop(q, k, v, o, d_o, lse, key_mask, dq, dk, dv, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, B, H,
                                    Sq, Sk, dh, causal, scale, nullptr, 0, 0.f, stream);
}
