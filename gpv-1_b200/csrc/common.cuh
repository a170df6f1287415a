// Shared device-side primitives for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers,
// warp reductions, bf16 packing. Everything here is raw PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpv {

typedef __nv_bfloat16 bf16;

#define GPV_DEVINL __device__ __forceinline__

GPV_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  ptxas recognises elect.sync: code guarded by it is known to run in a single thread, so the
// operands of warp-level (uniform-datapath) instructions such as UTCHMMA / UTMALDG move to uniform registers with one R2UR each
// instead of a per-active-lane "waterfall" loop (which is what `if (lane == 0)` compiles to).
GPV_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
GPV_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
GPV_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GPV_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
GPV_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GPV_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
GPV_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error at the next sync), never as a
// hung GPU. ~4e9 cycles is about two seconds at boost clock.
GPV_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------- TMA (tiled, 4-D)
GPV_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
GPV_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- CTA pair (cta_group::2: the two SMs of a TPC, cluster of 2)
GPV_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
GPV_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in the pair's leader (CTA rank 0)
GPV_DEVINL uint32_t leader_smem_addr(uint32_t cta_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(cta_addr));
  return r;
}
// This CTA's part of a pair operand into its OWN shared memory; the bytes are counted on `bar_cluster`, a
// shared::cluster address (the leader's full barrier).
GPV_DEVINL void tma_load_4d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
GPV_DEVINL void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
GPV_DEVINL void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
GPV_DEVINL void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
GPV_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA over the pair, issued by ONE thread of the leader: each CTA's tensor core reads its own 128 rows of A and
// both halves of B (its own and, through the pair's shared memory, the peer's) and accumulates into its own TMEM.
GPV_DEVINL void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have completed.
GPV_DEVINL void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
GPV_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
GPV_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
GPV_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
GPV_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
GPV_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
GPV_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
GPV_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Same with the A operand in tensor memory (lane = row, 32-bit column c = elements 2c (low half), 2c + 1 of the row): no
// shared-memory read for A, which is what bounds narrow-N SS-mode MMAs (A is 4 KB per 128 x N x 16 step).
GPV_DEVINL void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: thread t writes lane (base_lane + t), 32 consecutive 32-bit columns
GPV_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
GPV_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
GPV_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
GPV_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread t gets lane (base_lane + t), columns [col, col+16).
GPV_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns.
GPV_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation
// (bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor).
GPV_DEVINL uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 1u << 7;                      // a_format = BF16
  d |= 1u << 10;                     // b_format = BF16
  d |= (uint32_t)(a_mn & 1) << 15;   // a_major: 0 K, 1 MN
  d |= (uint32_t)(b_mn & 1) << 16;   // b_major
  d |= (uint32_t)(N >> 3) << 17;     // n_dim
  d |= (uint32_t)(M >> 4) << 24;     // m_dim
  return d;
}
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field set.
GPV_DEVINL uint64_t make_sdesc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // LayoutType::SWIZZLE_128B
  return d;
}

// ---------------------------------------------------------------- programmatic dependent launch
// First statement of every kernel launched through launch_k(): wait until the previous kernel of the stream has
// completed (no-op without the launch attribute), then let the next kernel start its own prologue.
GPV_DEVINL void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- dropout masks (train mode)
// Counter-based, so that backward regenerates the forward's mask instead of storing it.  Element (row, col) of a
// logical [rows, N] tensor at dropout site `site` of training step `seed`:
//     pair = row * ceil(N / 2) + (col >> 1);  bits = drop_bits(drop_key(seed, site), pair)
//     keep = ((col & 1) ? bits >> 16 : bits & 0xFFFF) >= thresh16,   thresh16 = round(p * 65536)
// (two elements per 32-bit hash; p = 0.1 -> 0.100006).  Kept elements are scaled by 1 / (1 - p) like nn.Dropout.
// The same formula serves GEMM epilogues (row = output row), LayerNorm (row = token) and attention probabilities
// (row = (b*H + h)*Sq + q, col = key).  The reference draws its masks from torch's Philox stream; masks cannot be
// bit-compatible across implementations, so parity runs with dropout off and dropout is tested against autograd with
// the masks exported by gpvb200_dropout_mask (tests/test_dropout_gpu.py).
struct DropArgs {
  const unsigned long long* seed;   // device scalar, bumped once per training step; nullptr = no dropout
  uint32_t site, thresh16;
  float scale;                      // 1 / (1 - p)
};
GPV_DEVINL uint32_t drop_key(unsigned long long seed, uint32_t site) {
  uint32_t x = (uint32_t)seed * 0x9E3779B1u ^ (uint32_t)(seed >> 32) ^ (site * 0x85EBCA6Bu + 0x6C62272Eu);
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  return x;
}
GPV_DEVINL uint32_t drop_bits(uint32_t key, uint32_t pair) {
  uint32_t x = pair * 0x9E3779B1u + key;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
// Applies the mask to the two elements of one pair.
GPV_DEVINL void drop_pair(float& a, float& b, uint32_t key, uint32_t pair, uint32_t thresh16, float scale) {
  const uint32_t bits = drop_bits(key, pair);
  a = ((bits & 0xFFFFu) >= thresh16) ? a * scale : 0.0f;
  b = ((bits >> 16) >= thresh16) ? b * scale : 0.0f;
}

// ---------------------------------------------------------------- misc math
GPV_DEVINL float ex2_approx(float x) {      // one MUFU.EX2 (exp2f() without fast-math adds a denormal-range rescale around it)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- packed fp32 pairs (FFMA2 / FADD2 on sm_100: one issue slot for two lanes of work)
GPV_DEVINL uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
GPV_DEVINL void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
GPV_DEVINL uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
GPV_DEVINL uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^t for two values t <= 0 on the FMA pipe instead of the SFU (which retires only 4 lanes per clock per scheduler and bounds a
// softmax over short head dimensions): j = round(t) by the 1.5 * 2^23 trick, f = t - j in [-0.5, 0.5], a cubic minimax fit of 2^f
// (relative error 7.5e-5, far below the bf16 rounding of the probabilities), and j added into the exponent field.
GPV_DEVINL void exp2_poly2(float t0, float t1, float& p0, float& p1) {
  t0 = fmaxf(t0, -125.0f);
  t1 = fmaxf(t1, -125.0f);
  const uint64_t T = pk2(t0, t1);
  const uint64_t R = fadd2(T, pk2(12582912.0f, 12582912.0f));
  const uint64_t J = fadd2(R, pk2(-12582912.0f, -12582912.0f));
  const uint64_t F = ffma2(J, pk2(-1.0f, -1.0f), T);
  uint64_t P = ffma2(F, pk2(0.055171459913253784f, 0.055171459913253784f), pk2(0.2426108568906784f, 0.2426108568906784f));
  P = ffma2(P, F, pk2(0.6932609677314758f, 0.6932609677314758f));
  P = ffma2(P, F, pk2(0.9999281167984009f, 0.9999281167984009f));
  float r0, r1, q0, q1;
  upk2(R, r0, r1);
  upk2(P, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}
GPV_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
GPV_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
GPV_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
GPV_DEVINL float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
GPV_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
GPV_DEVINL float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x);
}

// ---------------------------------------------------------------- shared-memory slabs, cp.async, bf16 vectors
// (used by the epilogues of gemm_umma.cu and layer_umma.cu; see the coalesced-epilogue note in gemm_umma.cu)
GPV_DEVINL uint4 ldg_u4(const bf16* ptr) { return __ldg(reinterpret_cast<const uint4*>(ptr)); }

GPV_DEVINL int slab_slot(int row, int c) { return row * 4 + (c ^ ((row >> 1) & 3)); }
GPV_DEVINL void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
GPV_DEVINL uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
GPV_DEVINL void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
GPV_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
GPV_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
GPV_DEVINL void cp_async_wait_n(int n) {   // n is a compile-time value after unrolling
  if (n <= 0) cp_async_wait<0>();
  else if (n == 1) cp_async_wait<1>();
  else if (n == 2) cp_async_wait<2>();
  else cp_async_wait<3>();
}
GPV_DEVINL void prefetch_l2_bulk(const void* g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
GPV_DEVINL void prefetch_l2_line(const void* g) { asm volatile("prefetch.global.L2 [%0];" ::"l"(g) : "memory"); }
// One row slice of `bytes` bytes into L2: mode 1 = one bulk (TMA-engine) request, mode 2 = one LSU prefetch per 128-byte line.
GPV_DEVINL void prefetch_l2_row(const bf16* g, uint32_t bytes, int mode) {
  if (mode == 1) {
    prefetch_l2_bulk(g, bytes);
  } else {
    for (uint32_t b = 0; b < bytes; b += 128) prefetch_l2_line(reinterpret_cast<const uint8_t*>(g) + b);
  }
}

GPV_DEVINL void unpack8(const uint4& q, float* v, bool add) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    if (add) {
      v[2 * j] += f.x;
      v[2 * j + 1] += f.y;
    } else {
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  }
}
GPV_DEVINL uint4 pack8(const float* v) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}


}  // namespace gpv
