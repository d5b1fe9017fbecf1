// tc_common.cuh -- device helpers shared by the tcgen05 / TMEM kernels (sm_100a): activations, TF32 split, UMMA
// descriptors, TMEM loads/stores, mbarriers, bulk-copy (TMA) gathers, packed FMAs.
#pragma once
#include "kernels.cuh"

namespace b2m {

// sigmoid on MUFU.EX2 + MUFU.RCP with flush-to-zero: the same bits as __fdividef(1, 1 + __expf(-x)) wherever the result
// is a normal number (|x| < 87), without the three range fix-up instructions (FSETP + 2 predicated FMUL) the non-ftz
// forms carry -- activations are 35-45 % of the instructions of the fused tile kernels (profiles/r01_stalls_*.txt).
__device__ __forceinline__ float sigm_(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float silu_(float x) { return x * sigm_(x); }
__device__ __forceinline__ float dsilu_(float x) {
  const float sg = sigm_(x);
  return sg * (1.f + x * (1.f - sg));
}
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tf32_hi_bits(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// tf32 x tf32 -> f32, A from TMEM, B from smem descriptor, M=128
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = s_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// named barrier of one 256-thread row group (group g of a 512-thread CTA)
__device__ __forceinline__ void gbar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ float ipow_(float x, int n) {
  float r = 1.f;
  for (int i = 0; i < n; i++) r *= x;
  return r;
}
__device__ __forceinline__ float rbf_env_val(float d, float freq, const RadialParams& rp) {
  const float invd = 1.f / d;
  float s, c;
  sincosf(d * (freq / rp.rc), &s, &c);
  const float rbf = rp.norm * s * invd;
  const int p = rp.p;
  const float c1 = -(p + 1) * (p + 2) * 0.5f, c2 = (float)(p * (p + 2)), c3 = -p * (p + 1) * 0.5f;
  const float rho = rbf / rp.rc;
  const float r0 = ipow_(rho, p), r1 = r0 * rho, r2 = r1 * rho;
  const float env = 1.f + c1 * r0 + c2 * r1 + c3 * r2;
  return rbf <= rp.rc ? env * rbf : 0.f;
}

// Tile-interleaved layout of kernel-private row tensors ([rows][W] logically, W = 4*NC4 floats):
//   float4 index ((tile * NC4 + c/4) * 128 + r)  ->  a warp's thread=row access touches 512 contiguous bytes
// instead of 32 different cache lines (the LSU wavefront count of the row-major layout was the bottleneck).
template <int NC4>
__device__ __forceinline__ size_t tl4(int64_t tile, int r, int c) {
  return ((size_t)(tile * NC4 + (c >> 2)) * 128 + r);
}

// vectorised reduction: one L2 RED operation for 4 consecutive floats (sm_90+)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Streaming operands of the NEXT tile of a persistent CTA (saved u|v, silu', radial basis, index blocks: contiguous
// per tile in the tile-interleaved layouts) are pulled into L2 while the current tile computes: with two tiles in
// flight per SM there is not enough parallelism to hide a DRAM round trip behind other warps.  Pure hint.
__device__ __forceinline__ void l2_prefetch(const void* base, int bytes, int tid, int nthr) {
  const char* p = reinterpret_cast<const char*>(base);
  for (int o = tid * 128; o < bytes; o += nthr * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}

// packed fp32 FMA (FFMA2): (d0, d1) += a * (b0, b1), each half an IEEE fma -- same bits as two fmaf, one issue slot
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  uint64_t A, B, C, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
// w[0..15] = sum_k b[k] * WT[k][col0 .. col0+15]   (WT: [9][64] k-major in shared memory, col0 % 16 == 0)
__device__ __forceinline__ void radial_dot16(const float* WT, int col0, const float (&b)[9], float (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; k++) {
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++) {
      const float4 x = *reinterpret_cast<const float4*>(WT + k * 64 + col0 + c4 * 4);
      ffma2(w[4 * c4], w[4 * c4 + 1], b[k], x.x, x.y);
      ffma2(w[4 * c4 + 2], w[4 * c4 + 3], b[k], x.z, x.w);
    }
  }
}

__device__ __forceinline__ void radial_dot4(const float* WT, int col0, const float (&b)[9], float (&w)[4]) {
  w[0] = w[1] = w[2] = w[3] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const float4 x = *reinterpret_cast<const float4*>(WT + k * 64 + col0);
    ffma2(w[0], w[1], b[k], x.x, x.y);
    ffma2(w[2], w[3], b[k], x.z, x.w);
  }
}

// instruction descriptors: c=F32, a=b=TF32, K-major, M=128
constexpr uint32_t kIdescN64 = (1u << 4) | (2u << 7) | (2u << 10) | (8u << 17) | (8u << 24);
constexpr uint32_t kIdescN128 = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);

__device__ __forceinline__ void rbf_env_both(float d, float freq, const RadialParams& rp, float& be, float& dbe) {
  const float invd = 1.f / d;
  const float wq = freq / rp.rc;
  float s, c;
  sincosf(d * wq, &s, &c);
  const float rbf = rp.norm * s * invd;
  const float drbf = rp.norm * (wq * c * invd - s * invd * invd);
  const int p = rp.p;
  const float c1 = -(p + 1) * (p + 2) * 0.5f, c2 = (float)(p * (p + 2)), c3 = -p * (p + 1) * 0.5f;
  const float rho = rbf / rp.rc;
  const float rm1 = ipow_(rho, p - 1);
  const float r0 = rm1 * rho, r1 = r0 * rho, r2 = r1 * rho;
  const float env = 1.f + c1 * r0 + c2 * r1 + c3 * r2;
  const float denv = (c1 * p * rm1 + c2 * (p + 1) * r0 + c3 * (p + 2) * r1) / rp.rc;
  const bool ok = rbf <= rp.rc;
  be = ok ? env * rbf : 0.f;
  dbe = ok ? (env + rbf * denv) * drbf : 0.f;
}

__device__ __forceinline__ void mbar_expect_tx_(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   s_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}


__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr)
               : "memory");
}
// cp.async (LDGSTS) of 16 bytes, L2 only; completion is reported to an mbarrier by cp_async_arrive_ (one arrival per
// issuing thread, counted against the barrier's initial count: .noinc)
__device__ __forceinline__ void cp_async16_(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}

// packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE operations per issue slot
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float ex2_(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
__device__ __forceinline__ float rcp_(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
}  // namespace b2m
