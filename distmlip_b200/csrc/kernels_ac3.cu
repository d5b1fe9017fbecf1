// kernels_ac3.cu -- third generation of the atom-conv ("edge-gather") kernels, sm_100a.
//
// What the round-2 measurements said about the first two generations (profiles/r02b_*):
//   * k_atomconv_fwd_tc / _v2 run at 37-41 % issue utilisation with 16 warps per SM: they are LATENCY bound -- every tile
//     is a serial chain  be/idx loads -> GEMM1 -> gather wait -> epilogue -> GEMM2 -> epilogue -> GEMM2 -> output
//     and only two tiles fit an SM (256 TMEM columns each);
//   * bulk copies (UBLKCP) cost 17-22 SM clocks EACH whatever their size (128 x 512 B: 2795 clk per tile, 256 x 256 B:
//     4464 clk), per-lane LDG.128 of thread-owned rows 4161 clk, cp.async (LDGSTS) of coalesced rows 2275 clk
//     (profiles/r02b_gather_probe.txt) -- so replacing LDGs by bulk copies alone bought nothing (v2 == v1);
//   * the MUFU pipe is 30 % busy (0.33 ms of work per launch at 97 k atoms) and ~5000 instructions are issued per edge:
//     both are within 2x of the kernel's duration, i.e. they are the floor once the latency chain is hidden.
// This generation therefore keeps the tcgen05 data flow (A operands written to TMEM from registers, 3xTF32, weights
// staged once per persistent CTA, two 256-thread row groups per SM) and removes exposed latency:
//   * A[src] half rows (256 B per edge and branch) arrive by cp.async, 16 lanes per row, into TWO padded stages per group,
//     each refilled for the NEXT tile as soon as its branch has been consumed -> a gather has a whole tile to land;
//     completion through cp.async.mbarrier.arrive (no thread waits on its own copies);
//   * radial basis rows and indices of the next tile are requested one tile ahead and parked in registers;
//   * the epilogue of the second branch computes its first 16 columns while the tensor core still reads the H operand of
//     the first branch; the shared bond weights w_ab are computed while the last GEMM runs;
//   * the second-layer pre-activations are NOT saved (the backward recomputes them) unless uv_save is given;
//   * sigmoid(v) silu(u) = u / ((1 + 2^-u')(1 + 2^-v')) with ONE reciprocal; no per-element `valid` selects (rows past the
//     end of the edge list carry dst = -1 and are dropped by the segmented sum; tensor-core rows are independent);
//   * segmented sum over destination runs on a [128][68] message tile (aliases the drained stage), 16-byte LDS and one
//     red.global.add.v4.f32 per run end.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b2m {

struct Ac3Smem {
  static constexpr int kBar = 0;                 // per group 8 mbarriers: [0..5] MMA, [6] stage0, [7] stage1; tmem ptr at float 48
  static constexpr int kW2 = 64;                 // 4 x 4096 (W2L hi, W2L lo, W2G hi, W2G lo), canonical K-major
  static constexpr int kM = kW2 + 4 * 4096;      // 2 x 2048 (M hi, lo), N=128, K=16
  static constexpr int kWab = kM + 2 * 2048;     // 576, k-major [9][64]
  static constexpr int kB2 = kWab + 576;         // 128
  static constexpr int kGrp = kB2 + 128;         // per group:
  static constexpr int kPitch = 68;              //   stage0 [128][68] (G half rows), stage1 [128][68] (L half rows | message tile)
  static constexpr int kStage = 128 * kPitch;
  static constexpr int kIdxOff = 2 * kStage;     //   idx [2][3][128] ints
  static constexpr int kGrpSize = kIdxOff + 2 * 3 * 128;
  static constexpr int kTotal = kGrp + 2 * kGrpSize;
  static constexpr size_t bytes = (size_t)kTotal * 4;
};
static_assert(Ac3Smem::bytes <= 232448, "shared memory budget");

constexpr float kNegLog2e = -1.4426950408889634f;

// 16 columns of the first layer: pre = D + A[src] + C[dst] (| Q[bond] + A + C), hid = silu(pre) -> tf32 hi / lo
__device__ __forceinline__ void ac3_first16(uint32_t taddr, const float* Ast, const float* Crow, const float* Qrow,
                                            uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  float4 av[4], cv[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    av[i] = *reinterpret_cast<const float4*>(Ast + 4 * i);
    cv[i] = *reinterpret_cast<const float4*>(Crow + 4 * i);
  }
  tc_wait_ld();
  if (Qrow != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 x = *reinterpret_cast<const float4*>(Qrow + 4 * i);
      v[4 * i] = __float_as_uint(x.x), v[4 * i + 1] = __float_as_uint(x.y);
      v[4 * i + 2] = __float_as_uint(x.z), v[4 * i + 3] = __float_as_uint(x.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float a4[4] = {av[i].x, av[i].y, av[i].z, av[i].w}, c4[4] = {cv[i].x, cv[i].y, cv[i].z, cv[i].w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float p = (__uint_as_float(v[4 * i + j]) + a4[j]) + c4[j];
      const float hval = p * rcp_(1.f + ex2_(p * kNegLog2e));
      const uint32_t h = tf32_hi_bits(hval);
      hi[4 * i + j] = h;
      lo[4 * i + j] = __float_as_uint(hval - __uint_as_float(h));
    }
  }
}

__device__ __forceinline__ void radial_dot16p(const float* WT, int col0, const float (&b)[9], float* w) {
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

__global__ void __launch_bounds__(512, 1) k_atomconv_fwd_v3(const AtomConvArgs a, const AtomConvTcW w) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x;
  const int g = tid >> 8, gt = tid & 255;
  const int warp = gt >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;  // my row (TMEM lane)
  const int c0 = half * 32;     // my columns inside each 64-wide branch
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + Ac3Smem::kBar) + g * 8;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + Ac3Smem::kBar + 48);
  float* W2s = smem + Ac3Smem::kW2;
  float* Ms = smem + Ac3Smem::kM;
  float* wabW = smem + Ac3Smem::kWab;
  float* b2s = smem + Ac3Smem::kB2;
  float* grp = smem + Ac3Smem::kGrp + g * Ac3Smem::kGrpSize;
  float* stage0 = grp;                      // G-branch half rows of A[src]
  float* stage1 = grp + Ac3Smem::kStage;    // L-branch half rows; message tile of the segmented sum once drained
  int* idx = reinterpret_cast<int*>(grp + Ac3Smem::kIdxOff);  // [buf][src|dst|bond][128]
  constexpr int PITCH = Ac3Smem::kPitch;
  const bool useQ = a.Qproj != nullptr;

  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (gt == 0) {
    for (int i = 0; i < 6; i++) mbar_init_(&mbar[i], 1);
    mbar_init_(&mbar[6], 256);  // one cp.async arrival per thread of the group
    mbar_init_(&mbar[7], 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * 1024; i += 512) reinterpret_cast<float4*>(W2s)[i] = reinterpret_cast<const float4*>(w.W2can)[i];
  for (int i = tid; i < 2 * 512; i += 512) reinterpret_cast<float4*>(Ms)[i] = reinterpret_cast<const float4*>(w.Mcan)[i];
  for (int i = tid; i < 576; i += 512) wabW[(i % 9) * 64 + i / 9] = a.Wabw[i];
  if (tid < 128) b2s[tid] = a.b2[tid];

  const int64_t ntiles = (a.E + 127) / 128;
  const int64_t tstride = 2 * (int64_t)gridDim.x;
  const int64_t t_first = 2 * (int64_t)blockIdx.x + g;
  // one 64-column half (br) of A[src] for the rows of a tile: 16 lanes x 16 B per row, 8 rows per thread
  auto issue_gather = [&](const int* ib, int nv, int br, float* st, uint64_t* gb) {
    const int cc = (gt & 15) * 4;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int row = (gt >> 4) + 16 * j;
      if (row < nv) cp_async16_(st + row * PITCH + cc, a.Aproj + (size_t)ib[row] * D2 + br * 64 + cc);
    }
    cp_async_arrive_(gb);
  };
  auto load_be = [&](int64_t t, float4 (&b)[3]) {
    const float4* bp = reinterpret_cast<const float4*>(a.be);
    b[0] = bp[tl4<3>(t, r, 0)], b[1] = bp[tl4<3>(t, r, 4)], b[2] = bp[tl4<3>(t, r, 8)];
  };
  float4 benext[3];
  if (t_first < ntiles) {
    if (gt < 128) {
      int src = 0, dst = -1, bond = -1;
      const int64_t e = t_first * 128 + gt;
      if (e < a.E) src = a.e_src[e], dst = a.e_dst[e], bond = a.e_bond[e];
      idx[gt] = src, idx[128 + gt] = dst, idx[256 + gt] = bond;
    }
    load_be(t_first, benext);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr + (uint32_t)g * 256u;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t COL_H = 0, COL_D = 128;
  const uint32_t w2_addr = s_u32(W2s), m_addr = s_u32(Ms);
  uint32_t phase = 0;
  if (t_first < ntiles) {
    const int nv = (int)min((int64_t)128, a.E - t_first * 128);
    issue_gather(idx, nv, 1, stage0, &mbar[6]);
    issue_gather(idx, nv, 0, stage1, &mbar[7]);
  }

  int buf = 0;
  for (int64_t t = t_first; t < ntiles; t += tstride, buf ^= 1) {
    const int64_t e0 = t * 128;
    const int nvalid = (int)min((int64_t)128, a.E - e0);
    const int64_t tn = t + tstride;
    const bool have_next = tn < ntiles;
    const int nvnext = have_next ? (int)min((int64_t)128, a.E - tn * 128) : 0;
    const int* s_src = idx + buf * 384;
    const int* s_dst = s_src + 128;
    const int* s_bond = s_src + 256;
    int* ibn = idx + (buf ^ 1) * 384;
    // indices of my next tile: requested now, written to the other index buffer after the first epilogue
    int nsrc = 0, ndst = -1, nbond = -1;
    if (have_next && gt < 128) {
      const int64_t e = tn * 128 + gt;
      if (e < a.E) nsrc = a.e_src[e], ndst = a.e_dst[e], nbond = a.e_bond[e];
    }
    float bek[9];
    bek[0] = benext[0].x, bek[1] = benext[0].y, bek[2] = benext[0].z, bek[3] = benext[0].w, bek[4] = benext[1].x;
    bek[5] = benext[1].y, bek[6] = benext[1].z, bek[7] = benext[1].w, bek[8] = benext[2].x;
    if (r >= nvalid) {
#pragma unroll
      for (int k = 0; k < 9; k++) bek[k] = 0.f;
    }
    if (half == 0) {  // be -> TMEM operand (K = 16: 9 values + zero pad), hi in cols 0..15, lo in 16..31
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const float x = k < 9 ? bek[k] : 0.f;
        const uint32_t h = tf32_hi_bits(x);
        hi[k] = h;
        lo[k] = __float_as_uint(x - __uint_as_float(h));
      }
      tmem_st16(tlane + COL_H, hi);
      tmem_st16(tlane + COL_H + 16, lo);
    }
    tc_wait_st();
    tc_fence_before();
    gbar(g);
    if (gt == 0) {  // GEMM1: D[128 x 128] = be[128 x 16] . M^T   (hi*hi + lo*hi + hi*lo)
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 16u : 0u;
        const uint32_t bsel = m_addr + (term == 2 ? 2048u * 4u : 0u);
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          umma_ts(tbase + COL_D, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 4096u, 2048u, 128u), kIdescN128, acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[0]);
    }
    const int dst = s_dst[r], bond = s_bond[r];
    const bool valid = r < nvalid;
    const bool viaQ = useQ && bond >= 0;
    const float* Crow = a.Cproj + (size_t)(valid ? dst : 0) * D2 + c0;
    const float* Qrow = viaQ ? a.Qproj + (size_t)bond * D2 + c0 : nullptr;

    // ---------------- gate branch (first-layer columns 64..127) ----------------
    mbar_wait_(&mbar[6], phase);  // stage0 has landed (requested one tile ago)
    mbar_wait_(&mbar[0], phase);  // GEMM1
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t hi[16], lo[16];
      ac3_first16(tlane + COL_D + 64 + c0 + ch * 16, stage0 + r * PITCH + c0 + ch * 16, Crow + 64 + ch * 16,
                  viaQ ? Qrow + 64 + ch * 16 : nullptr, hi, lo);
      tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
      tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
    }
    if (have_next && gt < 128) ibn[gt] = nsrc, ibn[128 + gt] = ndst, ibn[256 + gt] = nbond;
    tc_wait_st();
    tc_fence_before();
    gbar(g);  // H complete, stage0 drained, next indices visible
    if (gt == 0) {  // GEMM2 (gates): D[:, 64..127] = hid . W2G^T
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 64u : 0u;
        const uint32_t bsel = w2_addr + (uint32_t)(2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          umma_ts(tbase + COL_D + 64, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64,
                  acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[1]);
    }
    if (have_next) issue_gather(ibn, nvnext, 1, stage0, &mbar[6]);  // gate half of my NEXT tile

    // ---------------- layer branch (first-layer columns 0..63) ----------------
    mbar_wait_(&mbar[7], phase);  // stage1
    {
      uint32_t hi[16], lo[16];
      ac3_first16(tlane + COL_D + c0, stage1 + r * PITCH + c0, Crow, viaQ ? Qrow : nullptr, hi, lo);
      mbar_wait_(&mbar[1], phase);  // the tensor core has finished reading H (gate branch)
      tc_fence_after();
      tmem_st16(tlane + COL_H + c0, hi);
      tmem_st16(tlane + COL_H + 64 + c0, lo);
      ac3_first16(tlane + COL_D + c0 + 16, stage1 + r * PITCH + c0 + 16, Crow + 16, viaQ ? Qrow + 16 : nullptr, hi, lo);
      tmem_st16(tlane + COL_H + c0 + 16, hi);
      tmem_st16(tlane + COL_H + 64 + c0 + 16, lo);
    }
    tc_wait_st();
    tc_fence_before();
    gbar(g);  // H complete, stage1 drained (it becomes the message tile)
    if (gt == 0) {  // GEMM2 (layers): D[:, 0..63] = hid . W2L^T
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 64u : 0u;
        const uint32_t bsel = w2_addr + (uint32_t)(term == 2 ? 1 : 0) * 4096u * 4u;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          umma_ts(tbase + COL_D, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64, acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[2]);
    }
    // while the tensor core works: radial basis of my next tile, shared bond weights of this one
    if (have_next) load_be(tn, benext);
    float wab[32];
    radial_dot16p(wabW, c0, bek, wab);
    radial_dot16p(wabW, c0 + 16, bek, wab + 16);
    mbar_wait_(&mbar[2], phase);
    tc_fence_after();
    float* msg = stage1;
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t u[16], v[16];
      tmem_ld16(tlane + COL_D + c0 + ch * 16, u);
      tmem_ld16(tlane + COL_D + 64 + c0 + ch * 16, v);
      tc_wait_ld();
      float uu[16], vv[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int c = c0 + ch * 16 + i;
        uu[i] = __uint_as_float(u[i]) + b2s[c];
        vv[i] = __uint_as_float(v[i]) + b2s[64 + c];
      }
      if (a.uv_save != nullptr && valid) {
        float4* puv = reinterpret_cast<float4*>(a.uv_save);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          puv[tl4<32>(t, r, c0 + ch * 16 + 4 * i)] = make_float4(uu[4 * i], uu[4 * i + 1], uu[4 * i + 2], uu[4 * i + 3]);
          puv[tl4<32>(t, r, 64 + c0 + ch * 16 + 4 * i)] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        }
      }
      float mv[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        // silu(u) sigmoid(v) w = u w / ((1 + 2^(-u log2e)) (1 + 2^(-v log2e))): one reciprocal; an overflowing product
        // gives rcp(inf) = 0, the correct limit
        const float den = (1.f + ex2_(uu[i] * kNegLog2e)) * (1.f + ex2_(vv[i] * kNegLog2e));
        mv[i] = uu[i] * wab[ch * 16 + i] * rcp_(den);
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
        *reinterpret_cast<float4*>(msg + r * PITCH + c0 + ch * 16 + 4 * i) = make_float4(mv[4 * i], mv[4 * i + 1], mv[4 * i + 2], mv[4 * i + 3]);
    }
    tc_fence_before();
    gbar(g);
    {  // segmented sum over destination runs: 16 threads x 4 columns per row, 16 parts x 8 rows
      const int c4 = (gt & 15) * 4, part = gt >> 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur = -1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = part * 8 + i;
        const int k = s_dst[row];
        if (k != cur) {
          if (cur >= 0) red_add_v4(&a.agg[(size_t)cur * D + c4], s.x, s.y, s.z, s.w);
          cur = k;
          s = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (k >= 0) {
          const float4 m = *reinterpret_cast<const float4*>(msg + row * PITCH + c4);
          s.x += m.x, s.y += m.y, s.z += m.z, s.w += m.w;
        }
      }
      if (cur >= 0) red_add_v4(&a.agg[(size_t)cur * D + c4], s.x, s.y, s.z, s.w);
    }
    gbar(g);  // message tile drained: stage1 may be refilled
    if (have_next) issue_gather(ibn, nvnext, 0, stage1, &mbar[7]);
    phase ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tptr), "r"(512u));
}

void launch_atomconv_fwd_v3(cudaStream_t st, const AtomConvArgs& a, const AtomConvTcW& w, int num_sms) {
  if (a.E <= 0) return;
  static PerDeviceOnce attr;
  if (attr.first()) {
    B2M_CK(cudaFuncSetAttribute(k_atomconv_fwd_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ac3Smem::bytes));
  }
  const int64_t ntiles = (a.E + 127) / 128;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, (int64_t)num_sms);
  k_atomconv_fwd_v3<<<grid, 512, Ac3Smem::bytes, st>>>(a, w);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

}  // namespace b2m
