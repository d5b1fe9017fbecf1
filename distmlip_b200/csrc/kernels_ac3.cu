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
    b[0] = __ldcs(bp + tl4<3>(t, r, 0)), b[1] = __ldcs(bp + tl4<3>(t, r, 4)), b[2] = __ldcs(bp + tl4<3>(t, r, 8));  // streamed once per launch
  };
  float4 benext[3];
  if (t_first < ntiles) {
    if (gt < 128) {
      int src = 0, dst = -1, bond = -1;
      const int64_t e = t_first * 128 + gt;
      if (e < a.E) src = __ldcs(a.e_src + e), dst = __ldcs(a.e_dst + e), bond = __ldcs(a.e_bond + e);
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
      if (e < a.E) nsrc = __ldcs(a.e_src + e), ndst = __ldcs(a.e_dst + e), nbond = __ldcs(a.e_bond + e);
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
    if (have_next && gt < 128) {
      ibn[gt] = nsrc, ibn[128 + gt] = ndst, ibn[256 + gt] = nbond;
      if (w.l2pf) {  // C[dst] rows of my next tile into L1 (one request per destination run)
        const int pd = __shfl_up_sync(0xffffffffu, ndst, 1);
        if (ndst >= 0 && (lane == 0 || pd != ndst)) {
          const char* cp = reinterpret_cast<const char*>(a.Cproj + (size_t)ndst * D2);
#pragma unroll
          for (int i = 0; i < 4; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(cp + 128 * i));
        }
      }
    }
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
          // streaming stores (evict-first): 512 B/edge written once here and read once by the backward must not push
          // the gathered A / C rows out of L2
          __stcs(&puv[tl4<32>(t, r, c0 + ch * 16 + 4 * i)], make_float4(uu[4 * i], uu[4 * i + 1], uu[4 * i + 2], uu[4 * i + 3]));
          __stcs(&puv[tl4<32>(t, r, 64 + c0 + ch * 16 + 4 * i)], make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]));
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

// ============================================================================================
// atom conv backward, third generation.  Same staging as the forward (two cp.async stages per row group, refilled for the
// next tile as soon as they are drained; be / dbe / saved u|v of the next tile pulled into L2 one tile ahead), plus:
//   * silu'(pre) of a branch is parked IN PLACE of the staged A half row it was computed from (each thread owns its 32
//     columns of its row), turned into gpre = silu'(pre) * ghid in place once the tensor core has delivered ghid, and the
//     SAME tile is then the scatter staging ([128][68], 16-byte LDS): no separate scatter buffer, no 32 live registers
//     across the tensor-core waits;
//   * the saved u | v are read ONCE: both second-layer gradients come out of one pass (sigmoid(u), sigmoid(v) from one
//     reciprocal), gv goes to the tensor core at once and gu waits in registers; the layer branch's silu' is computed
//     while the tensor core multiplies the gate branch, and the gate branch's scatter runs while it multiplies the layers;
//   * dE/dd through the radial first-layer term, sum_j gpre[j] (M dbe)[j], takes M dbe from one more K = 16 tensor-core
//     product (dbe . M^T) instead of a 9-term dot per column: no row-major copy of M in shared memory (the budget is
//     W2^T 64 KB + M 16 KB + 4 stages 136 KB of 227 KB) and 27 % fewer instructions.
// Reads the second-layer pre-activations u|v saved by the forward (see DESIGN.md for why they are not recomputed: W2 and
// W2^T do not fit shared memory together next to the stages, and tcgen05 refused the MN-major view of W2, profiles/r02b).
struct Ac3BwdSmem {
  static constexpr int kBar = 0;                 // per group 8 mbarriers: [0] GEMM1, [1] GEMM3 G, [2] GEMM3 L, [6] stage0, [7] stage1
  static constexpr int kW2T = 64;                // 4 x 4096 (W2L^T hi, lo, W2G^T hi, lo)
  static constexpr int kM = kW2T + 4 * 4096;     // 2 x 2048 (M hi, lo)
  static constexpr int kWab = kM + 2 * 2048;     // 576, k-major [9][64]
  static constexpr int kGrp = kWab + 576;        // per group:
  static constexpr int kPitch = 68;
  static constexpr int kStage = 128 * kPitch;    //   stage0 (G), stage1 (L)
  static constexpr int kIdxOff = 2 * kStage;     //   idx [2][3][128] ints
  static constexpr int kGdOff = kIdxOff + 2 * 3 * 128;  // gd partials [256]
  static constexpr int kGrpSize = kGdOff + 256;
  static constexpr int kTotal = kGrp + 2 * kGrpSize;
  static constexpr size_t bytes = (size_t)kTotal * 4;
};
static_assert(Ac3BwdSmem::bytes <= 232448, "shared memory budget");

// 16 columns of silu'(first-layer pre-activation), parked IN PLACE of the staged A half row they were computed from
__device__ __forceinline__ void ac3_ds16(uint32_t taddr, float* Ast, const float4 (&cv)[4], const float* Qrow) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  float4 av[4];
#pragma unroll
  for (int i = 0; i < 4; i++) av[i] = *reinterpret_cast<const float4*>(Ast + 4 * i);
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
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float p = (__uint_as_float(v[4 * i + j]) + a4[j]) + c4[j];
      const float sg = rcp_(1.f + ex2_(p * kNegLog2e));
      d[j] = sg * (1.f + p * (1.f - sg));  // silu'(p)
    }
    *reinterpret_cast<float4*>(Ast + 4 * i) = make_float4(d[0], d[1], d[2], d[3]);
  }
}

// second-layer gradients of 16 columns from ONE pass over the saved u | v: gv (gate branch) -> tf32 hi / lo for the tensor
// core now, gu (layer branch) kept in fp32 for the second product; dE/dd through the shared bond weights w_ab
__device__ __forceinline__ void ac3_second16(const float4* uv4, size_t uvo, const float* gmrow, const float* wabW, int c,
                                             const float (&bek)[9], const float (&dbek)[9], float& gdpart, float* gu,
                                             uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float4 u4 = __ldcs(uv4 + uvo + (size_t)i * 128);  // read once: streaming
    const float4 v4 = __ldcs(uv4 + uvo + (size_t)(16 + i) * 128);
    const float4 g4 = *reinterpret_cast<const float4*>(gmrow + 4 * i);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float wab4[4], wabp4[4];
    radial_dot4(wabW, c + 4 * i, bek, wab4);
    radial_dot4(wabW, c + 4 * i, dbek, wabp4);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      // sigmoid(u), sigmoid(v) from ONE reciprocal; the exponentials are clamped so that the product stays finite
      const float eu = fminf(ex2_(uu[j] * kNegLog2e), 1e18f), ev = fminf(ex2_(vv[j] * kNegLog2e), 1e18f);
      const float rr = rcp_((1.f + eu) * (1.f + ev));
      const float su = rr * (1.f + ev), sv = rr * (1.f + eu);
      const float oL = uu[j] * su, gw = gg[j] * wab4[j];
      gu[4 * i + j] = gw * sv * (su * (1.f + uu[j] * (1.f - su)));
      gdpart = fmaf(gg[j] * oL * sv, wabp4[j], gdpart);
      const float gv = gw * oL * sv * (1.f - sv);
      const uint32_t h = tf32_hi_bits(gv);
      hi[4 * i + j] = h;
      lo[4 * i + j] = __float_as_uint(gv - __uint_as_float(h));
    }
  }
}

__global__ void __launch_bounds__(512, 1) k_atomconv_bwd_v3(const AtomConvArgs a, const AtomConvTcW w) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x;
  const int g = tid >> 8, gt = tid & 255;
  const int warp = gt >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;
  const int c0 = half * 32;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + Ac3BwdSmem::kBar) + g * 8;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + Ac3BwdSmem::kBar + 48);
  float* W2Ts = smem + Ac3BwdSmem::kW2T;
  float* Ms = smem + Ac3BwdSmem::kM;
  float* wabW = smem + Ac3BwdSmem::kWab;
  float* grp = smem + Ac3BwdSmem::kGrp + g * Ac3BwdSmem::kGrpSize;
  float* stage0 = grp;
  float* stage1 = grp + Ac3BwdSmem::kStage;
  int* idx = reinterpret_cast<int*>(grp + Ac3BwdSmem::kIdxOff);
  float* gdb = grp + Ac3BwdSmem::kGdOff;
  constexpr int PITCH = Ac3BwdSmem::kPitch;
  const bool useQ = a.Qproj != nullptr;
  const bool need_gx = a.gA != nullptr;

  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (gt == 0) {
    for (int i = 0; i < 6; i++) mbar_init_(&mbar[i], 1);
    mbar_init_(&mbar[6], 256);
    mbar_init_(&mbar[7], 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * 1024; i += 512) reinterpret_cast<float4*>(W2Ts)[i] = reinterpret_cast<const float4*>(w.W2Tcan)[i];
  for (int i = tid; i < 2 * 512; i += 512) reinterpret_cast<float4*>(Ms)[i] = reinterpret_cast<const float4*>(w.Mcan)[i];
  for (int i = tid; i < 576; i += 512) wabW[(i % 9) * 64 + i / 9] = a.Wabw[i];

  const int64_t ntiles = (a.E + 127) / 128;
  const int64_t tstride = 2 * (int64_t)gridDim.x;
  const int64_t t_first = 2 * (int64_t)blockIdx.x + g;
  auto issue_gather = [&](const int* ib, int nv, int br, float* st, uint64_t* gb) {
    const int cc = (gt & 15) * 4;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int row = (gt >> 4) + 16 * j;
      if (row < nv) cp_async16_(st + row * PITCH + cc, a.Aproj + (size_t)ib[row] * D2 + br * 64 + cc);
    }
    cp_async_arrive_(gb);
  };
  auto load_be = [&](int64_t t, float4 (&b)[3], float4 (&d)[3]) {
    const float4* bp = reinterpret_cast<const float4*>(a.be);
    const float4* dp = reinterpret_cast<const float4*>(a.dbe);
    b[0] = __ldcs(bp + tl4<3>(t, r, 0)), b[1] = __ldcs(bp + tl4<3>(t, r, 4)), b[2] = __ldcs(bp + tl4<3>(t, r, 8));  // streamed once per launch
    d[0] = __ldcs(dp + tl4<3>(t, r, 0)), d[1] = __ldcs(dp + tl4<3>(t, r, 4)), d[2] = __ldcs(dp + tl4<3>(t, r, 8));
  };
  float4 benext[3], dbenext[3];
  if (t_first < ntiles) {
    if (gt < 128) {
      int src = 0, dst = -1, bond = -1;
      const int64_t e = t_first * 128 + gt;
      if (e < a.E) src = __ldcs(a.e_src + e), dst = __ldcs(a.e_dst + e), bond = __ldcs(a.e_bond + e);
      idx[gt] = src, idx[128 + gt] = dst, idx[256 + gt] = bond;
    }
    load_be(t_first, benext, dbenext);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr + (uint32_t)g * 256u;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t COL_H = 0, COL_D = 128;
  const uint32_t w2t_addr = s_u32(W2Ts), m_addr = s_u32(Ms);
  uint32_t phase = 0;
  if (t_first < ntiles) {
    const int nv = (int)min((int64_t)128, a.E - t_first * 128);
    issue_gather(idx, nv, 1, stage0, &mbar[6]);
    issue_gather(idx, nv, 0, stage1, &mbar[7]);
  }
  const float4* uv4 = reinterpret_cast<const float4*>(a.uv);

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
    int nsrc = 0, ndst = -1, nbond = -1;
    if (have_next && gt < 128) {
      const int64_t e = tn * 128 + gt;
      if (e < a.E) nsrc = __ldcs(a.e_src + e), ndst = __ldcs(a.e_dst + e), nbond = __ldcs(a.e_bond + e);
    }
    if (have_next) {  // streamed operands of my next tile: DRAM -> L2 while this tile computes
      l2_prefetch(a.uv + tn * (128 * 128), 128 * 128 * 4, gt, 256);
      l2_prefetch(a.be + tn * (128 * 12), 128 * 12 * 4, gt, 256);
      l2_prefetch(a.dbe + tn * (128 * 12), 128 * 12 * 4, gt, 256);
    }
    float bek[9], dbek[9];
    bek[0] = benext[0].x, bek[1] = benext[0].y, bek[2] = benext[0].z, bek[3] = benext[0].w, bek[4] = benext[1].x;
    bek[5] = benext[1].y, bek[6] = benext[1].z, bek[7] = benext[1].w, bek[8] = benext[2].x;
    dbek[0] = dbenext[0].x, dbek[1] = dbenext[0].y, dbek[2] = dbenext[0].z, dbek[3] = dbenext[0].w, dbek[4] = dbenext[1].x;
    dbek[5] = dbenext[1].y, dbek[6] = dbenext[1].z, dbek[7] = dbenext[1].w, dbek[8] = dbenext[2].x;
    if (r >= nvalid) {
#pragma unroll
      for (int k = 0; k < 9; k++) bek[k] = dbek[k] = 0.f;
    }
    if (half == 0) {
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
    if (gt == 0) {  // GEMM1: D[128 x 128] = be . M^T
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
    const float* gmrow = a.gagg + (size_t)(valid ? dst : 0) * D + c0;
    const size_t uvo = tl4<32>(t, r, c0);  // float4 index of (my row, column c0) of the u block; +16*128 per 64 columns
    float gdpart = 0.f;
    float4 cv0[4], cv1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) cv0[i] = *reinterpret_cast<const float4*>(Crow + 64 + 4 * i);

    // ---------------- first-layer derivatives of the gate branch -> stage0 (in place) ----------------
    mbar_wait_(&mbar[6], phase);
    mbar_wait_(&mbar[0], phase);
    tc_fence_after();
    ac3_ds16(tlane + COL_D + 64 + c0, stage0 + r * PITCH + c0, cv0, viaQ ? Qrow + 64 : nullptr);
#pragma unroll
    for (int i = 0; i < 4; i++) cv1[i] = *reinterpret_cast<const float4*>(Crow + 64 + 16 + 4 * i);  // same line as cv0: L1
    ac3_ds16(tlane + COL_D + 64 + c0 + 16, stage0 + r * PITCH + c0 + 16, cv1, viaQ ? Qrow + 64 + 16 : nullptr);
    // ---------------- second layers: one pass over the saved u | v; gv -> H now, gu parked in registers ----------------
    float gu[32];
    {
      uint32_t hi[16], lo[16];
      ac3_second16(uv4, uvo, gmrow, wabW, c0, bek, dbek, gdpart, gu, hi, lo);
      tmem_st16(tlane + COL_H + c0, hi);
      tmem_st16(tlane + COL_H + 64 + c0, lo);
      ac3_second16(uv4, uvo + 4 * 128, gmrow + 16, wabW, c0 + 16, bek, dbek, gdpart, gu + 16, hi, lo);
      tmem_st16(tlane + COL_H + c0 + 16, hi);
      tmem_st16(tlane + COL_H + 64 + c0 + 16, lo);
    }
    if (have_next && gt < 128) {
      ibn[gt] = nsrc, ibn[128 + gt] = ndst, ibn[256 + gt] = nbond;
      // C[dst] and gagg[dst] rows of my next tile into L1 (one request per destination run)
      const int pd = __shfl_up_sync(0xffffffffu, ndst, 1);
      if (ndst >= 0 && (lane == 0 || pd != ndst)) {
        const char* cp = reinterpret_cast<const char*>(a.Cproj + (size_t)ndst * D2);
        const char* gp = reinterpret_cast<const char*>(a.gagg + (size_t)ndst * D);
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(cp + 128 * i));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(gp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(gp + 128));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) cv0[i] = *reinterpret_cast<const float4*>(Crow + 4 * i);  // layer branch, ahead of the barrier
    tc_wait_st();
    tc_fence_before();
    gbar(g);
    if (gt == 0) {  // GEMM3 (gates): D[:, 64..127] = gv . W2G   (B operand = W2G^T, canonical)
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 64u : 0u;
        const uint32_t bsel = w2t_addr + (uint32_t)(2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          umma_ts(tbase + COL_D + 64, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64,
                  acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[1]);
    }

    // ---------------- first-layer derivatives of the layer branch -> stage1, while the tensor core multiplies ----------------
    mbar_wait_(&mbar[7], phase);
    ac3_ds16(tlane + COL_D + c0, stage1 + r * PITCH + c0, cv0, viaQ ? Qrow : nullptr);
#pragma unroll
    for (int i = 0; i < 4; i++) cv1[i] = *reinterpret_cast<const float4*>(Crow + 16 + 4 * i);
    ac3_ds16(tlane + COL_D + c0 + 16, stage1 + r * PITCH + c0 + 16, cv1, viaQ ? Qrow + 16 : nullptr);
    mbar_wait_(&mbar[1], phase);  // GEMM3 (gates) done: H is free, ghid of the gate branch is in D[:, 64..127]
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t h = tf32_hi_bits(gu[ch * 16 + i]);
        hi[i] = h;
        lo[i] = __float_as_uint(gu[ch * 16 + i] - __uint_as_float(h));
      }
      tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
      tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
    }
    tc_wait_st();
    tc_fence_before();
    gbar(g);
    if (gt == 0) {  // GEMM3 (layers): D[:, 0..63] = gu . W2L
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 64u : 0u;
        const uint32_t bsel = w2t_addr + (uint32_t)(term == 2 ? 1 : 0) * 4096u * 4u;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          umma_ts(tbase + COL_D, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64, acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[2]);
    }

    // ---------------- post: gpre = silu'(pre) * ghid in place, scatter, dE/dd; gate branch first ----------------
    // dE/dbe . dbe/dd = sum_j gpre[j] (M dbe)[j]: the vector M dbe of every edge is one more K = 16 tensor-core product
    // (dbe . M^T, like GEMM1) instead of a 9-term dot per column on the FMA pipe (27 % of this kernel's instructions in
    // its first build, profiles/r02e)
    auto scatter = [&](const float* st, int br) {
      if (!need_gx) return;
      const int c4 = (gt & 15) * 4, part = gt >> 4;
      const int col = br * 64 + c4;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur = -1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = part * 8 + i;
        const int k = s_dst[row];
        if (k != cur) {
          if (cur >= 0) red_add_v4(&a.gC[(size_t)cur * D2 + col], s4.x, s4.y, s4.z, s4.w);
          cur = k;
          s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (k >= 0) {
          const float4 m = *reinterpret_cast<const float4*>(st + row * PITCH + c4);
          s4.x += m.x, s4.y += m.y, s4.z += m.z, s4.w += m.w;
          red_add_v4(&a.gA[(size_t)s_src[row] * D2 + col], m.x, m.y, m.z, m.w);
          const int bnd = s_bond[row];
          if (useQ && bnd >= 0) *reinterpret_cast<float4*>(&a.gQ[(size_t)bnd * D2 + col]) = m;
        }
      }
      if (cur >= 0) red_add_v4(&a.gC[(size_t)cur * D2 + col], s4.x, s4.y, s4.z, s4.w);
    };
    float gpG[32];  // gate-branch gpre of my columns: kept for the dE/dd dot after its stage has been refilled
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t gh[16];
      tmem_ld16(tlane + COL_D + 64 + c0 + ch * 16, gh);
      float4 ds4[4];
#pragma unroll
      for (int i = 0; i < 4; i++) ds4[i] = *reinterpret_cast<const float4*>(stage0 + r * PITCH + c0 + ch * 16 + 4 * i);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float* gp = gpG + ch * 16 + 4 * i;
        gp[0] = ds4[i].x * __uint_as_float(gh[4 * i]), gp[1] = ds4[i].y * __uint_as_float(gh[4 * i + 1]);
        gp[2] = ds4[i].z * __uint_as_float(gh[4 * i + 2]), gp[3] = ds4[i].w * __uint_as_float(gh[4 * i + 3]);
        *reinterpret_cast<float4*>(stage0 + r * PITCH + c0 + ch * 16 + 4 * i) = make_float4(gp[0], gp[1], gp[2], gp[3]);
      }
    }
    tc_fence_before();
    gbar(g);
    scatter(stage0, 1);
    gbar(g);  // stage0 drained: refill it for my next tile
    if (have_next) issue_gather(ibn, nvnext, 1, stage0, &mbar[6]);

    mbar_wait_(&mbar[2], phase);  // GEMM3 (layers): ghid in D[:, 0..63]; H is free
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t gh[16];
      tmem_ld16(tlane + COL_D + c0 + ch * 16, gh);
      float4 ds4[4];
#pragma unroll
      for (int i = 0; i < 4; i++) ds4[i] = *reinterpret_cast<const float4*>(stage1 + r * PITCH + c0 + ch * 16 + 4 * i);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; i++)
        *reinterpret_cast<float4*>(stage1 + r * PITCH + c0 + ch * 16 + 4 * i) =
            make_float4(ds4[i].x * __uint_as_float(gh[4 * i]), ds4[i].y * __uint_as_float(gh[4 * i + 1]),
                        ds4[i].z * __uint_as_float(gh[4 * i + 2]), ds4[i].w * __uint_as_float(gh[4 * i + 3]));
    }
    if (half == 0) {  // dbe -> TMEM operand for the M.dbe product
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const float x = k < 9 ? dbek[k] : 0.f;
        const uint32_t h = tf32_hi_bits(x);
        hi[k] = h;
        lo[k] = __float_as_uint(x - __uint_as_float(h));
      }
      tmem_st16(tlane + COL_H, hi);
      tmem_st16(tlane + COL_H + 16, lo);
    }
    tc_wait_st();
    tc_fence_before();
    gbar(g);  // every ghid has been read: D is free
    if (gt == 0) {  // D[128 x 128] = dbe[128 x 16] . M^T
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
      umma_commit(&mbar[3]);
    }
    scatter(stage1, 0);
    if (have_next) load_be(tn, benext, dbenext);
    mbar_wait_(&mbar[3], phase);
    tc_fence_after();
    {  // tcgen05.ld is warp-collective (.sync.aligned): every lane runs the loads, bond rows (e through Q, not through
       // the radial term) drop the result
      float dotp = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t mg[16], ml[16];
        tmem_ld16(tlane + COL_D + 64 + c0 + ch * 16, mg);
        tmem_ld16(tlane + COL_D + c0 + ch * 16, ml);
        float4 gl[4];
#pragma unroll
        for (int i = 0; i < 4; i++) gl[i] = *reinterpret_cast<const float4*>(stage1 + r * PITCH + c0 + ch * 16 + 4 * i);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) dotp = fmaf(gpG[ch * 16 + i], __uint_as_float(mg[i]), dotp);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          dotp = fmaf(gl[i].x, __uint_as_float(ml[4 * i]), dotp);
          dotp = fmaf(gl[i].y, __uint_as_float(ml[4 * i + 1]), dotp);
          dotp = fmaf(gl[i].z, __uint_as_float(ml[4 * i + 2]), dotp);
          dotp = fmaf(gl[i].w, __uint_as_float(ml[4 * i + 3]), dotp);
        }
      }
      if (!viaQ) gdpart += dotp;
    }
    gdb[gt] = gdpart;
    tc_fence_before();
    gbar(g);  // stage1 drained by the scatter and the dots; D read
    if (gt < nvalid) a.gd[e0 + gt] += gdb[gt] + gdb[128 + gt];
    if (have_next) issue_gather(ibn, nvnext, 0, stage1, &mbar[7]);
    phase ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tptr), "r"(512u));
}

void launch_atomconv_bwd_v3(cudaStream_t st, const AtomConvArgs& a, const AtomConvTcW& w, int num_sms) {
  if (a.E <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_)
    B2M_CK(cudaFuncSetAttribute(k_atomconv_bwd_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ac3BwdSmem::bytes));
  const int64_t ntiles = (a.E + 127) / 128;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, (int64_t)num_sms);
  k_atomconv_bwd_v3<<<grid, 512, Ac3BwdSmem::bytes, st>>>(a, w);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

void launch_atomconv_fwd_v3(cudaStream_t st, const AtomConvArgs& a, const AtomConvTcW& w, int num_sms) {
  if (a.E <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_atomconv_fwd_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ac3Smem::bytes));
  }
  const int64_t ntiles = (a.E + 127) / 128;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, (int64_t)num_sms);
  AtomConvTcW wl = w;
  static const int pfl1 = [] {
    const char* v = getenv("B2M_AC3_L1PF");
    return v ? atoi(v) : 1;  // measured 0.861 -> 0.855 ms per launch (profiles/r02g)
  }();
  wl.l2pf = pfl1;
  k_atomconv_fwd_v3<<<grid, 512, Ac3Smem::bytes, st>>>(a, wl);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

}  // namespace b2m
