// kernels_tc.cu -- tcgen05 / TMEM versions of the fused tile kernels (sm_100a): the line-graph kernels (bond conv, angle
// update: default path), the row GEMMs (default path) and the FIRST generation of the atom-conv kernels, which the third
// generation in kernels_ac3.cu replaced as the default in round 2 (kept selectable with B2M_ATOMCONV=1 for A/B checks).
//
// Design of the first-generation atom conv forward:
//   * persistent CTAs, 256 threads, 2 CTAs per SM, 256 TMEM columns each: [H operand 128 | D accum 128]
//   * thread <-> row mapping fixed by TMEM: warp w owns lanes 32*(w%4)..+31, half = w/4 picks 32 of 64 columns
//   * A operands live in TMEM (tcgen05.st from registers, hi/lo tf32 split) -- no shared-memory A tiles
//   * B operands (weights) staged ONCE per CTA in shared memory, already in canonical UMMA layout
//   * first-layer radial term  M.be(d)  : one K=16 MMA chain (N=128)
//     second layers  hid(64) x W2(64x64): two K=64 chains (N=64), gate product in the epilogue
//   * segmented sum over destination rows through a shared-memory message tile, RED to HBM on run ends
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b2m {

// B2M_L2_PREFETCH: 0 = no hints, 1 (default) = atom-conv kernels only, 2 = line-graph kernels too.
// Measured (profiles/r01c_*): in the atom-conv kernels the hints cost ~0.1 GB of extra DRAM reads per launch and buy
// ~2 %; in the line-graph kernels (160 KB per tile, 47 MB in flight next to 1.2 GB of streamed writes) most prefetched
// lines are evicted before use -- DRAM reads of k_line_bwd_tc<1> went from 2.06 to 3.33 GB per launch -- so they are off.
static int l2pf_level() {
  static const int v = [] {
    const char* e = getenv("B2M_L2_PREFETCH");
    return e ? atoi(e) : 1;
  }();
  return v;
}


// be[e][0..8], dbe[e][0..8] (padded to 12) for every edge: the same radial basis feeds all atom-conv layers
__global__ void k_edge_basis(int64_t E, const float4* __restrict__ e_vec, RadialParams rp, float* __restrict__ be,
                             float* __restrict__ dbe) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * 3) return;
  const int64_t e = i / 3;
  const int part = (int)(i % 3);
  const float d = e_vec[e].w;
  float b[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < 4; j++) {
    const int k = part * 4 + j;
    if (k < 9) rbf_env_both(d, rp.freq[k], rp, b[j], db[j]);
  }
  const size_t o = tl4<3>(e >> 7, (int)(e & 127), part * 4);
  reinterpret_cast<float4*>(be)[o] = make_float4(b[0], b[1], b[2], b[3]);
  reinterpret_cast<float4*>(dbe)[o] = make_float4(db[0], db[1], db[2], db[3]);
}
void launch_edge_basis(cudaStream_t st, int64_t E, const float4* e_vec, RadialParams rp, float* be, float* dbe) {
  if (E <= 0) return;
  k_edge_basis<<<cdiv(E * 3, 256), 256, 0, st>>>(E, e_vec, rp, be, dbe);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

struct FwdTcSmem {
  // float offsets
  static constexpr int kBar = 0;               // 4 mbarriers + tmem ptr (64 floats reserved)
  static constexpr int kW2 = 64;                // 4 x 4096
  static constexpr int kM = kW2 + 4 * 4096;     // 2 x 2048
  static constexpr int kMsg = kM + 2 * 2048;    // [64][65] (half tile staging)
  static constexpr int kBe = kMsg + 64 * 65;    // [128][12]
  static constexpr int kWab = kBe + 128 * 12;   // 576
  static constexpr int kB2 = kWab + 576;        // 128
  static constexpr int kD = kB2 + 128;          // [128]
  static constexpr int kIdx = kD + 128;         // 3 x 128 ints
  static constexpr int kTotal = kIdx + 3 * 128;
  static constexpr size_t bytes = (size_t)kTotal * 4;
};

// NTHR = 256: two threads per row (32 columns of each branch per thread, <=128 registers);
// NTHR = 512: four threads per row (16 columns, <=64 registers) -> twice the resident warps for the same TMEM/smem.
// PF = true: the gathers of A[src] / C[dst] for the next 16-column block are issued one block ahead (the first
// block before the wait for GEMM1), so their L2 latency overlaps the tensor-core wait and the activation math of
// the current block instead of following it (the gather wait was 16 % of the samples: profiles/r01_stalls_*.txt).
template <int NTHR, bool PF>
__global__ void __launch_bounds__(NTHR, 2) k_atomconv_fwd_tc(const AtomConvArgs a, const AtomConvTcW w) {
  constexpr int NP = NTHR / 128;   // threads per row
  constexpr int CPT = 64 / NP;     // columns of each branch per thread
  constexpr int NCH = CPT / 16;    // 16-column chunks per thread
  extern __shared__ __align__(1024) float smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + FwdTcSmem::kBar);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + FwdTcSmem::kBar + 16);
  float* W2s = smem + FwdTcSmem::kW2;
  float* Ms = smem + FwdTcSmem::kM;
  float* msg = smem + FwdTcSmem::kMsg;
  float* wabW = smem + FwdTcSmem::kWab;
  float* b2s = smem + FwdTcSmem::kB2;
  int* s_src = reinterpret_cast<int*>(smem + FwdTcSmem::kIdx);
  int* s_dst = s_src + 128;
  int* s_bond = s_dst + 128;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;  // half = column part 0..NP-1
  const int r = q * 32 + lane;   // my row (TMEM lane)
  const int c0 = half * CPT;     // my columns inside each 64-wide branch
  const bool useQ = a.Qproj != nullptr;

  // ---- one-time setup: TMEM, barriers, weights ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init_(&mbar[0], 1);
    mbar_init_(&mbar[1], 1);
    mbar_init_(&mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * 1024; i += NTHR) reinterpret_cast<float4*>(W2s)[i] = reinterpret_cast<const float4*>(w.W2can)[i];
  for (int i = tid; i < 2 * 512; i += NTHR) reinterpret_cast<float4*>(Ms)[i] = reinterpret_cast<const float4*>(w.Mcan)[i];
  for (int i = tid; i < 576; i += NTHR) wabW[(i % 9) * 64 + i / 9] = a.Wabw[i];  // k-major [9][64]
  if (tid < 128) b2s[tid] = a.b2[tid];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);  // my lane quarter
  constexpr uint32_t COL_H = 0, COL_D = 128;
  const uint32_t w2_addr = s_u32(W2s), m_addr = s_u32(Ms);
  uint32_t phase = 0;

  const int64_t ntiles = (a.E + 127) / 128;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t e0 = t * 128;
    const int nvalid = (int)min((int64_t)128, a.E - e0);
    if (tid < 128) {
      int src = 0, dst = -1, bond = -1;
      if (tid < nvalid) {
        const int64_t e = e0 + tid;
        src = a.e_src[e];
        dst = a.e_dst[e];
        bond = a.e_bond[e];
      }
      s_src[tid] = src;
      s_dst[tid] = dst;
      s_bond[tid] = bond;
    }
    if (w.l2pf) {
      const int64_t tn = t + gridDim.x;
      if (tn < ntiles) {
        l2_prefetch(a.be + tn * (128 * 12), 128 * 12 * 4, tid, NTHR);
        if (tn * 128 + 128 <= a.E && tid >= 128 && tid < 140) {
          const int which = (tid - 128) >> 2, line = (tid - 128) & 3;
          const int* base = which == 0 ? a.e_src : which == 1 ? a.e_dst : a.e_bond;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + tn * 128 + line * 32));
        }
      }
    }
    __syncthreads();
    float bek[9];  // radial basis of my row (precomputed once per step: launch_edge_basis)
    {
      const float4* bp = reinterpret_cast<const float4*>(a.be);
      const float4 b0 = bp[tl4<3>(t, r, 0)], b1 = bp[tl4<3>(t, r, 4)], b2 = bp[tl4<3>(t, r, 8)];
      bek[0] = b0.x, bek[1] = b0.y, bek[2] = b0.z, bek[3] = b0.w, bek[4] = b1.x, bek[5] = b1.y, bek[6] = b1.z,
      bek[7] = b1.w, bek[8] = b2.x;
      if (r >= nvalid) {
#pragma unroll
        for (int k = 0; k < 9; k++) bek[k] = 0.f;
      }
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
    __syncthreads();
    if (tid == 0) {  // GEMM1: D[128 x 128] = be[128 x 16] . M^T    (hi*hi + lo*hi + hi*lo)
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
    const int src = s_src[r], dst = s_dst[r], bond = s_bond[r];
    const bool valid = r < nvalid;
    const bool viaQ = useQ && bond >= 0;
    const float* Arow = a.Aproj + (size_t)src * D2;
    const float* Crow = a.Cproj + (size_t)(valid ? dst : 0) * D2;
    const float* Qrow = viaQ ? a.Qproj + (size_t)bond * D2 : nullptr;
    // ---- first layer epilogue + second layer, branch by branch (0: "layers", 1: "gates") ----
    float4 pa[4], pc[4];  // PF: gathered rows of the next 16-column block
    if (PF) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        pa[i] = *reinterpret_cast<const float4*>(Arow + c0 + i * 4);
        pc[i] = *reinterpret_cast<const float4*>(Crow + c0 + i * 4);
      }
    }
#pragma unroll 1
    for (int br = 0; br < 2; br++) {
      mbar_wait_(&mbar[br], phase);
      tc_fence_after();
      const int cb = br * 64 + c0;  // column in the 128-wide first layer
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        uint32_t v[16], hi[16], lo[16];
        tmem_ld16(tlane + COL_D + cb + ch * 16, v);
        tc_wait_ld();
        float av[16], cv[16];
        if (PF) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            av[4 * i] = pa[i].x, av[4 * i + 1] = pa[i].y, av[4 * i + 2] = pa[i].z, av[4 * i + 3] = pa[i].w;
            cv[4 * i] = pc[i].x, cv[4 * i + 1] = pc[i].y, cv[4 * i + 2] = pc[i].z, cv[4 * i + 3] = pc[i].w;
          }
          // next block: (br, ch + 1), or the first block of the other branch
          const int nb = ch + 1 < NCH ? cb + (ch + 1) * 16 : 64 + c0;
          if (ch + 1 < NCH || br == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              pa[i] = *reinterpret_cast<const float4*>(Arow + nb + i * 4);
              pc[i] = *reinterpret_cast<const float4*>(Crow + nb + i * 4);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 x = *reinterpret_cast<const float4*>(Arow + cb + ch * 16 + i * 4);
            const float4 y = *reinterpret_cast<const float4*>(Crow + cb + ch * 16 + i * 4);
            av[4 * i] = x.x, av[4 * i + 1] = x.y, av[4 * i + 2] = x.z, av[4 * i + 3] = x.w;
            cv[4 * i] = y.x, cv[4 * i + 1] = y.y, cv[4 * i + 2] = y.z, cv[4 * i + 3] = y.w;
          }
        }
        if (viaQ) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 x = *reinterpret_cast<const float4*>(Qrow + cb + ch * 16 + i * 4);
            v[4 * i] = __float_as_uint(x.x), v[4 * i + 1] = __float_as_uint(x.y);
            v[4 * i + 2] = __float_as_uint(x.z), v[4 * i + 3] = __float_as_uint(x.w);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float p = __uint_as_float(v[i]) + av[i] + cv[i];
          const float hval = valid ? silu_(p) : 0.f;
          const uint32_t h = tf32_hi_bits(hval);
          hi[i] = h;
          lo[i] = __float_as_uint(hval - __uint_as_float(h));
        }
        tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
        tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {  // GEMM2 for this branch: D[:, br*64 .. +64] = hid[128 x 64] . W2^T
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 3; term++) {
          const uint32_t acol = term == 1 ? 64u : 0u;
          const uint32_t bsel = w2_addr + (uint32_t)(br * 2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            umma_ts(tbase + COL_D + br * 64, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u),
                    kIdescN64, acc);
            acc = 1;
          }
        }
        umma_commit(&mbar[br + 1]);
      }
    }
    mbar_wait_(&mbar[2], phase);
    tc_fence_after();
    phase ^= 1;
    // ---- gate product, shared weights, segmented sum over dst (two half-tiles through smem) ----
    float mv[CPT];
    {
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        uint32_t u[16], g[16];
        tmem_ld16(tlane + COL_D + c0 + ch * 16, u);
        tmem_ld16(tlane + COL_D + 64 + c0 + ch * 16, g);
        float wab[16];  // shared bond weights w_ab = be . Wabw^T for my 16 columns (packed FMAs while the TMEM loads fly)
        radial_dot16(wabW, c0 + ch * 16, bek, wab);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int c = c0 + ch * 16 + i;
          const float uu = __uint_as_float(u[i]) + b2s[c], vv = __uint_as_float(g[i]) + b2s[64 + c];
          u[i] = __float_as_uint(uu);
          g[i] = __float_as_uint(vv);
          const float oL = silu_(uu);
          const float oG = sigm_(vv);
          mv[ch * 16 + i] = valid ? oL * oG * wab[i] : 0.f;
        }
        if (a.uv_save != nullptr && valid) {
          float4* puv = reinterpret_cast<float4*>(a.uv_save);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            puv[tl4<32>(t, r, c0 + ch * 16 + 4 * i)] = make_float4(__uint_as_float(u[4 * i]), __uint_as_float(u[4 * i + 1]),
                                                                  __uint_as_float(u[4 * i + 2]), __uint_as_float(u[4 * i + 3]));
            puv[tl4<32>(t, r, 64 + c0 + ch * 16 + 4 * i)] = make_float4(__uint_as_float(g[4 * i]), __uint_as_float(g[4 * i + 1]),
                                                                       __uint_as_float(g[4 * i + 2]), __uint_as_float(g[4 * i + 3]));
          }
        }
      }
    }
    tc_fence_before();
#pragma unroll 1
    for (int hp = 0; hp < 2; hp++) {
      if ((q >> 1) == hp) {
        const int rr = r - hp * 64;
#pragma unroll
        for (int i = 0; i < CPT; i++) msg[rr * 65 + c0 + i] = mv[i];
      }
      __syncthreads();
      {
        constexpr int RPP = 64 / (NTHR / 64);  // rows per reducing thread
        const int c = tid & 63, part = tid >> 6;
        float sum = 0.f;
        int cur = -1;
        const int rbeg = part * RPP;
        for (int rr = rbeg; rr < rbeg + RPP; rr++) {
          const int k = s_dst[hp * 64 + rr];
          if (k != cur) {
            if (cur >= 0) atomicAdd(&a.agg[(size_t)cur * D + c], sum);
            cur = k;
            sum = 0.f;
          }
          if (k >= 0) sum += msg[rr * 65 + c];
        }
        if (cur >= 0) atomicAdd(&a.agg[(size_t)cur * D + c], sum);
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u));
}

// ============================================================================================
// atom conv backward on tcgen05
//   recomputes only the first-layer pre-activation (M.be on the tensor core + gathers); second-layer
//   pre-activations (u,v) come from the forward; ghid = [gu . W2L | gv . W2G] on the tensor core;
//   gpre = ghid * silu'(pre) is scattered: segmented sums -> gC[dst], RED -> gA[src], store -> gQ[bond],
//   and contracted with M and dbe/dd for dE/dd_e.
// ============================================================================================
struct BwdTcSmem {
  static constexpr int kBar = 0;
  static constexpr int kW2T = 64;                  // 4 x 4096
  static constexpr int kM = kW2T + 4 * 4096;       // 2 x 2048 (canonical)
  static constexpr int kStage = kM + 2 * 2048;     // [64][65]
  static constexpr int kMrow = kStage + 64 * 65;   // [128][12] row-major copy of M (k >= 9 zero)
  static constexpr int kWab = kMrow + 128 * 12;    // 576, k-major [9][64]
  static constexpr int kIdx = kWab + 576;          // 3 x 128 ints
  static constexpr int kTotal = kIdx + 3 * 128;
  static constexpr size_t bytes = (size_t)kTotal * 4;
};

template <int NTHR>
__global__ void __launch_bounds__(NTHR, 2) k_atomconv_bwd_tc(const AtomConvArgs a, const AtomConvTcW w) {
  constexpr int NP = NTHR / 128;   // threads per row
  constexpr int CPT = 64 / NP;     // columns of each branch per thread
  constexpr int NCH = CPT / 16;
  extern __shared__ __align__(1024) float smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + BwdTcSmem::kBar);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + BwdTcSmem::kBar + 16);
  float* W2Ts = smem + BwdTcSmem::kW2T;
  float* Ms = smem + BwdTcSmem::kM;
  float* stage = smem + BwdTcSmem::kStage;
  float* Mrow = smem + BwdTcSmem::kMrow;  // M [128][12] in plain row-major fp32 (dE/dbe contraction)
  float* wabW = smem + BwdTcSmem::kWab;
  int* s_src = reinterpret_cast<int*>(smem + BwdTcSmem::kIdx);
  int* s_dst = s_src + 128;
  int* s_bond = s_dst + 128;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;
  const int c0 = half * CPT;
  const bool useQ = a.Qproj != nullptr;
  const bool need_gx = a.gA != nullptr;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init_(&mbar[0], 1);
    mbar_init_(&mbar[1], 1);
    mbar_init_(&mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * 1024; i += NTHR) reinterpret_cast<float4*>(W2Ts)[i] = reinterpret_cast<const float4*>(w.W2Tcan)[i];
  for (int i = tid; i < 2 * 512; i += NTHR) reinterpret_cast<float4*>(Ms)[i] = reinterpret_cast<const float4*>(w.Mcan)[i];
  for (int i = tid; i < 576; i += NTHR) wabW[(i % 9) * 64 + i / 9] = a.Wabw[i];
  for (int i = tid; i < 128 * 12; i += NTHR) Mrow[i] = (i % 12) < 9 ? a.M[(i / 12) * 9 + i % 12] : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t COL_H = 0, COL_D = 128;
  const uint32_t w2t_addr = s_u32(W2Ts), m_addr = s_u32(Ms);
  uint32_t phase = 0;

  const int64_t ntiles = (a.E + 127) / 128;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t e0 = t * 128;
    const int nvalid = (int)min((int64_t)128, a.E - e0);
    if (tid < 128) {
      int src = 0, dst = -1, bond = -1;
      if (tid < nvalid) {
        const int64_t e = e0 + tid;
        src = a.e_src[e];
        dst = a.e_dst[e];
        bond = a.e_bond[e];
      }
      s_src[tid] = src;
      s_dst[tid] = dst;
      s_bond[tid] = bond;
    }
    if (w.l2pf) {
      const int64_t tn = t + gridDim.x;
      if (tn < ntiles) {
        l2_prefetch(a.uv + tn * (128 * 128), 128 * 128 * 4, tid, NTHR);
        l2_prefetch(a.be + tn * (128 * 12), 128 * 12 * 4, tid, NTHR);
        l2_prefetch(a.dbe + tn * (128 * 12), 128 * 12 * 4, tid, NTHR);
        if (tn * 128 + 128 <= a.E && tid >= 128 && tid < 140) {
          const int which = (tid - 128) >> 2, line = (tid - 128) & 3;
          const int* base = which == 0 ? a.e_src : which == 1 ? a.e_dst : a.e_bond;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + tn * 128 + line * 32));
        }
      }
    }
    __syncthreads();
    float bek[9], dbek[9];
    {
      const float4* bp = reinterpret_cast<const float4*>(a.be);
      const float4* dp4 = reinterpret_cast<const float4*>(a.dbe);
      const float4 b0 = bp[tl4<3>(t, r, 0)], b1 = bp[tl4<3>(t, r, 4)], b2 = bp[tl4<3>(t, r, 8)];
      const float4 d0 = dp4[tl4<3>(t, r, 0)], d1 = dp4[tl4<3>(t, r, 4)], d2 = dp4[tl4<3>(t, r, 8)];
      bek[0] = b0.x, bek[1] = b0.y, bek[2] = b0.z, bek[3] = b0.w, bek[4] = b1.x, bek[5] = b1.y, bek[6] = b1.z,
      bek[7] = b1.w, bek[8] = b2.x;
      dbek[0] = d0.x, dbek[1] = d0.y, dbek[2] = d0.z, dbek[3] = d0.w, dbek[4] = d1.x, dbek[5] = d1.y, dbek[6] = d1.z,
      dbek[7] = d1.w, dbek[8] = d2.x;
      if (r >= nvalid) {
#pragma unroll
        for (int k = 0; k < 9; k++) bek[k] = dbek[k] = 0.f;
      }
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
    __syncthreads();
    if (tid == 0) {
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
    const int src = s_src[r], dst = s_dst[r], bond = s_bond[r];
    const bool valid = r < nvalid;
    const bool viaQ = useQ && bond >= 0;
    const float* Arow = a.Aproj + (size_t)src * D2;
    const float* Crow = a.Cproj + (size_t)(valid ? dst : 0) * D2;
    const float* Qrow = viaQ ? a.Qproj + (size_t)bond * D2 : nullptr;
    const float4* uv4 = reinterpret_cast<const float4*>(a.uv);
    const float* gmrow = a.gagg + (size_t)(valid ? dst : 0) * D;
    float gdpart = 0.f;   // dE/dd_e contribution of this thread's columns
    float gbeM[12];  // k >= 9 stay zero (padding of the packed FMAs)
#pragma unroll
    for (int k = 0; k < 12; k++) gbeM[k] = 0.f;

#pragma unroll 1
    for (int br = 0; br < 2; br++) {
      mbar_wait_(&mbar[br], phase);
      tc_fence_after();
      const int cb = br * 64 + c0;
      float ds[CPT];  // silu'(pre) for my columns of this branch
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        uint32_t v[16], hi[16], lo[16];
        tmem_ld16(tlane + COL_D + cb + ch * 16, v);
        tc_wait_ld();
        // (issuing these gathers one block ahead, as the forward does, was measured slower here: the kernel sits at
        //  the 128-register limit and the prefetched rows spill)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float4 x = *reinterpret_cast<const float4*>(Arow + cb + ch * 16 + i * 4);
          const float4 y = *reinterpret_cast<const float4*>(Crow + cb + ch * 16 + i * 4);
          float4 t4 = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                  __uint_as_float(v[4 * i + 3]));
          if (viaQ) t4 = *reinterpret_cast<const float4*>(Qrow + cb + ch * 16 + i * 4);
          const float p0 = t4.x + x.x + y.x, p1 = t4.y + x.y + y.y, p2 = t4.z + x.z + y.z, p3 = t4.w + x.w + y.w;
          float sg;
          sg = sigm_(p0), ds[ch * 16 + 4 * i] = sg * (1.f + p0 * (1.f - sg));
          sg = sigm_(p1), ds[ch * 16 + 4 * i + 1] = sg * (1.f + p1 * (1.f - sg));
          sg = sigm_(p2), ds[ch * 16 + 4 * i + 2] = sg * (1.f + p2 * (1.f - sg));
          sg = sigm_(p3), ds[ch * 16 + 4 * i + 3] = sg * (1.f + p3 * (1.f - sg));
        }
        // gradient w.r.t. this branch's second-layer pre-activation (gu for br=0, gv for br=1)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int c = c0 + ch * 16 + i * 4;
          const float4 u4 = uv4[tl4<32>(t, r, c)];
          const float4 v4 = uv4[tl4<32>(t, r, 64 + c)];
          const float4 g4 = *reinterpret_cast<const float4*>(gmrow + c);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
          float wab4[4], wabp4[4];  // w_ab and d(w_ab)/dd for these 4 columns (packed FMAs)
          radial_dot4(wabW, c, bek, wab4);
          if (br == 0) radial_dot4(wabW, c, dbek, wabp4);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float wab = wab4[j];
            const float su = sigm_(uu[j]), oG = sigm_(vv[j]);
            const float oL = uu[j] * su;
            float g;
            if (br == 0) {
              g = gg[j] * oG * wab * (su * (1.f + uu[j] * (1.f - su)));
              // d/d w_ab -> d/d d_e through d(be)/dd (only once per column: do it in the br == 0 pass)
              gdpart = fmaf(gg[j] * oL * oG, wabp4[j], gdpart);
            } else {
              g = gg[j] * oL * wab * oG * (1.f - oG);
            }
            if (!valid) g = 0.f;
            const uint32_t h = tf32_hi_bits(g);
            hi[4 * i + j] = h;
            lo[4 * i + j] = __float_as_uint(g - __uint_as_float(h));
          }
        }
        tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
        tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {  // ghid[:, br*64 .. +64] = g[128 x 64] . W2 (B operand = W2^T, canonical)
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 3; term++) {
          const uint32_t acol = term == 1 ? 64u : 0u;
          const uint32_t bsel = w2t_addr + (uint32_t)(br * 2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            umma_ts(tbase + COL_D + br * 64, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u),
                    kIdescN64, acc);
            acc = 1;
          }
        }
        umma_commit(&mbar[br + 1]);
      }
      // the next branch's first-layer columns live in the other half of D: wait for this GEMM before
      // touching H again; meanwhile nothing else to do for this tile.
      mbar_wait_(&mbar[br + 1], phase);
      tc_fence_after();
      // gpre = ghid * silu'(pre)
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        uint32_t v[16];
        tmem_ld16(tlane + COL_D + br * 64 + c0 + ch * 16, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) ds[ch * 16 + i] *= __uint_as_float(v[i]);
      }
      // dE/d be through the radial first-layer term (not for bond rows fed by Q)
      if (!viaQ) {
#pragma unroll
        for (int i = 0; i < CPT; i++) {
          const float4* Mj = reinterpret_cast<const float4*>(Mrow + (cb + i) * 12);  // warp-uniform: smem broadcast
#pragma unroll
          for (int k4 = 0; k4 < 3; k4++) {
            const float4 m = Mj[k4];
            ffma2(gbeM[4 * k4], gbeM[4 * k4 + 1], ds[i], m.x, m.y);
            ffma2(gbeM[4 * k4 + 2], gbeM[4 * k4 + 3], ds[i], m.z, m.w);
          }
        }
      }
      if (need_gx) {
#pragma unroll 1
        for (int hp = 0; hp < 2; hp++) {
          if ((q >> 1) == hp) {
            const int rr = r - hp * 64;
#pragma unroll
            for (int i = 0; i < CPT; i++) stage[rr * 65 + c0 + i] = ds[i];
          }
          __syncthreads();
          {
            // 4 columns per thread so that every L2 reduction is a 16-byte vector RED
            constexpr int RPP = 64 / (NTHR / 16);  // rows per reducing thread
            const int c4 = tid & 15, part = tid >> 4;
            const int col = br * 64 + c4 * 4;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int cur = -1;
            const int rbeg = part * RPP;
            for (int rr = rbeg; rr < rbeg + RPP; rr++) {
              const int row = hp * 64 + rr;
              const int k = s_dst[row];
              const float v0 = stage[rr * 65 + c4 * 4], v1 = stage[rr * 65 + c4 * 4 + 1], v2 = stage[rr * 65 + c4 * 4 + 2],
                          v3 = stage[rr * 65 + c4 * 4 + 3];
              if (k != cur) {
                if (cur >= 0) red_add_v4(&a.gC[(size_t)cur * D2 + col], s0, s1, s2, s3);
                cur = k;
                s0 = s1 = s2 = s3 = 0.f;
              }
              if (k >= 0) {
                s0 += v0, s1 += v1, s2 += v2, s3 += v3;
                red_add_v4(&a.gA[(size_t)s_src[row] * D2 + col], v0, v1, v2, v3);
                const int bnd = s_bond[row];
                if (useQ && bnd >= 0) *reinterpret_cast<float4*>(&a.gQ[(size_t)bnd * D2 + col]) = make_float4(v0, v1, v2, v3);
              }
            }
            if (cur >= 0) red_add_v4(&a.gC[(size_t)cur * D2 + col], s0, s1, s2, s3);
          }
          __syncthreads();
        }
      }
      // mbar indices: br=0 used mbar[0] (GEMM1) and mbar[1]; br=1 must wait on "GEMM of branch 0 done",
      // which it already has (above); re-arm by pointing the loop's first wait at an already-passed barrier.
    }
    phase ^= 1;
    {
      float s = gdpart;
#pragma unroll
      for (int k = 0; k < 9; k++) s = fmaf(gbeM[k], dbek[k], s);
      stage[tid] = valid ? s : 0.f;
      tc_fence_before();
      __syncthreads();
      if (tid < nvalid) {
        float tot = 0.f;
#pragma unroll
        for (int pp = 0; pp < NP; pp++) tot += stage[pp * 128 + tid];
        a.gd[e0 + tid] += tot;
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u));
}

void launch_atomconv_bwd_tc(cudaStream_t st, const AtomConvArgs& a, const AtomConvTcW& w, int num_sms) {
  if (a.E <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_atomconv_bwd_tc<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdTcSmem::bytes));
    B2M_CK(cudaFuncSetAttribute(k_atomconv_bwd_tc<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdTcSmem::bytes));
  }
  const int64_t ntiles = (a.E + 127) / 128;
  const int grid = (int)std::min<int64_t>(ntiles, 2 * (int64_t)num_sms);
  static const int nthr = [] {
    const char* v = getenv("B2M_BWD_THREADS");
    return (v && atoi(v) == 512) ? 512 : 256;  // 512 (4 threads/row, 64 regs) measured slower: LSU-bound, not warp-bound
  }();
  AtomConvTcW wl = w;
  wl.l2pf = l2pf_level() >= 1;
  if (nthr == 512)
    k_atomconv_bwd_tc<512><<<grid, 512, BwdTcSmem::bytes, st>>>(a, wl);
  else
    k_atomconv_bwd_tc<256><<<grid, 256, BwdTcSmem::bytes, st>>>(a, wl);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

constexpr bool kFwdPrefetchDefault = true;  // measured: 1.087 vs 1.122 ms per launch; B2M_FWD_PREFETCH=0 switches it off
void launch_atomconv_fwd_tc(cudaStream_t st, const AtomConvArgs& a, const AtomConvTcW& w, int num_sms) {
  if (a.E <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_atomconv_fwd_tc<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdTcSmem::bytes));
    B2M_CK(cudaFuncSetAttribute(k_atomconv_fwd_tc<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdTcSmem::bytes));
    B2M_CK(cudaFuncSetAttribute(k_atomconv_fwd_tc<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdTcSmem::bytes));
  }
  const int64_t ntiles = (a.E + 127) / 128;
  const int grid = (int)std::min<int64_t>(ntiles, 2 * (int64_t)num_sms);
  static const int nthr = [] {
    const char* v = getenv("B2M_FWD_THREADS");
    return (v && atoi(v) == 512) ? 512 : 256;  // 512 (4 threads/row, 64 regs) measured slower: LSU-bound, not warp-bound
  }();
  static const bool pf = [] {
    const char* v = getenv("B2M_FWD_PREFETCH");
    return v ? atoi(v) != 0 : kFwdPrefetchDefault;
  }();
  AtomConvTcW wl = w;
  wl.l2pf = l2pf_level() >= 1;
  if (nthr == 512)
    k_atomconv_fwd_tc<512, false><<<grid, 512, FwdTcSmem::bytes, st>>>(a, wl);
  else if (pf)
    k_atomconv_fwd_tc<256, true><<<grid, 256, FwdTcSmem::bytes, st>>>(a, wl);
  else
    k_atomconv_fwd_tc<256, false><<<grid, 256, FwdTcSmem::bytes, st>>>(a, wl);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

}  // namespace b2m

// ============================================================================================
// line-graph kernels on tcgen05: bond conv (HIDDEN) and angle update (!HIDDEN)
//   one CTA per SM, 512 threads = two independent 256-thread row groups (each: one 128-angle tile in
//   flight, 256 TMEM columns, its own named barrier) sharing one copy of the weights in shared memory.
// ============================================================================================
namespace b2m {


struct LineTcSmem {
  static constexpr int kBar = 0;                    // 2 groups x 6 mbarriers, tmem ptr   (64 floats)
  static constexpr int kWA = 64;                    // 16384 floats: fwd: Wg can   | bwd: WgT can (4 x 4096)
  static constexpr int kWB = kWA + 16384;           // 16384 floats: fwd: W2 can   | bwd: W2T can
  static constexpr int kB2 = kWB + 16384;           // 128
  static constexpr int kGrp = kB2 + 128;            // per group: stage [64][65] + 2 x 3 x 128 ints (double-buffered)
  static constexpr int kGrpSize = 64 * 65 + 2 * 3 * 128;
  static constexpr int kTotal = kGrp + 2 * kGrpSize;
  static constexpr size_t bytes = (size_t)kTotal * 4;
};

template <bool HIDDEN>
__global__ void __launch_bounds__(512, 1) k_line_fwd_tc(const LineArgs a, const LineTcW w) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x;
  const int g = tid >> 8, gt = tid & 255;
  const int warp = gt >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane, c0 = half * 32;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + LineTcSmem::kBar) + g * 6;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + LineTcSmem::kBar + 48);
  float* WAs = smem + LineTcSmem::kWA;
  float* WBs = smem + LineTcSmem::kWB;
  float* b2s = smem + LineTcSmem::kB2;
  float* stage = smem + LineTcSmem::kGrp + g * LineTcSmem::kGrpSize;
  int* s_a = reinterpret_cast<int*>(stage + 64 * 65);
  int* s_b = s_a + 128;
  int* s_c = s_b + 128;

  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (gt == 0) {
    for (int i = 0; i < 6; i++) mbar_init_(&mbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (gt < 128) {  // indices of my first tile (later tiles: prefetched one tile ahead inside the loop)
    const int64_t e = (2 * (int64_t)blockIdx.x + g) * 128 + gt;
    const bool ok = e < a.A;
    s_a[gt] = ok ? a.a_in[e] : 0, s_b[gt] = ok ? a.a_out[e] : -1, s_c[gt] = ok ? a.a_ctr[e] : 0;
  }
  for (int i = tid; i < 4096; i += 512) reinterpret_cast<float4*>(WAs)[i] = reinterpret_cast<const float4*>(w.Wgcan)[i];
  if (HIDDEN)
    for (int i = tid; i < 4096; i += 512) reinterpret_cast<float4*>(WBs)[i] = reinterpret_cast<const float4*>(w.W2can)[i];
  if (tid < 128) b2s[tid] = (HIDDEN && a.b2) ? a.b2[tid] : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr + (uint32_t)g * 256u;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t COL_H = 0, COL_D = 128;
  const uint32_t wa_addr = s_u32(WAs), wb_addr = s_u32(WBs);
  uint32_t phase = 0;

  const int64_t ntiles = (a.A + 127) / 128;
  const int64_t tstride = 2 * (int64_t)gridDim.x;
  int* idxb = s_a;  // [2][a | b | c][128]
  int buf = 0;
  for (int64_t t = 2 * (int64_t)blockIdx.x + g; t < ntiles; t += tstride, buf ^= 1) {
    const int64_t r0 = t * 128;
    const int nvalid = (int)min((int64_t)128, a.A - r0);
    const int64_t tn = t + tstride;
    s_a = idxb + buf * 384, s_b = s_a + 128, s_c = s_a + 256;
    // indices of my next tile: requested now, parked in registers, published at the end of this tile (their DRAM latency
    // used to sit in front of every tile)
    int nia = 0, nib = -1, nic = 0;
    if (tn < ntiles && gt < 128 && tn * 128 + gt < a.A) nia = a.a_in[tn * 128 + gt], nib = a.a_out[tn * 128 + gt], nic = a.a_ctr[tn * 128 + gt];
    if (tn < ntiles) {
      // the next tile's angle rows are one contiguous 32 KB block (tile-interleaved layout): DRAM -> L2 now
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.ang + tn * (128 * 64)) + gt * 128));
    }
    const bool valid = r < nvalid;
    const int ia = s_a[r], ib = s_b[r], ic = s_c[r];
    const float4* ang4 = reinterpret_cast<const float4*>(a.ang);  // tile-interleaved: (tile, c/4, row) -> coalesced
    float angv[32];
    {  // my half of the angle row -> TMEM operand (hi | lo)
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float4 x = ang4[tl4<16>(t, r, c0 + ch * 16 + i * 4)];
          if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            angv[ch * 16 + 4 * i + j] = xv[j];
            const uint32_t h = tf32_hi_bits(xv[j]);
            hi[4 * i + j] = h;
            lo[4 * i + j] = __float_as_uint(xv[j] - __uint_as_float(h));
          }
        }
        tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
        tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
      }
    }
    tc_wait_st();
    tc_fence_before();
    gbar(g);
    if (gt == 0) {  // GEMM1: D[128 x 128] = ang[128 x 64] . Wg^T
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int term = 0; term < 3; term++) {
        const uint32_t acol = term == 1 ? 64u : 0u;
        const uint32_t bsel = wa_addr + (term == 2 ? 8192u * 4u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          umma_ts(tbase + COL_D, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 4096u, 2048u, 128u), kIdescN128, acc);
          acc = 1;
        }
      }
      umma_commit(&mbar[0]);
    }
    const float* Harow = a.Ha + (size_t)ia * D2;
    const float* Hbrow = a.Hb + (size_t)(valid ? ib : 0) * D2;
    const float* Xcrow = a.Xc + (size_t)ic * D2;
    if (HIDDEN) {
#pragma unroll 1
      for (int br = 0; br < 2; br++) {
        mbar_wait_(&mbar[br], phase);
        tc_fence_after();
        const int cb = br * 64 + c0;
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          uint32_t v[16], hi[16], lo[16];
          tmem_ld16(tlane + COL_D + cb + ch * 16, v);
          tc_wait_ld();
          float dsv[16];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 x = *reinterpret_cast<const float4*>(Harow + cb + ch * 16 + i * 4);
            const float4 y = *reinterpret_cast<const float4*>(Hbrow + cb + ch * 16 + i * 4);
            const float4 z = *reinterpret_cast<const float4*>(Xcrow + cb + ch * 16 + i * 4);
            const float p4[4] = {__uint_as_float(v[4 * i]) + x.x + y.x + z.x, __uint_as_float(v[4 * i + 1]) + x.y + y.y + z.y,
                                 __uint_as_float(v[4 * i + 2]) + x.z + y.z + z.z,
                                 __uint_as_float(v[4 * i + 3]) + x.w + y.w + z.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float sg = sigm_(p4[j]);
              const float hval = valid ? p4[j] * sg : 0.f;
              dsv[4 * i + j] = sg * (1.f + p4[j] * (1.f - sg));
              const uint32_t h = tf32_hi_bits(hval);
              hi[4 * i + j] = h;
              lo[4 * i + j] = __float_as_uint(hval - __uint_as_float(h));
            }
          }
          if (a.ds_save != nullptr && valid) {
            float4* pd = reinterpret_cast<float4*>(a.ds_save);
#pragma unroll
            for (int i = 0; i < 4; i++)
              __stcs(&pd[tl4<32>(t, r, cb + ch * 16 + 4 * i)], make_float4(dsv[4 * i], dsv[4 * i + 1], dsv[4 * i + 2], dsv[4 * i + 3]));
          }
          tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
          tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
        }
        tc_wait_st();
        tc_fence_before();
        gbar(g);
        if (gt == 0) {
          tc_fence_after();
          uint32_t acc = 0;
#pragma unroll
          for (int term = 0; term < 3; term++) {
            const uint32_t acol = term == 1 ? 64u : 0u;
            const uint32_t bsel = wb_addr + (uint32_t)(br * 2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
              umma_ts(tbase + COL_D + br * 64, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u),
                      kIdescN64, acc);
              acc = 1;
            }
          }
          umma_commit(&mbar[br + 1]);
        }
      }
      mbar_wait_(&mbar[2], phase);
      tc_fence_after();
    } else {
      mbar_wait_(&mbar[0], phase);
      tc_fence_after();
    }
    // ---- last-layer pre-activations (u | v): HIDDEN from GEMM2 (+bias), else first layer + gathers ----
    float mv[32];
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      uint32_t u[16], v[16];
      tmem_ld16(tlane + COL_D + c0 + ch * 16, u);
      tmem_ld16(tlane + COL_D + 64 + c0 + ch * 16, v);
      tc_wait_ld();
      float uf[16], vf[16];
      if (HIDDEN) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          uf[i] = __uint_as_float(u[i]) + b2s[c0 + ch * 16 + i];
          vf[i] = __uint_as_float(v[i]) + b2s[64 + c0 + ch * 16 + i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int cL = c0 + ch * 16 + i * 4, cG = 64 + cL;
          const float4 x = *reinterpret_cast<const float4*>(Harow + cL);
          const float4 y = *reinterpret_cast<const float4*>(Hbrow + cL);
          const float4 z = *reinterpret_cast<const float4*>(Xcrow + cL);
          const float4 x2 = *reinterpret_cast<const float4*>(Harow + cG);
          const float4 y2 = *reinterpret_cast<const float4*>(Hbrow + cG);
          const float4 z2 = *reinterpret_cast<const float4*>(Xcrow + cG);
          uf[4 * i] = __uint_as_float(u[4 * i]) + x.x + y.x + z.x;
          uf[4 * i + 1] = __uint_as_float(u[4 * i + 1]) + x.y + y.y + z.y;
          uf[4 * i + 2] = __uint_as_float(u[4 * i + 2]) + x.z + y.z + z.z;
          uf[4 * i + 3] = __uint_as_float(u[4 * i + 3]) + x.w + y.w + z.w;
          vf[4 * i] = __uint_as_float(v[4 * i]) + x2.x + y2.x + z2.x;
          vf[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + x2.y + y2.y + z2.y;
          vf[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + x2.z + y2.z + z2.z;
          vf[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + x2.w + y2.w + z2.w;
        }
      }
      if (a.uv_save != nullptr && valid) {
        float4* puv = reinterpret_cast<float4*>(a.uv_save);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          __stcs(&puv[tl4<32>(t, r, c0 + ch * 16 + 4 * i)], make_float4(uf[4 * i], uf[4 * i + 1], uf[4 * i + 2], uf[4 * i + 3]));
          __stcs(&puv[tl4<32>(t, r, 64 + c0 + ch * 16 + 4 * i)], make_float4(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3]));
        }
      }
#pragma unroll
      for (int i = 0; i < 16; i++) mv[ch * 16 + i] = valid ? silu_(uf[i]) * sigm_(vf[i]) : 0.f;
    }
    phase ^= 1;
    tc_fence_before();
    if (HIDDEN) {
#pragma unroll 1
      for (int hp = 0; hp < 2; hp++) {
        if ((q >> 1) == hp) {
          const int rr = r - hp * 64;
#pragma unroll
          for (int i = 0; i < 32; i++) stage[rr * 65 + c0 + i] = mv[i];
        }
        gbar(g);
        {
          const int c = gt & 63, part = gt >> 6;
          float sum = 0.f;
          int cur = -1;
          const int rbeg = part * 16;
          for (int rr = rbeg; rr < rbeg + 16; rr++) {
            const int k = s_b[hp * 64 + rr];
            if (k != cur) {
              if (cur >= 0) atomicAdd(&a.aggB[(size_t)cur * D + c], sum);
              cur = k;
              sum = 0.f;
            }
            if (k >= 0) sum += stage[rr * 65 + c];
          }
          if (cur >= 0) atomicAdd(&a.aggB[(size_t)cur * D + c], sum);
        }
        if (hp == 1 && gt < 128) {
          int* nb_ = idxb + (buf ^ 1) * 384;
          nb_[gt] = nia, nb_[128 + gt] = nib, nb_[256 + gt] = nic;
        }
        gbar(g);
      }
    } else {
      if (valid) {
        float4* po = reinterpret_cast<float4*>(a.ang_out);
#pragma unroll
        for (int i = 0; i < 8; i++)
          po[tl4<16>(t, r, c0 + 4 * i)] = make_float4(angv[4 * i] + mv[4 * i], angv[4 * i + 1] + mv[4 * i + 1], angv[4 * i + 2] + mv[4 * i + 2],
                              angv[4 * i + 3] + mv[4 * i + 3]);
      }
      if (gt < 128) {
        int* nb_ = idxb + (buf ^ 1) * 384;
        nb_[gt] = nia, nb_[128 + gt] = nib, nb_[256 + gt] = nic;
      }
      gbar(g);
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tptr), "r"(512u));
}

// backward: gm -> g (per branch) -> [HIDDEN: ghid = g . W2 ; gpre = ghid * ds] -> scatter gpre, gang += gpre . Wg
template <bool HIDDEN>
__global__ void __launch_bounds__(512, 1) k_line_bwd_tc(const LineArgs a, const LineTcW w) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x;
  const int g = tid >> 8, gt = tid & 255;
  const int warp = gt >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane, c0 = half * 32;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + LineTcSmem::kBar) + g * 6;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + LineTcSmem::kBar + 48);
  float* WAs = smem + LineTcSmem::kWA;  // WgT can
  float* WBs = smem + LineTcSmem::kWB;  // W2T can
  float* stage = smem + LineTcSmem::kGrp + g * LineTcSmem::kGrpSize;
  int* s_a = reinterpret_cast<int*>(stage + 64 * 65);
  int* s_b = s_a + 128;
  int* s_c = s_b + 128;

  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (gt == 0) {
    for (int i = 0; i < 6; i++) mbar_init_(&mbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (gt < 128) {  // indices of my first tile (later tiles: prefetched one tile ahead inside the loop)
    const int64_t e = (2 * (int64_t)blockIdx.x + g) * 128 + gt;
    const bool ok = e < a.A;
    s_a[gt] = ok ? a.a_in[e] : -1, s_b[gt] = ok ? a.a_out[e] : -1, s_c[gt] = ok ? a.a_ctr[e] : -1;
  }
  for (int i = tid; i < 4096; i += 512) reinterpret_cast<float4*>(WAs)[i] = reinterpret_cast<const float4*>(w.WgTcan)[i];
  if (HIDDEN)
    for (int i = tid; i < 4096; i += 512) reinterpret_cast<float4*>(WBs)[i] = reinterpret_cast<const float4*>(w.W2Tcan)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr + (uint32_t)g * 256u;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t COL_H = 0, COL_D = 128, COL_G = 192;  // D: ghid of the current branch; G: gang accumulator
  const uint32_t wa_addr = s_u32(WAs), wb_addr = s_u32(WBs);
  uint32_t phase = 0;

  const int64_t ntiles = (a.A + 127) / 128;
  const int64_t tstride = 2 * (int64_t)gridDim.x;
  int* idxb = s_a;  // [2][a | b | c][128]
  int buf = 0;
  for (int64_t t = 2 * (int64_t)blockIdx.x + g; t < ntiles; t += tstride, buf ^= 1) {
    const int64_t r0 = t * 128;
    const int nvalid = (int)min((int64_t)128, a.A - r0);
    const int64_t tn = t + tstride;
    s_a = idxb + buf * 384, s_b = s_a + 128, s_c = s_a + 256;
    int nia = -1, nib = -1, nic = -1;  // indices of my next tile, published at the end of this one
    if (tn < ntiles && gt < 128 && tn * 128 + gt < a.A) nia = a.a_in[tn * 128 + gt], nib = a.a_out[tn * 128 + gt], nic = a.a_ctr[tn * 128 + gt];
    if (tn < ntiles) {
      // my next tile's adjoint rows: one contiguous 32 KB block (tile-interleaved layout), DRAM -> L2 now
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.gang + tn * (128 * 64)) + gt * 128));
      if (w.l2pf) {
        l2_prefetch(a.uv + tn * (128 * 128), 128 * 128 * 4, gt, 256);
        if (HIDDEN) l2_prefetch(a.ds + tn * (128 * 128), 128 * 128 * 4, gt, 256);
      }
    }
    const bool valid = r < nvalid;
    const int ib = s_b[r];
    const float4* uv4 = reinterpret_cast<const float4*>(a.uv);
    const float* gmrow = a.gaggB + (size_t)(valid && HIDDEN ? ib : 0) * D;       // HIDDEN: upstream grad of my out-bond
    const float4* gang4 = reinterpret_cast<const float4*>(a.gang);             // !HIDDEN: my own gang row (tile-interleaved)
#pragma unroll 1
    for (int br = 0; br < 2; br++) {
      float gp[32];
      const int cb = br * 64 + c0;
      if (br == 1) {  // H is still being read by the previous branch's gang GEMM
        mbar_wait_(&mbar[2], phase);
        tc_fence_after();
      }
      // g for this branch
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int c = c0 + ch * 16 + i * 4;
          const float4 u4 = uv4[tl4<32>(t, r, c)];
          const float4 v4 = uv4[tl4<32>(t, r, 64 + c)];
          const float4 g4 = HIDDEN ? *reinterpret_cast<const float4*>(gmrow + c) : gang4[tl4<16>(t, r, c)];
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float su = sigm_(uu[j]), oG = sigm_(vv[j]);
            float gval;
            if (br == 0)
              gval = gg[j] * oG * (su * (1.f + uu[j] * (1.f - su)));
            else
              gval = gg[j] * (uu[j] * su) * oG * (1.f - oG);
            if (!valid) gval = 0.f;
            gp[ch * 16 + 4 * i + j] = gval;
            const uint32_t h = tf32_hi_bits(gval);
            hi[4 * i + j] = h;
            lo[4 * i + j] = __float_as_uint(gval - __uint_as_float(h));
          }
        }
        if (HIDDEN) {
          tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
          tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
        }
      }
      if (HIDDEN) {
        tc_wait_st();
        tc_fence_before();
        gbar(g);
        if (gt == 0) {  // ghid (64 cols) = g . W2_br
          tc_fence_after();
          uint32_t acc = 0;
#pragma unroll
          for (int term = 0; term < 3; term++) {
            const uint32_t acol = term == 1 ? 64u : 0u;
            const uint32_t bsel = wb_addr + (uint32_t)(br * 2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
              umma_ts(tbase + COL_D, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64, acc);
              acc = 1;
            }
          }
          umma_commit(&mbar[br]);
        }
        // silu'(pre) of this branch streams in while the tensor core runs (loads issued before the wait)
        const float4* ds4 = reinterpret_cast<const float4*>(a.ds);
        float4 dsr[8];
#pragma unroll
        for (int i = 0; i < 8; i++) dsr[i] = __ldcs(ds4 + tl4<32>(t, r, cb + i * 4));  // read once: streaming
        mbar_wait_(&mbar[br], phase);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          uint32_t v[16];
          tmem_ld16(tlane + COL_D + c0 + ch * 16, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 d4 = dsr[ch * 4 + i];
            gp[ch * 16 + 4 * i] = __uint_as_float(v[4 * i]) * d4.x;
            gp[ch * 16 + 4 * i + 1] = __uint_as_float(v[4 * i + 1]) * d4.y;
            gp[ch * 16 + 4 * i + 2] = __uint_as_float(v[4 * i + 2]) * d4.z;
            gp[ch * 16 + 4 * i + 3] = __uint_as_float(v[4 * i + 3]) * d4.w;
          }
        }
      }
      // gpre (32 cols of this branch) -> TMEM operand, then gang += gpre . Wg_br
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float x = valid ? gp[ch * 16 + i] : 0.f;
          const uint32_t h = tf32_hi_bits(x);
          hi[i] = h;
          lo[i] = __float_as_uint(x - __uint_as_float(h));
        }
        tmem_st16(tlane + COL_H + c0 + ch * 16, hi);
        tmem_st16(tlane + COL_H + 64 + c0 + ch * 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      gbar(g);
      if (gt == 0) {
        tc_fence_after();
        uint32_t acc = br == 0 ? 0u : 1u;
#pragma unroll
        for (int term = 0; term < 3; term++) {
          const uint32_t acol = term == 1 ? 64u : 0u;
          const uint32_t bsel = wa_addr + (uint32_t)(br * 2 + (term == 2 ? 1 : 0)) * 4096u * 4u;
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            umma_ts(tbase + COL_G, tbase + COL_H + acol + ks * 8, umma_desc(bsel + ks * 2048u, 1024u, 128u), kIdescN64, acc);
            acc = 1;
          }
        }
        umma_commit(&mbar[2 + br]);
      }
      // scatter gpre of this branch: Sigma_b -> gHb, Sigma_c -> gXc (runs), RED -> gHa
#pragma unroll 1
      for (int hp = 0; hp < 2; hp++) {
        if ((q >> 1) == hp) {
          const int rr = r - hp * 64;
#pragma unroll
          for (int i = 0; i < 32; i++) stage[rr * 65 + c0 + i] = gp[i];
        }
        gbar(g);
        {
          // 4 columns per thread: every reduction below is a 16-byte vector RED
          const int c4 = gt & 15, part = gt >> 4;  // 16 parts x 4 rows
          const int col = br * 64 + c4 * 4;
          float sb0 = 0.f, sb1 = 0.f, sb2 = 0.f, sb3 = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f, sc3 = 0.f;
          int curb = -1, curc = -1;
          const int rbeg = part * 4;
          for (int rr = rbeg; rr < rbeg + 4; rr++) {
            const int rowi = hp * 64 + rr;
            const int kb = s_b[rowi], kc = s_c[rowi];
            const float v0 = stage[rr * 65 + c4 * 4], v1 = stage[rr * 65 + c4 * 4 + 1], v2 = stage[rr * 65 + c4 * 4 + 2],
                        v3 = stage[rr * 65 + c4 * 4 + 3];
            if (kb != curb) {
              if (curb >= 0) red_add_v4(&a.gHb[(size_t)curb * D2 + col], sb0, sb1, sb2, sb3);
              curb = kb;
              sb0 = sb1 = sb2 = sb3 = 0.f;
            }
            if (kc != curc) {
              if (curc >= 0) red_add_v4(&a.gXc[(size_t)curc * D2 + col], sc0, sc1, sc2, sc3);
              curc = kc;
              sc0 = sc1 = sc2 = sc3 = 0.f;
            }
            if (kb >= 0) {
              sb0 += v0, sb1 += v1, sb2 += v2, sb3 += v3;
              sc0 += v0, sc1 += v1, sc2 += v2, sc3 += v3;
              red_add_v4(&a.gHa[(size_t)s_a[rowi] * D2 + col], v0, v1, v2, v3);
            }
          }
          if (curb >= 0) red_add_v4(&a.gHb[(size_t)curb * D2 + col], sb0, sb1, sb2, sb3);
          if (curc >= 0) red_add_v4(&a.gXc[(size_t)curc * D2 + col], sc0, sc1, sc2, sc3);
        }
        gbar(g);
      }
    }
    {
      // gang += (accumulated gpre . Wg): the read half of the read-modify-write is issued BEFORE waiting for the
      // last GEMM (nobody else touches these 32 floats of this row during the kernel), so its latency hides behind it
      float4* pg = reinterpret_cast<float4*>(a.gang);
      float4 o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = valid ? pg[tl4<16>(t, r, c0 + 4 * i)] : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait_(&mbar[3], phase);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t v[16];
        tmem_ld16(tlane + COL_G + c0 + ch * 16, v);
        tc_wait_ld();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float4 x = o[ch * 4 + i];
            x.x += __uint_as_float(v[4 * i]), x.y += __uint_as_float(v[4 * i + 1]);
            x.z += __uint_as_float(v[4 * i + 2]), x.w += __uint_as_float(v[4 * i + 3]);
            pg[tl4<16>(t, r, c0 + ch * 16 + 4 * i)] = x;
          }
        }
      }
    }
    phase ^= 1;
    if (gt < 128) {
      int* nb_ = idxb + (buf ^ 1) * 384;
      nb_[gt] = nia, nb_[128 + gt] = nib, nb_[256 + gt] = nic;
    }
    tc_fence_before();
    gbar(g);
  }
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tptr), "r"(512u));
}

void launch_line_fwd_tc(cudaStream_t st, const LineArgs& a, const LineTcW& w, bool hidden, int num_sms) {
  if (a.A <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_line_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LineTcSmem::bytes));
    B2M_CK(cudaFuncSetAttribute(k_line_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LineTcSmem::bytes));
  }
  const int64_t ntiles = (a.A + 127) / 128;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, (int64_t)num_sms);
  LineTcW wl = w;
  wl.l2pf = l2pf_level() >= 2;
  if (hidden)
    k_line_fwd_tc<true><<<grid, 512, LineTcSmem::bytes, st>>>(a, wl);
  else
    k_line_fwd_tc<false><<<grid, 512, LineTcSmem::bytes, st>>>(a, wl);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}
void launch_line_bwd_tc(cudaStream_t st, const LineArgs& a, const LineTcW& w, bool hidden, int num_sms) {
  if (a.A <= 0) return;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_line_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LineTcSmem::bytes));
    B2M_CK(cudaFuncSetAttribute(k_line_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LineTcSmem::bytes));
  }
  const int64_t ntiles = (a.A + 127) / 128;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, (int64_t)num_sms);
  LineTcW wl = w;
  wl.l2pf = l2pf_level() >= 2;
  if (hidden)
    k_line_bwd_tc<true><<<grid, 512, LineTcSmem::bytes, st>>>(a, wl);
  else
    k_line_bwd_tc<false><<<grid, 512, LineTcSmem::bytes, st>>>(a, wl);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

}  // namespace b2m

// ============================================================================================
// node-level row GEMM on tcgen05 (projections and their transposes): HBM-bound streaming kernel
// ============================================================================================
namespace b2m {

// A tiles are read and C tiles written COOPERATIVELY (coalesced 512 B per warp instruction) through a padded
// shared-memory tile [128][68]; the thread=row view needed by tcgen05.st / tcgen05.ld is taken from shared memory
// (LDS.128 / STS.128 with a 272 B pitch are conflict free).  K = 128 is processed as two accumulating K = 64 halves,
// so every shape needs 128 (A hi|lo) + N TMEM columns <= 256 and two CTAs share an SM.
template <int K, int N>
__global__ void __launch_bounds__(256, 2)
    k_gemm_tc(const float* __restrict__ A, int lda, const float* __restrict__ Bcan, float* __restrict__ C, int ldc, int M,
              const float* __restrict__ bias, const float* __restrict__ R, int ldr, int accum) {
  constexpr uint32_t COL_D = 128;
  constexpr uint32_t LBO = (N / 8) * 128;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  constexpr int PITCH = 68;
  extern __shared__ __align__(1024) float smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 4);
  float* Bs = smem + 64;
  float* stg = Bs + 2 * N * K;  // [128][68]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init_(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * N * K / 4; i += 256) reinterpret_cast<float4*>(Bs)[i] = reinterpret_cast<const float4*>(Bcan)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  const uint32_t b_addr = s_u32(Bs);
  uint32_t phase = 0;
  const int ntiles = (M + 127) / 128;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int row0 = t * 128;
#pragma unroll 1
    for (int kh = 0; kh < K / 64; kh++) {
      // cooperative, coalesced load of A[row0 .. +128, kh*64 .. +64] into the padded tile
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int idx = tid + 256 * i;
        const int rr = idx >> 4, c4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + rr < M) v = *reinterpret_cast<const float4*>(A + (size_t)(row0 + rr) * lda + kh * 64 + c4 * 4);
        *reinterpret_cast<float4*>(stg + rr * PITCH + c4 * 4) = v;
      }
      __syncthreads();
      // thread = row view: my 32 of the 64 columns -> tf32 hi/lo -> TMEM
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float4 x = *reinterpret_cast<const float4*>(stg + r * PITCH + half * 32 + ch * 16 + i * 4);
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t h = tf32_hi_bits(xv[j]);
            hi[4 * i + j] = h;
            lo[4 * i + j] = __float_as_uint(xv[j] - __uint_as_float(h));
          }
        }
        tmem_st16(tlane + half * 32 + ch * 16, hi);
        tmem_st16(tlane + 64 + half * 32 + ch * 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        uint32_t acc = kh == 0 ? 0u : 1u;
#pragma unroll
        for (int term = 0; term < 3; term++) {
          const uint32_t acol = term == 1 ? 64u : 0u;
          const uint32_t bsel = b_addr + (term == 2 ? (uint32_t)(N * K) * 4u : 0u) + (uint32_t)kh * 16u * LBO;
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            umma_ts(tbase + COL_D, tbase + acol + ks * 8, umma_desc(bsel + ks * 2 * LBO, LBO, 128u), IDESC, acc);
            acc = 1;
          }
        }
        umma_commit(mbar);
      }
      mbar_wait_(mbar, phase);
      phase ^= 1;
      tc_fence_after();
    }
    // epilogue in 64-column slabs: TMEM -> padded tile (thread = row) -> cooperative coalesced global store
#pragma unroll 1
    for (int nh = 0; nh < N / 64; nh++) {
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t v[16];
        tmem_ld16(tlane + COL_D + nh * 64 + half * 32 + ch * 16, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<float4*>(stg + r * PITCH + half * 32 + ch * 16 + i * 4) =
              make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                          __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before();
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int idx = tid + 256 * i;
        const int rr = idx >> 4, c4 = idx & 15;
        const int row = row0 + rr, col = nh * 64 + c4 * 4;
        if (row < M) {
          float4 o = *reinterpret_cast<const float4*>(stg + rr * PITCH + c4 * 4);
          if (bias) {
            const float4 b = *reinterpret_cast<const float4*>(bias + col);
            o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
          }
          if (R) {
            const float4 x = *reinterpret_cast<const float4*>(R + (size_t)row * ldr + col);
            o.x += x.x, o.y += x.y, o.z += x.z, o.w += x.w;
          }
          float4* cp = reinterpret_cast<float4*>(C + (size_t)row * ldc + col);
          if (accum) {
            const float4 c = *cp;
            o.x += c.x, o.y += c.y, o.z += c.z, o.w += c.w;
          }
          *cp = o;
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u));
}

// Software-pipelined variant (default since round 2: A/B parity green and 25.72 -> 24.81 ms/step at 97 k atoms,
// profiles/r02a_bench_97k*.json; B2M_GEMM_PIPE=0 selects the first version): the A chunk of the NEXT (tile, K-half) is loaded into registers while the
// current one is converted, multiplied and written back, and the accumulate / residual rows of an output slab are
// requested before the TMEM read-back, so that no global load is waited for right after it is issued.
// EPI (TensorNet edge MLP): 1 = keep the pre-activation in Cpre and write SiLU(value); 2 = value *= SiLU'(Pre[row][col])
// (reverse pass; applied after the accumulate, so the last K-chunk of a split product carries it).
template <int K, int N, int EPI>
__global__ void __launch_bounds__(256, 2)
    k_gemm_tc_pipe(const float* __restrict__ A, int lda, const float* __restrict__ Bcan, float* __restrict__ C, int ldc,
                   int M, const float* __restrict__ bias, const float* __restrict__ R, int ldr, int accum,
                   float* __restrict__ Cpre, const float* __restrict__ Pre, int ldp) {
  constexpr uint32_t COL_D = 128;
  constexpr uint32_t LBO = (N / 8) * 128;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  constexpr int PITCH = 68;
  constexpr int KH = K / 64;
  extern __shared__ __align__(1024) float smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 4);
  float* Bs = smem + 64;
  float* stg = Bs + 2 * N * K;  // [128][68]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tptr)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init_(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int ntiles = (M + 127) / 128;
  // my cooperative-copy coordinates: 8 float4 per thread, rows rr0 + 16*i, 16-byte column c4
  const int rr0 = tid >> 4, c4 = tid & 15;
  float4 pre[8];
  auto load_chunk = [&](int t, int kh) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int row = t * 128 + rr0 + 16 * i;
      pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < M) pre[i] = *reinterpret_cast<const float4*>(A + (size_t)row * lda + kh * 64 + c4 * 4);
    }
  };
  if ((int)blockIdx.x < ntiles) load_chunk(blockIdx.x, 0);  // in flight while the weights are staged
  for (int i = tid; i < 2 * N * K / 4; i += 256) reinterpret_cast<float4*>(Bs)[i] = reinterpret_cast<const float4*>(Bcan)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
  const uint32_t b_addr = s_u32(Bs);
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int row0 = t * 128;
#pragma unroll 1
    for (int kh = 0; kh < KH; kh++) {
#pragma unroll
      for (int i = 0; i < 8; i++) *reinterpret_cast<float4*>(stg + (rr0 + 16 * i) * PITCH + c4 * 4) = pre[i];
      // next chunk: the other K half of this tile, or the first half of my next tile
      if (kh + 1 < KH)
        load_chunk(t, kh + 1);
      else if (t + (int)gridDim.x < ntiles)
        load_chunk(t + gridDim.x, 0);
      __syncthreads();
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float4 x = *reinterpret_cast<const float4*>(stg + r * PITCH + half * 32 + ch * 16 + i * 4);
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t h = tf32_hi_bits(xv[j]);
            hi[4 * i + j] = h;
            lo[4 * i + j] = __float_as_uint(xv[j] - __uint_as_float(h));
          }
        }
        tmem_st16(tlane + half * 32 + ch * 16, hi);
        tmem_st16(tlane + 64 + half * 32 + ch * 16, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        uint32_t acc = kh == 0 ? 0u : 1u;
#pragma unroll
        for (int term = 0; term < 3; term++) {
          const uint32_t acol = term == 1 ? 64u : 0u;
          const uint32_t bsel = b_addr + (term == 2 ? (uint32_t)(N * K) * 4u : 0u) + (uint32_t)kh * 16u * LBO;
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            umma_ts(tbase + COL_D, tbase + acol + ks * 8, umma_desc(bsel + ks * 2 * LBO, LBO, 128u), IDESC, acc);
            acc = 1;
          }
        }
        umma_commit(mbar);
      }
      mbar_wait_(mbar, phase);
      phase ^= 1;
      tc_fence_after();
    }
#pragma unroll 1
    for (int nh = 0; nh < N / 64; nh++) {
      // rows this slab adds to the product (residual R, or C itself when accumulating): requested before the read-back
      const float* addsrc = R ? R : (accum ? C : nullptr);
      const int addld = R ? ldr : ldc;
      float4 add[8];
      if (addsrc) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = row0 + rr0 + 16 * i;
          add[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < M) add[i] = *reinterpret_cast<const float4*>(addsrc + (size_t)row * addld + nh * 64 + c4 * 4);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        uint32_t v[16];
        tmem_ld16(tlane + COL_D + nh * 64 + half * 32 + ch * 16, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<float4*>(stg + r * PITCH + half * 32 + ch * 16 + i * 4) =
              make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                          __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before();
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int rr = rr0 + 16 * i;
        const int row = row0 + rr, col = nh * 64 + c4 * 4;
        if (row < M) {
          float4 o = *reinterpret_cast<const float4*>(stg + rr * PITCH + c4 * 4);
          if (bias) {
            const float4 b = *reinterpret_cast<const float4*>(bias + col);
            o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
          }
          if (addsrc) o.x += add[i].x, o.y += add[i].y, o.z += add[i].z, o.w += add[i].w;
          float4* cp = reinterpret_cast<float4*>(C + (size_t)row * ldc + col);
          if (R && accum) {  // both at once (not used by the engine): the accumulate read stays in place
            const float4 c = *cp;
            o.x += c.x, o.y += c.y, o.z += c.z, o.w += c.w;
          }
          if constexpr (EPI == 1) {
            *reinterpret_cast<float4*>(Cpre + (size_t)row * ldc + col) = o;
            o = make_float4(silu_(o.x), silu_(o.y), silu_(o.z), silu_(o.w));
          } else if constexpr (EPI == 2) {
            const float4 pv = *reinterpret_cast<const float4*>(Pre + (size_t)row * ldp + col);
            o.x *= dsilu_(pv.x), o.y *= dsilu_(pv.y), o.z *= dsilu_(pv.z), o.w *= dsilu_(pv.w);
          }
          *cp = o;
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u));
}

template <int K, int N, int EPI>
static void launch_gemm_tc_epi_t(cudaStream_t st, const float* A, int lda, const float* Bcan, float* C, int ldc, int M,
                                 const float* bias, bool accum, float* Cpre, const float* Pre, int ldp, int num_sms) {
  constexpr size_t bytes = (size_t)(64 + 2 * N * K + 128 * 68) * 4;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_gemm_tc_pipe<K, N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  const int ntiles = (M + 127) / 128;
  const int grid = std::min(ntiles, 2 * num_sms);
  k_gemm_tc_pipe<K, N, EPI><<<grid, 256, bytes, st>>>(A, lda, Bcan, C, ldc, M, bias, nullptr, 0, accum ? 1 : 0, Cpre, Pre,
                                                      ldp);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}
template <int EPI>
static void launch_gemm_tc_epi_s(cudaStream_t st, const float* A, int lda, const float* Bcan, float* C, int ldc, int M,
                                 int N, int K, const float* bias, bool accum, float* Cpre, const float* Pre, int ldp,
                                 int num_sms) {
  if (K == 64 && N == 128)
    launch_gemm_tc_epi_t<64, 128, EPI>(st, A, lda, Bcan, C, ldc, M, bias, accum, Cpre, Pre, ldp, num_sms);
  else if (K == 64 && N == 64)
    launch_gemm_tc_epi_t<64, 64, EPI>(st, A, lda, Bcan, C, ldc, M, bias, accum, Cpre, Pre, ldp, num_sms);
  else if (K == 128 && N == 64)
    launch_gemm_tc_epi_t<128, 64, EPI>(st, A, lda, Bcan, C, ldc, M, bias, accum, Cpre, Pre, ldp, num_sms);
  else
    throw Error(B2M_ERR_INVALID, "gemm_tc shape");
}
void launch_gemm_tc_epi(cudaStream_t st, const float* A, int lda, const float* Bcan, float* C, int ldc, int M, int N,
                        int K, const float* bias, bool accum, int epi, float* Cpre, const float* Pre, int ldp,
                        int num_sms) {
  if (M <= 0) return;
  B2M_REQUIRE(epi == 1 ? Cpre != nullptr : (epi == 2 ? Pre != nullptr : epi == 0), B2M_ERR_INVALID, "gemm_tc epilogue");
  if (epi == 1) launch_gemm_tc_epi_s<1>(st, A, lda, Bcan, C, ldc, M, N, K, bias, accum, Cpre, Pre, ldp, num_sms);
  else if (epi == 2) launch_gemm_tc_epi_s<2>(st, A, lda, Bcan, C, ldc, M, N, K, bias, accum, Cpre, Pre, ldp, num_sms);
  else launch_gemm_tc_epi_s<0>(st, A, lda, Bcan, C, ldc, M, N, K, bias, accum, Cpre, Pre, ldp, num_sms);
}

template <int K, int N>
static void launch_gemm_tc_t(cudaStream_t st, const float* A, int lda, const float* Bcan, float* C, int ldc, int M,
                             const float* bias, const float* R, int ldr, bool accum, int num_sms) {
  constexpr size_t bytes = (size_t)(64 + 2 * N * K + 128 * 68) * 4;
  static PerDeviceOnce attr;
  if (auto once_ = attr.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_gemm_tc<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  static const bool pipe = [] {
    const char* v = getenv("B2M_GEMM_PIPE");
    return v ? atoi(v) != 0 : true;
  }();
  static PerDeviceOnce attr_pipe;
  if (auto once_ = attr_pipe.first(); once_) {
    B2M_CK(cudaFuncSetAttribute(k_gemm_tc_pipe<K, N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  const int ntiles = (M + 127) / 128;
  const int grid = std::min(ntiles, 2 * num_sms);
  if (pipe)
    k_gemm_tc_pipe<K, N, 0><<<grid, 256, bytes, st>>>(A, lda, Bcan, C, ldc, M, bias, R, ldr, accum ? 1 : 0, nullptr, nullptr, 0);
  else
    k_gemm_tc<K, N><<<grid, 256, bytes, st>>>(A, lda, Bcan, C, ldc, M, bias, R, ldr, accum ? 1 : 0);
  B2M_CK(cudaGetLastError());
  g_launch_count++;
}

void launch_gemm_tc(cudaStream_t st, const float* A, int lda, const float* Bcan, float* C, int ldc, int M, int N, int K,
                    const float* bias, const float* R, int ldr, bool accum, int num_sms) {
  if (M <= 0) return;
  if (K == 64 && N == 128)
    launch_gemm_tc_t<64, 128>(st, A, lda, Bcan, C, ldc, M, bias, R, ldr, accum, num_sms);
  else if (K == 64 && N == 64)
    launch_gemm_tc_t<64, 64>(st, A, lda, Bcan, C, ldc, M, bias, R, ldr, accum, num_sms);
  else if (K == 128 && N == 64)
    launch_gemm_tc_t<128, 64>(st, A, lda, Bcan, C, ldc, M, bias, R, ldr, accum, num_sms);
  else
    throw Error(B2M_ERR_INVALID, "gemm_tc shape");
}

}  // namespace b2m
