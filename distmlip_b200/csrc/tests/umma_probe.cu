// umma_probe.cu -- standalone probe of the tcgen05 TF32 path used by the fused kernels:
//   D[128x64] (TMEM, fp32) = A[128x64] * B[64x64]^T   with B = W[n][k] (K-major, no swizzle) in smem and
//   A either in smem (SS) or in TMEM written with tcgen05.st (TS).  Also the 3xTF32 split variant.
// usage: umma_probe <mode: 0=SS 1=TS> <swap lbo/sbo: 0|1> <split: 0|1> [bmajor lbo_b sbo_b kstep_b]
//   bmajor=1: the SAME smem image of W (canonical K-major [n][k]) is read as an MN-major B operand, i.e. the product
//   D = A * W (W^T as the B operand without a second, transposed copy of the weights); lbo_b / sbo_b / kstep_b are the
//   descriptor byte offsets to try (expected: 128, 1024, 128).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);       \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ Afmt,
                                             const float* __restrict__ Bfmt, float* __restrict__ D, uint32_t lbo_b,
                                             uint32_t sbo_b, uint32_t lbo_a, uint32_t sbo_a, int use_ts, int split, int bmajor,
                                             uint32_t kstep_b) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* Bs = reinterpret_cast<float*>(smem);              // 2 x 16 KB (hi, lo)
  float* As = reinterpret_cast<float*>(smem + 32768);      // 2 x 32 KB (hi, lo)   (SS mode)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 32768 + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 32768 + 65536 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < 2 * 4096; i += 128) Bs[i] = Bfmt[i];
  if (!use_ts)
    for (int i = tid; i < 2 * 8192; i += 128) As[i] = Afmt[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tptr;
  const uint32_t col_d = 0, col_ahi = 64, col_alo = 128;
  if (use_ts) {
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int k = 0; k < 64; k += 8) {
      uint32_t hi[8], lo[8];
      for (int j = 0; j < 8; j++) {
        float x = A[tid * 64 + k + j];
        float h = split ? tf32_hi(x) : x;
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(x - h);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       lane_addr + col_ahi + k),
                   "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7])
                   : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       lane_addr + col_alo + k),
                   "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: c=F32 (1<<4), a=TF32 (2<<7), b=TF32 (2<<10), K-major both, N=64 (8<<17), M=128 (8<<24)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (8u << 17) | (8u << 24) | (bmajor ? (1u << 16) : 0u);
    const uint32_t bs_hi = smem_u32(Bs), bs_lo = smem_u32(Bs + 4096);
    const uint32_t as_hi = smem_u32(As), as_lo = smem_u32(As + 8192);
    int first = 1;
    const int nterm = split ? 3 : 1;
    for (int term = 0; term < nterm; term++) {
      // term 0: Ahi*Bhi, 1: Alo*Bhi, 2: Ahi*Blo
      const uint32_t bsel = term == 2 ? bs_lo : bs_hi;
      const uint32_t acol = term == 1 ? col_alo : col_ahi;
      const uint32_t asel = term == 1 ? as_lo : as_hi;
      for (int ks = 0; ks < 8; ks++) {
        const uint64_t bdesc = make_desc(bsel + ks * kstep_b, lbo_b, sbo_b);
        const uint32_t acc = first ? 0u : 1u;
        first = 0;
        if (use_ts) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase + col_d),
              "r"(tbase + acol + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc)
              : "memory");
        } else {
          const uint64_t adesc = make_desc(asel + ks * 2 * lbo_a, lbo_a, sbo_a);
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + col_d),
              "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                 : "memory");
  }
  {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(mbar);
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(addr), "r"(0u)
                   : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
          "[%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(lane_addr + col_d + c)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; j++) D[tid * 64 + c + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u));
}

static float tf32_round(float x) {  // round-to-nearest-away on 13 dropped bits (cvt.rna)
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

int main(int argc, char** argv) {
  const int use_ts = argc > 1 ? atoi(argv[1]) : 1;
  const int swap = argc > 2 ? atoi(argv[2]) : 0;
  const int split = argc > 3 ? atoi(argv[3]) : 0;
  const int bmajor = argc > 4 ? atoi(argv[4]) : 0;
  std::vector<float> A(128 * 64), W(64 * 64), Dref(128 * 64), D(128 * 64);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < 64; n++) {
      double s = 0;
      for (int k = 0; k < 64; k++) s += (double)A[m * 64 + k] * (double)(bmajor ? W[k * 64 + n] : W[n * 64 + k]);
      Dref[m * 64 + n] = (float)s;
    }
  // canonical K-major no-swizzle: element (r, k) at ((k/4) * (R/8) + r/8) * 32 floats + (r%8)*4 + k%4
  auto fmt = [&](const std::vector<float>& src, int R, std::vector<float>& hi, std::vector<float>& lo) {
    hi.assign(R * 64, 0.f);
    lo.assign(R * 64, 0.f);
    for (int r = 0; r < R; r++)
      for (int k = 0; k < 64; k++) {
        size_t o = ((size_t)(k / 4) * (R / 8) + r / 8) * 32 + (r % 8) * 4 + (k % 4);
        float x = src[r * 64 + k];
        float h = split ? tf32_round(x) : x;
        hi[o] = h;
        lo[o] = x - h;
      }
  };
  std::vector<float> Bhi, Blo, Ahi, Alo;
  fmt(W, 64, Bhi, Blo);
  fmt(A, 128, Ahi, Alo);
  std::vector<float> Bf(Bhi), Af(Ahi);
  Bf.insert(Bf.end(), Blo.begin(), Blo.end());
  Af.insert(Af.end(), Alo.begin(), Alo.end());
  float *dA, *dAf, *dBf, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dAf, Af.size() * 4));
  CK(cudaMalloc(&dBf, Bf.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAf, Af.data(), Af.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBf, Bf.data(), Bf.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  // B: 8 row-groups -> LBO (between K chunks) = 8*128 B, SBO (between row groups) = 128 B ; A: 16 row groups
  uint32_t lbo_b = 8 * 128, sbo_b = 128, lbo_a = 16 * 128, sbo_a = 128;
  if (swap) {
    // alternative reading of the descriptor fields: swap roles (keeps the same memory image)
    uint32_t t = lbo_b;
    lbo_b = sbo_b;
    sbo_b = t;
    t = lbo_a;
    lbo_a = sbo_a;
    sbo_a = t;
  }
  uint32_t kstep_b = 2 * lbo_b;
  if (argc > 7) lbo_b = atoi(argv[5]), sbo_b = atoi(argv[6]), kstep_b = atoi(argv[7]);
  const size_t smem = 32768 + 65536 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(dA, dAf, dBf, dD, lbo_b, sbo_b, lbo_a, sbo_a, use_ts, split, bmajor, kstep_b);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); i++) {
    maxerr = fmax(maxerr, fabs((double)D[i] - (double)Dref[i]));
    maxref = fmax(maxref, fabs((double)Dref[i]));
  }
  printf("bmajor=%d lbo_b=%u sbo_b=%u kstep_b=%u ", bmajor, lbo_b, sbo_b, kstep_b);
  printf("mode=%s swap=%d split=%d  max|err|=%.3e  max|ref|=%.3e  D[0]=%f ref=%f  D[last]=%f ref=%f\n",
         use_ts ? "TS" : "SS", swap, split, maxerr, maxref, D[0], Dref[0], D.back(), Dref.back());
  return 0;
}
