// gather_probe.cu -- microbenchmarks behind the staging design of the edge-gather kernels (sm_100a):
// how fast can ONE SM bring 128 gathered 512-byte rows (one edge tile of A[src]) into shared memory, and how fast are
// the MUFU-bound activations, measured in SM clocks per 128-row tile with one 512-thread CTA per SM.
//   mode 0: cp.async.bulk (UBLKCP), one 512 B copy per row, issued by 128 threads, mbarrier completion, 2 stages
//   mode 1: cp.async.bulk, two 256 B copies per row (the round-1 "v2" pattern)
//   mode 2: per-lane LDG.128 of thread-owned rows (thread = row, 32 different lines per warp instruction)
//   mode 3: warp-per-row coalesced LDG.128 -> STS.128 into a padded stage, then thread = row LDS.128 reads
//   mode 4: warp-per-row cp.async (LDGSTS.128, 16 B per lane), commit/wait groups, 2 stages, then thread = row reads
//   mode 5: sigmoid throughput (ex2 + rcp per element), 256 per row like one edge of the atom conv
//   mode 6: same with one rcp per PAIR of sigmoids
// usage: gather_probe <mode> [tiles_per_cta]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                \
  do {                                                                       \
    cudaError_t e = (x);                                                     \
    if (e != cudaSuccess) {                                                  \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); \
      exit(2);                                                               \
    }                                                                        \
  } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(s_u32(bar)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

constexpr int PITCH = 132;  // floats per staged row (528 B: thread = row LDS.128 is conflict-free)

__global__ void __launch_bounds__(512, 1) k_probe(int mode, int tiles, const float* __restrict__ table,
                                                  const int* __restrict__ idx, float* __restrict__ out,
                                                  long long* __restrict__ clk) {
  extern __shared__ __align__(1024) float smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);  // 2 barriers
  float* stage0 = smem + 16;
  float* stage1 = stage0 + 128 * PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int* myidx = idx + (size_t)blockIdx.x * tiles * 128;
  float acc = 0.f;
  const long long t0 = clock64();
  if (mode == 0 || mode == 1) {
    auto issue = [&](int t) {
      float* st = (t & 1) ? stage1 : stage0;
      if (tid == 0) expect_tx(&bar[t & 1], 128 * 512);
      if (tid < 128) {
        const float* src = table + (size_t)myidx[t * 128 + tid] * 128;
        if (mode == 0) {
          bulk(st + tid * PITCH, src, 512, &bar[t & 1]);
        } else {
          bulk(st + tid * PITCH, src, 256, &bar[t & 1]);
          bulk(st + tid * PITCH + 64, src + 64, 256, &bar[t & 1]);
        }
      }
    };
    issue(0);
    uint32_t ph[2] = {0, 0};
    for (int t = 0; t < tiles; t++) {
      if (t + 1 < tiles) issue(t + 1);
      mbar_wait(&bar[t & 1], ph[t & 1]);
      ph[t & 1] ^= 1;
      const float* st = (t & 1) ? stage1 : stage0;
      if (tid < 256) {  // thread = (row, half): read my 64 columns
        const float4* p = reinterpret_cast<const float4*>(st + (tid & 127) * PITCH + (tid >> 7) * 64);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 v = p[i];
          acc += v.x + v.y + v.z + v.w;
        }
      }
      __syncthreads();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  } else if (mode == 2) {
    for (int t = 0; t < tiles; t++) {
      if (tid < 256) {
        const float4* p = reinterpret_cast<const float4*>(table + (size_t)myidx[t * 128 + (tid & 127)] * 128 + (tid >> 7) * 64);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 v = p[i];
          acc += v.x + v.y + v.z + v.w;
        }
      }
    }
  } else if (mode == 3) {
    for (int t = 0; t < tiles; t++) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {  // 16 warps x 8 rows
        const int row = warp * 8 + j;
        v[j] = *reinterpret_cast<const float4*>(table + (size_t)myidx[t * 128 + row] * 128 + lane * 4);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(stage0 + (warp * 8 + j) * PITCH + lane * 4) = v[j];
      __syncthreads();
      if (tid < 256) {
        const float4* p = reinterpret_cast<const float4*>(stage0 + (tid & 127) * PITCH + (tid >> 7) * 64);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 x = p[i];
          acc += x.x + x.y + x.z + x.w;
        }
      }
      __syncthreads();
    }
  } else if (mode == 4) {
    auto issue = [&](int t) {
      float* st = (t & 1) ? stage1 : stage0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int row = warp * 8 + j;
        const float* src = table + (size_t)myidx[t * 128 + row] * 128 + lane * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(st + row * PITCH + lane * 4)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    for (int t = 0; t < tiles; t++) {
      if (t + 1 < tiles) {
        issue(t + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const float* st = (t & 1) ? stage1 : stage0;
      if (tid < 256) {
        const float4* p = reinterpret_cast<const float4*>(st + (tid & 127) * PITCH + (tid >> 7) * 64);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 x = p[i];
          acc += x.x + x.y + x.z + x.w;
        }
      }
      __syncthreads();
    }
  } else {
    // activations: 128 rows x 256 sigmoids per tile, 512 threads -> 64 per thread per tile
    float x = 0.001f * tid;
    for (int t = 0; t < tiles; t++) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float e0, e1, r0, r1;
        const float a = x + i * 0.01f, b = x - i * 0.02f;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a * -1.4426950408889634f));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(b * -1.4426950408889634f));
        if (mode == 5) {
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.f + e0));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.f + e1));
          acc += a * r0 * r1;
        } else {
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((1.f + e0) * (1.f + e1)));
          acc += a * r0;
        }
      }
      x += acc * 1e-20f;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 512 + tid] = acc;
  if (tid == 0) clk[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int tiles = argc > 2 ? atoi(argv[2]) : 144;
  const int nrows = 100000, grid = 148;
  std::vector<int> idx((size_t)grid * tiles * 128);
  srand(7);
  // edge-like locality: 28 consecutive "edges" share a destination; their sources are within +-2000 rows of it
  for (size_t i = 0; i < idx.size(); i++) {
    const long long centre = (long long)(i / 28) * 28 * nrows / (long long)idx.size();
    long long v = centre + (rand() % 4001) - 2000;
    if (v < 0) v += nrows;
    if (v >= nrows) v -= nrows;
    idx[i] = (int)v;
  }
  float *table, *out;
  int* didx;
  long long* clk;
  CK(cudaMalloc(&table, (size_t)nrows * 512));
  CK(cudaMemset(table, 0, (size_t)nrows * 512));
  CK(cudaMalloc(&didx, idx.size() * 4));
  CK(cudaMemcpy(didx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, grid * 512 * 4));
  CK(cudaMalloc(&clk, grid * 8));
  const size_t smem = 64 + 2 * 128 * PITCH * 4;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    k_probe<<<grid, 512, smem>>>(mode, tiles, table, didx, out, clk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), clk, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto c : h) mx = c > mx ? c : mx;
  printf("mode %d tiles/CTA %d: %.3f ms  (%.0f clk per 128-row tile, slowest CTA %lld clk)  -> %.2f M rows/ms chip-wide\n",
         mode, tiles, best, (double)mx / tiles, mx, (double)grid * tiles * 128 / best / 1e3);
  return 0;
}
