// Developer microbenchmark (not product code): pure issue-rate of tcgen05.mma from straight-line
// code.  One (or two) issuing thread(s), descriptors precomputed in registers, 16 MMAs per loop
// trip.  ND = number of independent accumulators the 16 MMAs rotate over; NW = issuing warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_issue tools/umma_issue.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, int ND, int NW, int CG>
__global__ void __launch_bounds__(128, 1) k_issue(int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const uint32_t a_base = base, b_base = base + 4 * 16384;
  const uint32_t bar = b_base + 2 * 32768, slot = bar + 16;
  for (uint32_t i = threadIdx.x; i < (4 * 16384 + 2 * 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u + ((i * 2654435761u) >> 20 & 0x00FF00FFu);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t rank = 0;
  if (CG == 2) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (warp == 3) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  if (warp < NW && rank == 0) {
    if (elect_one_sync()) {
      uint64_t ad[16], bd[4];
#pragma unroll
      for (int i = 0; i < 16; ++i) ad[i] = mk_desc(((a_base + (uint32_t)(i >> 2) * 16384u) >> 4) + 2 * (i & 3), DESC_HI);
#pragma unroll
      for (int i = 0; i < 4; ++i) bd[i] = mk_desc(((b_base + (uint32_t)warp * 32768u) >> 4) + 2 * i, DESC_HI);
      const uint32_t d0 = tmem + (uint32_t)(warp * ND * N);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mma<CG>(d0 + (uint32_t)((i % ND) * N), ad[i], bd[i & 3], idesc);
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * warp) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * warp) : "memory");
      uint32_t ok = 0, spins = 0;
      while (!ok && ++spins < (1u << 24))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar + 8 * warp) : "memory");
      const long long t1 = clock64();
      out[blockIdx.x * 2 + warp] = (unsigned long long)(t1 - t0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N, int ND, int NW, int CG>
static void run(unsigned long long* d_out) {
  const int smem = 4 * 16384 + 2 * 32768 + 2048, iters = 500;
  cudaFuncSetAttribute(k_issue<N, ND, NW, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d_out, 0, 4096 * 8);
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_issue<N, ND, NW, CG>, iters, d_out);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  unsigned long long h[296];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; int n = 0;
  for (int i = 0; i < 296; ++i) if (h[i]) { s += (double)h[i]; ++n; }
  const double per_warp = s / n / (iters * 16.0);
  printf("CG=%d N=%3d accumulators=%d issuers=%d : %7.1f cycles per MMA per issuer, %7.1f per MMA overall (tensor floor %d)\n", CG, N, ND, NW,
         per_warp, per_warp / NW, N / 2);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 4096 * 8);
  run<32, 1, 1, 1>(d_out); run<64, 1, 1, 1>(d_out); run<96, 1, 1, 1>(d_out); run<128, 1, 1, 1>(d_out); run<192, 1, 1, 1>(d_out); run<256, 1, 1, 1>(d_out);
  run<64, 2, 2, 1>(d_out);
  run<32, 1, 1, 2>(d_out); run<64, 1, 1, 2>(d_out); run<96, 1, 1, 2>(d_out); run<128, 1, 1, 2>(d_out); run<192, 1, 1, 2>(d_out); run<256, 1, 1, 2>(d_out);
  run<64, 4, 1, 2>(d_out); run<128, 2, 1, 2>(d_out);
  return 0;
}
