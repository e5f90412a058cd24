// Developer microbenchmark (not product code): how many SM cycles does one bf16 tcgen05.mma
// (M = 128 per CTA, K = 16, both operands in 128B-swizzled shared memory) take as a function of
// N, for cta_group::1 and cta_group::2, with and without concurrent LDS traffic on the same SM?
// Used to decide whether the conv kernel is bound by the tensor pipe or by shared-memory operand
// bandwidth ("what bounds conv_tc_kernel").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_bw tools/umma_bw.cu -lcuda
#include <cuda.h>
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
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(192, 1) umma_bw(int N, int iters, int noise, int a_stages, int d_stages, int M,
                                                  unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_base = base, b_base = base + 8 * 16384;
  const uint32_t bar = b_base + 2 * 16384 * 2, slot = bar + 8;
  uint32_t rank = 0;
  if (CG == 2) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // fill operands with small bf16 values
  for (uint32_t i = threadIdx.x; i < (8 * 16384 + 4 * 16384) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u + ((i * 2654435761u) >> 20 & 0x00FF00FFu);
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile uint32_t*>(gen + (slot - base) + 8) = 0u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((M * CG) >> 4) << 24);

  if (warp == 1 && rank == 0) {
    if (lane == 0) {
      const long long t0 = clock64();
      // each iteration: 4 K-steps x d_stages independent accumulators, interleaved so that consecutive
      // MMAs never target the same accumulator when d_stages > 1
      for (int it = 0; it < iters; ++it) {
        const uint32_t a_lo = (a_base + (uint32_t)(it % a_stages) * 16384u) >> 4;
        const uint32_t b_lo = (b_base + (uint32_t)(it & 1) * 32768u) >> 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          for (int s = 0; s < d_stages; ++s)
            mma<CG>(tmem + (uint32_t)(s * N), mk_desc(a_lo + 2 * k, DESC_HI), mk_desc(b_lo + 2 * k, DESC_HI), idesc,
                    (uint32_t)(k | it));
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      uint32_t ok = 0, spins = 0;
      while (!ok && ++spins < (1u << 24))      // bounded: a bad encoding must not hang the box
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar) : "memory");
      const long long t1 = clock64();
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
      *reinterpret_cast<volatile uint32_t*>(gen + (slot - base) + 8) = 1u;
      if (CG == 2) {   // release the peer CTA's traffic warps too
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote) : "r"(slot + 8));
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
      }
    }
  } else if (warp >= 2 && warp - 2 < noise) {
    // concurrent shared-memory traffic until the MMA thread is done: conflict-free LDS.128
    // (512 B per warp instruction = 4 wavefronts); `noise` = number of warps doing this
    const uint4* p = reinterpret_cast<const uint4*>(gen) + (warp - 2) * 1024 + lane;
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(gen + (slot - base) + 8);
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned long long n = 0;
    while (*flag == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p + j * 32)));
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
      }
      n += 8;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) out[1000] = 1;
    if (lane == 0 && warp == 2) out[256 + blockIdx.x] = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int CG>
static double run(int N, int iters, int noise, int a_stages, int d_stages, int M, unsigned long long* d_out, double* lds_bpc) {
  const int smem = 8 * 16384 + 4 * 16384 + 2048;
  cudaFuncSetAttribute(umma_bw<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaMemset(d_out, 0, 2048 * 8);
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_bw<CG>, N, iters, noise, a_stages, d_stages, M, d_out);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  unsigned long long h[512];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0, l = 0; int n = 0;
  for (int i = 0; i < 148; ++i) if (h[i]) { s += (double)h[i]; l += (double)h[256 + i] * 512.0 * noise / (double)h[i]; ++n; }
  *lds_bpc = l / n;
  return s / n / (iters * 4.0 * d_stages);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 2048 * 8);
  const int iters = 1000;
  printf("cycles per tcgen05.mma (K=16, bf16, SS); d_st = independent accumulators interleaved\n");
  printf("%3s %4s %4s %5s %5s | %9s %9s %11s %11s\n", "CG", "M", "N", "d_st", "noise", "cyc/MMA", "ideal", "opB/clk/SM", "ldsB/clk/SM");
  const int Ns[6] = {32, 64, 96, 128, 192, 256};
  for (int cg = 1; cg <= 2; ++cg)
    for (int M = (cg == 1 ? 64 : 128); M <= 128; M += 64)
      for (int ni = 0; ni < 6; ++ni)
        for (int ds = 1; ds <= 8; ds *= 2) {
          const int N = Ns[ni];
          if (ds * N > 512) continue;
          for (int noise = 0; noise <= 4; noise += 4) {
            if (noise && !(ds == 1 || ds * N == 512 || ds == 4)) continue;
            double lds = 0;
            const double c = cg == 1 ? run<1>(N, iters, noise, 8, ds, M, d_out, &lds) : run<2>(N, iters, noise, 8, ds, M, d_out, &lds);
            printf("%3d %4d %4d %5d %5d | %9.1f %9.1f %11.1f %11.1f\n", cg, M, N, ds, noise, c, (M < 128 ? 128 : M) * N / 256.0,
                   (M * 32.0 + 32.0 * N / cg) / c, lds);
          }
        }
  return 0;
}
