// NHWC bf16 implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (cta_group::1,
// M=128, N=Cout<=256, K=16) -> fp32 accumulators in TMEM -> tcgen05.ld -> fused epilogue.
//
// Serves every multi-channel conv on the hot path:
// 3x3 stride-1 (reflect or zero ring), 3x3 stride-2, ConvTranspose2d k=2S/stride S as S*S
// sub-pixel phases, 1x1, with up to three K sources (the never-materialised torch.cat).
//
// Geometry.  Activations are padded NHWC [B][H+2][W+2][ct] bf16; all images are viewed as
// ONE stack of B*(H+2) rows, so an M tile is 16 stacked rows x 8 columns = 128 output
// positions and may straddle images (ring rows become masked garbage rows: efficiency
// H/(H+2) instead of rounding H up to 16 per image).  Per (source, 64-channel chunk, kx) the
// producer issues ONE TMA box {64ch, 8px, 18 rows}; the three ky taps are the same smem box
// read at row offsets 0/1/2 (1 row = 8 px * 128 B = one 1024 B swizzle atom, so every UMMA
// descriptor stays 1024 B aligned).  A traffic = 3 * 18/16 = 3.4x instead of 9x.  Weights
// stream through their own ring as [Npad][64] K-major tiles.  MT (1|2) M tiles share every
// weight tile to halve B traffic.
//
// Pipeline.  warp 0: TMA producer (1 lane).  warp 1: MMA issuer (1 lane) + TMEM owner.
// warps 2-5: epilogue (thread = output pixel = TMEM lane; channels in 16-column chunks).
// Rings: A full/empty, B full/empty, TMEM full/empty (2 accumulator stages) so the epilogue
// of tile i overlaps the MMAs of tile i+1.  Persistent CTAs, static round-robin tile order.
#include <cuda.h>
#include <string.h>
#include <new>
#include "epilogue.cuh"

namespace srd {

constexpr int TC_THREADS = 192;
constexpr int TILE_W = 8, TILE_H = 16;
constexpr int BOX_ROWS = TILE_H + 2;
constexpr int A_BYTES = BOX_ROWS * TILE_W * 128;   // 18432: one {64ch x 8px x 18rows} bf16 box
constexpr int MAX_STAGES = 8;

struct TcArgs {
  CUtensorMap tmA[SRD_MAX_SRC];
  CUtensorMap tmB;
  int n_src;
  int nchunk[SRD_MAX_SRC];
  int H_m, W_m;            // output positions per image (per phase)
  int period, voff, vstart, n_vrows;
  int n_colblk, tiles_per_phase, items_per_phase, n_items;
  int stride, nky, nkx, ky0, kx0, up;
  int Npad, MT;
  int a_stages, b_stages;
  int kb_per_phase;
  uint32_t idesc;
  Epi ep;
  const float* head_w;
  const float* head_b;
  int* err_flag;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hang costs the whole box).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, code);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct Ring {
  int stage, phase, n;
  __device__ __forceinline__ void next() { if (++stage == n) { stage = 0; phase ^= 1; } }
};

// taps served by one A box load: stride 1 -> all nky taps at row offset ky0+j;
// stride 2 -> load 0 = even input rows (ky 0 -> off 0, ky 2 -> off 1), load 1 = odd rows (ky 1).
__device__ __forceinline__ int n_loads(const TcArgs& a) { return a.stride == 2 ? 2 : 1; }
__device__ __forceinline__ int n_taps(const TcArgs& a, int ld, int ky0) {
  (void)ky0;
  return a.stride == 2 ? (ld == 0 ? 2 : 1) : a.nky;
}
__device__ __forceinline__ int tap_off(const TcArgs& a, int ld, int j, int ky0) {
  return a.stride == 2 ? (ld == 0 ? j : 0) : ky0 + j;
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B_BYTES = a.Npad * 128;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)a.a_stages * a.MT * A_BYTES;
  const uint32_t bar_base = b_base + (uint32_t)a.b_stages * B_BYTES;
  // barrier layout (8 B each): A_full[8] A_empty[8] B_full[8] B_empty[8] T_full[2] T_empty[2]
  auto A_full = [&](int s) { return bar_base + 8u * s; };
  auto A_empty = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto B_full = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto B_empty = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + s); };
  auto T_full = [&](int s) { return bar_base + 8u * (4 * MAX_STAGES + s); };
  auto T_empty = [&](int s) { return bar_base + 8u * (4 * MAX_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * MAX_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.a_stages; ++s) { mbar_init(A_full(s), 1); mbar_init(A_empty(s), 1); }
    for (int s = 0; s < a.b_stages; ++s) { mbar_init(B_full(s), 1); mbar_init(B_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(T_full(s), 1); mbar_init(T_empty(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < a.n_src; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmA[s]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int S = a.up;
  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      Ring ra{0, 0, a.a_stages}, rb{0, 0, a.b_stages};
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int phase = item / a.items_per_phase;
        const int t0 = (item % a.items_per_phase) * a.MT;
        const int pr = phase / S, ps = phase % S;
        int ky0 = a.ky0, kx0 = a.kx0;
        if (S > 1) { ky0 = (pr >= S / 2) ? 1 : 0; kx0 = (ps >= S / 2) ? 1 : 0; }
        int kb = phase * a.kb_per_phase;
        for (int s = 0; s < a.n_src; ++s) {
          for (int ch = 0; ch < a.nchunk[s]; ++ch) {
            for (int kxi = 0; kxi < a.nkx; ++kxi) {
              const int kxp = kx0 + kxi;
              for (int ld = 0; ld < n_loads(a); ++ld) {
                mbar_wait(A_empty(ra.stage), ra.phase ^ 1, a.err_flag, 1);
                mbar_expect_tx(A_full(ra.stage), (uint32_t)(a.MT * A_BYTES));
                for (int mt = 0; mt < a.MT; ++mt) {
                  const int tt = t0 + mt;
                  const int rowblk = tt / a.n_colblk, colblk = tt % a.n_colblk;
                  const int V0 = rowblk * TILE_H + a.vstart, x0 = colblk * TILE_W;
                  const uint32_t dst = a_base + (uint32_t)(ra.stage * a.MT + mt) * A_BYTES;
                  if (a.stride == 1)
                    tma_load_3d(dst, &a.tmA[s], A_full(ra.stage), ch * 64, x0 + kxp, V0 - 1);
                  else
                    tma_load_5d(dst, &a.tmA[s], A_full(ra.stage), ch * 64, kxp & 1, x0 + (kxp >> 1),
                                ld, V0);
                }
                ra.next();
                const int nt = n_taps(a, ld, ky0);
                for (int j = 0; j < nt; ++j) {
                  mbar_wait(B_empty(rb.stage), rb.phase ^ 1, a.err_flag, 2);
                  mbar_expect_tx(B_full(rb.stage), (uint32_t)B_BYTES);
                  tma_load_2d(b_base + (uint32_t)rb.stage * B_BYTES, &a.tmB, B_full(rb.stage), 0, kb * a.Npad);
                  rb.next();
                  ++kb;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      Ring ra{0, 0, a.a_stages}, rb{0, 0, a.b_stages}, rt{0, 0, 2};
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int phase = item / a.items_per_phase;
        const int pr = phase / S;
        int ky0 = a.ky0;
        if (S > 1) ky0 = (pr >= S / 2) ? 1 : 0;
        mbar_wait(T_empty(rt.stage), rt.phase ^ 1, a.err_flag, 3);
        tc_fence_after();
        uint32_t kcount = 0;
        for (int s = 0; s < a.n_src; ++s) {
          for (int ch = 0; ch < a.nchunk[s]; ++ch) {
            for (int kxi = 0; kxi < a.nkx; ++kxi) {
              for (int ld = 0; ld < n_loads(a); ++ld) {
                mbar_wait(A_full(ra.stage), ra.phase, a.err_flag, 4);
                tc_fence_after();
                const int nt = n_taps(a, ld, ky0);
                for (int j = 0; j < nt; ++j) {
                  mbar_wait(B_full(rb.stage), rb.phase, a.err_flag, 5);
                  tc_fence_after();
                  const uint32_t off = (uint32_t)tap_off(a, ld, j, ky0) * (TILE_W * 128);
                  const uint32_t b_addr = b_base + (uint32_t)rb.stage * B_BYTES;
                  for (int mt = 0; mt < a.MT; ++mt) {
                    const uint32_t a_addr = a_base + (uint32_t)(ra.stage * a.MT + mt) * A_BYTES + off;
                    const uint32_t d = tmem_base + (uint32_t)((rt.stage * a.MT + mt) * a.Npad);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                      tc_mma_bf16(d, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(b_addr + k4 * 32),
                                  a.idesc, (kcount | (uint32_t)k4) != 0u);
                  }
                  tc_commit(B_empty(rb.stage));
                  rb.next();
                  ++kcount;
                }
                tc_commit(A_empty(ra.stage));
                ra.next();
              }
            }
          }
        }
        tc_commit(T_full(rt.stage));
        rt.next();
      }
    }
  } else {
    // ================================================================= epilogue (warps 2..5)
    Ring rt{0, 0, 2};
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int m = quarter * 32 + lane;
    const int g = m >> 3, px = m & 7;
    const bool has_head = a.head_w != nullptr;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int phase = item / a.items_per_phase;
      const int t0 = (item % a.items_per_phase) * a.MT;
      const int pr = phase / S, ps = phase % S;
      mbar_wait(T_full(rt.stage), rt.phase, a.err_flag, 6);
      tc_fence_after();
      for (int mt = 0; mt < a.MT; ++mt) {
        const int tt = t0 + mt;
        const int rowblk = tt / a.n_colblk, colblk = tt % a.n_colblk;
        const int V = rowblk * TILE_H + a.vstart + g;
        const int x = colblk * TILE_W + px;
        bool valid = (tt < a.tiles_per_phase) && (V < a.n_vrows) && (x < a.W_m);
        int b = 0, y = 0;
        if (valid) {
          b = V / a.period;
          y = V % a.period - a.voff;
          valid = (y >= 0) && (y < a.H_m);
        }
        const int oy = y * S + pr, ox = x * S + ps;
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                               (uint32_t)((rt.stage * a.MT + mt) * a.Npad);
        float head[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 < a.Npad; c0 += 16) {
          uint32_t r[16];
          tc_ld16(tbase + (uint32_t)c0, r);
          tc_ld_wait();
          if (valid && c0 < a.ep.N) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            epilogue_math<__nv_bfloat16, 16, 8, false>(a.ep, b, oy, ox, c0, v);
            epilogue_store<__nv_bfloat16, 16, 8>(a.ep, b, oy, ox, c0, v, !has_head);
            if (has_head) {
              const int nv = min(16, a.ep.N - c0);
              for (int jh = 0; jh < a.ep.planar_c; ++jh) {
                float acc = head[jh];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nv) acc = fmaf(__ldg(a.head_w + (size_t)jh * a.ep.N + c0 + i), v[i], acc);
                head[jh] = acc;
              }
            }
          }
        }
        if (has_head && valid) {
          for (int jh = 0; jh < a.ep.planar_c; ++jh)
            a.ep.planar[((size_t)(b * a.ep.planar_c + jh) * a.ep.H_out + oy) * a.ep.W_out + ox] =
                head[jh] + __ldg(a.head_b + jh);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(T_empty(rt.stage));
      rt.next();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

int tc_supported(const SrdConvDesc& d, const char** why) {
  const char* w = nullptr;
  if (d.dtype != SRD_BF16) w = "dtype is not bf16";
  else if (!d.w_tc) w = "w_tc is NULL";
  else if (d.N < 1 || d.N > 256) w = "N out of range";
  else if (d.stride == 2 && (d.up != 1 || d.nky != 3 || d.nkx != 3 || (d.H_in % 2) || (d.W_in % 2)))
    w = "stride-2 needs a plain 3x3 on even dims";
  else if (d.stride != 1 && d.stride != 2) w = "stride";
  else if (d.head_w && (d.planar_c > 4 || !d.planar)) w = "head needs planar_c <= 4";
  else {
    for (int i = 0; i < d.n_src; ++i) {
      const SrdSrc& s = d.src[i];
      if ((s.ct % 8) || (s.c_off % 8) || s.cin < 1 || ((uintptr_t)s.ptr % 16)) { w = "source alignment"; break; }
    }
  }
  if (why) *why = w;
  return w == nullptr;
}

struct TcPlan {
  TcArgs args;
  int grid, smem;
};

int tc_plan_build(const SrdConvDesc& d, TcPlan* plan, int* err_flag) {
  const char* why = nullptr;
  SRD_CHECK(tc_supported(d, &why), "conv_tc: unsupported descriptor: %s", why ? why : "?");
  EncodeTiledFn enc = get_encode_fn();
  SRD_CHECK(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled not available (no CUDA driver?)");
  TcArgs& a = plan->args;
  memset(&a, 0, sizeof(a));
  const int Hp = d.H_in + 2, Wp = d.W_in + 2;
  a.n_src = d.n_src;
  a.stride = d.stride; a.nky = d.nky; a.nkx = d.nkx; a.ky0 = d.ky0; a.kx0 = d.kx0; a.up = d.up;
  a.H_m = d.H_in / d.stride; a.W_m = d.W_in / d.stride;
  a.Npad = (d.N + 15) / 16 * 16;
  if (d.stride == 1) { a.period = Hp; a.voff = 1; a.vstart = 1; }
  else { a.period = Hp / 2; a.voff = 0; a.vstart = 0; }
  a.n_vrows = d.B * a.period;
  const int n_rowblk = (a.n_vrows - a.vstart + TILE_H - 1) / TILE_H;
  a.n_colblk = (a.W_m + TILE_W - 1) / TILE_W;
  a.tiles_per_phase = n_rowblk * a.n_colblk;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int n_phase = d.up * d.up;
  a.MT = (2 * 2 * a.Npad <= 512 && (long long)n_phase * a.tiles_per_phase >= 2LL * n_sm) ? 2 : 1;
  a.items_per_phase = (a.tiles_per_phase + a.MT - 1) / a.MT;
  a.n_items = n_phase * a.items_per_phase;
  int kb = 0;
  const int taps = d.nky * d.nkx;
  for (int i = 0; i < d.n_src; ++i) { a.nchunk[i] = (d.src[i].cin + 63) / 64; kb += a.nchunk[i] * taps; }
  a.kb_per_phase = kb;
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.Npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  a.ep = make_epi(d, nullptr);
  a.head_w = d.head_w; a.head_b = d.head_b;
  a.err_flag = err_flag;
  // shared memory: A ring + B ring + barriers
  const int B_BYTES = a.Npad * 128;
  const int budget = 216 * 1024;
  a.a_stages = (a.MT == 2) ? 3 : 4;
  int bst = (budget - a.a_stages * a.MT * A_BYTES) / B_BYTES;
  a.b_stages = bst > MAX_STAGES ? MAX_STAGES : bst;
  SRD_CHECK(a.b_stages >= 2, "conv_tc: not enough shared memory for the weight ring");
  plan->smem = a.a_stages * a.MT * A_BYTES + a.b_stages * B_BYTES + 8 * (4 * MAX_STAGES + 4) + 16 + 1024;
  plan->grid = a.n_items < n_sm ? a.n_items : n_sm;

  // ---- tensor maps
  for (int i = 0; i < d.n_src; ++i) {
    const SrdSrc& s = d.src[i];
    void* base = (void*)((const __nv_bfloat16*)s.ptr + s.c_off);
    CUresult r;
    if (d.stride == 1) {
      cuuint64_t dims[3] = {(cuuint64_t)s.cin, (cuuint64_t)Wp, (cuuint64_t)d.B * Hp};
      cuuint64_t strides[2] = {(cuuint64_t)s.ct * 2, (cuuint64_t)Wp * s.ct * 2};
      cuuint32_t box[3] = {64, TILE_W, BOX_ROWS};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&a.tmA[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      // [rows/2][row parity][cols/2][col parity][c]
      cuuint64_t dims[5] = {(cuuint64_t)s.cin, 2, (cuuint64_t)Wp / 2, 2, (cuuint64_t)d.B * Hp / 2};
      cuuint64_t strides[4] = {(cuuint64_t)s.ct * 2, (cuuint64_t)2 * s.ct * 2, (cuuint64_t)Wp * s.ct * 2,
                               (cuuint64_t)2 * Wp * s.ct * 2};
      cuuint32_t box[5] = {64, 1, TILE_W, 1, BOX_ROWS};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      r = enc(&a.tmA[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A%d) failed with %d", i, (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)n_phase * kb * a.Npad};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)a.Npad};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d.w_tc, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  static bool attr_set = false;
  if (!attr_set) {
    SRD_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  return 0;
}

TcPlan* tc_plan_new() { return new (std::nothrow) TcPlan; }
void tc_plan_delete(TcPlan* p) { delete p; }

int launch_conv_tc(const TcPlan& plan, const float* temb, cudaStream_t stream) {
  TcArgs a = plan.args;
  if (temb) a.ep.temb = temb;
  conv_tc_kernel<<<plan.grid, TC_THREADS, plan.smem, stream>>>(a);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace srd
