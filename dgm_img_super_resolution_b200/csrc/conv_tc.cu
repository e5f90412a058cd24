// NHWC bf16 implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma.cta_group::2 (CTA pairs, M = 256 = two
// 128-pixel tiles, N = Cout <= 256, K = 16) -> fp32 accumulators in TMEM -> tcgen05.ld -> fused epilogue -> TMA stores.
//
// Serves every multi-channel conv on the hot path: 3x3 stride-1 (reflect or zero
// ring), 3x3 stride-2, ConvTranspose2d k=2S/stride S as S*S sub-pixel phases, 1x1, up to three K sources (the
// never-materialised torch.cat), an optional second 1x1 output (res_conv) and a fused 1x1 head (final_conv.1).
//
// Geometry.  Activations are padded NHWC [B][H+2][W+2][ct] bf16; all images are viewed as ONE stack of B*(H+2) rows,
// so an M tile is 16 stacked rows x 8 columns = 128 output positions and may straddle images (ring rows become
// masked rows: efficiency H/(H+2) instead of rounding H up to 16 per image).  Stride 1: per (source, 64-channel
// chunk) ONE halo box {64 ch, 10 px, 18 rows} per tile; the nine taps are that box addressed through UMMA descriptor
// start offsets of (ky*10 + kx)*128 B with SBO = 1280 B.  Stride 2: one box per (kx, row parity) of a parity-split
// 5-D map.  Weights are [k-block][Npad][64] K-major tiles, resident in shared memory when they fit, else streamed
// through a ring; each CTA of a pair holds half of every weight tile.  MT (1|2) tile pairs share every weight tile.
//
// Roles (608 threads): warps 0-15 epilogue (4 groups x the 4 TMEM lane quarters; thread = output pixel = TMEM lane,
// 16-column chunks; per-warp swizzled staging buffer + cp.async.bulk.tensor stores, residual prefetched into it),
// warp 16 activation-TMA producer, warp 17 weight-TMA producer, warp 18 MMA issuer + TMEM owner (one thread each).
// Rings: A full/empty, B full/empty, 2-4 TMEM accumulator stages full/empty.  Persistent CTAs, static item order.
//
// Two entry points share the role code: conv_tc_kernel (one layer per launch) and conv_chain_kernel (a list of
// layers in ONE launch; CTAs move from layer to layer on their own, ordered by per-row-block completion counters
// instead of kernel boundaries -- see the "layer chain" section).
#include <stdlib.h>
#include "conv_tc.cuh"

namespace srd {

struct TcLaunch {
  TcMaps m;
  TcArgs a;
};
template <int CG>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcLaunch lp) {
  const TcArgs& a = lp.a;
  const TcMaps& m_ = lp.m;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  kern_stamp(a, 0);
  kern_stamp(a, 6, true);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)a.a_stages * a.MT * a.a_stage_bytes;
  const uint32_t stg_base = b_base + (uint32_t)a.b_region_bytes;        // staging: [epilogue warp][buf][4 KB]
  const uint32_t stg_bytes = a.epi_tma ? (uint32_t)(EPI_WARPS * a.epi_nbuf) * 4096u : 0u;
  const uint32_t bar_base = stg_base + stg_bytes;
  // barrier layout (8 B each): A_full[8] A_empty[8] B_full[8] B_empty[8] T_full[2] T_empty[2]
  auto A_full = [&](int s) { return bar_base + 8u * s; };
  auto A_empty = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto B_full = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto B_empty = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + s); };
  auto T_full = [&](int s) { return bar_base + 8u * (4 * MAX_STAGES + s); };
  auto T_empty = [&](int s) { return bar_base + 8u * (4 * MAX_STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * MAX_STAGES + 8);
  const uint32_t B_res = bar_base + 8u * (4 * MAX_STAGES + 9);
  const uint32_t R_full0 = bar_base + 8u * (4 * MAX_STAGES + 10);       // [epilogue warp][2] -> 32 barriers
  uint8_t* gen_base = smem_raw + (bar_base - smem_u32(smem_raw));       // generic pointer to bar_base
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + 8 * (4 * MAX_STAGES + 8));
  float* vec = reinterpret_cast<float*>(gen_base + BAR_BYTES);
  float* s_bias = vec;                 // [512]: bias, then (second output) bias2 at [Npad, 2 Npad)
  float* s_temb = vec + 512;           // [512]
  float* s_head = vec + 1024;          // [4][256]
  float* s_headb = vec + 1024 + 1024;  // [4]
  float* s_hpart = nullptr;

  if (warp == WARP_TMA && lane == 0) {
    for (int s = 0; s < a.a_stages; ++s) { mbar_init(A_full(s), 1); mbar_init(A_empty(s), 1); }
    for (int s = 0; s < a.b_stages; ++s) { mbar_init(B_full(s), 1); mbar_init(B_empty(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(T_full(s), 1); mbar_init(T_empty(s), (uint32_t)(4 * a.n_share * CG)); }
    mbar_init(B_res, 1);
    for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(R_full0 + 8u * s, 1);
    for (int s = 0; s < SCHED_SLOTS; ++s) mbar_init(bar_base + SCHED_BAR_OFF + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < a.n_src; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmA[s]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmB) : "memory");
    if (a.epi_tma) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmO) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmOrow) : "memory");
      if (a.ep.res1) asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmR) : "memory");
      if (a.dual) asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmO2) : "memory");
    }
  }
  if (warp == WARP_MMA) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {      // issued by the same warp of both CTAs of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  // stage the per-channel epilogue vectors
  for (int i = threadIdx.x; i < a.Npad * a.nsplit; i += TC_THREADS) {
    s_bias[i] = (i < a.ep.N) ? __ldg(a.ep.bias + i) : 0.f;
    s_temb[i] = (a.ep.temb && a.ep.temb_bstride == 0 && i < a.ep.N) ? __ldg(a.ep.temb + i) : 0.f;
  }
  if (a.head_w) {
    for (int i = threadIdx.x; i < 4 * a.Npad; i += TC_THREADS) {
      const int jh = i / a.Npad, c = i % a.Npad;
      s_head[i] = (jh < a.ep.planar_c && c < a.ep.N) ? __ldg(a.head_w + (size_t)jh * a.ep.N + c) : 0.f;
    }
    if (threadIdx.x < 4) s_headb[threadIdx.x] = (threadIdx.x < a.ep.planar_c) ? __ldg(a.head_b + threadIdx.x) : 0.f;
  }
  if (a.dual) {       // second output: its bias after the first one, no time embedding
    for (int i = threadIdx.x; i < a.Npad; i += TC_THREADS) {
      s_bias[a.Npad + i] = (i < a.ep.N) ? __ldg(a.bias2 + i) : 0.f;
      s_temb[a.Npad + i] = 0.f;
    }
  }
  if (a.lin_w) {      // linear-residual weights, transposed to [4][Npad] (shares the head-weight storage)
    for (int i = threadIdx.x; i < 4 * a.Npad; i += TC_THREADS) {
      const int c = i / a.Npad, n = i % a.Npad;
      s_head[i] = (c < a.lin_c && n < a.ep.N) ? __ldg(a.lin_w + (size_t)n * a.lin_c + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // both CTAs' barriers are initialised before any remote arrive / TMA lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map and
  // epilogue-vector staging) touched only launch constants, so it may overlap the tail of the previous
  // layer.  From here on activations are read / written: wait for the previous grid -- except the weight
  // producer, whose loads are layer constants and may pre-fill the weight ring meanwhile.
  kern_stamp(a, 1);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp != WARP_TMB) asm volatile("griddepcontrol.wait;" ::: "memory");
  kern_stamp(a, 2);

  if (warp == WARP_TMA) {
    if (elect_one_sync()) producer_role<CG>(a, m_, a_base, bar_base);
  } else if (warp == WARP_TMB) {
    if (elect_one_sync()) weight_producer_role<CG>(a, m_, b_base, bar_base);
  } else if (warp == WARP_MMA || (SRD_DUAL_ISSUE && warp == WARP_MMA2)) {
    const int issuer = warp - WARP_MMA;
    if ((CG == 1 || cluster_ctarank() == 0) && elect_one_sync()) {
      // specialised issue loops: taps unrolled (descriptor offsets become immediates), MT and the weight
      // mode resolved at compile time
      const int taps = a.a_mode == 2 ? TAPS_S2 : (a.a_mode ? a.nkx * a.nky : 0);
#define SRD_MMA_CASE(T)                                                                      \
  do {                                                                                       \
    if (a.b_resident && a.MT == 2) mma_role<T, 2, true, CG>(a, tmem_base, a_base, b_base, bar_base, issuer); \
    else if (a.b_resident) mma_role<T, 1, true, CG>(a, tmem_base, a_base, b_base, bar_base, issuer); \
    else if (a.MT == 2) mma_role<T, 2, false, CG>(a, tmem_base, a_base, b_base, bar_base, issuer);   \
    else mma_role<T, 1, false, CG>(a, tmem_base, a_base, b_base, bar_base, issuer);                  \
  } while (0)
      if (a.dual) {
        if (a.b_resident && a.MT == 2) mma_role<9, 2, true, CG, true>(a, tmem_base, a_base, b_base, bar_base, issuer);
        else if (a.b_resident) mma_role<9, 1, true, CG, true>(a, tmem_base, a_base, b_base, bar_base, issuer);
        else if (a.MT == 2) mma_role<9, 2, false, CG, true>(a, tmem_base, a_base, b_base, bar_base, issuer);
        else mma_role<9, 1, false, CG, true>(a, tmem_base, a_base, b_base, bar_base, issuer);
      } else if (taps == 9 && a.nkx == 3) SRD_MMA_CASE(9);
      else if (taps == 4 && a.nkx == 2) SRD_MMA_CASE(4);
      else if (taps == 1) SRD_MMA_CASE(1);
      else if (taps == TAPS_S2) SRD_MMA_CASE(TAPS_S2);
      else SRD_MMA_CASE(0);
#undef SRD_MMA_CASE
    }
  } else {
    // ================================================================= epilogue (warps 2..9)
    EpiSmem es;
    const Sched sc = make_sched<CG>(a, bar_base);
    es.bias = s_bias; es.temb = s_temb;
    es.head_w = a.head_w ? s_head : nullptr; es.head_b = s_headb; es.hpart = s_hpart;
    es.lin_w = a.lin_w ? s_head : nullptr;
    if (a.epi_tma) {
      const int ew = warp - WARP_EPI0;                 // epilogue warp index
      StageSmem sg;
      sg.base = stg_base + (uint32_t)(ew * a.epi_nbuf) * 4096u;
      sg.gen = smem_raw + (sg.base - smem_u32(smem_raw));
      sg.r_full0 = R_full0 + 8u * (uint32_t)(ew * 2);
      if (a.ep.act == SRD_ACT_MISH)
        epilogue_tma<SRD_ACT_MISH, CG>(a, m_, es, sg, sc, tmem_base, T_full(0), T_empty(0), warp, lane);
      else
        epilogue_tma<SRD_ACT_NONE, CG>(a, m_, es, sg, sc, tmem_base, T_full(0), T_empty(0), warp, lane);
    } else if (a.ep.act == SRD_ACT_MISH)
      epilogue_role<SRD_ACT_MISH, CG>(a, es, sc, tmem_base, T_full(0), T_empty(0), warp, lane);
    else if (a.ep.act == SRD_ACT_LRELU02)
      epilogue_role<SRD_ACT_LRELU02, CG>(a, es, sc, tmem_base, T_full(0), T_empty(0), warp, lane);
    else
      epilogue_role<SRD_ACT_NONE, CG>(a, es, sc, tmem_base, T_full(0), T_empty(0), warp, lane);
  }

  kern_stamp(a, 3);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // the peer's shared memory / TMEM stay alive until the pair's MMAs are done
  kern_stamp(a, 4);
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  kern_stamp(a, 5);
  kern_stamp(a, 7, true);
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

// tuning / validation switches (srd_tc_set_mode): A-operand mode for stride-1 convs, resident weights
static int g_a_mode = 1;
// Stride-2 convs: halo boxes per (row parity, column parity) (1, default) or boxes per (kx, row parity) (0).  The two
// modes consume the weight k-blocks in different tap orders; engine.py packs for the mode srd_tc_s2_halo() reports.
int tc_s2_halo() {
  static const int v = [] { const char* e = getenv("SRD_S2_HALO"); return e ? (atoi(e) ? 1 : 0) : 1; }();
  return v;
}
static int g_b_resident = 1;
static int g_epi_tma = 1;
static int g_cg = 2;                     // CTAs per MMA: 2 = cta_group::2 pairs (default), 1 = single-CTA MMAs
static int g_pdl = 1;                    // programmatic dependent launch of consecutive conv layers
static int g_res_mt2 = 1;                // MT = 2 also with resident weights
static int g_small_mt2 = 0;              // MT = 2 with a single TMEM stage on small, wide layers (measured slower: off)
static int g_dyn_sched = 0;              // ticket-based item scheduling (see Sched; measured 2 % slower: off)
static unsigned long long* g_dbg = nullptr;
void tc_set_debug(unsigned long long* p) { g_dbg = p; }
void tc_set_mode(int a_mode, int b_resident) {
  g_a_mode = a_mode & 1;
  g_epi_tma = (a_mode & 2) ? 0 : 1;      // bit 1 of a_mode disables the TMA-staged epilogue
  g_cg = (a_mode & 4) ? 1 : 2;           // bit 2 of a_mode: single-CTA MMAs instead of CTA pairs
  g_pdl = (a_mode & 8) ? 0 : 1;          // bit 3 of a_mode: no programmatic dependent launch
  g_res_mt2 = (a_mode & 16) ? 0 : 1;     // bit 4 of a_mode: resident weights always with MT = 1
  g_small_mt2 = (a_mode & 32) ? 1 : 0;   // bit 5 of a_mode: trade the second TMEM stage for MT = 2 on small, wide layers
  g_dyn_sched = (a_mode & 64) ? 1 : 0;   // bit 6 of a_mode: ticket-based item scheduling instead of round-robin
  g_b_resident = b_resident;
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
  else if (d.head_w && d.lin_src) w = "fused head and linear residual are exclusive";
  else if (d.N2 && (d.N2 != d.N || (d.N % 64) || d.N > 256 || d.stride != 1 || d.up != 1 || d.nky != 3 || d.nkx != 3 ||
                    d.res1 || d.res2 || d.lin_src || d.head_w || d.out2 || d.planar || !d.out || (d.dst2_ct % 8) ||
                    (d.dst2_coff % 8)))
    w = "second output needs a plain 3x3 stride-1 conv with N2 == N in {64, 128, 192, 256}, one NHWC output, no residuals";
  else if (d.n_split != 0 && d.n_split != 1 &&
           (d.n_split != 2 || (d.N % 128) || d.stride != 1 || d.up != 1 || d.N2 || d.head_w || d.lin_src || d.planar || d.out2 ||
            !d.out || d.H_out < 4 || d.temb_bstride != 0 || (d.out_ct % 8) || (d.out_coff % 8) || d.act == SRD_ACT_LRELU02))
    w = "n_split = 2 needs a stride-1 conv with N % 128 == 0, one NHWC output through the TMA-staged epilogue";
  else {
    for (int i = 0; i < d.n_src; ++i) {
      const SrdSrc& s = d.src[i];
      if ((s.ct % 8) || (s.c_off % 8) || s.cin < 1 || ((uintptr_t)s.ptr % 16)) { w = "source alignment"; break; }
      if ((s.flags & SRD_SRC_WIN8) && (s.ct != 8 || s.c_off != 0 || s.cin != 64 || d.stride != 1 || d.up != 1 || d.nkx != 1)) {
        w = "SRD_SRC_WIN8 needs ct == 8, c_off == 0, cin == 64, a plain stride-1 conv and nkx == 1";
        break;
      }
    }
  }
  if (why) *why = w;
  return w == nullptr;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: set it once per device ordinal, not once
// per process (a second GPU in the same process would otherwise fail every launch that needs > 48 KB).
static int tc_set_func_attrs(int dev);


int tc_plan_build(const SrdConvDesc& d, TcPlan* plan, int* err_flag) {
  const char* why = nullptr;
  SRD_CHECK(tc_supported(d, &why), "conv_tc: unsupported descriptor: %s", why ? why : "?");
  EncodeTiledFn enc = get_encode_fn();
  SRD_CHECK(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled not available (no CUDA driver?)");
  TcArgs& a = plan->args;
  TcMaps& mp = plan->maps;
  memset(&a, 0, sizeof(a));
  memset(&mp, 0, sizeof(mp));
  const int Hp = d.H_in + 2, Wp = d.W_in + 2;
  a.n_src = d.n_src;
  a.stride = d.stride; a.nky = d.nky; a.nkx = d.nkx; a.ky0 = d.ky0; a.kx0 = d.kx0; a.up = d.up;
  a.H_m = d.H_in / d.stride; a.W_m = d.W_in / d.stride;
  a.nsplit = d.n_split == 2 ? 2 : 1;
  a.Npad = (d.N / a.nsplit + 15) / 16 * 16;        // output channels (padded) of ONE channel group
  if (d.stride == 1) { a.period = Hp; a.voff = 1; a.vstart = 1; }
  else { a.period = Hp / 2; a.voff = 0; a.vstart = 0; }
  a.n_vrows = d.B * a.period;
  const int n_rowblk = (a.n_vrows - a.vstart + TILE_H - 1) / TILE_H;
  a.n_colblk = (a.W_m + TILE_W - 1) / TILE_W;
  a.tiles_per_phase = n_rowblk * a.n_colblk;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int n_phase = d.up * d.up * a.nsplit;      // sub-pixel phases x channel groups: "phases" of the item space
  int kb = 0;
  const int taps = d.nky * d.nkx;
  a.dual = d.N2 ? 1 : 0;
  a.bias2 = d.bias2;
  for (int i = 0; i < d.n_src; ++i) { a.nchunk[i] = (d.src[i].cin + 63) / 64; kb += a.nchunk[i] * (taps + a.dual); }
  a.kb_per_phase = kb;
  a.kb_total = n_phase * kb;
  const int cg = (n_sm % 2 == 0) ? g_cg : 1;
  a.cg = cg;
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.Npad >> 3) << 17) | ((uint32_t)((128 * cg) >> 4) << 24);
  a.ep = make_epi(d, nullptr);
  a.head_w = d.head_w; a.head_b = d.head_b;
  a.lin_src = static_cast<const __nv_bfloat16*>(d.lin_src); a.lin_w = d.lin_src ? d.lin_w : nullptr;
  a.lin_ct = d.lin_ct; a.lin_c = d.lin_c;
  a.err_flag = err_flag;
  a.dbg = g_dbg;
  {
    auto ok8 = [](const void* p, int ct, int coff) { return !p || ((ct % 8) == 0 && (coff % 8) == 0); };
    a.fast = (d.N % 16 == 0) && d.temb_bstride == 0 &&
             ok8(d.out, d.out_ct, d.out_coff) && ok8(d.out2, d.out2_ct, d.out2_coff) &&
             ok8(d.res1, d.res1_ct, d.res1_coff) && ok8(d.res2, d.res2_ct, d.res2_coff) &&
             (!d.planar || d.head_w) && (d.act != SRD_ACT_MISH || d.alpha == 1.0f);
  }
  a.epi_geo = (d.stride == 1 && d.up == 1) ? 0 : 1;
  a.epi_tma = g_epi_tma && a.fast && (d.up == 1 || (d.up == 2 && d.stride == 1)) && (d.N % 64 == 0) && d.out &&
              !d.out2 && !d.planar && !d.head_w && d.act != SRD_ACT_LRELU02 && d.H_out >= 4 &&
              (a.epi_geo == 0 || !d.res1) && (d.up == 1 || a.W_m % TILE_W == 0);
  // ---- shared-memory plan: A ring + weights (ring or resident) + staging + barriers + epilogue vectors
  const int B_BYTES = a.Npad * 128 / cg;      // per CTA: each CTA of a pair holds Npad / cg rows of a weight k-block
  const int budget = 227 * 1024 - BAR_BYTES - VEC_BYTES - 1024;
  a.a_mode = (d.stride == 1) ? g_a_mode : (tc_s2_halo() ? 2 : 0);
  a.a_stage_bytes = a.a_mode == 2 ? A_S2_STAGE : (a.a_mode ? A_HALO_STAGE : A_BYTES);
  const long long res_bytes = (long long)a.kb_total * B_BYTES;
  const long long n_tiles_all = (long long)n_phase * a.tiles_per_phase;
  // MT = 2 halves the weight traffic of an item (two M tiles share every weight k-block).  Normally it needs
  // two TMEM stages (2 x 2 x Npad <= 512 columns) and enough tiles for every SM; small layers (at most ~2
  // item rounds, e.g. the 20x20 level) also take it with ONE stage: they are bound by streaming the whole
  // weight tensor from L2 once per item, and with one round there is no next item to overlap with anyway.
  const int acc_cols = a.Npad * (1 + a.dual);          // TMEM columns of one M tile's accumulator(s)
  const bool small_wide = g_small_mt2 && (2 * acc_cols <= 512) && (2 * 2 * acc_cols > 512) &&
                          n_tiles_all >= n_sm && n_tiles_all <= 2LL * n_sm;
  const bool mt2_ok = ((2 * 2 * acc_cols <= 512) && n_tiles_all >= 2LL * n_sm) || small_wide;
  const int nbuf_max = EPI_WARPS <= 8 ? 2 : 1;   // 4 KB staging buffers per epilogue warp
  bool planned = false;
  int b_bytes_total = 0, stage_bytes = 0;
  // preference: staging depth 2 before 1; resident weights (MT = 1: nothing is re-streamed) before a ring;
  // for rings MT = 2 (each weight tile shared by two M tiles) before MT = 1.
  for (int nbuf = a.epi_tma ? nbuf_max : 0; !planned; --nbuf) {
    const int stg = EPI_WARPS * nbuf * 4096;
    // passes: resident weights with MT = 2 (large layers: per-item overheads of the MMA thread and of the
    // epilogue hand-over amortise over twice the work), resident MT = 1, ring MT = 2, ring MT = 1
    for (int pass = -1; pass < 3 && !planned; ++pass) {
      const bool resident = (pass <= 0);
      const int MT = (pass == -1 || pass == 1) ? 2 : 1;
      if (pass == -1 && !(g_res_mt2 && 4 * 2 * acc_cols <= 512 && n_tiles_all >= 16LL * n_sm)) continue;
      if (resident && !(g_b_resident && res_bytes < (1 << 20))) continue;
      if (pass == 1 && !mt2_ok) continue;
      const int sb = MT * a.a_stage_bytes;
      if (resident) {
        const long long left = (long long)budget - stg - res_bytes;
        if (left < 2LL * sb) continue;
        const int ast = (int)(left / sb);
        const int ast_max = a.a_mode == 2 ? 6 : 4;         // stride-2 halo: four boxes per chunk
        a.a_stages = ast > ast_max ? ast_max : ast;
        a.b_stages = 1;
        b_bytes_total = (int)res_bytes;
      } else {
        // deepest A ring that still leaves >= 4 weight stages; at least 2 (halo) / 3 (per-kx) A stages and 3 weight stages
        const int ast_min = a.a_mode == 1 ? 2 : 3;
        int ast = 4;
        for (; ast > ast_min; --ast)
          if (budget - stg - ast * sb >= 4 * B_BYTES) break;
        const int left = budget - stg - ast * sb;
        if (left < 3 * B_BYTES) continue;
        const int bst = left / B_BYTES;
        a.a_stages = ast;
        a.b_stages = bst > MAX_STAGES ? MAX_STAGES : bst;
        b_bytes_total = a.b_stages * B_BYTES;
      }
      a.b_resident = resident ? 1 : 0;
      a.MT = MT;
      a.epi_nbuf = nbuf;
      stage_bytes = sb;
      planned = true;
    }
    if (planned) break;
    if (nbuf <= (a.epi_tma ? 1 : 0)) {
      if (!a.epi_tma) break;
      a.epi_tma = 0;              // staging does not fit: register epilogue
      nbuf = 1;                   // loop decrements to 0
    }
  }
  SRD_CHECK(planned, "conv_tc: no shared-memory plan fits (N=%d)", d.N);
  a.items_per_phase = (a.tiles_per_phase + a.MT * cg - 1) / (a.MT * cg);
  a.n_items = n_phase * a.items_per_phase;
  SRD_CHECK(a.n_items < (1 << 24) && a.n_vrows < (1 << 24), "conv_tc: problem too large for the 24-bit tile decode");
  a.fd_ipp = make_fastdiv((uint32_t)a.items_per_phase);
  a.fd_colblk = make_fastdiv((uint32_t)a.n_colblk);
  a.fd_period = make_fastdiv((uint32_t)a.period);
  a.fd_up = make_fastdiv((uint32_t)d.up);
  a.n_chunks = 0;
  for (int i = 0; i < d.n_src; ++i) a.n_chunks += a.nchunk[i];
  if (SRD_DUAL_ISSUE && a.b_resident && a.a_mode == 1 && a.n_chunks == 1 && a.a_stages == 3) a.a_stages = 2;

  a.b_region_bytes = b_bytes_total;
  const int stg_total = a.epi_tma ? EPI_WARPS * a.epi_nbuf * 4096 : 0;
  // TMEM accumulator stages, and how the epilogue groups are dealt: with >= 3 (2) sub-tiles per item all
  // (two) groups split one item, so an accumulator stage is drained -- and free for the next MMAs -- in
  // the time of ONE sub-tile; otherwise every group drains its own item.
  a.n_stages = (EPI_GROUPS >= 4 && 4 * a.MT * acc_cols <= 512) ? 4 : (2 * a.MT * acc_cols <= 512 ? 2 : 1);
  a.stage_shift = a.n_stages == 4 ? 2 : (a.n_stages == 2 ? 1 : 0);
  a.dual_issue = (SRD_DUAL_ISSUE && a.b_resident && a.a_mode == 1 && a.n_chunks == 1 && a.a_stages % 2 == 0 &&
                  a.n_stages % 2 == 0) ? 1 : 0;
  {
    const int n_sub = a.MT * ((a.Npad + 63) / 64) * (1 + a.dual);
    int share = d.head_w ? 1 : (n_sub >= 3 ? 4 : n_sub);
    if (share > EPI_GROUPS) share = EPI_GROUPS;
    while (EPI_GROUPS / share > a.n_stages) share *= 2;
    a.n_share = share;
  }
  SRD_CHECK(!d.head_w || a.n_share == 1, "conv_tc: fused head needs N * MT <= 128");
  SRD_CHECK(!d.N2 || (a.epi_tma && a.a_mode == 1 && a.epi_geo == 0),
            "conv_tc: the second output needs the halo A mode and the TMA-staged epilogue");
  SRD_CHECK(a.nsplit == 1 || (a.epi_tma && a.a_mode == 1 && a.epi_geo == 0),
            "conv_tc: n_split needs the halo A mode and the TMA-staged epilogue");
  SRD_CHECK(!d.lin_src || a.epi_tma, "conv_tc: the linear residual needs the TMA-staged epilogue "
            "(stride-1 conv, N %% 64 == 0, NHWC output, batch-uniform temb)");
  plan->smem = a.a_stages * stage_bytes + b_bytes_total + stg_total + BAR_BYTES + VEC_BYTES + 1024;
  plan->grid = (a.n_items < n_sm / cg ? a.n_items : n_sm / cg) * cg;
  // dynamic item scheduling (see Sched) once a pair has more than one item
  a.dyn_sched = (g_dyn_sched && a.n_items > 2 * (plan->grid / cg)) ? 1 : 0;
  a.reverse_items = (d.reverse_items && !a.dyn_sched) ? 1 : 0;
  if (a.dyn_sched) {
    if (!plan->counter) SRD_CUDA(cudaMalloc((void**)&plan->counter, sizeof(unsigned)));
    SRD_CUDA(cudaMemset(plan->counter, 0, sizeof(unsigned)));
    a.sched_counter = plan->counter;
  }

  // ---- tensor maps
  for (int i = 0; i < d.n_src; ++i) {
    const SrdSrc& s = d.src[i];
    void* base = (void*)((const __nv_bfloat16*)s.ptr + s.c_off);
    CUresult r;
    if (s.flags & SRD_SRC_WIN8) {
      // overlapping windows: element (k, x, row) = buffer[row][x + k / 8][k % 8]; consecutive x positions are
      // one pixel (16 B) apart while the 64-element inner box spans 8 pixels (128 B)
      SRD_CHECK(a.a_mode == 1, "conv_tc: SRD_SRC_WIN8 needs the halo A mode");
      cuuint64_t dims[3] = {64, (cuuint64_t)Wp, (cuuint64_t)d.B * Hp};
      cuuint64_t strides[2] = {16, (cuuint64_t)Wp * 16};
      cuuint32_t box[3] = {64, (cuuint32_t)HALO_W, BOX_ROWS};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&mp.tmA[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)s.ptr, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (d.stride == 1) {
      cuuint64_t dims[3] = {(cuuint64_t)s.cin, (cuuint64_t)Wp, (cuuint64_t)d.B * Hp};
      cuuint64_t strides[2] = {(cuuint64_t)s.ct * 2, (cuuint64_t)Wp * s.ct * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)(a.a_mode ? HALO_W : TILE_W), BOX_ROWS};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&mp.tmA[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      // [rows/2][row parity][cols/2][col parity][c]
      cuuint64_t dims[5] = {(cuuint64_t)s.cin, 2, (cuuint64_t)Wp / 2, 2, (cuuint64_t)d.B * Hp / 2};
      cuuint64_t strides[4] = {(cuuint64_t)s.ct * 2, (cuuint64_t)2 * s.ct * 2, (cuuint64_t)Wp * s.ct * 2,
                               (cuuint64_t)2 * Wp * s.ct * 2};
      cuuint32_t box[5] = {64, 1, TILE_W, 1, BOX_ROWS};
      if (a.a_mode == 2) { box[2] = S2_W; box[4] = S2_ROWS; }
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      r = enc(&mp.tmA[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A%d) failed with %d", i, (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)n_phase * kb * a.Npad};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(a.Npad / cg)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d.w_tc, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  if (a.epi_tma) {
    const int Hop = d.H_out + 2, Wop = d.W_out + 2;
    // The column dimension stops at padded column W_out (the right ring column is excluded): boxes of
    // the last, ragged column tile are clipped by the TMA unit instead of spilling into the ring.
    auto enc_out = [&](CUtensorMap* tm, const void* ptr, int ct, int coff, int rows) {
      void* base = (void*)((const __nv_bfloat16*)ptr + coff);
      cuuint64_t dims[3] = {(cuuint64_t)d.N, (cuuint64_t)Wop - 1, (cuuint64_t)d.B * Hop};
      cuuint64_t strides[2] = {(cuuint64_t)ct * 2, (cuuint64_t)Wop * ct * 2};
      cuuint32_t box[3] = {64, TILE_W, (cuuint32_t)rows};
      cuuint32_t es[3] = {1, 1, 1};
      return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    // 2x transposed conv: [c][col parity][col pair][row parity][row pair]; one phase writes one parity
    // plane.  The right ring column shares its pair with the last real column, so it cannot be clipped
    // away here: this map is only used when W_m % 8 == 0 (no ragged column tile).
    auto enc_out5 = [&](CUtensorMap* tm, const void* ptr, int ct, int coff, int rows) {
      void* base = (void*)((const __nv_bfloat16*)ptr + coff);
      cuuint64_t dims[5] = {(cuuint64_t)d.N, 2, (cuuint64_t)Wop / 2, 2, (cuuint64_t)d.B * Hop / 2};
      cuuint64_t strides[4] = {(cuuint64_t)ct * 2, (cuuint64_t)2 * ct * 2, (cuuint64_t)Wop * ct * 2,
                               (cuuint64_t)2 * Wop * ct * 2};
      cuuint32_t box[5] = {64, 1, TILE_W, 1, (cuuint32_t)rows};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    SRD_CHECK(((uintptr_t)d.out % 16) == 0, "conv_tc: out must be 16-byte aligned");
    CUresult r = (d.up == 2) ? enc_out5(&mp.tmO, d.out, d.out_ct, d.out_coff, 4) : enc_out(&mp.tmO, d.out, d.out_ct, d.out_coff, 4);
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    r = (d.up == 2) ? enc_out5(&mp.tmOrow, d.out, d.out_ct, d.out_coff, 1) : enc_out(&mp.tmOrow, d.out, d.out_ct, d.out_coff, 1);
    SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(out row) failed with %d", (int)r);
    if (d.res1) {
      r = enc_out(&mp.tmR, d.res1, d.res1_ct, d.res1_coff, 4);
      SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(res1) failed with %d", (int)r);
    }
    if (d.N2) {
      SRD_CHECK(((uintptr_t)d.dst2 % 16) == 0, "conv_tc: dst2 must be 16-byte aligned");
      r = enc_out(&mp.tmO2, d.dst2, d.dst2_ct, d.dst2_coff, 4);
      SRD_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(dst2) failed with %d", (int)r);
    }
  }
  SRD_CHECK(tc_set_func_attrs(dev) == 0, "conv_tc: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
  return 0;
}

static int tc_set_func_attrs(int dev) {
  static bool done[64] = {};
  if (dev < 0 || dev >= 64) return 1;
  if (done[dev]) return 0;
  if (cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return 1;
  done[dev] = true;
  return 0;
}

TcPlan* tc_plan_new() { return new (std::nothrow) TcPlan; }
void tc_plan_delete(TcPlan* p) { delete p; }

int launch_conv_tc(const TcPlan& plan, const float* temb, cudaStream_t stream) {
  TcLaunch lp;
  lp.m = plan.maps;
  lp.a = plan.args;
  TcArgs& a = lp.a;
  if (temb) a.ep.temb = temb;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  if (g_pdl) {      // this launch may begin while the previous kernel on the stream drains (see griddepcontrol.wait)
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  if (a.cg == 2) {
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = 2; at[n_at].val.clusterDim.y = 1; at[n_at].val.clusterDim.z = 1;
    ++n_at;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n_at;
  if (a.cg == 1) SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<1>, lp));
  else SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<2>, lp));
  return 0;
}

int tc_pdl_enabled() { return g_pdl; }

}  // namespace srd
