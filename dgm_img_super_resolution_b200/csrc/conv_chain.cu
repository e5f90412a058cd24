// Multi-layer launches of the tcgen05 convolution: conv_chain_kernel (a whole layer list, run-time
// layer index) and conv_seg_kernel<K> (up to K same-plan layers, compile-time unrolled), plus their host planner.
// The roles themselves live in conv_tc.cuh and are shared with the single-layer kernel (conv_tc.cu).
#include "conv_tc.cuh"

namespace srd {

// ------------------------------------------------------------------------------------------ layer chain kernel
// Shared-memory layout of the chain kernel: a fixed head (barriers, a DOUBLE-buffered bias / time-embedding vector
// pair so the epilogue warps can stage layer l+1's vectors while they run layer l, the fused-head / linear-residual
// weights) followed by the per-segment A | B | staging regions of the ordinary plan.
constexpr int CHAIN_BT_OFF = BAR_BYTES;                        // 2 x (bias[256], temb[256]) floats
constexpr int CHAIN_BT_BYTES = 2 * 256 * 4;
constexpr int CHAIN_HEAD_OFF = CHAIN_BT_OFF + 2 * CHAIN_BT_BYTES;   // [4][128] + [4] floats (head / lin layers: Npad <= 128)
constexpr int CHAIN_HEAD_FLOATS = 4 * 128;
// Shared-memory copies of the CURRENT layer's arguments, one per role (the three single-thread roles and the
// epilogue warps, which double-buffer theirs like the vectors): the layer list lives in parameter space, where a
// run-time layer index means register-indexed LDC loads -- cheap once per layer for a copy, far too slow for the role
// loops, which in the single-layer kernel read their arguments as constant-bank operands.
constexpr int CHAIN_ARG_SLOT = ((int)sizeof(ChainLayer) + 15) / 16 * 16;
constexpr int CHAIN_ARG_OFF = (CHAIN_HEAD_OFF + (CHAIN_HEAD_FLOATS + 4) * 4 + 15) / 16 * 16;
constexpr int CHAIN_ARG_TMA = 0, CHAIN_ARG_TMB = 1, CHAIN_ARG_MMA = 2, CHAIN_ARG_EPI = 3;   // EPI: two slots
constexpr int CHAIN_FIXED_BYTES = 11264;
static_assert(CHAIN_ARG_OFF + 5 * CHAIN_ARG_SLOT <= CHAIN_FIXED_BYTES, "chain head does not fit");
static_assert(CHAIN_FIXED_BYTES <= BAR_BYTES + VEC_BYTES && CHAIN_FIXED_BYTES % 1024 == 0,
              "chain head larger than the single-layer kernel's tail");

// bias / temb (+ the second output's bias, zero temb) of one layer -> bt[0..511]; head / linear-residual weights ->
// s_head (only one layer of a segment may use those: see same_segment)
// All layers' scalar arguments travel as ONE kernel parameter (CUDA 12.1+: up to 32,764 bytes), i.e. in the constant
// bank, exactly where the single-layer kernel finds its TcArgs: the role loops read them as instruction operands /
// uniform loads instead of global-memory loads that every asm memory clobber would force to be repeated.
constexpr int CHAIN_MAX_LAYERS = (32764 - 32) / (int)sizeof(ChainLayer);
struct ChainParams {
  const TcMaps* maps;      // [n_layers] in global memory
  const float* temb_row;
  int n_layers, pad_;
  ChainLayer layers[CHAIN_MAX_LAYERS];
};
static_assert(sizeof(ChainParams) <= 32764, "chain parameters exceed the kernel parameter space");

__device__ __forceinline__ void chain_stage_vecs(const ChainLayer& L, const float* temb_row, float* bt, float* s_head,
                                                 int tid, int nthr) {
  const TcArgs& a = L.a;
  float* s_bias = bt;
  float* s_temb = bt + 256;
  const float* temb = L.temb_off >= 0 ? temb_row + L.temb_off : a.ep.temb;
  const int Npad = a.Npad, N = a.ep.N;
  for (int i = tid; i < Npad * a.nsplit; i += nthr) {
    s_bias[i] = (i < N) ? __ldg(a.ep.bias + i) : 0.f;
    s_temb[i] = (temb && a.ep.temb_bstride == 0 && i < N) ? __ldg(temb + i) : 0.f;
  }
  if (a.dual) {
    for (int i = tid; i < Npad; i += nthr) {
      s_bias[Npad + i] = (i < N) ? __ldg(a.bias2 + i) : 0.f;
      s_temb[Npad + i] = 0.f;
    }
  }
  if (a.head_w) {
    for (int i = tid; i < 4 * Npad; i += nthr) {
      const int jh = i / Npad, c = i % Npad;
      s_head[i] = (jh < a.ep.planar_c && c < N) ? __ldg(a.head_w + (size_t)jh * N + c) : 0.f;
    }
    if (tid < 4) s_head[CHAIN_HEAD_FLOATS + tid] = (tid < a.ep.planar_c) ? __ldg(a.head_b + tid) : 0.f;
  }
  if (a.lin_w) {
    for (int i = tid; i < 4 * Npad; i += nthr) {
      const int c = i / Npad, n = i % Npad;
      s_head[i] = (c < a.lin_c && n < N) ? __ldg(a.lin_w + (size_t)n * a.lin_c + c) : 0.f;
    }
  }
}

#ifndef SRD_CHAIN_ARGS_SMEM
#define SRD_CHAIN_ARGS_SMEM 0      // 1: roles read their layer's arguments from shared-memory copies (measured slower)
#endif
// word-wise copy of one layer's arguments from parameter space into a role's shared-memory slot
__device__ __forceinline__ void chain_copy_args(const ChainLayer& src, ChainLayer* dst, int tid, int nthr) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int i = tid; i < (int)(sizeof(ChainLayer) / 4); i += nthr) d[i] = s[i];
}

template <int CG>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainParams P) {
  const ChainLayer* const layers = P.layers;
  const TcMaps* const maps = P.maps;
  const int n_layers = P.n_layers;
  const float* const temb_row = P.temb_row;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen0 = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base;
  const Bars bars{bar_base};
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen0 + 8 * (4 * MAX_STAGES + 8));
  float* const bt0 = reinterpret_cast<float*>(gen0 + CHAIN_BT_OFF);
  float* const s_head = reinterpret_cast<float*>(gen0 + CHAIN_HEAD_OFF);
  auto arg_slot = [&](int i) { return reinterpret_cast<ChainLayer*>(gen0 + CHAIN_ARG_OFF + i * CHAIN_ARG_SLOT); };
  const uint32_t a_base = smem_base + CHAIN_FIXED_BYTES;

  if (warp == WARP_MMA) {       // the 512 TMEM columns are allocated once for the whole chain
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars.tmem_slot()), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars.tmem_slot()), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  const PairId pid = pair_id<CG>();
  ChainRt cx;
  cx.temb_row = temb_row;
  cx.exp_mult = 1u;
  uint32_t tmem_base = 0;
  bool first = true;
  int l0 = 0;
  while (l0 < n_layers) {
    // ---- a SEGMENT = consecutive layers with the same shared-memory / TMEM / barrier plan: inside it every role runs
    // from layer to layer without ever meeting the others; between segments the CTA pair drains and re-plans.
    int l1 = l0 + 1;
    while (l1 < n_layers && !layers[l1].seg_first) ++l1;
    const TcArgs& a0 = layers[l0].a;
    const uint32_t b_base = a_base + (uint32_t)a0.a_stages * a0.MT * a0.a_stage_bytes;
    const uint32_t stg_base = b_base + (uint32_t)a0.b_region_bytes;
    if (warp == WARP_TMA && lane == 0) {
      // the same set of barriers is (re)built for every segment; a live mbarrier object must be invalidated before
      // mbarrier.init may be applied to its location again
      if (!first) {
        for (int s = 0; s < MAX_STAGES; ++s) {
          mbar_inval(bars.A_full(s)); mbar_inval(bars.A_empty(s)); mbar_inval(bars.B_full(s)); mbar_inval(bars.B_empty(s));
        }
        for (int s = 0; s < 4; ++s) { mbar_inval(bars.T_full(s)); mbar_inval(bars.T_empty(s)); }
        mbar_inval(bars.B_res());
        mbar_inval(bars.M_done());
        for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_inval(bars.R_full0() + 8u * s);
      }
      for (int s = 0; s < MAX_STAGES; ++s) {
        mbar_init(bars.A_full(s), 1); mbar_init(bars.A_empty(s), 1); mbar_init(bars.B_full(s), 1); mbar_init(bars.B_empty(s), 1);
      }
      for (int s = 0; s < 4; ++s) { mbar_init(bars.T_full(s), 1); mbar_init(bars.T_empty(s), (uint32_t)(4 * a0.n_share * CG)); }
      mbar_init(bars.B_res(), 1);
      mbar_init(bars.M_done(), 1);
      for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(bars.R_full0() + 8u * s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    chain_stage_vecs(layers[l0], temb_row, bt0, s_head, (int)threadIdx.x, TC_THREADS);
    if (SRD_CHAIN_ARGS_SMEM && warp >= WARP_EPI0 && warp < WARP_EPI0 + EPI_WARPS)
      chain_copy_args(layers[l0], arg_slot(CHAIN_ARG_EPI), (int)threadIdx.x - WARP_EPI0 * 32, EPI_WARPS * 32);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    if (first) {
      tmem_base = *tmem_slot_ptr;
      // everything so far touched only launch constants; activations follow (see conv_tc_kernel)
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      if (warp != WARP_TMB) asm volatile("griddepcontrol.wait;" ::: "memory");
      first = false;
    }
    cx.it_base = 0; cx.k = 0; cx.n_seg = l1 - l0;
    cx.ra = Ring{0, 0, a0.a_stages}; cx.rb = Ring{0, 0, a0.b_stages}; cx.rt = Ring{0, 0, a0.n_stages};

    if (warp == WARP_TMA) {
      for (int l = l0; l < l1; ++l) {
#if SRD_CHAIN_ARGS_SMEM
        ChainLayer* const mine = arg_slot(CHAIN_ARG_TMA);
        __syncwarp();
        chain_copy_args(layers[l], mine, lane, 32);
        __syncwarp();
#else
        const ChainLayer* const mine = &layers[l];
#endif
        if (elect_one_sync()) {
          const TcArgs& a = mine->a;
          const TcMaps& m_ = maps[l];
          cx.j = l - l0;
          for (int s = 0; s < a.n_src; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmA[s]) : "memory");
          producer_role<CG, true>(a, m_, a_base, bar_base, &cx, mine);
        }
      }
    } else if (warp == WARP_TMB) {
      for (int l = l0; l < l1; ++l) {
#if SRD_CHAIN_ARGS_SMEM
        ChainLayer* const mine = arg_slot(CHAIN_ARG_TMB);
        __syncwarp();
        chain_copy_args(layers[l], mine, lane, 32);
        __syncwarp();
#else
        const ChainLayer* const mine = &layers[l];
#endif
        if (elect_one_sync()) {
          const TcArgs& a = mine->a;
          const TcMaps& m_ = maps[l];
          cx.j = l - l0;
          asm volatile("prefetch.tensormap [%0];" ::"l"(&m_.tmB) : "memory");
          weight_producer_role<CG, true>(a, m_, b_base, bar_base, &cx);
        }
      }
    } else if (warp == WARP_MMA) {
      for (int l = l0; l < l1; ++l) {
#if SRD_CHAIN_ARGS_SMEM
        ChainLayer* const mine = arg_slot(CHAIN_ARG_MMA);
        __syncwarp();
        chain_copy_args(layers[l], mine, lane, 32);
        __syncwarp();
#else
        const ChainLayer* const mine = &layers[l];
#endif
        if ((CG == 1 || pid.rank == 0) && elect_one_sync()) {
          const TcArgs& a = mine->a;
          cx.j = l - l0;
          const int taps = a.a_mode == 2 ? TAPS_S2 : (a.a_mode ? a.nkx * a.nky : 0);
#define SRD_MMA_CASE(T, D)                                                                                             \
  do {                                                                                                                 \
    if (a.b_resident && a.MT == 2) mma_role<T, 2, true, CG, D, true>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);  \
    else if (a.b_resident) mma_role<T, 1, true, CG, D, true>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);          \
    else if (a.MT == 2) mma_role<T, 2, false, CG, D, true>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);            \
    else mma_role<T, 1, false, CG, D, true>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);                           \
  } while (0)
          if (a.dual) SRD_MMA_CASE(9, true);
          else if (taps == 9 && a.nkx == 3) SRD_MMA_CASE(9, false);
          else if (taps == 4 && a.nkx == 2) SRD_MMA_CASE(4, false);
          else if (taps == 1) SRD_MMA_CASE(1, false);
          else if (taps == TAPS_S2) SRD_MMA_CASE(TAPS_S2, false);
          else SRD_MMA_CASE(0, false);
#undef SRD_MMA_CASE
        }
      }
    } else if (warp >= WARP_EPI0 && warp < WARP_EPI0 + EPI_WARPS) {
      const int etid = (int)threadIdx.x - WARP_EPI0 * 32;
      const int ew = warp - WARP_EPI0;
      for (int l = l0; l < l1; ++l) {
        const int j = l - l0;
#if SRD_CHAIN_ARGS_SMEM
        const ChainLayer* const Lc = arg_slot(CHAIN_ARG_EPI + (j & 1));
#else
        const ChainLayer* const Lc = &layers[l];
#endif
        const TcArgs& a = Lc->a;
        const TcMaps& m_ = maps[l];
        cx.j = j;
        // every epilogue warp has left layer l-1: its vector / argument buffers may be refilled, and layer l's
        // (staged while layer l-1 ran) are visible to all
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
        if (l + 1 < l1) {
          chain_stage_vecs(layers[l + 1], temb_row, bt0 + ((j + 1) & 1) * 512, s_head, etid, EPI_WARPS * 32);
          if (SRD_CHAIN_ARGS_SMEM) chain_copy_args(layers[l + 1], arg_slot(CHAIN_ARG_EPI + ((j + 1) & 1)), etid, EPI_WARPS * 32);
        }
        EpiSmem es;
        const Sched sc = make_sched<CG>(a, bar_base);
        es.bias = bt0 + (j & 1) * 512; es.temb = es.bias + 256;
        es.head_w = a.head_w ? s_head : nullptr; es.head_b = s_head + CHAIN_HEAD_FLOATS; es.hpart = nullptr;
        es.lin_w = a.lin_w ? s_head : nullptr;
        if (a.epi_tma) {
          StageSmem sg;
          sg.base = stg_base + (uint32_t)(ew * a.epi_nbuf) * 4096u;
          sg.gen = smem_raw + (sg.base - smem_u32(smem_raw));
          sg.r_full0 = bars.R_full0() + 8u * (uint32_t)(ew * 2);
          if (a.ep.act == SRD_ACT_MISH)
            epilogue_tma<SRD_ACT_MISH, CG, true>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, Lc);
          else
            epilogue_tma<SRD_ACT_NONE, CG, true>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, Lc);
        } else if (a.ep.act == SRD_ACT_MISH)
          epilogue_role<SRD_ACT_MISH, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, Lc);
        else if (a.ep.act == SRD_ACT_LRELU02)
          epilogue_role<SRD_ACT_LRELU02, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, Lc);
        else
          epilogue_role<SRD_ACT_NONE, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, Lc);
        cx.it_base += (pid.pair < a.n_items) ? (uint32_t)((a.n_items - 1 - pid.pair) / pid.n_pairs + 1) : 0u;
      }
    }
    // segment end: all roles of the pair are done (stores issued and read; completion was awaited when publishing)
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    l0 = l1;
  }
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ fused segment kernel
// conv_seg_kernel<CG, K>: up to K consecutive layers that share one shared-memory / TMEM / barrier plan (one
// "segment" of the chain planner) in ONE launch.  Same role code and the same completion-counter protocol as
// conv_chain_kernel, but the layer loop is unrolled at compile time, so every layer's arguments sit at a fixed offset
// of the kernel parameter block: the role loops keep their constant-bank operands and uniform-datapath control flow
// (what the run-time-indexed chain kernel loses).  The counters are never reset: a device-side launch epoch scales the
// expected values instead, so no memset node sits in front of the launch and programmatic dependent launch still
// overlaps its prologue with the previous layer's tail.
constexpr unsigned SEG_EPOCH_MAX = 1u << 20;
#ifndef SRD_SEG_K
#define SRD_SEG_K 2
#endif
constexpr int SEG_K = SRD_SEG_K;       // layers one fused-segment launch can hold (the role code is instantiated K times)
template <int K>
struct SegParams {
  const TcMaps* maps;      // [n] in global memory
  const float* temb_row;
  unsigned* epoch;         // completed launches of this segment since the counters were last zeroed
  unsigned* done;          // CTAs that have finished the current launch
  unsigned* counters;
  int n_counters, n;
  ChainLayer layers[K];
};

#ifndef SRD_SEG_PLAIN_ROLES
#define SRD_SEG_PLAIN_ROLES 0     // diagnosis (single-layer segments only): bit 0 producer, 1 weights, 2 MMA, 3 epilogue
#endif                            // run the NON-chain instantiation of that role inside the fused-segment kernel
template <int CG, int K>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_seg_kernel(const __grid_constant__ SegParams<K> P) {
  constexpr bool CH_P = !(SRD_SEG_PLAIN_ROLES & 1), CH_W = !(SRD_SEG_PLAIN_ROLES & 2), CH_M = !(SRD_SEG_PLAIN_ROLES & 4),
                 CH_E = !(SRD_SEG_PLAIN_ROLES & 8);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen0 = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base;
  const Bars bars{bar_base};
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen0 + 8 * (4 * MAX_STAGES + 8));
  float* const bt0 = reinterpret_cast<float*>(gen0 + CHAIN_BT_OFF);
  float* const s_head = reinterpret_cast<float*>(gen0 + CHAIN_HEAD_OFF);
  const TcArgs& a0 = P.layers[0].a;
  const uint32_t a_base = smem_base + CHAIN_FIXED_BYTES;
  const uint32_t b_base = a_base + (uint32_t)a0.a_stages * a0.MT * a0.a_stage_bytes;
  const uint32_t stg_base = b_base + (uint32_t)a0.b_region_bytes;
  const int n = P.n;

  if (warp == WARP_TMA && lane == 0) {
    for (int s = 0; s < a0.a_stages; ++s) { mbar_init(bars.A_full(s), 1); mbar_init(bars.A_empty(s), 1); }
    for (int s = 0; s < a0.b_stages; ++s) { mbar_init(bars.B_full(s), 1); mbar_init(bars.B_empty(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(bars.T_full(s), 1); mbar_init(bars.T_empty(s), (uint32_t)(4 * a0.n_share * CG)); }
    mbar_init(bars.B_res(), 1);
    mbar_init(bars.M_done(), 1);
    for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(bars.R_full0() + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (j < n) {
        for (int s = 0; s < P.layers[j].a.n_src; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[j].tmA[s]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[j].tmB) : "memory");
      }
    }
  }
  if (warp == WARP_MMA) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars.tmem_slot()), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars.tmem_slot()), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  // the vectors of layer 0; the per-launch time-embedding row is a kernel parameter (launch constant)
  chain_stage_vecs(P.layers[0], P.temb_row, bt0, s_head, (int)threadIdx.x, TC_THREADS);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp != WARP_TMB) asm volatile("griddepcontrol.wait;" ::: "memory");

  const PairId pid = pair_id<CG>();
  ChainRt cx;
  cx.temb_row = P.temb_row;
  cx.it_base = 0; cx.k = 0; cx.n_seg = n; cx.j = 0;
  cx.ra = Ring{0, 0, a0.a_stages}; cx.rb = Ring{0, 0, a0.b_stages}; cx.rt = Ring{0, 0, a0.n_stages};
  unsigned ep = 0;
  if (warp != WARP_TMB) {          // (the previous launch of this segment finished at least one kernel boundary ago)
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ep) : "l"(P.epoch) : "memory");
  }
  cx.exp_mult = ep + 1u;

  if (warp == WARP_TMA) {
    if (elect_one_sync()) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < n) {
          cx.j = j;
          producer_role<CG, CH_P>(P.layers[j].a, P.maps[j], a_base, bar_base, &cx, &P.layers[j]);
        }
      }
    }
  } else if (warp == WARP_TMB) {
    if (elect_one_sync()) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < n) {
          cx.j = j;
          weight_producer_role<CG, CH_W>(P.layers[j].a, P.maps[j], b_base, bar_base, &cx);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    if ((CG == 1 || pid.rank == 0) && elect_one_sync()) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < n) {
          const TcArgs& a = P.layers[j].a;
          cx.j = j;
          const int taps = a.a_mode == 2 ? TAPS_S2 : (a.a_mode ? a.nkx * a.nky : 0);
#define SRD_MMA_CASE(T, D)                                                                                             \
  do {                                                                                                                 \
    if (a.b_resident && a.MT == 2) mma_role<T, 2, true, CG, D, CH_M>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);  \
    else if (a.b_resident) mma_role<T, 1, true, CG, D, CH_M>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);          \
    else if (a.MT == 2) mma_role<T, 2, false, CG, D, CH_M>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);            \
    else mma_role<T, 1, false, CG, D, CH_M>(a, tmem_base, a_base, b_base, bar_base, 0, &cx);                           \
  } while (0)
          if (a.dual) SRD_MMA_CASE(9, true);
          else if (taps == 9 && a.nkx == 3) SRD_MMA_CASE(9, false);
          else if (taps == 4 && a.nkx == 2) SRD_MMA_CASE(4, false);
          else if (taps == 1) SRD_MMA_CASE(1, false);
          else if (taps == TAPS_S2) SRD_MMA_CASE(TAPS_S2, false);
          else SRD_MMA_CASE(0, false);
#undef SRD_MMA_CASE
        }
      }
    }
  } else if (warp >= WARP_EPI0 && warp < WARP_EPI0 + EPI_WARPS) {
    const int etid = (int)threadIdx.x - WARP_EPI0 * 32;
    const int ew = warp - WARP_EPI0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (j < n) {
        const TcArgs& a = P.layers[j].a;
        const TcMaps& m_ = P.maps[j];
        cx.j = j;
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
        if (j + 1 < K && j + 1 < n)
          chain_stage_vecs(P.layers[j + 1 < K ? j + 1 : j], P.temb_row, bt0 + ((j + 1) & 1) * 512, s_head, etid, EPI_WARPS * 32);
        EpiSmem es;
        const Sched sc = make_sched<CG>(a, bar_base);
        es.bias = bt0 + (j & 1) * 512; es.temb = es.bias + 256;
        es.head_w = a.head_w ? s_head : nullptr; es.head_b = s_head + CHAIN_HEAD_FLOATS; es.hpart = nullptr;
        es.lin_w = a.lin_w ? s_head : nullptr;
        if (a.epi_tma) {
          StageSmem sg;
          sg.base = stg_base + (uint32_t)(ew * a.epi_nbuf) * 4096u;
          sg.gen = smem_raw + (sg.base - smem_u32(smem_raw));
          sg.r_full0 = bars.R_full0() + 8u * (uint32_t)(ew * 2);
          if (a.ep.act == SRD_ACT_MISH)
            epilogue_tma<SRD_ACT_MISH, CG, CH_E>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
          else
            epilogue_tma<SRD_ACT_NONE, CG, CH_E>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        } else if (a.ep.act == SRD_ACT_MISH)
          epilogue_role<SRD_ACT_MISH, CG, CH_E>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        else if (a.ep.act == SRD_ACT_LRELU02)
          epilogue_role<SRD_ACT_LRELU02, CG, CH_E>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        else
          epilogue_role<SRD_ACT_NONE, CG, CH_E>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        cx.it_base += (pid.pair < a.n_items) ? (uint32_t)((a.n_items - 1 - pid.pair) / pid.n_pairs + 1) : 0u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  // launch epoch: the CTA that finishes last advances it (every CTA has read it by then); long before the 32-bit
  // counters could wrap it zeroes them and starts over
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(P.done, 1u);
    if (t == gridDim.x - 1) {
      *P.done = 0u;
      if (ep + 1u >= SEG_EPOCH_MAX) {
        for (int i = 0; i < P.n_counters; ++i) P.counters[i] = 0u;
        *P.epoch = 0u;
      } else {
        *P.epoch = ep + 1u;
      }
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------ layer chain: host
// (per-device attribute, like tc_set_func_attrs in conv_tc.cu; only the CTA-pair instantiations exist)
static int chain_set_func_attrs(int dev) {
  static bool done[64] = {};
  if (dev < 0 || dev >= 64) return 1;
  if (done[dev]) return 0;
  if (cudaFuncSetAttribute(conv_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_seg_kernel<2, SEG_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return 1;
  done[dev] = true;
  return 0;
}

struct TcChain {
  ChainParams params;      // host copy; temb_row is filled in per launch
  int fused = 0;           // 1: conv_seg_kernel (<= SEG_K layers of one segment, epoch-scaled counters, no memset)
  TcMaps* d_maps = nullptr;
  unsigned* d_counters = nullptr;  // counters, then (fused) the epoch and done words
  size_t counter_bytes = 0;
  long long n_counters = 0;
  int n_layers = 0, n_segments = 0, grid = 0, smem = 0, cg = 2;
  ~TcChain() {
    if (d_maps) cudaFree(d_maps);
    if (d_counters) cudaFree(d_counters);
  }
};

// Two consecutive layers may share a segment when every role can carry its ring / barrier state across: identical
// shared-memory partition, TMEM stage layout and epilogue organisation.  A layer that keeps head / linear-residual
// weights in the shared vector block is a segment of its own (that block is single-buffered).
static bool same_segment(const TcArgs& x, const TcArgs& y) {
  return x.cg == y.cg && x.a_mode == y.a_mode && x.a_stage_bytes == y.a_stage_bytes && x.MT == y.MT &&
         x.a_stages == y.a_stages && x.b_resident == y.b_resident && x.Npad == y.Npad && x.b_stages == y.b_stages &&
         x.b_region_bytes == y.b_region_bytes && x.n_stages == y.n_stages && x.n_share == y.n_share && x.dual == y.dual &&
         x.epi_tma == y.epi_tma && x.epi_nbuf == y.epi_nbuf && !x.head_w && !y.head_w && !x.lin_w && !y.lin_w;
}

// links[i]: for every K source / the res1 tensor of layer i, the chain layer that produces it (-1: written before the
// launch) and which of its outputs (0: out, 1: second output); temb_off >= 0: offset of the layer's time-embedding
// vector in the per-launch row.
struct TcChainLink {
  int src_layer[SRD_MAX_SRC], src_which[SRD_MAX_SRC];
  int res1_layer, res1_which;
  int temb_off;
};

int tc_chain_build(const TcPlan* const* plans, const TcChainLink* links, int n, TcChain* chain, int flags) {
  const int g_single_segments = flags & 1;
  chain->fused = (flags & 2) ? 1 : 0;
  SRD_CHECK(n >= 1 && n <= CHAIN_MAX_LAYERS, "conv_chain: %d layers (at most %d fit the kernel parameter space)", n,
            CHAIN_MAX_LAYERS);
  SRD_CHECK(!chain->fused || n <= SEG_K, "conv_chain: a fused segment holds at most %d layers (got %d)", SEG_K, n);
  memset(&chain->params, 0, sizeof(chain->params));
  ChainLayer* L = chain->params.layers;
  std::vector<TcMaps> maps((size_t)n);
  // counters: one array per (layer, output) that has a reader inside the chain
  std::vector<long long> cnt_off((size_t)n * 2, -1);
  std::vector<int> n_rowblk((size_t)n);
  long long total = 0;
  auto need = [&](int p, int which) {
    if (cnt_off[(size_t)p * 2 + which] < 0) { cnt_off[(size_t)p * 2 + which] = total; total += n_rowblk[p]; }
  };
  int smem = 0, cg = plans[0]->args.cg;
  for (int i = 0; i < n; ++i) {
    const TcArgs& a = plans[i]->args;
    SRD_CHECK(a.cg == cg, "conv_chain: layers disagree on the CTA-pair mode");
#if defined(SRD_TC_DEBUG)
    SRD_CHECK(!a.dyn_sched, "conv_chain: layer %d uses the dynamic scheduler", i);      // debug builds keep the stamps
#else
    SRD_CHECK(!a.dyn_sched && !a.dbg, "conv_chain: layer %d uses the dynamic scheduler / debug stamps", i);
#endif
    SRD_CHECK(a.up == 1 || a.up == 2, "conv_chain: layer %d has up = %d (only 1 and 2 can be chained)", i, a.up);
    SRD_CHECK(!(a.head_w || a.lin_w) || a.Npad <= 128, "conv_chain: layer %d: fused head / linear residual need N <= 128", i);
    SRD_CHECK(!a.dual || a.Npad <= 128, "conv_chain: layer %d: a second output needs N <= 128 inside a chain", i);
    n_rowblk[i] = (a.n_vrows - a.vstart + TILE_H - 1) / TILE_H;
    SRD_CHECK(n_rowblk[i] < 65535, "conv_chain: layer %d has too many row blocks for the packed dependency cache", i);
    const int layer_smem = plans[i]->smem - (BAR_BYTES + VEC_BYTES) + CHAIN_FIXED_BYTES;
    if (layer_smem > smem) smem = layer_smem;
  }
  SRD_CHECK(smem <= 227 * 1024, "conv_chain: shared-memory plan too large (%d)", smem);
  for (int i = 0; i < n; ++i) {
    const TcChainLink& k = links[i];
    for (int s = 0; s < plans[i]->args.n_src; ++s)
      if (k.src_layer[s] >= 0) {
        SRD_CHECK(k.src_layer[s] < i && (k.src_which[s] == 0 || k.src_which[s] == 1), "conv_chain: bad source link of layer %d", i);
        need(k.src_layer[s], k.src_which[s]);
      }
    if (k.res1_layer >= 0) {
      SRD_CHECK(k.res1_layer < i && (k.res1_which == 0 || k.res1_which == 1), "conv_chain: bad res1 link of layer %d", i);
      need(k.res1_layer, k.res1_which);
    }
  }
  if (total == 0) total = 1;
  chain->n_counters = total;
  SRD_CUDA(cudaMalloc((void**)&chain->d_counters, (size_t)(total + 2) * sizeof(unsigned)));
  SRD_CUDA(cudaMemset(chain->d_counters, 0, (size_t)(total + 2) * sizeof(unsigned)));
  chain->counter_bytes = (size_t)total * sizeof(unsigned);
  auto make_dep = [&](int p, int which, ChainDep* d) {
    const TcArgs& pa = plans[p]->args;
    const int Ho = pa.ep.H_out;
    d->cnt = chain->d_counters + cnt_off[(size_t)p * 2 + which];
    const int n_phase = pa.up * pa.up * pa.nsplit;
    d->expect = pa.epi_tma ? (unsigned)(pa.n_colblk * (pa.Npad / 64) * 4 * n_phase)
                           : (unsigned)(pa.n_colblk * 4 * pa.n_share * n_phase);
    d->n_rowblk = n_rowblk[p];
    d->hp = Ho + 2; d->H = Ho;
    d->n_rows = (pa.n_vrows / pa.period) * (Ho + 2);
    d->up_shift = pa.up == 2 ? 1 : 0;
    d->period = pa.period; d->voff = pa.voff; d->vstart = pa.vstart;
    d->fd_hp = make_fastdiv((uint32_t)(Ho + 2));
  };
  int n_seg = 0;
  for (int i = 0; i < n; ++i) {
    const TcArgs& a = plans[i]->args;
    const TcChainLink& k = links[i];
    L[i].a = a;
    maps[(size_t)i] = plans[i]->maps;
    L[i].temb_off = k.temb_off;
    SRD_CHECK(k.temb_off < 0 || (a.fast && a.ep.temb_bstride == 0), "conv_chain: layer %d cannot take a per-launch temb row", i);
    for (int s = 0; s < a.n_src; ++s)
      if (k.src_layer[s] >= 0) make_dep(k.src_layer[s], k.src_which[s], &L[i].dep_src[s]);
    if (k.res1_layer >= 0) {
      SRD_CHECK(a.epi_tma || a.epi_geo == 0, "conv_chain: layer %d reads an in-chain residual in a strided register epilogue", i);
      make_dep(k.res1_layer, k.res1_which, &L[i].dep_res1);
    }
    if (cnt_off[(size_t)i * 2] >= 0) L[i].out_cnt = chain->d_counters + cnt_off[(size_t)i * 2];
    if (cnt_off[(size_t)i * 2 + 1] >= 0) {
      SRD_CHECK(a.dual, "conv_chain: layer %d has no second output", i);
      L[i].out2_cnt = chain->d_counters + cnt_off[(size_t)i * 2 + 1];
    }
    L[i].seg_first = (i == 0 || g_single_segments || !same_segment(plans[i - 1]->args, a)) ? 1 : 0;
    SRD_CHECK(!chain->fused || i == 0 || !L[i].seg_first, "conv_chain: layer %d does not share the segment plan of layer 0", i);
    n_seg += L[i].seg_first;
  }
  SRD_CUDA(cudaMalloc((void**)&chain->d_maps, sizeof(TcMaps) * (size_t)n));
  SRD_CUDA(cudaMemcpy(chain->d_maps, maps.data(), sizeof(TcMaps) * (size_t)n, cudaMemcpyHostToDevice));
  chain->params.maps = chain->d_maps;
  chain->params.n_layers = n;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  chain->n_layers = n; chain->n_segments = n_seg; chain->smem = smem; chain->cg = cg;
  chain->grid = (n_sm / cg) * cg;
  SRD_CHECK(cg == 2, "conv_chain: multi-layer launches are built for CTA pairs (cta_group::2) only");
  SRD_CHECK(chain_set_func_attrs(dev) == 0, "conv_chain: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
  return 0;
}

TcChain* tc_chain_new() { return new (std::nothrow) TcChain; }
void tc_chain_delete(TcChain* c) { delete c; }
int tc_chain_segments(const TcChain* c) { return c->n_segments; }
int tc_chain_max_layers() { return CHAIN_MAX_LAYERS; }

int tc_same_segment(const TcPlan* x, const TcPlan* y) { return same_segment(x->args, y->args) ? 1 : 0; }
int tc_seg_k() { return SEG_K; }

int launch_chain(const TcChain& c, const float* temb_row, cudaStream_t stream) {
  if (!c.fused) SRD_CUDA(cudaMemsetAsync(c.d_counters, 0, c.counter_bytes, stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)c.grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)c.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  if (tc_pdl_enabled()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  if (c.cg == 2) {
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = 2; at[n_at].val.clusterDim.y = 1; at[n_at].val.clusterDim.z = 1;
    ++n_at;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n_at;
  if (c.fused) {
    SegParams<SEG_K> S;
    memset(&S, 0, sizeof(S));
    S.maps = c.d_maps; S.temb_row = temb_row;
    S.counters = c.d_counters; S.n_counters = (int)c.n_counters;
    S.epoch = c.d_counters + c.n_counters; S.done = c.d_counters + c.n_counters + 1;
    S.n = c.n_layers;
    for (int i = 0; i < c.n_layers; ++i) S.layers[i] = c.params.layers[i];
    SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_seg_kernel<2, SEG_K>, S));
    return 0;
  }
  ChainParams P = c.params;
  P.temb_row = temb_row;
  SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<2>, P));
  return 0;
}

}  // namespace srd
