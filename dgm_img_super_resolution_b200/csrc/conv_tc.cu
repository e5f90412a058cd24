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
#include <cuda.h>
#include <string.h>
#include <new>
#include <vector>
#include "epilogue.cuh"

namespace srd {

#ifndef SRD_EPI_WARPS
#define SRD_EPI_WARPS 16
#endif
#ifndef SRD_ROLES_HIGH
#define SRD_ROLES_HIGH 1
#endif
constexpr int EPI_WARPS = SRD_EPI_WARPS; // groups of 4 warps (one warp per TMEM lane quarter)
constexpr int EPI_GROUPS = EPI_WARPS / 4;
// Experiment kept behind a macro: a second MMA-issuing thread taking the odd items of resident-weight
// layers.  Measured twice on B200: no gain (229 vs 230 images/s) -- those layers are balanced between the
// MMA thread, the Mish epilogue and HBM, not bound by MMA issue alone.
#ifndef SRD_DUAL_ISSUE
#define SRD_DUAL_ISSUE 0
#endif
// activation-TMA warp + weight-TMA warp + MMA warp(s) + epilogue warps
constexpr int TC_THREADS = 96 + 32 * SRD_DUAL_ISSUE + 32 * EPI_WARPS;
// Warp roles.  The SMSP arbiter prefers the highest warp id, so the two single-thread roles
// (TMA producer, MMA issuer) sit ABOVE the issue-heavy epilogue warps: an MMA that is ready to
// issue must never queue behind Mish arithmetic.
constexpr int WARP_EPI0 = SRD_ROLES_HIGH ? 0 : 3 + SRD_DUAL_ISSUE;
constexpr int WARP_TMA = SRD_ROLES_HIGH ? EPI_WARPS : 0;        // activation (A operand) producer
constexpr int WARP_TMB = SRD_ROLES_HIGH ? EPI_WARPS + 1 : 1;    // weight (B operand) producer
constexpr int WARP_MMA = SRD_ROLES_HIGH ? EPI_WARPS + 2 : 2;    // MMA issuer (+ TMEM owner)
constexpr int WARP_MMA2 = WARP_MMA + 1;   // SRD_DUAL_ISSUE: second MMA issuer (odd local items of dual_issue layers)
constexpr int TILE_W = 8, TILE_H = 16;
constexpr int BOX_ROWS = TILE_H + 2;
constexpr int A_BYTES = BOX_ROWS * TILE_W * 128;   // 18432: one {64ch x 8px x 18rows} bf16 box
constexpr int HALO_W = TILE_W + 2;
constexpr int A_HALO_BYTES = BOX_ROWS * HALO_W * 128;      // 23040: one {64ch x 10px x 18rows} box (all 9 taps)
constexpr int A_HALO_STAGE = (A_HALO_BYTES + 1023) / 1024 * 1024;
constexpr int MAX_STAGES = 8;
constexpr int DBG_ITEMS = 64;
constexpr int BAR_BYTES = 1024;          // mbarriers + TMEM slot
constexpr int VEC_BYTES = (512 + 1024 + 4 + 2 * 128 * 4) * 4;   // epilogue vectors (see conv_tc_kernel); the layout uses
                                                                // 512 + 512 + 1024 + 4 floats of it

// Exact n / d for n < 2^24 as one 32x32->64 multiply and a shift (the single-thread roles and the
// per-tile epilogue decode must not run ~40-instruction integer divisions).
struct FastDiv {
  uint32_t mul, shr;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  uint32_t s = 0;
  while ((1u << s) < d) ++s;
  FastDiv f;
  f.shr = 24 + s;
  f.mul = (uint32_t)((((uint64_t)1 << f.shr) + d - 1) / d);
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return (int)(((uint64_t)(uint32_t)n * f.mul) >> f.shr);
}

// The TMA tensor maps of one layer.  Kept apart from the scalar arguments: the single-layer kernel receives both as
// __grid_constant__ parameters, the chain kernel receives the scalar arguments of ALL its layers in parameter space
// (constant-bank operands, like the single-layer kernel's) and finds the maps in global memory.
struct TcMaps {
  CUtensorMap tmA[SRD_MAX_SRC];
  CUtensorMap tmB;
  CUtensorMap tmO;         // output, box {64 ch, 8 px, 4 rows} = one epilogue warp's rows (TMA-staged epilogue)
  CUtensorMap tmOrow;      // output, box {64 ch, 8 px, 1 row}
  CUtensorMap tmR;         // residual res1, box {64 ch, 8 px, 4 rows}
  CUtensorMap tmO2;        // second output (dual), box {64 ch, 8 px, 4 rows}
};
struct TcArgs {
  int n_src;
  int nchunk[SRD_MAX_SRC];
  int H_m, W_m;            // output positions per image (per phase)
  int period, voff, vstart, n_vrows;
  int n_colblk, tiles_per_phase, items_per_phase, n_items;
  FastDiv fd_ipp, fd_colblk, fd_period, fd_up;   // dividers by items_per_phase, n_colblk, period, up
  int n_chunks;            // 64-channel K chunks per item over all sources
  int dual_issue;          // SRD_DUAL_ISSUE build: two MMA-issuing threads take alternate items (see mma_role)
  int dual;                // 1: second output = 1x1 conv of the same sources (SrdConvDesc.N2): its accumulator sits
                           //    MT*Npad columns after the first one, its weights are the 10th k-block of every chunk
  const float* bias2;
  int cg;                  // CTAs per MMA (1, or 2 = cta_group::2 pair: M = 256, each CTA holds half of every weight tile)
  int stride, nky, nkx, ky0, kx0, up;
  int Npad, MT;
  int a_stages, b_stages;
  int kb_per_phase;
  uint32_t idesc;
  Epi ep;
  const float* head_w;
  const float* head_b;
  const __nv_bfloat16* lin_src;   // linear residual: v += lin_w[n][c] * lin_src[pixel][c], c < lin_c (<= 4)
  const float* lin_w;
  int lin_ct, lin_c;
  int* err_flag;
  int fast;                // epilogue fast path eligible (N % 16 == 0, uniform temb, aligned windows)
  int a_mode;              // 0: one box per (chunk, kx), ky by 1024 B row offsets (always-aligned descriptors)
                           // 1|2: ONE halo box per chunk, (ky,kx) by 128 B-granular descriptor start offsets
                           //      (2: descriptor base_offset = (addr >> 7) & 7)
  int a_stage_bytes;       // bytes of one A tile slot inside a stage
  int b_resident;          // all weight k-blocks stay in shared memory for the whole kernel
  int kb_total;
  int b_region_bytes;      // bytes of the weight region (ring or resident)
  int epi_tma;             // TMA-staged epilogue (see epilogue_tma)
  int epi_geo;             // 0: stride-1 conv (virtual row == padded output row); 1: stride-2 conv or 2x transposed
                           //    conv phase: output rows / columns are mapped per tile row (tmOrow is 5-D when up == 2)
  int epi_nbuf;            // staging buffers per epilogue warp (1)
  int n_stages;            // TMEM accumulator stages (2 | 4, power of two); local item i accumulates in stage i % n_stages
  int n_share;             // epilogue groups that split ONE item's sub-tiles (1 | 2 | 4); the EPI_GROUPS / n_share
                           // group sets work on different items concurrently (<= n_stages of them)
  int stage_shift;         // log2(n_stages)
  unsigned long long* dbg; // optional timeline buffer [cta][DBG_ITEMS][8] of clock64 stamps (debug)
  unsigned* sched_counter; // dynamic item scheduler (see Sched): per-plan ticket counter, 0 between launches
  int dyn_sched;           // 1: items beyond each pair's first are handed out by tickets, not round-robin
  int reverse_items;       // static deal only: item order reversed (see SrdConvDesc.reverse_items)
  int nsplit;              // 1 | 2: the N output channels are computed as `nsplit` groups of Npad channels, each group a
                           // "phase" of its own (own weight k-blocks, own items): N = 256 runs as two 128-column halves
                           // with TWO tile pairs per weight block (MT = 2) in the TMEM space one 256-column tile needs,
                           // which halves the weight bytes streamed from L2 per MMA (SrdConvDesc.n_split)
};

// ------------------------------------------------------------------------------------------ layer chain
// conv_chain_kernel runs a LIST of conv layers in one launch.  Every CTA pair walks the list on its own; what orders
// a consumer layer after its producer is not a kernel boundary but a per-row-block completion counter in global
// memory: every epilogue warp bumps `cnt[row block of its tile]` once the TMA stores of a sub-tile have completed,
// and whoever is about to read rows of that tensor (the activation-TMA producer for the K sources, the epilogue warps
// for the residual prefetch) first waits until the row blocks covering those rows -- plus the rows their mirrored
// ring copies come from -- have reached the tensor's full count.  Row block r of a tensor = the producer's M tiles
// [r * n_colblk, (r + 1) * n_colblk) of every sub-pixel phase.
struct ChainDep {
  const unsigned* cnt;     // [n_rowblk] completion counters of the producing layer's output; nullptr = written before the launch
  unsigned expect;         // value of a complete row block
  int n_rowblk, n_rows;    // row blocks / padded stacked rows (B * (H + 2)) of the tensor
  int hp, H;               // padded / unpadded rows per image
  int up_shift;            // producer's M row = output row >> up_shift (1 for the 2x transposed conv)
  int period, voff, vstart;   // producer's stacked-row geometry (TcArgs)
  FastDiv fd_hp;
};
struct ChainLayer {
  TcArgs a;
  ChainDep dep_src[SRD_MAX_SRC];
  ChainDep dep_res1;
  unsigned* out_cnt;       // counters this layer's epilogue bumps for `out` (nullptr: nobody in the chain reads it)
  unsigned* out2_cnt;      // ... for the second output (SrdConvDesc.N2)
  int temb_off;            // >= 0: time-embedding vector = temb_row + temb_off (per-launch row of the time table)
  int seg_first;           // 1: first layer of a segment (see conv_chain_kernel)
};
// Role state handed from layer to layer inside a segment.
struct Ring {
  int stage, phase, n;
  __device__ __forceinline__ void next() { if (++stage == n) { stage = 0; phase ^= 1; } }
  __device__ __forceinline__ void skip(int k) {          // k slots consumed by another thread
    stage += k;
    while (stage >= n) { stage -= n; phase ^= 1; }
  }
};
struct ChainRt {
  int j, n_seg;            // the current layer's index inside its segment / layers in the segment
  uint32_t it_base;        // items this pair has processed in the segment's earlier layers (TMEM stage / parity continuity)
  Ring ra, rb, rt;
  uint32_t k;              // epilogue staging-buffer counter
  uint32_t exp_mult;       // completion counters are compared with expect * exp_mult (1, or launch epoch + 1 when the
                           // counters are never reset: conv_seg_kernel)
  const float* temb_row;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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
// Bounded wait: a protocol bug must not hang the GPU (a hang costs the whole box).  The bound is TIME
// (%globaltimer, 20 s), not a spin count: under a profiler, time-slicing or a PDL-overlapped predecessor a
// legitimate wait (e.g. the weight producer parked behind the previous grid) can spin arbitrarily often.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifndef SRD_WATCHDOG_MS
#define SRD_WATCHDOG_MS 20000
#endif
constexpr unsigned long long WATCHDOG_NS = (unsigned long long)SRD_WATCHDOG_MS * 1000 * 1000;
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  const unsigned long long t0 = global_ns();
  while (!mbar_try(bar, parity)) {
    __nanosleep(64);
    if (global_ns() - t0 > WATCHDOG_NS) {
      if (err_flag) atomicExch(err_flag, code);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 16)) {                 // ~tens of ms of hardware try_wait windows: switch to the slow path
      mbar_wait_slow(bar, parity, err_flag, code);
      return;
    }
  }
}
// ---- layer-chain dependencies (see ChainDep) --------------------------------------------------------------------
// Producer row block that writes padded stacked row R of a tensor (ring rows map to the adjacent interior row).
__device__ __forceinline__ int dep_rowblk(const ChainDep& d, int R) {
  R = R < 0 ? 0 : (R >= d.n_rows ? d.n_rows - 1 : R);
  const int b = (int)(((uint64_t)(uint32_t)R * d.fd_hp.mul) >> d.fd_hp.shr);
  int oy = R - b * d.hp - 1;
  oy = oy < 0 ? 0 : (oy >= d.H ? d.H - 1 : oy);
  const int r = (b * d.period + (oy >> d.up_shift) + d.voff - d.vstart) >> 4;
  return r >= d.n_rowblk ? d.n_rowblk - 1 : r;
}
__device__ __noinline__ void dep_poll(const unsigned* p, unsigned expect, int* err_flag) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (v >= expect) return;
  const unsigned long long t0 = global_ns();
  for (;;) {
    __nanosleep(40);
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= expect) return;
    if (global_ns() - t0 > WATCHDOG_NS) {
      if (err_flag) atomicExch(err_flag, 11);
      __trap();
    }
  }
}
// Wait until padded stacked rows [Ra, Rb) of the tensor -- and the rows two above / below, where mirrored ring rows
// come from -- are complete.  [vlo, vhi] caches the row blocks this thread has already seen complete.
__device__ __forceinline__ bool dep_cached(const ChainDep& d, int Ra, int Rb, int vlo, int vhi) {
  return dep_rowblk(d, Ra - 2) >= vlo && dep_rowblk(d, Rb + 1) <= vhi;
}
__device__ __forceinline__ void dep_wait(const ChainDep& d, int Ra, int Rb, int& vlo, int& vhi, int* err_flag,
                                         uint32_t exp_mult) {
  const int lo = dep_rowblk(d, Ra - 2), hi = dep_rowblk(d, Rb + 1);
  if (lo >= vlo && hi <= vhi) return;
  for (int r = lo; r <= hi; ++r)
    if (r < vlo || r > vhi) dep_poll(d.cnt + r, d.expect * exp_mult, err_flag);
  if (vhi >= vlo && lo <= vhi + 1 && hi >= vlo - 1) { vlo = lo < vlo ? lo : vlo; vhi = hi > vhi ? hi : vhi; }
  else { vlo = lo; vhi = hi; }
  // the counters were read with acquire loads; the loads that follow belong to the async proxy
  asm volatile("fence.proxy.async;" ::: "memory");
}
// One epilogue warp has finished a sub-tile: its bulk stores are COMPLETE (not merely read), the per-thread ring
// stores of the other lanes are ordered before this by the preceding __syncwarp.  Lane 0 only.
// The bump itself is a RELAXED reduction: a gpu-scope release would make lane 0 (and with it the warp) wait for a
// full memory round trip per sub-tile, which the 64-channel layers' epilogue cannot afford.  What it has to be ordered
// after is (a) the warp's bulk stores -- complete, i.e. performed at the L2, once wait_group returns -- and (b) the
// few per-thread ring-column stores of edge tiles, whose lanes fence them themselves before the __syncwarp.
__device__ __forceinline__ void dep_publish(unsigned* cnt) {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
}

// In a CTA pair (CG == 2) the shared::cluster address of a CTA-local object carries the CTA rank in
// bit 24; clearing it names the same offset in the LEADER (even) CTA.  The pair's operand loads
// (.cta_group::2) complete on the leader's mbarrier, where the single MMA-issuing thread waits.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
template <int CG = 1>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  if (CG == 1)
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
template <int CG = 1>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  if (CG == 1)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int CG = 1>
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  if (CG == 1)
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (epilogue warps of both CTAs release the pair's accumulators)
template <int CG>
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  if (CG == 1)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  else
    // default (.release.cta) semantics: the accumulator hand-over is ordered by tcgen05.wait::ld +
    // tcgen05.fence::before_thread_sync; a .release.cluster arrive would add a MEMBAR.ALL.CTA per warp per item
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// CG == 2: the commit arrives on the barrier at the same offset in BOTH CTAs of the pair (each CTA's
// producer / epilogue waits on its own copy).
template <int CG = 1>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG = 1>
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

__device__ __forceinline__ uint64_t umma_desc_sw128_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7u) << 49) | (2ull << 61);
}

// Descriptor split into 32-bit halves: lo = start address >> 4 (LBO = 0); hi = SBO >> 4 | version 1 (bit 46)
// | SWIZZLE_128B (bits 61-63).  Advancing the start address is an add on the low word.
constexpr uint32_t DESC_HI_1024 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_HALO = ((uint32_t)(TILE_W + 2) * 128u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Elect exactly one lane of a converged warp.  Unlike `lane == 0`, the compiler knows the guarded
// region runs on a single thread, so tcgen05/TMA operands live in uniform registers without
// per-instruction ELECT/BRA.U.ANY waterfall loops.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Debug timeline (tools/timeline_conv.py); compiled in only with -DSRD_TC_DEBUG so the role loops
// of the product build carry no extra code.
__device__ __forceinline__ void dbg_stamp(const TcArgs& a, int item_local, int ev) {
#if defined(SRD_TC_DEBUG) && SRD_TC_DEBUG != 2
  if (a.dbg && item_local < DBG_ITEMS)
    a.dbg[((size_t)blockIdx.x * DBG_ITEMS + item_local) * 8 + ev] = (unsigned long long)clock64();
#else
  (void)a; (void)item_local; (void)ev;
#endif
}

// -DSRD_TC_DEBUG=2: the eight stamps of an item describe ONE epilogue warp (quarter 0 of the item's group):
// 0 loop top, 1 T_full acquired, 2 staging buffer ready, 3 chunk loop done, 4 proxy fence done,
// 5 TMA store(s) issued, 6 store read drained, 7 next buffer prepared.
__device__ __forceinline__ void epi_stamp(const TcArgs& a, int item_local, int ev) {
#if defined(SRD_TC_DEBUG) && SRD_TC_DEBUG == 2
  if (a.dbg && item_local < DBG_ITEMS)
    a.dbg[((size_t)blockIdx.x * DBG_ITEMS + item_local) * 8 + ev] = (unsigned long long)clock64();
#else
  (void)a; (void)item_local; (void)ev;
#endif
}

// -DSRD_TC_DEBUG=3: kernel-phase stamps of thread 0 in item slot DBG_ITEMS-1: 0 entry, 1 setup done (after
// the block / cluster sync), 2 previous grid complete (griddepcontrol.wait), 3 roles done (before the final
// sync), 4 after the final sync, 5 TMEM freed; plus globaltimer (ns) at entry in 6 and at exit in 7.
__device__ __forceinline__ void kern_stamp(const TcArgs& a, int ev, bool global_ns = false) {
#if defined(SRD_TC_DEBUG) && SRD_TC_DEBUG == 3
  if (a.dbg && threadIdx.x == 0) {
    unsigned long long v;
    if (global_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    else v = (unsigned long long)clock64();
    a.dbg[((size_t)blockIdx.x * DBG_ITEMS + DBG_ITEMS - 1) * 8 + ev] = v;
  }
#else
  (void)a; (void)ev; (void)global_ns;
#endif
}

// Where this CTA sits: work items are dealt to CTA pairs (CG == 2) or single CTAs (CG == 1); item i of
// a phase covers tiles [i*MT*CG, (i+1)*MT*CG), of which CTA `rank` owns tiles i*MT*CG + mt*CG + rank.
struct PairId {
  int rank, pair, n_pairs;
};
template <int CG>
__device__ __forceinline__ PairId pair_id() {
  PairId p;
  if (CG == 1) { p.rank = 0; p.pair = (int)blockIdx.x; p.n_pairs = (int)gridDim.x; }
  else { p.rank = (int)cluster_ctarank(); p.pair = (int)blockIdx.x >> 1; p.n_pairs = (int)gridDim.x >> 1; }
  return p;
}

// Work distribution.  Static (default): pair p runs items p, p + n_pairs, ...  Dynamic (a.dyn_sched, opt-in through
// srd_tc_set_mode bit 6): pairs progress at different speeds, and with a static deal the last pair exits 7-12 us
// after the first while the next launch waits -- so hand the items out by tickets instead.  The first two items of
// a pair stay static; every further one is a ticket: the leader CTA's activation-producer thread draws
// `atomicAdd(counter)` when it starts an item (the latency hides behind that item's loads) and publishes the item
// after next -- or -1 -- into a small ring in the shared memory of BOTH CTAs of the pair (st.async + mbarrier
// complete_tx, so the value is visible to whoever waits on the slot's barrier).  Every other role reads its items
// from that ring, in order.  Each processed item draws exactly one ticket, so a launch draws n_items tickets in
// total and the thread that gets the last one resets the counter.  The ring (32 slots) is deeper than the furthest
// any role can lag behind the producer (A ring <= 8 items + 2 x 4 accumulator stages + the trailing -1 entries the
// residual prefetch may look ahead into).
// MEASURED: correct (same-seed replays bit-identical) but 2 % SLOWER end to end than the static deal (221 vs 226
// images/s): the ring reads and the later start of the peer CTA's loads cost more than the evened-out tail wins,
// and the 20x20 layers (1.8 item rounds) cannot be helped by any scheduler.  Kept as a switch.
constexpr int SCHED_SLOTS = 32;
constexpr uint32_t SCHED_BAR_OFF = 8u * (4 * MAX_STAGES + 10 + 2 * EPI_WARPS);    // inside the BAR_BYTES block
constexpr uint32_t SCHED_RING_OFF = SCHED_BAR_OFF + 8u * SCHED_SLOTS;
static_assert(SCHED_RING_OFF + 4 * SCHED_SLOTS <= BAR_BYTES, "scheduler ring does not fit the barrier block");
struct Sched {
  uint32_t bar0, ring0;      // shared::cta addresses of the slot barriers / values
  int dyn, rev, pair, n_pairs, n_items;
  int* err;
  // the il-th item of this pair, -1 past the end
  __device__ __forceinline__ int get(int il) const {
    if (!dyn) {
      const int item = pair + il * n_pairs;
      if (item >= n_items) return -1;
      return rev ? n_items - 1 - item : item;          // SrdConvDesc.reverse_items: last tile first
    }
    const uint32_t slot = (uint32_t)il & (SCHED_SLOTS - 1);
    mbar_wait(bar0 + 8u * slot, ((uint32_t)il / SCHED_SLOTS) & 1u, err, 10);
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(ring0 + 4u * slot) : "memory");
    return v;
  }
  // scheduler thread only: entry il := item, in every CTA of the pair
  template <int CG>
  __device__ __forceinline__ void publish(int il, int item) const {
    const uint32_t slot = (uint32_t)il & (SCHED_SLOTS - 1);
#pragma unroll
    for (uint32_t r = 0; r < (uint32_t)CG; ++r) {
      uint32_t rbar, rval;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar0 + 8u * slot), "r"(r));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rval) : "r"(ring0 + 4u * slot), "r"(r));
      asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], 4;" ::"r"(rbar) : "memory");
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                   ::"r"(rval), "r"(item), "r"(rbar) : "memory");
    }
  }
};
template <int CG>
__device__ __forceinline__ Sched make_sched(const TcArgs& a, uint32_t bar_base) {
  const PairId pid = pair_id<CG>();
  Sched sc;
  sc.bar0 = bar_base + SCHED_BAR_OFF; sc.ring0 = bar_base + SCHED_RING_OFF;
  sc.dyn = a.dyn_sched; sc.rev = a.reverse_items; sc.pair = pid.pair; sc.n_pairs = pid.n_pairs; sc.n_items = a.n_items; sc.err = a.err_flag;
  return sc;
}

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

// ------------------------------------------------------------------------------------------ epilogue
// Per-launch constants of the fused epilogue staged in shared memory (float): bias[Npad],
// temb[Npad] (zeros when absent), head weights [4][Npad] + head bias[4], and the partial head
// sums exchanged between the two warps that share a TMEM lane quarter.
struct EpiSmem {
  const float* bias;
  const float* temb;
  const float* head_w;
  const float* head_b;
  float* hpart;     // [2][128][4]
  const float* lin_w;   // [4][Npad] (transposed, zero padded) or nullptr; shares the head_w storage
};

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// eight bf16 sums, each rounded once (HADD2.BF16_V2)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 add_bf16x8(const uint4& a, const uint4& b) {
  return make_uint4(add_bf16x2(a.x, b.x), add_bf16x2(a.y, b.y), add_bf16x2(a.z, b.z), add_bf16x2(a.w, b.w));
}

template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == SRD_ACT_MISH) return mish_fast(x);
  if (ACT == SRD_ACT_LRELU02) return fmaxf(x, 0.2f * x);
  return x;
}

// 16 bf16 channels = 32 B = one full DRAM sector.  When the destination is 32 B aligned, one 256-bit
// store (sm_100 STG.256) writes the whole sector; two 16 B stores would be two PARTIAL-sector writes.
__device__ __forceinline__ void st16_bf16(__nv_bfloat16* p, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
  b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]);
  b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  } else {
    reinterpret_cast<uint4*>(p)[0] = a;
    reinterpret_cast<uint4*>(p)[1] = b;
  }
}
// 32 B residual load (one sector) when aligned
__device__ __forceinline__ void ld16_bf16(const __nv_bfloat16* p, uint4 (&x)[2]) {
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(x[0].x), "=r"(x[0].y), "=r"(x[0].z), "=r"(x[0].w), "=r"(x[1].x), "=r"(x[1].y), "=r"(x[1].z), "=r"(x[1].w)
                 : "l"(p));
  } else {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    x[0] = q[0]; x[1] = q[1];
  }
}

// Write one pixel's 16-channel chunk, plus the mirrored ring copies this pixel owns (reflect).
// p points at the pixel's channel window (+ chunk offset); wp = padded pixels per row.
__device__ __forceinline__ void store_chunk(__nv_bfloat16* p, int ct, long long wp, uint32_t edge,
                                            const float (&v)[16]) {
  st16_bf16(p, v);
  if (edge) {   // rare: only pixels in row/col 1 or H-2/W-2 of a reflect-ringed output
    const long long dn = wp * ct;
    const bool top = edge & 1u, bot = edge & 2u, lef = edge & 4u, rig = edge & 8u;
    if (top) st16_bf16(p - 2 * dn, v);
    if (bot) st16_bf16(p + 2 * dn, v);
    if (lef) st16_bf16(p - 2 * ct, v);
    if (rig) st16_bf16(p + 2 * ct, v);
    if (top && lef) st16_bf16(p - 2 * dn - 2 * ct, v);
    if (top && rig) st16_bf16(p - 2 * dn + 2 * ct, v);
    if (bot && lef) st16_bf16(p + 2 * dn - 2 * ct, v);
    if (bot && rig) st16_bf16(p + 2 * dn + 2 * ct, v);
  }
}

// alpha * act(x) + t with the activation resolved at compile time.
template <int ACT>
__device__ __forceinline__ float act_axpy(float x, float alpha, float t) {
  if (ACT == SRD_ACT_MISH) {
    // mish(x) = x * g,  g = tanh(softplus(x)) = 1 - 2 / (n*(n+2) + 2),  n = e^x  (1 ex2 + 1 rcp).
    // No clamp on x: n overflows to +inf only where g == 1 anyway (rcp(inf) = 0), and there is no
    // inf * 0 on the way.  8 instructions per element including the bias add of the caller.
    float n, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(x * 1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(n, n + 2.f, 2.f)));
    const float g = fmaf(r, -2.f, 1.f);
    (void)alpha;                                   // the fast epilogues require alpha == 1 with Mish (tc_plan_build)
    return fmaf(x, g, t);
  }
  if (ACT == SRD_ACT_LRELU02) return fmaf(fmaxf(x, 0.2f * x), alpha, t);
  return fmaf(x, alpha, t);
}

// Per-tile, per-thread constants of the fast epilogue.
struct TilePix {
  bool valid;
  uint32_t edge;
  int b, oy, ox;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  const __nv_bfloat16* r1;
  const __nv_bfloat16* r2;
};

// Fast path: N % 16 == 0, batch-uniform temb, 16-byte aligned channel windows.
template <int ACT>
__device__ __forceinline__ void epi_chunk_fast(const TcArgs& a, const EpiSmem& es, int c0, const uint32_t (&r)[16],
                                               const TilePix& tp, long long dn, const uint4 (&q1)[2],
                                               const uint4 (&q2)[2], float (&head)[4]) {
  float v[16];
  const float4* sb = reinterpret_cast<const float4*>(es.bias + c0);
  const float4* st = reinterpret_cast<const float4*>(es.temb + c0);
  const float alpha = a.ep.alpha;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b4 = sb[q], t4 = st[q];
    v[4 * q + 0] = act_axpy<ACT>(__uint_as_float(r[4 * q + 0]) + b4.x, alpha, t4.x);
    v[4 * q + 1] = act_axpy<ACT>(__uint_as_float(r[4 * q + 1]) + b4.y, alpha, t4.y);
    v[4 * q + 2] = act_axpy<ACT>(__uint_as_float(r[4 * q + 2]) + b4.z, alpha, t4.z);
    v[4 * q + 3] = act_axpy<ACT>(__uint_as_float(r[4 * q + 3]) + b4.w, alpha, t4.w);
  }
  if (tp.r1) {
    float f[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      unpack_bf16x8(q1[h], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] = fmaf(a.ep.beta1, f[i], v[8 * h + i]);
    }
  }
  if (tp.r2) {
    float f[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      unpack_bf16x8(q2[h], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] = fmaf(a.ep.beta2, f[i], v[8 * h + i]);
    }
  }
  if (tp.valid) {
    if (tp.out) store_chunk(tp.out + c0, a.ep.out_ct, dn, tp.edge, v);
    if (tp.out2) store_chunk(tp.out2 + c0, a.ep.out2_ct, dn, tp.edge, v);   // out/out2 share H, W; dn uses out's ct
  }
  if (es.head_w) {
#pragma unroll
    for (int jh = 0; jh < 4; ++jh) {           // static indices: head[] stays in registers
      if (jh >= a.ep.planar_c) break;
      const float4* hw = reinterpret_cast<const float4*>(es.head_w + jh * a.Npad + c0);
      float acc = head[jh];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = hw[q];
        acc = fmaf(w4.x, v[4 * q + 0], acc); acc = fmaf(w4.y, v[4 * q + 1], acc);
        acc = fmaf(w4.z, v[4 * q + 2], acc); acc = fmaf(w4.w, v[4 * q + 3], acc);
      }
      head[jh] = acc;
    }
  }
}

// Decode one tile for this thread: (image, row, col) of its output pixel, without divisions per thread.
struct TileGeo {
  int V0, x0;          // first virtual stacked row / first column of the tile
  int b, y, x;         // this thread's image / row / col in the per-phase M grid
  bool valid;
};
__device__ __forceinline__ TileGeo tile_geo(const TcArgs& a, int tt, int g, int px) {
  TileGeo t;
  const int rowblk = fdiv(tt, a.fd_colblk), colblk = tt - rowblk * a.n_colblk;
  t.V0 = rowblk * TILE_H + a.vstart;
  t.x0 = colblk * TILE_W;
  int b = fdiv(t.V0, a.fd_period);
  int vv = t.V0 - b * a.period + g;
  while (vv >= a.period) { vv -= a.period; ++b; }
  t.b = b;
  t.y = vv - a.voff;
  t.x = t.x0 + px;
  t.valid = (tt < a.tiles_per_phase) && (t.V0 + g < a.n_vrows) && (t.x < a.W_m) && (t.y >= 0) && (t.y < a.H_m);
  return t;
}

// ---------------------------------------------------------------------------------------------
// Register epilogue: thread = pixel, direct 16 B global stores.  Used by the layers the TMA-staged
// epilogue does not take (transposed-conv phases, stride 2, N % 64 != 0, fused 1x1 head, ragged W).
// The two epilogue warpgroups (halves) ping-pong on the two TMEM accumulator stages: half h owns
// the items with local index = h (mod 2), i.e. TMEM stage h.
// ---------------------------------------------------------------------------------------------
template <int ACT, int CG, bool CHAIN = false>
__device__ __forceinline__ void epilogue_role(const TcArgs& a, const EpiSmem& es, const Sched& sc, uint32_t tmem_base,
                                              uint32_t t_full0, uint32_t t_empty0, int warp, int lane,
                                              ChainRt* cx = nullptr, const ChainLayer* L = nullptr) {
  const uint32_t it_base = CHAIN ? cx->it_base : 0u;
  unsigned* const out_cnt = CHAIN ? L->out_cnt : nullptr;
  int r_vlo = 0, r_vhi = -1;
  const int S = a.up;
  const PairId pid = pair_id<CG>();
  const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
  const int group = (warp - WARP_EPI0) >> 2;    // epilogue group
  const int n_conc = EPI_GROUPS / a.n_share;    // group sets working on different items
  const int set = group % n_conc, share = group / n_conc;
  const int m = quarter * 32 + lane;
  const int g = m >> 3, px = m & 7;
  const int nchunks = a.Npad >> 4;
  const bool has_head = es.head_w != nullptr;
  const int Ho = a.ep.H_out, Wo = a.ep.W_out;
  // groups sharing an item split the 16-column chunks round-robin
  const int c_first = share, c_step = a.n_share;
  for (int it = set;; it += n_conc) {
    const int item = sc.get(it);
    if (item < 0) break;
    const int stage = (int)((it_base + (uint32_t)it) & (uint32_t)(a.n_stages - 1));
    const uint32_t round = (it_base + (uint32_t)it) >> a.stage_shift;   // parity of this TMEM stage
    const uint32_t t_full = t_full0 + 8u * stage, t_empty = t_empty0 + 8u * stage;
    const int phase = fdiv(item, a.fd_ipp);
    const int t0 = (item - phase * a.items_per_phase) * (a.MT * CG) + pid.rank;
    const int pr = fdiv(phase, a.fd_up), ps = phase - pr * S;
    mbar_wait(t_full, round & 1u, a.err_flag, 6);
    tc_fence_after();
    bool released = false;
    for (int mt = 0; mt < a.MT; ++mt) {
      const TileGeo tg = tile_geo(a, t0 + mt * CG, g, px);

      const bool valid = tg.valid;
      const int b = tg.b, oy = tg.y * S + pr, ox = tg.x * S + ps;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((stage * a.MT + mt) * a.Npad);
      float head[4] = {0.f, 0.f, 0.f, 0.f};
      const bool last_tile = (mt == a.MT - 1);
      if (CHAIN && a.ep.res1 && L->dep_res1.cnt) {       // residual written inside the chain (epi_geo == 0 only)
        if (lane == 0) dep_wait(L->dep_res1, tg.V0 + quarter * 4, tg.V0 + quarter * 4 + 4, r_vlo, r_vhi, a.err_flag, cx->exp_mult);
        __syncwarp();
      }
      if (a.fast) {
        TilePix tp;
        tp.valid = valid; tp.b = b; tp.oy = oy; tp.ox = ox;
        const size_t pix = valid ? ((size_t)(b * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1) : 0;
        tp.edge = 0;
        if (valid && a.ep.border == SRD_BORDER_REFLECT)
          tp.edge = (oy == 1 ? 1u : 0u) | (oy == Ho - 2 ? 2u : 0u) | (ox == 1 ? 4u : 0u) | (ox == Wo - 2 ? 8u : 0u);
        tp.out = a.ep.out ? static_cast<__nv_bfloat16*>(a.ep.out) + pix * a.ep.out_ct + a.ep.out_coff : nullptr;
        tp.out2 = a.ep.out2 ? static_cast<__nv_bfloat16*>(a.ep.out2) + pix * a.ep.out2_ct + a.ep.out2_coff : nullptr;
        tp.r1 = (a.ep.res1 && valid) ? static_cast<const __nv_bfloat16*>(a.ep.res1) + pix * a.ep.res1_ct + a.ep.res1_coff : nullptr;
        tp.r2 = (a.ep.res2 && valid) ? static_cast<const __nv_bfloat16*>(a.ep.res2) + pix * a.ep.res2_ct + a.ep.res2_coff : nullptr;
        const long long wp = (long long)(Wo + 2);
#pragma unroll 1
        for (int c = c_first; c < nchunks; c += c_step) {
          uint32_t r[16];
          tc_ld16(tbase + (uint32_t)(c * 16), r);
          uint4 q1[2] = {}, q2[2] = {};
          if (tp.r1) ld16_bf16(tp.r1 + 16 * c, q1);
          if (tp.r2) ld16_bf16(tp.r2 + 16 * c, q2);
          tc_ld_wait();
          if (last_tile && c + c_step >= nchunks) {              // accumulator fully read: release the TMEM stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(t_empty);
            released = true;
          }
          epi_chunk_fast<ACT>(a, es, c * 16, r, tp, wp, q1, q2, head);
        }
        if (has_head && valid) {
#pragma unroll
          for (int jh = 0; jh < 4; ++jh)
            if (jh < a.ep.planar_c)
              a.ep.planar[((size_t)(b * a.ep.planar_c + jh) * Ho + oy) * Wo + ox] = head[jh] + es.head_b[jh];
        }
      } else {
        // generic path (ragged N, per-image temb rows, planar output): one group per stage does all chunks
        for (int c0 = 0; c0 < (share == 0 ? a.Npad : 0); c0 += 16) {
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
#pragma unroll
              for (int jh = 0; jh < 4; ++jh) {
                if (jh >= a.ep.planar_c) break;
                float acc = head[jh];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nv) acc = fmaf(es.head_w[jh * a.Npad + c0 + i], v[i], acc);
                head[jh] = acc;
              }
            }
          }
        }
        if (has_head && valid && share == 0) {
#pragma unroll
          for (int jh = 0; jh < 4; ++jh)
            if (jh < a.ep.planar_c)
              a.ep.planar[((size_t)(b * a.ep.planar_c + jh) * Ho + oy) * Wo + ox] = head[jh] + es.head_b[jh];
        }
      }
      if (CHAIN && out_cnt && t0 + mt * CG < a.tiles_per_phase) {   // this warp's part of the tile is in global memory
        __syncwarp();
        if (lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(out_cnt + ((tg.V0 - a.vstart) >> 4)), "r"(1u) : "memory");
      }
    }
    if (!released) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader<CG>(t_empty);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged epilogue (stride-1, up == 1, N % 64 == 0, W % 8 == 0, NHWC output):
//   * the 128 px x 64 ch output sub-tile is assembled in a 128B-swizzled 16 KB shared-memory
//     buffer and written with cp.async.bulk.tensor stores (full 128 B lines, no partial sectors);
//   * the residual (res1) sub-tile is TMA-loaded into the SAME buffer ahead of time (while the MMAs
//     of the tile still run) and updated in place, so its HBM latency is off the critical path;
//   * reflect ring: mirrored ROWS are extra 1-row TMA stores of the same smem rows, mirrored
//     COLUMNS (first / last column tile only) are per-thread 16 B stores;
//   * zero ring (border == KEEP): masked rows are written as zeros by one full-box store.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

struct StageSmem {
  uint32_t base;        // shared address of THIS WARP's staging buffers ([nbuf][4 KB])
  uint8_t* gen;         // generic pointer to the same
  uint32_t r_full0;     // this warp's mbarrier for buffer 0 (buffer i at +8*i)
};

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// Each epilogue WARP owns 32 TMEM lanes = 4 tile rows x 8 px and stages them in its own 4 KB
// swizzled buffer(s): no cross-warp synchronisation, one lane issues the warp's TMA traffic.
template <int ACT, int CG, bool CHAIN = false>
__device__ __forceinline__ void epilogue_tma(const TcArgs& a, const TcMaps& m_, const EpiSmem& es, const StageSmem& sg,
                                             const Sched& sc, uint32_t tmem_base, uint32_t t_full0, uint32_t t_empty0,
                                             int warp, int lane, ChainRt* cx = nullptr, const ChainLayer* L = nullptr) {
#ifndef SRD_CHAIN_EPI_MIN
#define SRD_CHAIN_EPI_MIN 0       // experiment: chain epilogue without dependency / publishing code (1-layer timing only)
#endif
  constexpr bool PUB = CHAIN && !SRD_CHAIN_EPI_MIN;
  const uint32_t it_base = CHAIN ? cx->it_base : 0u;
  unsigned* const out_cnt = PUB ? L->out_cnt : nullptr;
  unsigned* const out2_cnt = PUB ? L->out2_cnt : nullptr;
  int r_vlo = 0, r_vhi = -1;                      // residual producer's row blocks already seen complete (lane 0)
  // Layer chain: the counter of the sub-tile this warp stored last, not yet bumped (lane 0).  Waiting for the bulk
  // stores to COMPLETE right after issuing them would stall the warp for a DRAM / L2 round trip per sub-tile, so
  // the bump is deferred to the next point where the warp either would idle anyway (its next accumulator is not
  // ready, a dependency has to be polled) or is about to issue the next store (the old one has long completed).
  // It is always flushed before a wait that another CTA's progress could depend on.
  unsigned* pend = nullptr;
  auto flush = [&]() { if (PUB && pend) { dep_publish(pend); pend = nullptr; } };
  const PairId pid = pair_id<CG>();
  const int S = a.up;
  const int quarter = warp & 3, group = (warp - WARP_EPI0) >> 2;
  const int n_conc = EPI_GROUPS / a.n_share;     // group sets working on different items
  const int set = group % n_conc, share = group / n_conc;
  const int m = quarter * 32 + lane;
  const int g = m >> 3, px = m & 7;
  const int nsj = a.Npad >> 6;                   // 64-channel sub-tiles per tile
  const int n_sub1 = a.MT * nsj;                 // sub-tiles of the (first) output per item
  const int n_sub = n_sub1 * (1 + a.dual);       // + those of the second output (dual)
  const int stage_cols = a.MT * a.Npad * (1 + a.dual);
  const int nbuf = a.epi_nbuf;
  const int Ho = a.ep.H_out, Wo = a.ep.W_out;
  const bool reflect = (a.ep.border == SRD_BORDER_REFLECT);
  const bool has_r1 = a.ep.res1 != nullptr;
  // block2 / resampling layers: no vector after the activation
  const bool has_temb = a.ep.temb != nullptr || (CHAIN && L->temb_off >= 0);
  const bool r1_packed = (a.ep.beta1 == 1.0f), r2_packed = (a.ep.beta2 == 1.0f);
  const uint32_t swz = (uint32_t)(lane & 7);     // == m & 7
  uint8_t* my_row = sg.gen + lane * 128;         // this pixel's 128 B row inside a staging buffer
  const int g0 = quarter * 4;                    // first tile row of this warp

  const uint32_t nbuf_mask = (uint32_t)nbuf - 1u, nbuf_sh = (uint32_t)nbuf >> 1;   // nbuf is 1 or 2
  // lane 0: make buffer (kn % nbuf) ready for this warp's rows of its kn-th sub-tile, which is
  // sub-tile qn of local item itn (the successor of the one just finished: no divisions)
  auto prepare = [&](uint32_t kn, int itn, int qn) {
    if (qn >= n_sub) { itn += n_conc; qn = share; }
    if (qn >= n_sub) return;                                   // this group never has a sub-tile
    const int item = sc.get(itn);                              // may be a later item of this group: see Sched
    if (item < 0) return;
    if (qn >= n_sub1) qn -= n_sub1;                            // second output: same tiles again
    const int mt = (qn >= nsj) ? 1 : 0, j = qn - mt * nsj;     // MT <= 2
    const int ph = fdiv(item, a.fd_ipp);
    const int tt = (item - ph * a.items_per_phase) * (a.MT * CG) + mt * CG + pid.rank;
    const int rowblk = fdiv(tt, a.fd_colblk), colblk = tt - rowblk * a.n_colblk;
    const uint32_t buf = kn & nbuf_mask;
    const uint32_t bar = sg.r_full0 + 8u * buf;
    if (has_r1) {
      const int Vr = rowblk * TILE_H + a.vstart + g0;
      if (PUB && L->dep_res1.cnt && !dep_cached(L->dep_res1, Vr, Vr + 4, r_vlo, r_vhi)) {
        flush();
        dep_wait(L->dep_res1, Vr, Vr + 4, r_vlo, r_vhi, a.err_flag, cx->exp_mult);
      }
      mbar_expect_tx(bar, 4096u);
      tma_load_3d(sg.base + buf * 4096u, &m_.tmR, bar, j * 64 + (a.nsplit > 1 ? ph * a.Npad : 0), colblk * TILE_W + 1, Vr);
    } else {
      mbar_arrive(bar);
    }
  };

  uint32_t k = CHAIN ? cx->k : 0u;
  if (lane == 0) prepare(k, set, share);
  for (int it = set;; it += n_conc) {
    const int item = sc.get(it);
    if (item < 0) break;
    const int stage = (int)((it_base + (uint32_t)it) & (uint32_t)(a.n_stages - 1));
    const uint32_t round = (it_base + (uint32_t)it) >> a.stage_shift;   // parity of this TMEM stage
    const uint32_t t_full = t_full0 + 8u * stage, t_empty = t_empty0 + 8u * stage;
    const int phase = fdiv(item, a.fd_ipp);
    const int t0 = (item - phase * a.items_per_phase) * (a.MT * CG) + pid.rank;
    int pr = 0, ps = 0;                                          // sub-pixel phase of a transposed conv
    if (S > 1) { pr = fdiv(phase, a.fd_up); ps = phase - pr * S; }
    const int choff = a.nsplit > 1 ? phase * a.Npad : 0;         // first output channel of this item's channel group
    const bool stamp = (quarter == 0 && lane == 0);
    if (stamp) epi_stamp(a, it, 0);
    if (PUB) {
      if (!mbar_try(t_full, round & 1u)) {           // would idle: publish the last sub-tile meanwhile
        if (lane == 0) flush();
        mbar_wait(t_full, round & 1u, a.err_flag, 6);
      }
    } else {
      mbar_wait(t_full, round & 1u, a.err_flag, 6);
    }
    tc_fence_after();
    if (stamp) epi_stamp(a, it, 1);
    if (quarter == 0 && lane == 0) dbg_stamp(a, it, 5);
    bool released = false;
    int mt_cur = -1;
    TileGeo tg;
    bool valid = false;
    size_t pix = 0;
    uint32_t edge = 0, tbase = 0;
    __nv_bfloat16* outp = nullptr;
    const __nv_bfloat16* r2p = nullptr;
    float lin[4] = {0.f, 0.f, 0.f, 0.f};                         // this pixel's linear-residual inputs
    {
      for (int qq = share; qq < n_sub; qq += a.n_share, ++k) {
        const int which = (qq >= n_sub1) ? 1 : 0;                // 0: conv output, 1: second (1x1) output
        const int q1 = qq - which * n_sub1;
        const int mt = (q1 >= nsj) ? 1 : 0, j = q1 - mt * nsj;
        if (mt != mt_cur || which) {
          mt_cur = mt;
          tg = tile_geo(a, t0 + mt * CG, g, px);
          valid = tg.valid;
          const int oy = tg.y * S + pr, ox = tg.x * S + ps;      // output pixel of this thread
          pix = valid ? ((size_t)(tg.b * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1) : 0;
          edge = 0;
          if (valid && reflect)
            edge = (oy == 1 ? 1u : 0u) | (oy == Ho - 2 ? 2u : 0u) | (ox == 1 ? 4u : 0u) | (ox == Wo - 2 ? 8u : 0u);
          outp = static_cast<__nv_bfloat16*>(a.ep.out) + pix * a.ep.out_ct + a.ep.out_coff + choff;
          r2p = (a.ep.res2 && valid) ? static_cast<const __nv_bfloat16*>(a.ep.res2) + pix * a.ep.res2_ct + a.ep.res2_coff + choff : nullptr;
          tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                  (uint32_t)(stage * stage_cols + which * n_sub1 * 64 + mt * a.Npad);
          if (which) { edge = 0; r2p = nullptr; mt_cur = -1; }     // plain output: no ring mirrors, no residuals
          if (es.lin_w) {
            uint2 xv = make_uint2(0u, 0u);                         // 4 bf16 channels of the pixel (buffers have ct % 8 == 0)
            if (valid) xv = *reinterpret_cast<const uint2*>(a.lin_src + pix * a.lin_ct);
            lin[0] = __uint_as_float(xv.x << 16); lin[1] = __uint_as_float(xv.x & 0xFFFF0000u);
            lin[2] = __uint_as_float(xv.y << 16); lin[3] = __uint_as_float(xv.y & 0xFFFF0000u);
          }
        }
        const uint32_t buf = k & nbuf_mask;
        uint8_t* row = my_row + buf * 4096u;
        mbar_wait(sg.r_full0 + 8u * buf, (k >> nbuf_sh) & 1u, a.err_flag, 8);   // buffer free / residual landed
        if (stamp && qq == share) epi_stamp(a, it, 2);
        const bool last_sub = (qq + a.n_share >= n_sub);
        // Compact hot loop (one 16-channel chunk per iteration, NOT unrolled: the body must stay
        // resident in the instruction cache; rare ring / second-residual work lives outside).
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = j * 64 + cc * 16;
          const int cb = c0 + which * a.Npad + choff;            // bias2 / zero temb are staged after bias / temb
          uint32_t r[16];
          tc_ld16(tbase + (uint32_t)c0, r);
          uint4* s0 = reinterpret_cast<uint4*>(row + (((uint32_t)(2 * cc) ^ swz) << 4));
          uint4* s1 = reinterpret_cast<uint4*>(row + (((uint32_t)(2 * cc + 1) ^ swz) << 4));
          const float4* sb = reinterpret_cast<const float4*>(es.bias + cb);
          const float4* st = reinterpret_cast<const float4*>(es.temb + cb);
          const float alpha = a.ep.alpha;
          tc_ld_wait();
          if (last_sub && cc == 3) {                             // accumulator fully read: release the TMEM stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(t_empty);
            if (quarter == 0 && lane == 0) dbg_stamp(a, it, 6);
            released = true;
          }
          float v[16];
          if (which) {                                           // second output: bias only
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = sb[q];
              v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b4.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
              v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = sb[q];
              const float4 t4 = has_temb ? st[q] : make_float4(0.f, 0.f, 0.f, 0.f);   // saves a shared-memory wavefront
              v[4 * q + 0] = act_axpy<ACT>(__uint_as_float(r[4 * q + 0]) + b4.x, alpha, t4.x);
              v[4 * q + 1] = act_axpy<ACT>(__uint_as_float(r[4 * q + 1]) + b4.y, alpha, t4.y);
              v[4 * q + 2] = act_axpy<ACT>(__uint_as_float(r[4 * q + 2]) + b4.z, alpha, t4.z);
              v[4 * q + 3] = act_axpy<ACT>(__uint_as_float(r[4 * q + 3]) + b4.w, alpha, t4.w);
            }
          }
          // Residuals.  beta == 1 (every Unet residual): packed bf16x2 adds on the already rounded result --
          // one HADD2.BF16 per two channels instead of an unpack + FFMA per channel; the extra rounding is
          // far inside the tolerance (tools/accuracy_report.py).  Otherwise fp32 FMAs before rounding.
          uint4 q20 = make_uint4(0, 0, 0, 0), q21 = q20;
          if (r2p) {
            const uint4* q = reinterpret_cast<const uint4*>(r2p + c0);
            q20 = q[0]; q21 = q[1];
          }
          if (has_r1 && !r1_packed) {
            float f[8];
            unpack_bf16x8(*s0, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(a.ep.beta1, f[i], v[i]);
            unpack_bf16x8(*s1, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 + i] = fmaf(a.ep.beta1, f[i], v[8 + i]);
          }
          if (es.lin_w) {                                        // 1x1 res_conv over <= 4 channels, weights [c][Npad] in smem
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < a.lin_c) {
                const float4* lw = reinterpret_cast<const float4*>(es.lin_w + c * a.Npad + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 w4 = lw[q];
                  v[4 * q + 0] = fmaf(w4.x, lin[c], v[4 * q + 0]); v[4 * q + 1] = fmaf(w4.y, lin[c], v[4 * q + 1]);
                  v[4 * q + 2] = fmaf(w4.z, lin[c], v[4 * q + 2]); v[4 * q + 3] = fmaf(w4.w, lin[c], v[4 * q + 3]);
                }
              }
            }
          }
          if (r2p && !r2_packed) {
            float f[8];
            unpack_bf16x8(q20, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(a.ep.beta2, f[i], v[i]);
            unpack_bf16x8(q21, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 + i] = fmaf(a.ep.beta2, f[i], v[8 + i]);
          }
          uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;            // masked rows become the zero ring (KEEP)
          if (valid) {
            o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
            o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
            o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
            o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
            if (has_r1 && r1_packed) { o0 = add_bf16x8(o0, *s0); o1 = add_bf16x8(o1, *s1); }
            if (r2p && r2_packed) { o0 = add_bf16x8(o0, q20); o1 = add_bf16x8(o1, q21); }
          }
          *s0 = o0; *s1 = o1;
        }
        if (edge & 12u) {                                        // mirrored ring COLUMNS (+ corners): rare, cold
          const long long dn = (long long)(Wo + 2) * a.ep.out_ct;
          const int ct = a.ep.out_ct;
          const bool top = edge & 1u, bot = edge & 2u, lef = edge & 4u, rig = edge & 8u;
#pragma unroll 1
          for (uint32_t u = 0; u < 8; ++u) {
            const uint4 val = *reinterpret_cast<const uint4*>(row + ((u ^ swz) << 4));
            uint4* p = reinterpret_cast<uint4*>(outp + j * 64 + u * 8);
            if (lef) *(p - (2 * ct) / 8) = val;
            if (rig) *(p + (2 * ct) / 8) = val;
            if (top && lef) *(p - (2 * dn + 2 * ct) / 8) = val;
            if (top && rig) *(p - (2 * dn - 2 * ct) / 8) = val;
            if (bot && lef) *(p + (2 * dn - 2 * ct) / 8) = val;
            if (bot && rig) *(p + (2 * dn + 2 * ct) / 8) = val;
          }
          if (PUB) __threadfence();                              // see dep_publish
        }
        if (stamp && last_sub) epi_stamp(a, it, 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy
        __syncwarp();
        if (stamp && last_sub) epi_stamp(a, it, 4);
        if (lane == 0) {
          flush();                                                 // (the previous sub-tile's stores have long completed)
          const uint32_t src = sg.base + buf * 4096u;
          const int V0 = tg.V0 + g0, xc = tg.x0 + 1;               // this warp's 4 rows
          const int jc = j * 64 + choff;                           // first channel of the sub-tile in the tensor
          if (which) {                                             // second output: zero ring, one 4-row box
            tma_store_3d(&m_.tmO2, src, jc, xc, V0);
          } else if (a.epi_geo == 0) {
            bool whole = !reflect;
            int yp0 = 0;
            if (reflect) {                                         // stride 1: virtual row == padded stacked row
              const int b0 = fdiv(V0, a.fd_period);
              yp0 = V0 - b0 * a.period;
              whole = (yp0 >= 3) && (yp0 + 3 <= Ho - 2) && (V0 + 4 <= a.n_vrows);
            }
            if (whole) {
              tma_store_3d(&m_.tmO, src, jc, xc, V0);
            } else {
              int yp = yp0;
              for (int gg = 0; gg < 4; ++gg, ++yp) {
                if (yp >= a.period) yp -= a.period;
                const int V = V0 + gg;
                if (V >= a.n_vrows) break;
                if (yp < 1 || yp > Ho) continue;                     // ring row: written by its mirror source
                const uint32_t rs = src + (uint32_t)gg * 1024u;
                tma_store_3d(&m_.tmOrow, rs, jc, xc, V);
                if (yp == 2) tma_store_3d(&m_.tmOrow, rs, jc, xc, V - 2);          // oy == 1 -> ring row 0
                if (yp == Ho - 1) tma_store_3d(&m_.tmOrow, rs, jc, xc, V + 2);     // oy == H-2 -> ring row H+1
              }
            }
          } else {
            // stride-2 conv / 2x transposed-conv phase: every tile row maps to its own output row
            const int Cc = tg.x0 * S + ps + 1;                     // first padded output column (step S)
            const int hp2 = (Ho + 2) >> 1;
            // common case: the warp's four tile rows are four consecutive valid rows of one image that own no
            // mirrored ring row -> ONE 4-row store (output rows are consecutive in the row / row-pair dimension)
            const int bb0 = fdiv(V0, a.fd_period), y0 = V0 - bb0 * a.period - a.voff;
            const int oy0 = y0 * S + pr, oy3 = (y0 + 3) * S + pr;
            if (y0 >= 0 && y0 + 3 < a.H_m && V0 + 3 < a.n_vrows && (!reflect || (oy0 > 1 && oy3 < Ho - 2))) {
              if (S == 1) tma_store_3d(&m_.tmO, src, jc, xc, bb0 * (Ho + 2) + oy0 + 1);
              else tma_store_5d(&m_.tmO, src, jc, Cc & 1, Cc >> 1, (oy0 + 1) & 1, bb0 * hp2 + ((oy0 + 1) >> 1));
            } else
            for (int gg = 0; gg < 4; ++gg) {
              const int V = V0 + gg;
              if (V >= a.n_vrows) break;
              const int bb = fdiv(V, a.fd_period);
              const int y = V - bb * a.period - a.voff;
              if (y < 0 || y >= a.H_m) continue;
              const int oy = y * S + pr;
              const uint32_t rs = src + (uint32_t)gg * 1024u;
              auto put = [&](int Rl) {                               // Rl: padded row inside image bb
                if (S == 1) tma_store_3d(&m_.tmOrow, rs, jc, xc, bb * (Ho + 2) + Rl);
                else tma_store_5d(&m_.tmOrow, rs, jc, Cc & 1, Cc >> 1, Rl & 1, bb * hp2 + (Rl >> 1));
              };
              put(oy + 1);
              if (reflect) {
                if (oy == 1) put(0);
                if (oy == Ho - 2) put(Ho + 1);
              }
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (stamp && last_sub) epi_stamp(a, it, 5);
          // the buffer of sub-tile k+1 was last read by the store of sub-tile k+1-nbuf
          if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (PUB) {         // layer chain: the readers of this tensor learn about these rows at the next flush()
            unsigned* const cnt = which ? out2_cnt : out_cnt;
            if (cnt && (t0 + mt * CG) < a.tiles_per_phase) pend = cnt + ((tg.V0 - a.vstart) >> 4);
          }
          if (stamp && last_sub) epi_stamp(a, it, 6);
          prepare(k + 1, it, qq + a.n_share);
          if (stamp && last_sub) epi_stamp(a, it, 7);
        }
        __syncwarp();
      }
    }
    if (!released) {                                             // this group had no sub-tile in the item
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader<CG>(t_empty);
    }
    if (quarter == 0 && lane == 0) dbg_stamp(a, it, 7);
  }
  // Before the CTA exits the staging buffers must have been READ by the TMA unit; the global writes themselves
  // complete with the grid (and before a dependent grid's griddepcontrol.wait returns).
  if (lane == 0) {
    flush();
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  if (CHAIN) cx->k = k;
}


// ------------------------------------------------------------------------------------------ roles
// barrier layout (8 B each) at bar_base: A_full[8] A_empty[8] B_full[8] B_empty[8] T_full[4] T_empty[4]
// tmem_slot B_res R_full[2 * EPI_WARPS]
struct Bars {
  uint32_t base;
  __device__ __forceinline__ uint32_t A_full(int s) const { return base + 8u * s; }
  __device__ __forceinline__ uint32_t A_empty(int s) const { return base + 8u * (MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t B_full(int s) const { return base + 8u * (2 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t B_empty(int s) const { return base + 8u * (3 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t T_full(int s) const { return base + 8u * (4 * MAX_STAGES + s); }
  __device__ __forceinline__ uint32_t T_empty(int s) const { return base + 8u * (4 * MAX_STAGES + 4 + s); }
  __device__ __forceinline__ uint32_t tmem_slot() const { return base + 8u * (4 * MAX_STAGES + 8); }
  __device__ __forceinline__ uint32_t B_res() const { return base + 8u * (4 * MAX_STAGES + 9); }
  __device__ __forceinline__ uint32_t R_full0() const { return base + 8u * (4 * MAX_STAGES + 10); }
  // layer chain only (static item order, so the scheduler ring's slots are free): "all MMAs of the layer done"
  __device__ __forceinline__ uint32_t M_done() const { return base + SCHED_BAR_OFF; }
};

// Activation producer (one thread).  Per (source, 64-channel chunk): ONE halo box per M tile
// (stride 1) or one box per (kx, row parity) (stride 2), in the order the MMA role consumes them.
// Weights have their own producer thread (below): a single in-order thread would park the next
// activation box behind the weight ring's back-pressure and expose its full HBM latency.
template <int CG, bool CHAIN = false>
__device__ __forceinline__ void producer_role(const TcArgs& a, const TcMaps& m_, uint32_t a_base, uint32_t bar_base,
                                              ChainRt* cx = nullptr, const ChainLayer* L = nullptr) {
  const Bars bars{bar_base};
  const PairId pid = pair_id<CG>();
  const bool leader = (pid.rank == 0);
  const int n_items = a.n_items, G = pid.n_pairs, MT = a.MT, S = a.up;
  const int n_src = a.n_src, n_colblk = a.n_colblk, ipp = a.items_per_phase, vstart = a.vstart;
  const int a_mode = a.a_mode, nkx = a.nkx, stride = a.stride;
  const uint32_t a_stage_bytes = (uint32_t)a.a_stage_bytes;
  int* err = a.err_flag;
  Ring ra{0, 0, a.a_stages};
  if (CHAIN) ra = cx->ra;
  int d_vlo[SRD_MAX_SRC], d_vhi[SRD_MAX_SRC];                    // layer chain: source row blocks already seen complete
#pragma unroll
  for (int s = 0; s < SRD_MAX_SRC; ++s) { d_vlo[s] = 0; d_vhi[s] = -1; }
  const Sched sc = make_sched<CG>(a, bar_base);
  const bool scheduler = sc.dyn && leader;                       // this thread hands out the pair's items
  // the first two items of a pair are static; the ticket drawn while item il is loaded names item il + 2, so an
  // entry is in the ring a whole item before the peer CTA's producer needs it
  int next_item = pid.pair, next2 = (pid.pair + G < n_items) ? pid.pair + G : -1;
  if (scheduler) { sc.publish<CG>(0, next_item); sc.publish<CG>(1, next2); }
  for (int il = 0;; ++il) {
    int item;
    unsigned ticket = 0;
    if (scheduler) {
      item = next_item;
      if (item < 0) break;
      ticket = atomicAdd(a.sched_counter, 1u);                   // consumed after this item's loads are issued
    } else {
      item = sc.get(il);
      if (item < 0) break;
    }
    const int phase = fdiv(item, a.fd_ipp);
    const int t0 = (item - phase * ipp) * (MT * CG) + pid.rank;
    int kx0 = a.kx0;
    if (S > 1) { const int pr = fdiv(phase, a.fd_up); kx0 = (phase - pr * S >= S / 2) ? 1 : 0; }
    // tile origins of this item
    int V0[2], X0[2];
    for (int mt = 0; mt < MT; ++mt) {
      const int tt = t0 + mt * CG;
      const int rowblk = fdiv(tt, a.fd_colblk);
      V0[mt] = rowblk * TILE_H + vstart;
      X0[mt] = (tt - rowblk * n_colblk) * TILE_W;
    }
    if (CHAIN) {
      // the rows this item's boxes cover must have been written by the layers that produce them
#pragma unroll
      for (int s = 0; s < SRD_MAX_SRC; ++s) {
        if (s < n_src && L->dep_src[s].cnt) {
          for (int mt = 0; mt < MT; ++mt) {
            if (stride == 1) dep_wait(L->dep_src[s], V0[mt] - 1, V0[mt] + TILE_H + 1, d_vlo[s], d_vhi[s], err, cx->exp_mult);
            else dep_wait(L->dep_src[s], 2 * V0[mt], 2 * V0[mt] + 2 * BOX_ROWS, d_vlo[s], d_vhi[s], err, cx->exp_mult);
          }
        }
      }
    }
    for (int s = 0; s < n_src; ++s) {
      const CUtensorMap* tm = &m_.tmA[s];
      const int nch = a.nchunk[s];
      for (int ch = 0; ch < nch; ++ch) {
        if (a_mode) {
          mbar_wait(bars.A_empty(ra.stage), ra.phase ^ 1, err, 1);
          if (s == 0 && ch == 0) dbg_stamp(a, il, 0);
          if (leader) mbar_expect_tx(bars.A_full(ra.stage), (uint32_t)(CG * MT * A_HALO_BYTES));
          for (int mt = 0; mt < MT; ++mt)
            tma_load_3d<CG>(a_base + (uint32_t)(ra.stage * MT + mt) * a_stage_bytes, tm, bars.A_full(ra.stage),
                            ch * 64, X0[mt], V0[mt] - 1);
          ra.next();
          continue;
        }
        for (int kxi = 0; kxi < nkx; ++kxi) {
          const int kxp = kx0 + kxi;
          for (int ld = 0; ld < n_loads(a); ++ld) {
            mbar_wait(bars.A_empty(ra.stage), ra.phase ^ 1, err, 1);
            if (leader) mbar_expect_tx(bars.A_full(ra.stage), (uint32_t)(CG * MT * A_BYTES));
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t dst = a_base + (uint32_t)(ra.stage * MT + mt) * a_stage_bytes;
              if (stride == 1)
                tma_load_3d<CG>(dst, tm, bars.A_full(ra.stage), ch * 64, X0[mt] + kxp, V0[mt] - 1);
              else
                tma_load_5d<CG>(dst, tm, bars.A_full(ra.stage), ch * 64, kxp & 1, X0[mt] + (kxp >> 1), ld, V0[mt]);
            }
            ra.next();
          }
        }
      }
    }
    if (scheduler) {
      const bool ended = next2 < 0;                                       // entry il + 1 was already the end marker
      next_item = next2;
      next2 = 2 * G + (int)ticket;
      if (next2 >= n_items || ended) next2 = -1;
      if (ticket == (unsigned)(n_items - 1)) *a.sched_counter = 0u;     // the launch's last draw: ready for the next one
      if (!ended) {
        sc.publish<CG>(il + 2, next2);
        if (next2 < 0)
          for (int k = 3; k <= 10; ++k) sc.publish<CG>(il + k, -1);       // look-ahead room of the residual prefetch
      }
    }
  }
  if (CHAIN) cx->ra = ra;
}

// Weight producer (one thread): either all k-blocks once (resident) or, per item, kb_per_phase
// k-blocks of the item's phase streamed through the weight ring.  Each CTA of a pair holds
// Npad / CG rows of every k-block; the leader's barriers count the bytes of BOTH CTAs.
template <int CG, bool CHAIN = false>
__device__ __forceinline__ void weight_producer_role(const TcArgs& a, const TcMaps& m_, uint32_t b_base, uint32_t bar_base,
                                                     ChainRt* cx = nullptr) {
  const Bars bars{bar_base};
  const PairId pid = pair_id<CG>();
  const bool leader = (pid.rank == 0);
  const int Npad = a.Npad, b_rows = Npad / CG, b_row0 = pid.rank * b_rows;
  const uint32_t B_BYTES = (uint32_t)b_rows * 128u;
  int* err = a.err_flag;
  if (a.b_resident) {
    // layer chain: the previous layer of the segment has the resident region until its last MMA has completed
    if (CHAIN && cx->j > 0) mbar_wait(bars.M_done(), (uint32_t)(cx->j - 1) & 1u, err, 12);
    if (leader) mbar_expect_tx(bars.B_res(), (uint32_t)a.kb_total * B_BYTES * CG);
    for (int kb = 0; kb < a.kb_total; ++kb)
      tma_load_2d<CG>(b_base + (uint32_t)kb * B_BYTES, &m_.tmB, bars.B_res(), 0, kb * Npad + b_row0);
    return;
  }
  const int kb_per_phase = a.kb_per_phase;
  Ring rb{0, 0, a.b_stages};
  if (CHAIN) rb = cx->rb;
  const Sched sc = make_sched<CG>(a, bar_base);
  for (int il = 0;; ++il) {
    const int item = sc.get(il);
    if (item < 0) break;
    int kb = (a.up > 1 || a.nsplit > 1) ? fdiv(item, a.fd_ipp) * kb_per_phase : 0;
    for (int j = 0; j < kb_per_phase; ++j, ++kb) {
      mbar_wait(bars.B_empty(rb.stage), rb.phase ^ 1, err, 2);
      if (leader) mbar_expect_tx(bars.B_full(rb.stage), B_BYTES * CG);
      tma_load_2d<CG>(b_base + (uint32_t)rb.stage * B_BYTES, &m_.tmB, bars.B_full(rb.stage), 0, kb * Npad + b_row0);
      rb.next();
    }
  }
  if (CHAIN) cx->rb = rb;
}

// MMA issuer (one thread).  One tap = MT x 4 MMAs (K = 64) against one weight k-block.  The issuing
// thread is a critical resource: a K = 16 MMA occupies the tensor pipe for only max(N/2, 32 + N/4)
// cycles (measured, tools/umma_issue.cu), and this thread shares its scheduler with four epilogue
// warps.  So: every launch constant lives in a register, descriptor arithmetic is 32-bit adds on
// the low word (high words are constants), and for the 3x3 / 2x2 / 1x1 halo cases (TAPS = 9/4/1)
// the tap loops are unrolled so tap offsets are immediates.  TAPS = 0: generic runtime loops
// (stride 2, validation modes).
template <int TAPS, int MT, bool BRES, int CG, bool DUAL = false, bool CHAIN = false>
__device__ __forceinline__ void mma_role(const TcArgs& a, uint32_t tmem_base, uint32_t a_base, uint32_t b_base,
                                         uint32_t bar_base, int issuer = 0, ChainRt* cx = nullptr) {
  const Bars bars{bar_base};
  const PairId pid = pair_id<CG>();                  // only the leader CTA of a pair runs this role
  const int S = a.up, n_chunks = a.n_chunks;
  const uint32_t Npad = (uint32_t)a.Npad, idesc = a.idesc;
  const uint32_t bB16 = Npad * 8u / CG, aS16 = (uint32_t)a.a_stage_bytes >> 4;   // per-CTA weight tile = Npad/CG rows
  const uint32_t a_lo_base = a_base >> 4, b_lo_base = b_base >> 4;
  const uint32_t kbpp16 = (uint32_t)a.kb_per_phase * bB16;
  int* err = a.err_flag;
  Ring ra{0, 0, a.a_stages}, rb{0, 0, a.b_stages}, rt{0, 0, a.n_stages};
  if (CHAIN) { ra = cx->ra; rb = cx->rb; rt = cx->rt; }
  if (BRES) {
    mbar_wait(bars.B_res(), CHAIN ? ((uint32_t)cx->j & 1u) : 0u, err, 7);   // chain: one phase per resident layer
    tc_fence_after();
  }
  // SRD_DUAL_ISSUE: two issuing threads (different schedulers) take alternate items, skipping each other's
  // ring slots.  Only sound when every slot a thread waits on was last used by the SAME thread (a parity
  // wait must never run a whole phase ahead of its barrier): resident weights, one A slot per item and an
  // even A ring -- the host sets dual_issue accordingly.
  const int dual = SRD_DUAL_ISSUE ? a.dual_issue : 0;
  if (!dual && issuer != 0) return;
  int mma_il = 0;
  const Sched sc = make_sched<CG>(a, bar_base);
  for (int sched_il = 0;; ++sched_il) {
    const int item = sc.get(sched_il);
    if (item < 0) break;
    if (dual && (mma_il & 1) != issuer) {
      ra.skip(n_chunks);
      rt.next();
      ++mma_il;
      continue;
    }
    int ky0 = a.ky0, kx0 = a.kx0;
    uint32_t b_cur = b_lo_base;
    if (S > 1) {
      const int phase = fdiv(item, a.fd_ipp);
      const int pr = fdiv(phase, a.fd_up), ps = phase - pr * S;
      ky0 = (pr >= S / 2) ? 1 : 0;
      kx0 = (ps >= S / 2) ? 1 : 0;
      b_cur += (uint32_t)phase * kbpp16;
    } else if (a.nsplit > 1) {
      b_cur += (uint32_t)fdiv(item, a.fd_ipp) * kbpp16;      // resident weights: the item's channel group
    }
    dbg_stamp(a, mma_il, 1);
    mbar_wait(bars.T_empty(rt.stage), rt.phase ^ 1, err, 3);
    tc_fence_after();
    dbg_stamp(a, mma_il, 2);
    const uint32_t d0 = tmem_base + (uint32_t)(rt.stage * MT * (DUAL ? 2 : 1)) * Npad;
    uint32_t kcount = 0;     // 0 only for the first k-block of the item (accumulator overwrite)
    uint32_t kcount2 = 0;    // same for the second output's accumulator (DUAL)
    // DUAL: the chunk's 10th weight k-block (1x1 conv) against the centre tap, into the accumulator
    // MT*Npad columns further
    auto issue_tap2 = [&](uint32_t a_lo, uint32_t a_hi) {
      uint32_t b_lo;
      if (BRES) {
        b_lo = b_cur;
        b_cur += bB16;
      } else {
        mbar_wait(bars.B_full(rb.stage), rb.phase, err, 5);
        tc_fence_after();
        b_lo = b_lo_base + (uint32_t)rb.stage * bB16;
      }
      const uint32_t e0 = d0 + (uint32_t)MT * Npad;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t e = e0 + (uint32_t)mt * Npad, am = a_lo + (uint32_t)mt * aS16;
        tc_mma_bf16<CG>(e, mk_desc(am, a_hi), mk_desc(b_lo, DESC_HI_1024), idesc, kcount2);
        tc_mma_bf16<CG>(e, mk_desc(am + 2, a_hi), mk_desc(b_lo + 2, DESC_HI_1024), idesc, 1u);
        tc_mma_bf16<CG>(e, mk_desc(am + 4, a_hi), mk_desc(b_lo + 4, DESC_HI_1024), idesc, 1u);
        tc_mma_bf16<CG>(e, mk_desc(am + 6, a_hi), mk_desc(b_lo + 6, DESC_HI_1024), idesc, 1u);
      }
      if (!BRES) {
        tc_commit<CG>(bars.B_empty(rb.stage));
        rb.next();
      }
      kcount2 = 1u;
    };
    auto issue_tap = [&](uint32_t a_lo, uint32_t a_hi) {
      uint32_t b_lo;
      if (BRES) {
        b_lo = b_cur;
        b_cur += bB16;
      } else {
        mbar_wait(bars.B_full(rb.stage), rb.phase, err, 5);
        tc_fence_after();
        b_lo = b_lo_base + (uint32_t)rb.stage * bB16;
      }
      tc_mma_bf16<CG>(d0, mk_desc(a_lo, a_hi), mk_desc(b_lo, DESC_HI_1024), idesc, kcount);
      tc_mma_bf16<CG>(d0, mk_desc(a_lo + 2, a_hi), mk_desc(b_lo + 2, DESC_HI_1024), idesc, 1u);
      tc_mma_bf16<CG>(d0, mk_desc(a_lo + 4, a_hi), mk_desc(b_lo + 4, DESC_HI_1024), idesc, 1u);
      tc_mma_bf16<CG>(d0, mk_desc(a_lo + 6, a_hi), mk_desc(b_lo + 6, DESC_HI_1024), idesc, 1u);
      if (MT == 2) {
        const uint32_t d1 = d0 + Npad, a1 = a_lo + aS16;
        tc_mma_bf16<CG>(d1, mk_desc(a1, a_hi), mk_desc(b_lo, DESC_HI_1024), idesc, kcount);
        tc_mma_bf16<CG>(d1, mk_desc(a1 + 2, a_hi), mk_desc(b_lo + 2, DESC_HI_1024), idesc, 1u);
        tc_mma_bf16<CG>(d1, mk_desc(a1 + 4, a_hi), mk_desc(b_lo + 4, DESC_HI_1024), idesc, 1u);
        tc_mma_bf16<CG>(d1, mk_desc(a1 + 6, a_hi), mk_desc(b_lo + 6, DESC_HI_1024), idesc, 1u);
      }
      if (!BRES) {
        tc_commit<CG>(bars.B_empty(rb.stage));
        rb.next();
      }
      kcount = 1u;
    };
    if (TAPS > 0) {
      constexpr int NK = TAPS == 9 ? 3 : (TAPS == 4 ? 2 : 1);
      const uint32_t tap0 = (uint32_t)(ky0 * HALO_W + kx0) * 8u;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(bars.A_full(ra.stage), ra.phase, err, 4);
        tc_fence_after();
        if (c == 0) dbg_stamp(a, mma_il, 4);
        const uint32_t st_lo = a_lo_base + (uint32_t)(ra.stage * MT) * aS16 + tap0;
#pragma unroll
        for (int kxi = 0; kxi < NK; ++kxi)
#pragma unroll
          for (int kyi = 0; kyi < NK; ++kyi)
            issue_tap(st_lo + (uint32_t)(kyi * HALO_W + kxi) * 8u, DESC_HI_HALO);
        if (DUAL) issue_tap2(a_lo_base + (uint32_t)(ra.stage * MT) * aS16 + (uint32_t)(HALO_W + 1) * 8u, DESC_HI_HALO);
        tc_commit<CG>(bars.A_empty(ra.stage));
        ra.next();
      }
    } else {
      const int a_mode = a.a_mode, nkx = a.nkx, nky = a.nky;
      for (int c = 0; c < n_chunks; ++c) {
        if (a_mode) {
          mbar_wait(bars.A_full(ra.stage), ra.phase, err, 4);
          tc_fence_after();
          const uint32_t st_lo = a_lo_base + (uint32_t)(ra.stage * MT) * aS16 + (uint32_t)(ky0 * HALO_W + kx0) * 8u;
          for (int kxi = 0; kxi < nkx; ++kxi)
            for (int kyi = 0; kyi < nky; ++kyi)
              issue_tap(st_lo + (uint32_t)(kyi * HALO_W + kxi) * 8u, DESC_HI_HALO);
          tc_commit<CG>(bars.A_empty(ra.stage));
          ra.next();
          continue;
        }
        for (int kxi = 0; kxi < nkx; ++kxi) {
          for (int ld = 0; ld < n_loads(a); ++ld) {
            mbar_wait(bars.A_full(ra.stage), ra.phase, err, 4);
            tc_fence_after();
            const uint32_t st_lo = a_lo_base + (uint32_t)(ra.stage * MT) * aS16;
            const int nt = n_taps(a, ld, ky0);
            for (int j = 0; j < nt; ++j)
              issue_tap(st_lo + (uint32_t)tap_off(a, ld, j, ky0) * (TILE_W * 8u), DESC_HI_1024);
            tc_commit<CG>(bars.A_empty(ra.stage));
            ra.next();
          }
        }
      }
    }
    tc_commit<CG>(bars.T_full(rt.stage));
    dbg_stamp(a, mma_il, 3);
    ++mma_il;
    rt.next();
  }
  if (CHAIN) {
    // "every MMA of this layer has completed": the next layer's weight producer may overwrite resident weights, and
    // at the end of a segment this thread itself waits for it, so that no tcgen05.commit arrival is still in flight
    // when the barriers are re-initialised for the next segment
    tc_commit<CG>(bars.M_done());
    if (cx->j + 1 == cx->n_seg) mbar_wait(bars.M_done(), (uint32_t)cx->j & 1u, err, 13);
    cx->ra = ra; cx->rb = rb; cx->rt = rt;
  }
}

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
      const int taps = a.a_mode ? a.nkx * a.nky : 0;
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
          const int taps = a.a_mode ? a.nkx * a.nky : 0;
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
#define SRD_SEG_K 3
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

template <int CG, int K>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_seg_kernel(const __grid_constant__ SegParams<K> P) {
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
          producer_role<CG, true>(P.layers[j].a, P.maps[j], a_base, bar_base, &cx, &P.layers[j]);
        }
      }
    }
  } else if (warp == WARP_TMB) {
    if (elect_one_sync()) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < n) {
          cx.j = j;
          weight_producer_role<CG, true>(P.layers[j].a, P.maps[j], b_base, bar_base, &cx);
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
          const int taps = a.a_mode ? a.nkx * a.nky : 0;
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
            epilogue_tma<SRD_ACT_MISH, CG, true>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
          else
            epilogue_tma<SRD_ACT_NONE, CG, true>(a, m_, es, sg, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        } else if (a.ep.act == SRD_ACT_MISH)
          epilogue_role<SRD_ACT_MISH, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        else if (a.ep.act == SRD_ACT_LRELU02)
          epilogue_role<SRD_ACT_LRELU02, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
        else
          epilogue_role<SRD_ACT_NONE, CG, true>(a, es, sc, tmem_base, bars.T_full(0), bars.T_empty(0), warp, lane, &cx, &P.layers[j]);
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

struct TcPlan {
  TcMaps maps;
  TcArgs args;
  int grid, smem;
  unsigned* counter = nullptr;           // ticket counter of the dynamic item scheduler (device memory)
  ~TcPlan() { if (counter) cudaFree(counter); }
};

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
  a.a_mode = (d.stride == 1) ? g_a_mode : 0;
  a.a_stage_bytes = a.a_mode ? A_HALO_STAGE : A_BYTES;
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
        a.a_stages = ast > 4 ? 4 : ast;
        a.b_stages = 1;
        b_bytes_total = (int)res_bytes;
      } else {
        // deepest A ring that still leaves >= 4 weight stages; at least 2 (halo) / 3 (per-kx) A stages and 3 weight stages
        const int ast_min = a.a_mode ? 2 : 3;
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
      cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_seg_kernel<1, SEG_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv_seg_kernel<2, SEG_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
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

// ------------------------------------------------------------------------------------------ layer chain: host
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
    SRD_CHECK(!a.dyn_sched && !a.dbg, "conv_chain: layer %d uses the dynamic scheduler / debug stamps", i);
    SRD_CHECK(a.up == 1 || a.up == 2, "conv_chain: layer %d has up = %d (only 1 and 2 can be chained)", i, a.up);
    SRD_CHECK(!(a.head_w || a.lin_w) || a.Npad <= 128, "conv_chain: layer %d: fused head / linear residual need N <= 128", i);
    SRD_CHECK(!a.dual || a.Npad <= 128, "conv_chain: layer %d: a second output needs N <= 128 inside a chain", i);
    n_rowblk[i] = (a.n_vrows - a.vstart + TILE_H - 1) / TILE_H;
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
  SRD_CHECK(tc_set_func_attrs(dev) == 0, "conv_chain: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
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
  if (g_pdl) {
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
    if (c.cg == 1) SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_seg_kernel<1, SEG_K>, S));
    else SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_seg_kernel<2, SEG_K>, S));
    return 0;
  }
  ChainParams P = c.params;
  P.temb_row = temb_row;
  if (c.cg == 1) SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<1>, P));
  else SRD_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<2>, P));
  return 0;
}

}  // namespace srd
