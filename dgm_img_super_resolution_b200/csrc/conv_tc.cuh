// Shared device code of the tcgen05 convolution kernels (conv_tc.cu: one layer per launch; conv_chain.cu: layer
// chains and fused segments): argument structures, PTX wrappers, and the four roles (activation-TMA producer,
// weight-TMA producer, MMA issuer, epilogue warps).  See conv_tc.cu for the design overview.
#pragma once
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
// stride-2 convs (a_mode 2): one halo box per (input row parity, input column parity) of the parity-split view --
// {64ch x 9 half-columns x 17 half-rows}; tap (ky, kx) reads box (ky & 1, kx & 1) at half-row / half-column offset
// (ky >> 1, kx >> 1).  Four boxes (78 KB) per 64-channel chunk and M tile instead of six (kx, row parity) boxes
// (110 KB).
constexpr int S2_W = TILE_W + 1, S2_ROWS = TILE_H + 1;
constexpr int A_S2_BYTES = S2_ROWS * S2_W * 128;           // 19584
constexpr int A_S2_STAGE = (A_S2_BYTES + 1023) / 1024 * 1024;
constexpr int TAPS_S2 = 18;                                // mma_role<TAPS> code of this mode
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
  int a_mode;              // 2: stride-2 halo boxes per (row parity, column parity), see A_S2_BYTES;
                           // 0: one box per (chunk, kx), ky by 1024 B row offsets (always-aligned descriptors)
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
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* err_flag, int code) {
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
static __device__ __noinline__ void dep_poll(const unsigned* p, unsigned expect, int* err_flag) {
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
constexpr uint32_t DESC_HI_S2 = ((uint32_t)(TILE_W + 1) * 128u >> 4) | (1u << 14) | (2u << 29);
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
  // Chain state is kept to TWO registers across the hot chunk loop (the epilogue sits at the kernel's register cap):
  // `pend` = row block + 1 of the sub-tile this warp stored last and has not published yet (bit 30: second output;
  // 0 = nothing pending) -- the counter arrays are re-read from the layer arguments when the bump happens -- and the
  // [lo, hi] interval of residual row blocks already seen complete, packed into one word (lane 0 only).
  // Waiting for the bulk stores to COMPLETE right after issuing them would stall the warp for a DRAM / L2 round trip
  // per sub-tile, so the bump is deferred to the next point where the warp either would idle anyway (its next
  // accumulator is not ready, a dependency has to be polled) or starts its next sub-tile.  It is always flushed
  // before a wait that another CTA's progress could depend on.
  uint32_t pend = 0u;
  uint32_t r_seen = 0x0000FFFFu;                  // (hi << 16) | lo; lo = 0xFFFF marks the empty interval
  auto flush = [&]() {
    if (PUB && pend) {
      unsigned* const base = (pend & 0x40000000u) ? L->out2_cnt : L->out_cnt;
      dep_publish(base + ((pend & 0x3FFFFFFFu) - 1u));
      pend = 0u;
    }
  };
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
      if (PUB && L->dep_res1.cnt) {
        int r_vlo = (int)(r_seen & 0xFFFFu), r_vhi = (int)(r_seen >> 16);
        if (r_vlo == 0xFFFF) { r_vlo = 0; r_vhi = -1; }
        if (!dep_cached(L->dep_res1, Vr, Vr + 4, r_vlo, r_vhi)) {
          flush();
          dep_wait(L->dep_res1, Vr, Vr + 4, r_vlo, r_vhi, a.err_flag, cx->exp_mult);
          r_seen = ((uint32_t)r_vhi << 16) | (uint32_t)r_vlo;
        }
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
    bool released = false;
    int mt_cur = -1;
    TileGeo tg;
    bool valid = false;
    size_t pix = 0;
    uint32_t edge = 0, tbase = 0;
    __nv_bfloat16* outp = nullptr;
    const __nv_bfloat16* r2p = nullptr;
    float lin[4] = {0.f, 0.f, 0.f, 0.f};                         // this pixel's linear-residual inputs
    // per-tile constants of this thread (output pixel, pointers, ring flags, TMEM address): ~80 dependent integer
    // instructions, about 1000 cycles next to three other epilogue warps on the same scheduler
    auto setup_tile = [&](int mt, int which) {
      mt_cur = mt;
      tg = tile_geo(a, t0 + mt * CG, g, px);
      valid = tg.valid;
      const int oy = tg.y * S + pr, ox = tg.x * S + ps;          // output pixel of this thread
      pix = valid ? ((size_t)(tg.b * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1) : 0;
      edge = 0;
      if (valid && reflect)
        edge = (oy == 1 ? 1u : 0u) | (oy == Ho - 2 ? 2u : 0u) | (ox == 1 ? 4u : 0u) | (ox == Wo - 2 ? 8u : 0u);
      outp = static_cast<__nv_bfloat16*>(a.ep.out) + pix * a.ep.out_ct + a.ep.out_coff + choff;
      r2p = (a.ep.res2 && valid) ? static_cast<const __nv_bfloat16*>(a.ep.res2) + pix * a.ep.res2_ct + a.ep.res2_coff + choff : nullptr;
      tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) +
              (uint32_t)(stage * stage_cols + which * n_sub1 * 64 + mt * a.Npad);
      if (which) { edge = 0; r2p = nullptr; mt_cur = -1; }       // plain output: no ring mirrors, no residuals
      if (es.lin_w) {
        uint2 xv = make_uint2(0u, 0u);                           // 4 bf16 channels of the pixel (buffers have ct % 8 == 0)
        if (valid) xv = *reinterpret_cast<const uint2*>(a.lin_src + pix * a.lin_ct);
        lin[0] = __uint_as_float(xv.x << 16); lin[1] = __uint_as_float(xv.x & 0xFFFF0000u);
        lin[2] = __uint_as_float(xv.y << 16); lin[3] = __uint_as_float(xv.y & 0xFFFF0000u);
      }
    };
    // SRD_EPI_SETUP_EARLY=1 works the first sub-tile's constants out BEFORE the accumulator is waited for (the warp
    // idles in that wait in the MMA-bound layers, and a launch's drain starts ~1000 cycles earlier).  Measured on
    // B200, three alternating runs each: 230.1-230.6 images/s against 231.4-232.3 without it -- off.
#ifndef SRD_EPI_SETUP_EARLY
#define SRD_EPI_SETUP_EARLY 0
#endif
    if (SRD_EPI_SETUP_EARLY && share < n_sub) {
      const int which = (share >= n_sub1) ? 1 : 0;
      const int q1 = share - which * n_sub1;
      setup_tile((q1 >= nsj) ? 1 : 0, which);
    }
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
    {
      bool first_sub = true;
      for (int qq = share; qq < n_sub; qq += a.n_share, ++k) {
        const int which = (qq >= n_sub1) ? 1 : 0;                // 0: conv output, 1: second (1x1) output
        const int q1 = qq - which * n_sub1;
        const int mt = (q1 >= nsj) ? 1 : 0, j = q1 - mt * nsj;
        if ((first_sub && !SRD_EPI_SETUP_EARLY) || (!first_sub && (mt != mt_cur || which))) setup_tile(mt, which);
        first_sub = false;
        const uint32_t buf = k & nbuf_mask;
        uint8_t* row = my_row + buf * 4096u;
#ifdef SRD_EPI_STAMP_SPLIT
        if (stamp && qq == share) epi_stamp(a, it, 2);          // diagnosis: stamp 2 BEFORE the buffer wait
#endif
        mbar_wait(sg.r_full0 + 8u * buf, (k >> nbuf_sh) & 1u, a.err_flag, 8);   // buffer free / residual landed
#ifndef SRD_EPI_STAMP_SPLIT
        if (stamp && qq == share) epi_stamp(a, it, 2);
#endif
        // Layer chain: the previous sub-tile's stores were issued a T_full / buffer wait ago -- publish them NOW, before
        // this sub-tile's arithmetic: waiting for the next store point instead delays the consumer layer's first
        // items by a whole sub-tile (measured: a 10 us bubble per layer at 20x20, profiles/r02)
        if (PUB && lane == 0) flush();
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
            if ((which ? L->out2_cnt : L->out_cnt) && (t0 + mt * CG) < a.tiles_per_phase)
              pend = (uint32_t)(((tg.V0 - a.vstart) >> 4) + 1) | (which ? 0x40000000u : 0u);
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
    // tile origins (and first kx) of an item
    auto origins = [&](int it_, int (&V0_)[2], int (&X0_)[2], int& kx0_) {
      const int phase = fdiv(it_, a.fd_ipp);
      const int t0 = (it_ - phase * ipp) * (MT * CG) + pid.rank;
      kx0_ = a.kx0;
      if (S > 1) { const int pr = fdiv(phase, a.fd_up); kx0_ = (phase - pr * S >= S / 2) ? 1 : 0; }
      for (int mt = 0; mt < MT; ++mt) {
        const int tt = t0 + mt * CG;
        const int rowblk = fdiv(tt, a.fd_colblk);
        V0_[mt] = rowblk * TILE_H + vstart;
        X0_[mt] = (tt - rowblk * n_colblk) * TILE_W;
      }
    };
    int V0[2], X0[2], kx0;
    origins(item, V0, X0, kx0);
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
        if (a_mode == 2) {
          for (int p = 0; p < 4; ++p) {                      // (row parity, column parity) = (p >> 1, p & 1)
            mbar_wait(bars.A_empty(ra.stage), ra.phase ^ 1, err, 1);
            if (s == 0 && ch == 0 && p == 0) dbg_stamp(a, il, 0);
            if (leader) mbar_expect_tx(bars.A_full(ra.stage), (uint32_t)(CG * MT * A_S2_BYTES));
            for (int mt = 0; mt < MT; ++mt)
              tma_load_5d<CG>(a_base + (uint32_t)(ra.stage * MT + mt) * a_stage_bytes, tm, bars.A_full(ra.stage),
                              ch * 64, p & 1, X0[mt], p >> 1, V0[mt]);
            ra.next();
          }
          continue;
        }
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
    if (TAPS == TAPS_S2) {
      // stride-2 halo boxes; tap order == weight k-block order (engine._S2_TAPS):
      // box (even rows, even cols): (0,0) (0,2) (2,0) (2,2); (even, odd): (0,1) (2,1); (odd, even): (1,0) (1,2); (odd, odd): (1,1)
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          mbar_wait(bars.A_full(ra.stage), ra.phase, err, 4);
          tc_fence_after();
          if (c == 0 && p == 0) dbg_stamp(a, mma_il, 4);
          const uint32_t st_lo = a_lo_base + (uint32_t)(ra.stage * MT) * aS16;
#pragma unroll
          for (int hy = 0; hy < ((p >> 1) ? 1 : 2); ++hy)
#pragma unroll
            for (int hx = 0; hx < ((p & 1) ? 1 : 2); ++hx)
              issue_tap(st_lo + (uint32_t)(hy * S2_W + hx) * 8u, DESC_HI_S2);
          tc_commit<CG>(bars.A_empty(ra.stage));
          ra.next();
        }
      }
    } else if (TAPS > 0) {
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

// One layer plan: tensor maps + scalar arguments + launch geometry (built by tc_plan_build, conv_tc.cu).
struct TcPlan {
  TcMaps maps;
  TcArgs args;
  int grid, smem;
  unsigned* counter = nullptr;           // ticket counter of the dynamic item scheduler (device memory)
  ~TcPlan() { if (counter) cudaFree(counter); }
};

int tc_s2_halo();              // conv_tc.cu: stride-2 A-operand mode (SRD_S2_HALO, default 1)
int tc_pdl_enabled();          // conv_tc.cu: programmatic dependent launch switch (srd_tc_set_mode bit 3)

}  // namespace srd
