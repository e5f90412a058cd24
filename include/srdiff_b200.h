/*
 * srdiff_b200.h -- C ABI of libsrdiff_b200.so (sm_100a CUDA kernels for the SRDiff
 * reverse-diffusion inference hot path).
 *
 * The reference (aryaman1802/dgm_img_super_resolution) has NO native/FFI layer: its
 * boundary is the Python module API of srdiff/model.py.  This header is therefore the
 * interface that the Python mirror of that API (dgm_img_super_resolution_b200/model.py)
 * binds with ctypes; each entry point names the reference code whose arithmetic it
 * replaces (file:line in /root/reference/srdiff/model.py).  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Device memory is owned by the caller.
 *   - every function returns 0 on success, non-zero on failure; srd_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - `stream` is a cudaStream_t passed as void*.  Launch functions never allocate, never
 *     synchronise and never call back into the host, so they may be stream-captured into a
 *     CUDA graph.
 *   - Activations are "padded NHWC": [B][H+2][W+2][ct] elements (fp32 or bf16) whose 1-pixel
 *     border ring holds either zeros (zero padding, RRDB convs :674-678,714-723 and the
 *     transposed convs :487,759) or the reflected interior (nn.ReflectionPad2d(1) :444,498).
 *     The producer of a buffer writes the ring (desc.border), consumers never pad.
 */
#ifndef SRDIFF_B200_H_
#define SRDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRD_MAX_SRC 3

enum { SRD_F32 = 0, SRD_BF16 = 1 };                      /* activation element type            */
enum { SRD_KERNEL_AUTO = 0, SRD_KERNEL_DIRECT = 1, SRD_KERNEL_TC = 2 };
enum { SRD_ACT_NONE = 0, SRD_ACT_MISH = 1, SRD_ACT_LRELU02 = 2 };
enum { SRD_BORDER_KEEP = 0, SRD_BORDER_REFLECT = 1 };

/* One K-source of an implicit-GEMM convolution: channels [c_off, c_off+cin) of a padded NHWC
 * buffer with `ct` channels per pixel.  Several sources = torch.cat(..., dim=1) that is never
 * materialised (Unet skip concat :842, dense-block concats :686-689, cond concat :825). */
typedef struct {
  const void* ptr;
  int32_t ct;
  int32_t c_off;
  int32_t cin;
  int32_t flags;          /* SRD_SRC_* */
} SrdSrc;

/* SRD_SRC_WIN8 (tcgen05 kernel only, stride 1): the buffer has ct == 8 channels per pixel and the
 * source's 64 "channels" are the 8-pixel window [x, x+8) x 8 channels that starts at the tap's
 * column, i.e. K index = dx*8 + c.  A 3x3 conv over <= 8 real channels then needs nky = 3, nkx = 1
 * (three K = 64 blocks instead of nine): this is how the Unet's first conv (3 -> 64, :444-446) runs.
 * The window of the last pixels of a row runs into the next row (those K slots carry zero weights);
 * the buffer must therefore be readable for 112 bytes past its last pixel. */
enum { SRD_SRC_WIN8 = 1 };

/* Convolution descriptor.  Computes, for every output pixel p and channel n < N,
 *   acc      = sum_{src, ky, kx, c} in_src[p*stride + (ky,kx)][c] * W[...][n]
 *   v        = alpha * act(acc + bias[n]) + temb[b*temb_bstride + n]
 *              + beta1 * res1[p][n] + beta2 * res2[p][n]
 * and stores v (converted to the activation type) to out / out2 and/or to a planar fp32 NCHW
 * tensor.  Covers: Block :442-447 (reflect 3x3 + Mish), ResnetBlock tail :476-480, Downsample
 * :494-503 (stride 2), Upsample :483-491 and cond_proj :759-761 (ConvTranspose2d k=2S, stride S,
 * pad S/2 run as S*S sub-pixel phases of 2x2 taps), res_conv / final 1x1 :471,800, and the RRDB
 * convs :674-690,714-745 (zero padding, LeakyReLU 0.2, 0.2-scaled residuals). */
typedef struct {
  int32_t dtype;          /* SRD_F32 | SRD_BF16 : element type of src/res/out buffers            */
  int32_t kernel;         /* SRD_KERNEL_*                                                        */
  int32_t n_src;
  int32_t B, H_in, W_in;  /* unpadded input dims shared by all sources                           */
  SrdSrc src[SRD_MAX_SRC];
  int32_t stride;         /* 1 | 2 (input stride of the conv)                                    */
  int32_t nky, nkx;       /* taps per side: 3 (3x3), 2 (transposed-conv phase), 1 (1x1)          */
  int32_t ky0, kx0;       /* first tap, padded coords, when up == 1 (0 for 3x3, 1 for 1x1)       */
  int32_t up;             /* S: 1 = plain conv; 2|4 = ConvTranspose2d(k=2S, stride=S, pad=S/2)   */
  int32_t N;              /* output channels                                                     */
  const float* w_direct;  /* fp32 [up*up][K][N], K order (src, ky, kx, c)                        */
  const void* w_tc;       /* bf16 [up*up][KB][Npad][64] K-blocks in order (src, chunk, kx, ky*);
                           * stride 2 with srd_tc_s2_halo(): (src, chunk, tap) over the taps (ky, kx) =
                           * (0,0) (0,2) (2,0) (2,2) (0,1) (2,1) (1,0) (1,2) (1,1)               */
  const float* bias;      /* [N]                                                                 */
  int32_t act;
  float alpha;
  const float* temb;      /* optional [*, N] added after the activation (:476)                   */
  int32_t temb_bstride;   /* 0 = same row for every image, else elements between images          */
  const void* res1; int32_t res1_ct, res1_coff; float beta1;
  const void* res2; int32_t res2_ct, res2_coff; float beta2;
  void* out;  int32_t out_ct,  out_coff;
  void* out2; int32_t out2_ct, out2_coff;
  int32_t H_out, W_out;   /* unpadded output dims (H_in*up for up>1, H_in/stride otherwise)      */
  int32_t border;         /* SRD_BORDER_* : what the producer writes into out/out2's ring        */
  float* planar;          /* optional fp32 NCHW [B][planar_c][H_out][W_out] destination          */
  int32_t planar_c;
  int32_t planar_clamp01; /* 1: planar = clamp(v,0,1)*2-1 (RRDBNet tail :744-745)                */
  const float* head_w;    /* optional fused 1x1 head [planar_c][N] (final_conv.1 :800), tc only  */
  const float* head_b;
  /* optional linear residual (tc kernel, TMA-staged epilogue): v += sum_{c<lin_c} lin_w[n][c] *
   * lin_src[p][c] -- a 1x1 res_conv over <= 4 input channels (ResnetBlock.res_conv of the first
   * block, :471,480) evaluated in the epilogue instead of a separate launch; its bias rides on temb. */
  const void* lin_src; int32_t lin_ct, lin_c;
  const float* lin_w;     /* fp32 [N][lin_c] */
  /* optional SECOND output of the same launch (tc kernel, 3x3 stride-1 conv, TMA-staged epilogue, N % 64 == 0,
   * N <= 128): dst2 = conv1x1(sources) + bias2 with N2 == N output channels, no activation, zero ring -- the
   * ResnetBlock's res_conv (:471) computed from the activation tiles block1 (:474) has already loaded.  Its
   * weights ride in w_tc: per (source, 64-channel chunk) the nine 3x3 k-blocks are followed by ONE k-block
   * [Npad][64] of the 1x1 conv. */
  int32_t N2;             /* 0 = no second output */
  int32_t dst2_ct, dst2_coff, reserved2;
  const float* bias2;     /* [N2] */
  void* dst2;
  /* tc kernel: walk the output tiles from the last to the first.  A consumer that starts where its producer
   * just finished finds the freshest part of its input still in the 126 MB L2 (layers are launched back to back
   * and every 160x160 / 80x80 tensor of a 64-image batch is larger than L2), so consecutive layers alternate. */
  int32_t reverse_items;
  /* tc kernel, 0 | 1 | 2.  2: compute the N output channels as two groups of N/2 (N % 128 == 0; stride-1 conv, one
   * NHWC output, no second output / head / planar): w_tc then holds the k-blocks of group 0 (output channels
   * [0, N/2), each block [N/2][64]) followed by those of group 1.  Every weight block is then shared by TWO tile
   * pairs within the TMEM space one N-wide tile would need, halving the weight bytes streamed from L2 per MMA --
   * what bounds the 256-channel layers of the 20x20 Unet level. */
  int32_t n_split;
} SrdConvDesc;

typedef struct SrdConvPlan SrdConvPlan;

const char* srd_last_error(void);
int srd_version(void);
/* ABI self-check for FFI bindings: which = 0 -> sizeof(SrdConvDesc), 1 -> sizeof(SrdPSampleArgs),
 * 2 -> offsetof(SrdConvDesc, dst2), 3 -> offsetof(SrdPSampleArgs, seed_ptr), 4 -> sizeof(SrdGroupNormArgs),
 * 5 -> offsetof(SrdGroupNormArgs, stats). */
int srd_abi_sizeof(int32_t which);
/* 1 if the tcgen05/TMA path can serve this descriptor (bf16, channel/stride alignment...). */
int srd_conv_tc_supported(const SrdConvDesc* desc);
/* Validates the descriptor, picks the kernel, encodes the TMA tensor maps. Host-only work. */
int srd_conv_plan_create(const SrdConvDesc* desc, SrdConvPlan** out_plan);
int srd_conv_plan_destroy(SrdConvPlan* plan);
/* Which kernel the plan resolved to (SRD_KERNEL_DIRECT | SRD_KERNEL_TC). */
int srd_conv_plan_kernel(const SrdConvPlan* plan);
/* Launch.  `temb` overrides desc.temb when non-NULL (per-step row of the time table). */
int srd_conv_run(const SrdConvPlan* plan, const float* temb, void* stream);
/* ---- Layer chain: a LIST of tcgen05 conv plans executed by ONE persistent launch (the whole Unet.forward :820-847
 * of one denoiser call).  CTA pairs move from layer to layer on their own; a consumer layer is ordered after its
 * producer by per-row-block completion counters in global memory instead of kernel boundaries, so the per-launch
 * fill / drain of ~40 separate launches disappears.  links[i] names, for every K source and for res1 of layer i, the
 * chain layer that writes that tensor (-1: written before the launch) and which of its outputs (0: out, 1: the second
 * output dst2); res2, lin_src, temb and the weights must not be produced inside the chain.  temb_off >= 0: the
 * layer's time-embedding vector is temb_row + temb_off of the per-launch row handed to srd_chain_run (replaces the
 * per-launch `temb` of srd_conv_run).  flags bit 0: every layer its own segment (debug).  The buffers must all be
 * distinct (no in-place layers).
 * flags bit 1: FUSED SEGMENT -- at most srd_chain_fused_max_layers() layers that all share one segment plan
 * (srd_conv_plans_share_segment): the layer loop is unrolled at compile time so every layer's arguments stay
 * constant-bank operands (as fast per item as the single-layer kernel), the counters are scaled by a device-side
 * launch epoch instead of being cleared, and programmatic dependent launch is kept.  This is the mode the engine
 * uses for runs of same-shaped layers (ResnetBlock pairs, the 20x20 level). */
typedef struct SrdChain SrdChain;
typedef struct {
  int32_t src_layer[SRD_MAX_SRC], src_which[SRD_MAX_SRC];
  int32_t res1_layer, res1_which;
  int32_t temb_off;
} SrdChainLink;
int srd_chain_create(SrdConvPlan* const* plans, const SrdChainLink* links, int32_t n, int32_t flags, SrdChain** out);
int srd_chain_run(const SrdChain* chain, const float* temb_row, void* stream);
int srd_chain_destroy(SrdChain* chain);
int srd_chain_fused_max_layers(void);
/* 1 if two tcgen05 plans have the same shared-memory / TMEM / barrier plan, i.e. may sit in one segment */
int srd_conv_plans_share_segment(const SrdConvPlan* x, const SrdConvPlan* y);
/* most layers one chain can hold (their scalar arguments travel in the 32 KB kernel parameter space) */
int srd_chain_max_layers(void);
/* number of segments (runs of layers the roles cross without a CTA-level drain) the planner found */
int srd_chain_segments(const SrdChain* chain);

/* Tuning / validation switches for plans created afterwards (defaults: a_mode = 1, b_resident = 1).
 * a_mode bit 0: 1 = one halo TMA box per K chunk, taps addressed by 128 B-granular UMMA descriptor
 *               offsets (default); 0 = one box per (chunk, kx) with 1024 B-aligned descriptors.
 *        bit 1: disable the TMA-staged epilogue (register epilogue with direct global stores).
 *        bit 2: single-CTA MMAs (cta_group::1) instead of CTA pairs (cta_group::2, M = 256).
 *        bit 3: no programmatic dependent launch between consecutive conv layers.
 *        bit 4: resident weights always with one M tile per CTA and item (MT = 1).
 *        bit 5: give up the second TMEM accumulator stage for MT = 2 on small, wide layers (measured slower).
 *        bit 6: hand work items to CTA pairs through a ticket counter instead of round-robin (measured slower).
 * b_resident: keep all weight k-blocks in shared memory when they fit. */
int srd_tc_set_mode(int32_t a_mode, int32_t b_resident);
/* 1 when stride-2 convs read halo boxes per (row parity, column parity) -- the default; SRD_S2_HALO=0 in the
 * environment selects boxes per (kx, row parity).  Decides the tap order of a stride-2 conv's w_tc k-blocks
 * (see SrdConvDesc.w_tc). */
int srd_tc_s2_halo(void);
/* Debug: plans created afterwards record per-item clock64 stamps into buf [grid][64][8] (NULL = off). */
int srd_tc_set_debug(void* buf);
/* Non-zero after the tcgen05 pipeline watchdog fired (value = which wait timed out). */
int srd_watchdog_flag(void);

/* Time-embedding table: SinusoidalPosEmb :412-419 -> Unet.mlp :764-768 -> per-ResnetBlock
 * Mish+Linear :464-467 for n_t time values at once.  out[n_t][S] with S = sum of Cout over
 * blocks; w_blk [S][dim], b_blk [S] are the per-block Linear weights stacked; freqs[dim/2] =
 * exp(-j*ln(1e4)/(dim/2-1)) exactly as :415-416 evaluates it (host-provided). */
int srd_time_table(const float* t_values, int32_t n_t, int32_t dim, const float* freqs,
                   const float* w0, const float* b0, const float* w2, const float* b2,
                   const float* w_blk, const float* b_blk, int32_t S,
                   float* out, void* stream);

/* SinusoidalPosEmb.forward :412-419 stand-alone: out[n][dim] = cat(sin(t[i] * freqs), cos(t[i] * freqs)),
 * freqs[dim/2] host-provided as for srd_time_table. */
int srd_sinusoidal_emb(const float* t, int32_t n, int32_t dim, const float* freqs, float* out, void* stream);

/* p_sample update :1054-1060 (+ predict_start_from_noise :968-972, clamp :986-987, q_posterior
 * :974-981) as ONE vectorised elementwise kernel over [B][chw] fp32 tensors.
 *   coef: [T][5] fp32 rows {sqrt_recip_ac, sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2,
 *         post_log_variance_clipped};  t: [B] int32 timestep per image.
 *   noise: fp32 [B][chw] or NULL -> in-kernel Philox4x32-10 + Box-Muller keyed by
 *          (seed, step, image_offset + b, element).
 *   x0_out optional.  img_out optional: res2img :1118-1125 fused,
 *          img = clamp(x_prev)/res_rescale + img_lr_up (clamp iff clip_input).
 *   nhwc_out optional: x_prev re-laid out as padded NHWC (ct channels, reflect ring) in
 *          act_dtype for the next step's first conv. */
typedef struct {
  const float* x; const float* eps; const float* noise;
  const float* coef; const int32_t* t;
  int32_t B, C, H, W;
  int32_t clip_denoised;
  uint64_t seed; uint32_t step; uint32_t image_offset;
  float* x_out; float* x0_out;
  float* img_out; const float* img_lr_up; float res_rescale; int32_t clip_input;
  void* nhwc_out; int32_t nhwc_ct; int32_t act_dtype;
  const uint64_t* seed_ptr;   /* optional: Philox seed read from device memory (overrides `seed`), so a
                                 captured CUDA graph can be replayed with a new seed */
} SrdPSampleArgs;
int srd_p_sample_step(const SrdPSampleArgs* args, void* stream);

/* Standard-normal fill with the same Philox stream as srd_p_sample_step (x_T :1070). */
int srd_randn(float* out, int32_t B, int32_t chw, uint64_t seed, uint32_t step,
              uint32_t image_offset, void* stream);

/* fp32 NCHW [B][C][H][W] -> padded NHWC [B][H+2][W+2][ct] (channels >= C zeroed),
 * v = x*scale + shift ((x+1)/2 of RRDBNet.forward :729), ring = reflect | zero. */
int srd_nchw_to_act(const float* x, int32_t B, int32_t C, int32_t H, int32_t W,
                    float scale, float shift, void* out, int32_t ct, int32_t dtype,
                    int32_t border, void* stream);
/* padded NHWC channels [c_off, c_off+C) -> fp32 NCHW. */
int srd_act_to_nchw(const void* act, int32_t ct, int32_t c_off, int32_t dtype,
                    int32_t B, int32_t C, int32_t H, int32_t W, float* out, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest') :738-739 between padded NHWC buffers
 * (same ct), zero ring. */
int srd_upsample_nearest2x(const void* in, void* out, int32_t dtype, int32_t B, int32_t H,
                           int32_t W, int32_t ct, void* stream);
/* Mish :422-424 elementwise on fp32 (stand-alone module API). */
int srd_mish_f32(const float* x, float* y, int64_t n, void* stream);
/* res2img / img2res :1118-1134 elementwise on fp32. */
int srd_res2img(const float* x, const float* img_lr_up, float* out, int64_t n,
                float res_rescale, int32_t clip, int32_t inverse, void* stream);

/* ---- SURVEY 8f "next" rows: the sampler's remaining arithmetic and the driver's image handling ---- */

/* q_sample :1041-1048 on fp32 [B][chw]: out = (sqrt_ac[t]*x + sqrt_1m_ac[t]*noise) * [t >= 0] + x * [t < 0];
 * sqrt_ac / sqrt_1m_ac are the GaussianDiffusion buffers sqrt_alphas_cumprod / sqrt_one_minus_alphas_cumprod,
 * t: [B] int32. */
int srd_q_sample(const float* x, const float* noise, const float* sqrt_ac, const float* sqrt_1m_ac,
                 const int32_t* t, int32_t B, int64_t chw, float* out, void* stream);
/* out = a*x + b*y on fp32 (interpolate() :1109). */
int srd_axpby(const float* x, const float* y, float a, float b, int64_t n, float* out, void* stream);
/* F.interpolate(mode='bicubic', align_corners=True) (st_app1.py:1593-1597) on BC planes of H x W -> Ho x Wo. */
int srd_bicubic_upsample(const float* in, int32_t BC, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                         float* out, void* stream);
/* uint8 HWC [B][H][W][C] -> fp32 NCHW: (u/255 - mean[c]) / std[c] (ToTensor + Normalize, st_app1.py:1585-1589);
 * mean == std == NULL: u/127.5 - 1. */
int srd_u8hwc_to_nchw(const uint8_t* in, int32_t B, int32_t H, int32_t W, int32_t C, const float* mean,
                      const float* stdv, float* out, void* stream);
/* fp32 NCHW -> uint8 HWC: clamp(-1,1), (x+1)/2, clip(x*255, 0, 255), truncate (st_app1.py:1682-1684,1710-1711). */
int srd_nchw_to_u8hwc(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, uint8_t* out, void* stream);
/* Tile split (to_tiles = 1) / ImgMerger-style merge (to_tiles = 0, srdiff/model.py:133-145) between
 * image [n][C][nx*th][ny*tw] and tiles [n][nx*ny][C][th][tw] (tile (x, y) at index x*ny + y covers rows
 * [x*th, +th), cols [y*tw, +tw)). */
int srd_tiles(const float* src, float* dst, int32_t n, int32_t C, int32_t nx, int32_t ny, int32_t th,
              int32_t tw, int32_t to_tiles, void* stream);

/* ---- architecture variants and quality metrics (csrc/variant_ops.cu) ---- */

/* GroupNorm variant of `Block` (srdiff/model.py:448-454), applied to the output of a bias-only conv launch:
 *   out = act(GroupNorm(groups, C)(src)) + temb[b][c] + res      (ResnetBlock.forward :474-480)
 * on padded NHWC activations of `dtype`; statistics are per (image, group) over (C/groups, H, W), biased variance,
 * eps inside the square root (nn.GroupNorm).  border = SRD_BORDER_REFLECT fills the ring with the mirrored result
 * (the next launch is a ReflectionPad2d conv), anything else writes a zero ring.  src != out. */
typedef struct SrdGroupNormArgs {
  const void* src; int32_t src_ct, src_coff;
  int32_t dtype;                          /* SRD_BF16 / SRD_F32 of src, res and out */
  void* out; int32_t out_ct, out_coff;
  int32_t B, H, W, C, groups;
  int32_t act, border;
  float eps; int32_t temb_bstride;        /* temb row stride (floats) per image; 0 = one row for all */
  const float* gamma; const float* beta;  /* [C] GroupNorm weight / bias */
  const float* temb;                      /* optional [B][temb_bstride] or [C] */
  const void* res; int32_t res_ct, res_coff;
  int32_t reserved;
  float* stats;                           /* workspace, 2 * B * groups floats (mean, rstd) */
} SrdGroupNormArgs;
int srd_groupnorm_act(const SrdGroupNormArgs* args, void* stream);
/* Core of `LinearAttention.forward` (:514-522) between its two 1x1 convs: qkv holds 3*heads*dim_head channels in
 * (qkv, head, d) order; k is soft-maxed over the H*W positions, context = k v^T per head, out = context^T q, written
 * as heads*dim_head channels in (head, d) order.  dim_head must be 32. */
int srd_linear_attention(const void* qkv, int32_t qkv_ct, int32_t qkv_coff, int32_t dtype, int32_t B, int32_t H,
                         int32_t W, int32_t heads, int32_t dim_head, void* out, int32_t out_ct, int32_t out_coff,
                         void* stream);
/* y[B][N] = f(x[B][K]) . w[N][K]^T + bias[N], f = Mish when pre_mish: the ResnetBlock time MLP
 * `Sequential(Mish(), Linear(time_emb_dim, dim_out))` (:464-468) for stand-alone block calls. */
int srd_linear_f32(const float* x, const float* w, const float* bias, int32_t B, int32_t K, int32_t N,
                   int32_t pre_mish, float* y, void* stream);
/* torch SSIM (:146-222): per-image mean of the SSIM map of fp32 NCHW images after x*pre_scale + pre_shift
 * (`SSIM.forward` uses 0.5 / 0.5, the functional `ssim()` 1 / 0), depthwise window window1d (x) window1d with
 * zero padding window_size/2.  out: [B] floats; workspace: srd_ssim_workspace(...) floats. */
int srd_ssim_workspace(int32_t B, int32_t C, int32_t H, int32_t W);
int srd_ssim(const float* img1, const float* img2, int32_t B, int32_t C, int32_t H, int32_t W,
             int32_t window_size, const float* window1d, float pre_scale, float pre_shift, float* workspace,
             float* out, void* stream);
/* `Measure.measure` (:1503-1542): np.round((x + 1) * 127.5).clip(0, 255).astype(uint8), element order kept ... */
int srd_f32_to_u8_round(const float* x, int64_t n, uint8_t* out, void* stream);
/* ... and the exact sum of squared differences of two uint8 arrays (PSNR = 10 log10(255^2 * n / sum)). */
int srd_sq_err_u8(const uint8_t* a, const uint8_t* b, int64_t n, uint64_t* out_sum, void* stream);

/* One separable pass of `imresize` (:1162-1306, `imresizevec`): out[i] = sum_p weights[i][p] * in[indices[i][p]]
 * along dim (0: rows, 1: columns) of an [H][W][C] image; weights [out_len][P] float64 and indices [out_len][P] come
 * from the host-side `contributions` (:1196-1218).  is_u8: uint8 image, clipped and rounded half-to-even after the
 * pass; otherwise a float64 image. */
int srd_resize_dim(const void* in, int32_t is_u8, int32_t H, int32_t W, int32_t C, int32_t dim, int32_t out_len,
                   const double* weights, const int32_t* indices, int32_t P, void* out, void* stream);
/* `Measure.ssim` (:1537-1539) = skimage.metrics.structural_similarity(a, b, multichannel=True, data_range=255) on
 * uint8 [H][W][C] images: uniform win x win window (default 7), sample covariance, float64, mean over the map
 * cropped by (win-1)/2 and over channels.  workspace: 512 doubles; out: 1 double. */
int srd_ssim_uniform_u8(const uint8_t* a, const uint8_t* b, int32_t H, int32_t W, int32_t C, int32_t win,
                        double* workspace, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRDIFF_B200_H_ */
