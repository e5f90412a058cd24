// SURVEY 8f-3 / 8f-4 "next" rows around the hot path:
//  * GroupNorm variant of `Block` (srdiff/model.py:448-454): conv -> GroupNorm -> Mish, with the ResnetBlock's
//    time-embedding / residual adds (:474-480) folded into the same pass;
//  * the ResnetBlock time MLP `Linear(Mish(time_emb))` (:464-468,476) for stand-alone block calls;
//  * quality metrics: the torch SSIM (:146-222) and the uint8 PSNR of `Measure` (:1499-1542).
// All HBM-bound passes over padded NHWC activations / fp32 NCHW images.
#include "common.cuh"

namespace srd {

static inline int vgrid(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

// ---------------------------------------------------------------------------- GroupNorm (+ Mish + adds)
// nn.GroupNorm(groups, C) :452: per (image, group) mean and BIASED variance over (C/groups, H, W), eps inside the
// square root.  One block per (b, g); fp32 per-thread partial sums, double across the block.
template <typename T>
__global__ void gn_stats_kernel(const T* x, int ct, int c_off, int H, int W, int C, int G, float eps, float* stats) {
  __shared__ double sh[32];
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const long long n = (long long)H * W * cpg;
  float s = 0.f, ss = 0.f;
  int cnt = 0;
  double ds = 0.0, dss = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % cpg);
    const long long p = i / cpg;
    const int xx = (int)(p % W), y = (int)(p / W);
    const float v = Elem<T>::ld(x + ((size_t)(b * (H + 2) + y + 1) * (W + 2) + xx + 1) * ct + c_off + g * cpg + c);
    s += v; ss += v * v;
    if (++cnt == 64) { ds += s; dss += ss; s = ss = 0.f; cnt = 0; }   // bound the fp32 run length
  }
  ds += s; dss += ss;
  ds = block_sum(ds, sh);
  dss = block_sum(dss, sh);
  if (threadIdx.x == 0) {
    const double mean = ds / (double)n;
    double var = dss / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// thread = one (padded pixel, channel): y = act((x - mean) * rstd * gamma + beta) + temb[b][c] + res; ring per `border`.
template <typename T>
__global__ void gn_apply_kernel(SrdGroupNormArgs a) {
  const int Hp = a.H + 2, Wp = a.W + 2, C = a.C, cpg = a.C / a.groups;
  const T* src = static_cast<const T*>(a.src);
  const T* res = static_cast<const T*>(a.res);
  T* out = static_cast<T*>(a.out);
  const long long total = (long long)a.B * Hp * Wp * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pp = i / C;
    const int xp = (int)(pp % Wp), yp = (int)((pp / Wp) % Hp), b = (int)(pp / ((long long)Wp * Hp));
    int y = yp - 1, x = xp - 1;
    const bool ring = (y < 0 || y >= a.H || x < 0 || x >= a.W);
    float v = 0.f;
    if (!ring || a.border == SRD_BORDER_REFLECT) {
      if (y < 0) y = 1; if (y >= a.H) y = a.H - 2;
      if (x < 0) x = 1; if (x >= a.W) x = a.W - 2;
      const size_t pix = (size_t)(b * Hp + y + 1) * Wp + x + 1;
      const float* st = a.stats + 2 * (b * a.groups + c / cpg);
      v = Elem<T>::ld(src + pix * a.src_ct + a.src_coff + c);
      v = (v - st[0]) * st[1] * a.gamma[c] + a.beta[c];
      if (a.act == SRD_ACT_MISH) v = mish_precise(v);
      else if (a.act == SRD_ACT_LRELU02) v = v > 0.f ? v : 0.2f * v;
      if (a.temb) v += a.temb[(size_t)b * a.temb_bstride + c];
      if (res) v += Elem<T>::ld(res + pix * a.res_ct + a.res_coff + c);
    }
    Elem<T>::st(out + (size_t)pp * a.out_ct + a.out_coff + c, v);
  }
}

// ---------------------------------------------------------------------------- LinearAttention core (:514-522)
// qkv: padded NHWC activation with 3 * heads * D channels ordered (qkv, head, d) -- the reference's
// 'b (qkv heads c) h w'.  Per (image, head): k <- softmax over the n = H*W positions; context[d][e] = sum_n k[d][n] *
// v[e][n]; out[e][n] = sum_d context[d][e] * q[d][n], written to channel head*D + e.  One block per (image, head), D = 32.
template <typename T>
__global__ void linear_attention_kernel(const T* qkv, int ct, int c_off, int H, int W, int heads, T* out, int out_ct,
                                        int out_coff) {
  constexpr int D = 32, TN = 32;
  __shared__ float s_max[D], s_inv[D], s_ctx[D][D + 1], s_k[TN][D + 1], s_v[TN][D + 1];
  const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
  const int n = H * W, Wp = W + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const size_t img = (size_t)b * (H + 2) * Wp;
  auto pix = [&](int i) { return img + (size_t)(i / W + 1) * Wp + (i % W) + 1; };
  const int kq = c_off + hd * D, kk = c_off + (heads + hd) * D, kv = c_off + (2 * heads + hd) * D;
  // row statistics of k: lane = channel d, warps stride over the positions
  float m = -INFINITY;
  for (int i = warp; i < n; i += nwarp) m = fmaxf(m, Elem<T>::ld(qkv + pix(i) * ct + kk + lane));
  s_k[warp][lane] = m;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < nwarp; ++w) m = fmaxf(m, s_k[w][lane]);
    s_max[lane] = m;
  }
  __syncthreads();
  m = s_max[lane];
  float sum = 0.f;
  for (int i = warp; i < n; i += nwarp) sum += expf(Elem<T>::ld(qkv + pix(i) * ct + kk + lane) - m);
  __syncthreads();
  s_k[warp][lane] = sum;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < nwarp; ++w) sum += s_k[w][lane];
    s_inv[lane] = 1.f / sum;
  }
  __syncthreads();
  // context: thread owns (d, e) pairs; positions staged TN at a time
  float acc[4] = {0.f, 0.f, 0.f, 0.f};     // blockDim.x == 256: 1024 / 256 entries per thread
  for (int i0 = 0; i0 < n; i0 += TN) {
    for (int j = threadIdx.x; j < TN * D; j += blockDim.x) {
      const int p = j / D, c = j % D;
      float kvv = 0.f, vv = 0.f;
      if (i0 + p < n) {
        const size_t base = pix(i0 + p) * ct;
        kvv = expf(Elem<T>::ld(qkv + base + kk + c) - s_max[c]) * s_inv[c];
        vv = Elem<T>::ld(qkv + base + kv + c);
      }
      s_k[p][c] = kvv; s_v[p][c] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = threadIdx.x + r * 256, d = idx / D, e = idx % D;
      float a = acc[r];
      for (int p = 0; p < TN; ++p) a = fmaf(s_k[p][d], s_v[p][e], a);
      acc[r] = a;
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int idx = threadIdx.x + r * 256;
    s_ctx[idx / D][idx % D] = acc[r];
  }
  __syncthreads();
  // out[e][i]: lane = e, warps stride over the positions
  for (int i = warp; i < n; i += nwarp) {
    const size_t base = pix(i);
    const float qd = Elem<T>::ld(qkv + base * ct + kq + lane);
    float o = 0.f;
    for (int d = 0; d < D; ++d) o = fmaf(s_ctx[d][lane], __shfl_sync(0xffffffffu, qd, d), o);
    Elem<T>::st(out + base * out_ct + out_coff + hd * D + lane, o);
  }
}

// ---------------------------------------------------------------------------- time MLP
// y[b][n] = sum_k f(x[b][k]) * w[n][k] + bias[n], f = Mish when pre_mish (ResnetBlock.mlp :464-468)
__global__ void linear_kernel(const float* x, const float* w, const float* bias, int B, int K, int N, int pre_mish,
                              float* y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i % N;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float v = x[(size_t)b * K + k];
    if (pre_mish) v = mish_precise(v);
    acc = fmaf(v, w[(size_t)n * K + k], acc);
  }
  y[i] = acc + (bias ? bias[n] : 0.f);
}

// ---------------------------------------------------------------------------- SSIM (:146-222)
// `_ssim`: depthwise Gaussian window (zero padding ws/2) of img, img^2, img1*img2 after x*0.5+0.5; thread = one
// pixel; each block writes one partial sum of the SSIM map (summed per image by ssim_reduce_kernel).
__global__ void ssim_map_kernel(const float* img1, const float* img2, int H, int W, int ws, const float* win,
                                float pre_scale, float pre_shift, int blocks_per_plane, float* partial) {
  __shared__ double sh[32];
  const int plane = blockIdx.x / blocks_per_plane, blk = blockIdx.x % blocks_per_plane;
  const float* p1 = img1 + (size_t)plane * H * W;
  const float* p2 = img2 + (size_t)plane * H * W;
  const int half = ws / 2;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  double acc = 0.0;
  for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < (long long)H * W;
       i += (long long)blocks_per_plane * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int dy = 0; dy < ws; ++dy) {
      const int yy = y + dy - half;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < ws; ++dx) {
        const int xx = x + dx - half;
        if (xx < 0 || xx >= W) continue;
        const float w = win[dy] * win[dx];
        const float a = fmaf(p1[(size_t)yy * W + xx], pre_scale, pre_shift);
        const float b = fmaf(p2[(size_t)yy * W + xx], pre_scale, pre_shift);
        mu1 = fmaf(w, a, mu1); mu2 = fmaf(w, b, mu2);
        s11 = fmaf(w, a * a, s11); s22 = fmaf(w, b * b, s22); s12 = fmaf(w, a * b, s12);
      }
    }
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = s11 - mu1_sq, sg2 = s22 - mu2_sq, sg12 = s12 - mu12;
    const float v = ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
    acc += (double)v;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = (float)acc;
}

__global__ void ssim_reduce_kernel(const float* partial, int per_image, double inv_count, float* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < per_image; i += blockDim.x) acc += (double)partial[(size_t)blockIdx.x * per_image + i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(acc * inv_count);
}

// ---------------------------------------------------------------------------- PSNR (Measure :1503-1542)
// np.round((x + 1) * 127.5).clip(0, 255).astype(np.uint8): fp32 arithmetic, round-half-even
__global__ void f32_to_u8_round_kernel(const float* x, long long n, uint8_t* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = rintf(__fmul_rn(__fadd_rn(x[i], 1.f), 127.5f));
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = (uint8_t)v;
  }
}

// exact integer sum of squared differences of two uint8 arrays (one 64-bit atomic per block)
__global__ void sq_err_u8_kernel(const uint8_t* a, const uint8_t* b, long long n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)a[i] - (int)b[i];
    acc += (unsigned long long)(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}


// ---------------------------------------------------------------------------- imresize (:1162-1306)
// One separable pass of the MATLAB-compatible resize: out[i] = sum_p weights[i][p] * in[indices[i][p]] along `dim`
// of an [H][W][C] image, accumulated in float64 like the reference (`imresizevec`); uint8 images are clipped to
// [0, 255] and rounded half-to-even after EVERY pass (np.around), float64 images are returned as they are.
template <typename T>
__global__ void resize_dim_kernel(const T* in, int H, int W, int C, int dim, int out_len, const double* weights,
                                  const int* indices, int P, T* out) {
  const int Ho = dim == 0 ? out_len : H, Wo = dim == 1 ? out_len : W;
  const long long total = (long long)Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), x = (int)((i / C) % Wo), y = (int)(i / ((long long)C * Wo));
    const int o = dim == 0 ? y : x;
    double acc = 0.0;
    for (int p = 0; p < P; ++p) {
      const int src = indices[(size_t)o * P + p];
      const size_t at = dim == 0 ? ((size_t)src * W + x) * C + c : ((size_t)y * W + src) * C + c;
      acc = __dadd_rn(acc, __dmul_rn(weights[(size_t)o * P + p], (double)in[at]));   // no FMA: numpy multiplies, then sums
    }
    if (sizeof(T) == 1) out[i] = (T)rint(fmin(fmax(acc, 0.0), 255.0));
    else out[i] = (T)acc;
  }
}

// ---------------------------------------------------------------------------- skimage-style SSIM (Measure.ssim :1537-1539)
// structural_similarity(a, b, multichannel=True, data_range=255) of scikit-image: uniform 7x7 window, sample
// covariance (NP / (NP - 1)), float64, mean over the map cropped by (win - 1) / 2, then over channels.
__global__ void ssim_uniform_u8_kernel(const uint8_t* a, const uint8_t* b, int H, int W, int C, int win, double* partial) {
  __shared__ double sh[32];
  const int pad = (win - 1) / 2, Hc = H - 2 * pad, Wc = W - 2 * pad;
  const double NP = (double)win * win, cov_norm = NP / (NP - 1.0);
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  const long long total = (long long)Hc * Wc * C;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), x = (int)((i / C) % Wc), y = (int)(i / ((long long)C * Wc));
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int dy = 0; dy < win; ++dy)
      for (int dx = 0; dx < win; ++dx) {
        const size_t at = ((size_t)(y + dy) * W + x + dx) * C + c;
        const double p = (double)a[at], q = (double)b[at];
        sx += p; sy += q; sxx += p * p; syy += q * q; sxy += p * q;
      }
    const double ux = sx / NP, uy = sy / NP;
    const double vx = cov_norm * (sxx / NP - ux * ux), vy = cov_norm * (syy / NP - uy * uy);
    const double vxy = cov_norm * (sxy / NP - ux * uy);
    acc += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void sum_f64_kernel(const double* partial, int n, double scale, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) *out = acc * scale;
}

}  // namespace srd

using namespace srd;

extern "C" {

int srd_groupnorm_act(const SrdGroupNormArgs* args, void* stream) {
  SRD_CHECK(args, "srd_groupnorm_act: null args");
  const SrdGroupNormArgs& a = *args;
  SRD_CHECK(a.src && a.out && a.gamma && a.beta && a.stats, "srd_groupnorm_act: null pointer");
  SRD_CHECK(a.src != a.out, "srd_groupnorm_act: in-place is not supported (ring pixels re-read the interior)");
  SRD_CHECK(a.B > 0 && a.H > 0 && a.W > 0 && a.C > 0 && a.groups > 0 && a.C % a.groups == 0,
            "srd_groupnorm_act: C (%d) must be a positive multiple of groups (%d)", a.C, a.groups);
  SRD_CHECK(a.src_coff + a.C <= a.src_ct && a.out_coff + a.C <= a.out_ct && (!a.res || a.res_coff + a.C <= a.res_ct),
            "srd_groupnorm_act: channel window exceeds the buffer");
  SRD_CHECK(a.border != SRD_BORDER_REFLECT || (a.H >= 2 && a.W >= 2), "srd_groupnorm_act: reflect needs H,W >= 2");
  SRD_CHECK(a.dtype == SRD_BF16 || a.dtype == SRD_F32, "srd_groupnorm_act: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)a.B * (a.H + 2) * (a.W + 2) * a.C;
  if (a.dtype == SRD_BF16) {
    gn_stats_kernel<__nv_bfloat16><<<a.B * a.groups, 256, 0, s>>>((const __nv_bfloat16*)a.src, a.src_ct, a.src_coff, a.H,
                                                                a.W, a.C, a.groups, a.eps, a.stats);
    gn_apply_kernel<__nv_bfloat16><<<vgrid(total), 256, 0, s>>>(a);
  } else {
    gn_stats_kernel<float><<<a.B * a.groups, 256, 0, s>>>((const float*)a.src, a.src_ct, a.src_coff, a.H, a.W, a.C,
                                                        a.groups, a.eps, a.stats);
    gn_apply_kernel<float><<<vgrid(total), 256, 0, s>>>(a);
  }
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_linear_attention(const void* qkv, int32_t qkv_ct, int32_t qkv_coff, int32_t dtype, int32_t B, int32_t H,
                         int32_t W, int32_t heads, int32_t dim_head, void* out, int32_t out_ct, int32_t out_coff,
                         void* stream) {
  SRD_CHECK(qkv && out && B > 0 && H > 0 && W > 0 && heads > 0, "srd_linear_attention: bad arguments");
  SRD_CHECK(dim_head == 32, "srd_linear_attention: dim_head must be 32 (the reference's default), got %d", dim_head);
  SRD_CHECK(qkv_coff + 3 * heads * dim_head <= qkv_ct && out_coff + heads * dim_head <= out_ct,
            "srd_linear_attention: channel window exceeds the buffer");
  SRD_CHECK(dtype == SRD_BF16 || dtype == SRD_F32, "srd_linear_attention: bad dtype");
  if (dtype == SRD_BF16)
    linear_attention_kernel<__nv_bfloat16><<<B * heads, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv, qkv_ct, qkv_coff, H, W, heads, (__nv_bfloat16*)out, out_ct, out_coff);
  else
    linear_attention_kernel<float><<<B * heads, 256, 0, (cudaStream_t)stream>>>((const float*)qkv, qkv_ct, qkv_coff, H, W,
                                                                              heads, (float*)out, out_ct, out_coff);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_linear_f32(const float* x, const float* w, const float* bias, int32_t B, int32_t K, int32_t N,
                   int32_t pre_mish, float* y, void* stream) {
  SRD_CHECK(x && w && y && B > 0 && K > 0 && N > 0, "srd_linear_f32: bad arguments");
  linear_kernel<<<(B * N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, w, bias, B, K, N, pre_mish, y);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_ssim_workspace(int32_t B, int32_t C, int32_t H, int32_t W) {
  (void)H; (void)W;
  return B * C * 64;     // floats: 64 partial sums per plane
}

int srd_ssim(const float* img1, const float* img2, int32_t B, int32_t C, int32_t H, int32_t W, int32_t window_size,
             const float* window1d, float pre_scale, float pre_shift, float* workspace, float* out, void* stream) {
  SRD_CHECK(img1 && img2 && window1d && workspace && out, "srd_ssim: null pointer");
  SRD_CHECK(B > 0 && C > 0 && H > 0 && W > 0 && window_size > 0 && (window_size & 1), "srd_ssim: bad shape / even window");
  const int bpp = 64;
  ssim_map_kernel<<<B * C * bpp, 256, 0, (cudaStream_t)stream>>>(img1, img2, H, W, window_size, window1d, pre_scale,
                                                               pre_shift, bpp, workspace);
  ssim_reduce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(workspace, C * bpp, 1.0 / ((double)C * H * W), out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_resize_dim(const void* in, int32_t is_u8, int32_t H, int32_t W, int32_t C, int32_t dim, int32_t out_len,
                   const double* weights, const int32_t* indices, int32_t P, void* out, void* stream) {
  SRD_CHECK(in && out && weights && indices, "srd_resize_dim: null pointer");
  SRD_CHECK(H > 0 && W > 0 && C > 0 && out_len > 0 && P > 0 && (dim == 0 || dim == 1), "srd_resize_dim: bad arguments");
  const long long total = (long long)(dim == 0 ? out_len : H) * (dim == 1 ? out_len : W) * C;
  if (is_u8)
    resize_dim_kernel<uint8_t><<<vgrid(total), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, H, W, C, dim, out_len,
                                                                              weights, indices, P, (uint8_t*)out);
  else
    resize_dim_kernel<double><<<vgrid(total), 256, 0, (cudaStream_t)stream>>>((const double*)in, H, W, C, dim, out_len,
                                                                             weights, indices, P, (double*)out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_ssim_uniform_u8(const uint8_t* a, const uint8_t* b, int32_t H, int32_t W, int32_t C, int32_t win,
                        double* workspace, double* out, void* stream) {
  SRD_CHECK(a && b && workspace && out, "srd_ssim_uniform_u8: null pointer");
  SRD_CHECK(win >= 3 && (win & 1) && H >= win && W >= win && C > 0,
            "srd_ssim_uniform_u8: win_size (%d) must be odd and not exceed the image (%d x %d)", win, H, W);
  const int pad = (win - 1) / 2;
  const long long total = (long long)(H - 2 * pad) * (W - 2 * pad) * C;
  const int nb = total < 256LL * 512 ? (int)((total + 255) / 256) : 512;      // workspace: 512 doubles
  ssim_uniform_u8_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(a, b, H, W, C, win, workspace);
  sum_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, nb, 1.0 / (double)total, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_f32_to_u8_round(const float* x, int64_t n, uint8_t* out, void* stream) {
  SRD_CHECK(x && out && n >= 0, "srd_f32_to_u8_round: bad arguments");
  if (n == 0) return 0;
  f32_to_u8_round_kernel<<<vgrid(n), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_sq_err_u8(const uint8_t* a, const uint8_t* b, int64_t n, uint64_t* out_sum, void* stream) {
  SRD_CHECK(a && b && out_sum && n >= 0, "srd_sq_err_u8: bad arguments");
  SRD_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(uint64_t), (cudaStream_t)stream));
  if (n == 0) return 0;
  sq_err_u8_kernel<<<vgrid(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, (unsigned long long*)out_sum);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
