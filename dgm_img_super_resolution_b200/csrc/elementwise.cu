// HBM-bound elementwise kernels of the sampler: the fused p_sample update, Philox noise,
// layout converters, nearest x2 upsample, stand-alone Mish / res2img, and the time table.
#include "epilogue.cuh"

namespace srd {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) + Box-Muller.  One call -> 4 standard normals.
// Counter = (quad index, global image index, step, 0); key = seed.  Keyed by GLOBAL image
// index so results do not depend on how a batch is sharded over GPUs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ void normal4(uint64_t seed, uint32_t quad, uint32_t image, uint32_t step,
                                        float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(quad, image, step, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float k = 2.3283064365386963e-10f;  // 2^-32
  // Box-Muller on the MUFU pipe (lg2 / sqrt / sin / cos approximations, ~1e-6 absolute error on
  // arguments in [0, 2*pi]): the libm versions made this HBM-bound kernel ALU-bound.
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float u1 = ((float)r[2 * i] + 1.0f) * k;        // (0, 1]
    const float u2 = (float)r[2 * i + 1] * k;             // [0, 1]
    float rad;                                                                  // sqrt(-2 ln u1), MUFU sqrt
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * __log2f(fminf(u1, 1.0f))));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z[2 * i] = rad * c;
    z[2 * i + 1] = rad * s;
  }
}

__global__ void randn_kernel(float* out, int B, int chw, uint64_t seed, uint32_t step,
                             uint32_t image_offset) {
  const int nquad = (chw + 3) / 4;
  const long long total = (long long)B * nquad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / nquad), q = (int)(i % nquad);
    float z[4];
    normal4(seed, (uint32_t)q, image_offset + b, step, z);
    float* p = out + (size_t)b * chw + (size_t)q * 4;
    if (q * 4 + 3 < chw && (chw % 4) == 0) {
      *reinterpret_cast<float4*>(p) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (q * 4 + j < chw) p[j] = z[j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// p_sample update (reference :1054-1060, :968-990).  Thread = V consecutive pixels of one image,
// all C channels.  Arithmetic mirrors the reference's fp32 op order (no FMA contraction).
// Algorithmic HBM traffic: 16 B/element with a noise tensor (x, eps, z in; x' out), 12 B with
// in-kernel Philox (+ the small NHWC copy of x' when requested).
// ------------------------------------------------------------------------------------------
#ifndef SRD_PS_MINBLOCKS
#define SRD_PS_MINBLOCKS 3      // 4 (64 registers, spills) and 2 measured the same 24 us: the kernel is issue-bound
#endif
template <int V, typename T, bool SMALLC>
__global__ void __launch_bounds__(256, SRD_PS_MINBLOCKS) p_sample_kernel(const SrdPSampleArgs a) {
  const int HW = a.H * a.W;
  const int nq = HW / V;                       // V divides W (checked on the host)
  const uint64_t seed = a.seed_ptr ? __ldg(reinterpret_cast<const unsigned long long*>(a.seed_ptr)) : a.seed;
  const long long total = (long long)a.B * nq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((unsigned)i / (unsigned)nq), q = (int)((unsigned)i - (unsigned)b * (unsigned)nq);   // total < 2^31 (host check)
    const int t = a.t[b];
    const float* cf = a.coef + (size_t)t * 5;
    const float cA = __ldg(cf), cB = __ldg(cf + 1), c1 = __ldg(cf + 2), c2 = __ldg(cf + 3);
    const float sigma = (t == 0) ? 0.f : expf(0.5f * __ldg(cf + 4));   // nonzero_mask * exp(.5 logvar)
    const int pix = q * V;
    const int y = pix / a.W, x0p = pix % a.W;
    const size_t off0 = (size_t)b * a.C * HW + pix;        // channel c of this pixel group: off0 + c * HW
    constexpr int NHC = SMALLC ? 4 : 8;         // channels kept for the NHWC copy (the rest of the 8 are zero)
    float nh[V][NHC];
    if (a.nhwc_out) {
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < NHC; ++c) nh[v][c] = 0.f;
    }
    // All loads of the pixel group are issued before any store: x_out may alias x (in-place update), so
    // the compiler cannot hoist the next channel's loads above this channel's stores by itself, and a
    // thread would have only two 16 B loads in flight.
    float4 px4[4], pe4[4];
    if (SMALLC) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < a.C) {
          const size_t off = off0 + (size_t)c * HW;
          // plain (coherent) loads: x_out may alias x, and ld.global.nc on memory this kernel writes is undefined
          px4[c] = *reinterpret_cast<const float4*>(a.x + off);
          pe4[c] = *reinterpret_cast<const float4*>(a.eps + off);
        }
      }
    }
    // SMALLC (V == 4, C <= 4): four fully unrolled, predicated channel iterations over the preloaded
    // registers; otherwise one runtime loop over the channels.
#pragma unroll
    for (int cu = 0; cu < (SMALLC ? 4 : 1); ++cu)
    for (int c = SMALLC ? cu : 0; c < (SMALLC ? (cu < a.C ? cu + 1 : 0) : a.C); ++c) {
      const size_t off = off0 + (size_t)c * HW;
      float xv[V], ev[V], zv[V], ov[V], x0v[V];
      if (SMALLC) {
        const float4 t0 = px4[cu], t1 = pe4[cu];
        xv[0] = t0.x; xv[1] = t0.y; xv[2] = t0.z; xv[3] = t0.w;
        ev[0] = t1.x; ev[1] = t1.y; ev[2] = t1.z; ev[3] = t1.w;
      } else if (V == 4) {
        const float4 t0 = *reinterpret_cast<const float4*>(a.x + off);
        const float4 t1 = *reinterpret_cast<const float4*>(a.eps + off);
        xv[0] = t0.x; xv[1] = t0.y; xv[2] = t0.z; xv[3] = t0.w;
        ev[0] = t1.x; ev[1] = t1.y; ev[2] = t1.z; ev[3] = t1.w;
      } else {
        for (int v = 0; v < V; ++v) { xv[v] = a.x[off + v]; ev[v] = a.eps[off + v]; }
      }
      if (a.noise) {
        if (V == 4) {
          const float4 t2 = *reinterpret_cast<const float4*>(a.noise + off);
          zv[0] = t2.x; zv[1] = t2.y; zv[2] = t2.z; zv[3] = t2.w;
        } else {
          for (int v = 0; v < V; ++v) zv[v] = a.noise[off + v];
        }
      } else {
        // quad index over the image's [C][H][W] element order, 4 elements per Philox call
        float z4[4];
        if (V == 4) {
          normal4(seed, (uint32_t)((c * HW + pix) >> 2), a.image_offset + b, a.step, z4);
#pragma unroll
          for (int v = 0; v < V; ++v) zv[v] = z4[v];
        } else {
          for (int v = 0; v < V; ++v) {
            const int e = c * HW + pix + v;
            normal4(seed, (uint32_t)(e >> 2), a.image_offset + b, a.step, z4);
            zv[v] = z4[e & 3];
          }
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float x0 = __fsub_rn(__fmul_rn(cA, xv[v]), __fmul_rn(cB, ev[v]));
        if (a.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv[v]));
        ov[v] = __fadd_rn(mean, __fmul_rn(sigma, zv[v]));
        x0v[v] = x0;
        if (a.nhwc_out && c < NHC) nh[v][SMALLC ? cu : c] = ov[v];
      }
      if (V == 4) {
        *reinterpret_cast<float4*>(a.x_out + off) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        if (a.x0_out) *reinterpret_cast<float4*>(a.x0_out + off) = make_float4(x0v[0], x0v[1], x0v[2], x0v[3]);
      } else {
        for (int v = 0; v < V; ++v) { a.x_out[off + v] = ov[v]; if (a.x0_out) a.x0_out[off + v] = x0v[v]; }
      }
      if (a.img_out) {                          // res2img :1118-1125
        float iv[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float r = ov[v];
          if (a.clip_input) r = fminf(fmaxf(r, -1.f), 1.f);
          iv[v] = __fadd_rn(__fdiv_rn(r, a.res_rescale), a.img_lr_up[off + v]);
        }
        if (V == 4) *reinterpret_cast<float4*>(a.img_out + off) = make_float4(iv[0], iv[1], iv[2], iv[3]);
        else for (int v = 0; v < V; ++v) a.img_out[off + v] = iv[v];
      }
    }
    if (a.nhwc_out) {
      // interior pixel groups (all but the two outermost rows / columns, whose pixels also own mirrored ring
      // copies): the V pixels are V consecutive 8-channel slots -- plain vector stores, no ring logic
      // bf16, <= 4 channels, V == 4 (quads never straddle a row): the V pixels are V consecutive 8-channel slots.
      // The mirrored ring copies (pad row 0 <- row 1, pad row H+1 <- row H-2, same for columns, corners follow) are
      // straight-line predicated stores: nearly every warp holds a quad of the first / last column, so a generic
      // per-pixel ring loop would run -- for a few active lanes -- in almost every warp.
      if (SMALLC && sizeof(T) == 2 && V == 4 && a.H >= 4 && a.W >= 8) {
        const int Wp = a.W + 2;
        uint4* dst = reinterpret_cast<uint4*>(a.nhwc_out) + ((size_t)(b * (a.H + 2) + y + 1) * Wp + x0p + 1);
        uint4 val[V];
#pragma unroll
        for (int v = 0; v < V; ++v)
          val[v] = make_uint4(pack_bf16x2(nh[v][0], nh[v][1]), pack_bf16x2(nh[v][2], NHC > 3 ? nh[v][NHC > 3 ? 3 : 0] : 0.f), 0u, 0u);
        const bool lef = (x0p == 0), rig = (x0p + V == a.W);     // ox == 1 is val[1], ox == W - 2 is val[V - 2]
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = val[v];
        if (lef) dst[-1] = val[1];
        if (rig) dst[V] = val[V - 2];
        if (y == 1 || y == a.H - 2) {
          uint4* d2 = (y == 1) ? dst - 2 * Wp : dst + 2 * Wp;
#pragma unroll
          for (int v = 0; v < V; ++v) d2[v] = val[v];
          if (lef) d2[-1] = val[1];
          if (rig) d2[V] = val[V - 2];
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float o8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) o8[c] = c < NHC ? nh[v][c < NHC ? c : 0] : 0.f;
          store_nhwc<T, 8, (sizeof(T) == 2 ? 8 : 4)>(static_cast<T*>(a.nhwc_out), a.nhwc_ct, 0, b, y,
                                                     x0p + v, a.H, a.W, 0, 8, SRD_BORDER_REFLECT, o8);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout converters
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_act_kernel(const float* x, int B, int C, int H, int W, float scale,
                                   float shift, T* out, int ct, int border) {
  // thread = one padded pixel; writes all ct channels (ring included)
  const int Hp = H + 2, Wp = W + 2;
  const long long total = (long long)B * Hp * Wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / (Hp * Wp));
    const int r = (int)(i % (Hp * Wp));
    int yp = r / Wp, xp = r % Wp;
    int y = yp - 1, xx = xp - 1;
    bool ring = (y < 0 || y >= H || xx < 0 || xx >= W);
    bool zero = false;
    if (ring) {
      if (border == SRD_BORDER_REFLECT) {
        if (y < 0) y = 1; if (y >= H) y = H - 2;
        if (xx < 0) xx = 1; if (xx >= W) xx = W - 2;
      } else {
        zero = true;
      }
    }
    T* p = out + (size_t)i * ct;
    for (int c = 0; c < ct; ++c) {
      float v = 0.f;
      if (!zero && c < C) v = fmaf(x[((size_t)(b * C + c) * H + y) * W + xx], scale, shift);
      Elem<T>::st(p + c, v);
    }
  }
}

template <typename T>
__global__ void act_to_nchw_kernel(const T* act, int ct, int c_off, int B, int C, int H, int W,
                                   float* out) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    out[i] = Elem<T>::ld(act + ((size_t)(b * (H + 2) + y + 1) * (W + 2) + x + 1) * ct + c_off + c);
  }
}

template <typename T>
__global__ void upsample2x_kernel(const T* in, T* out, int B, int H, int W, int ct) {
  // thread = (output interior pixel, 8-channel group)
  const int Ho = 2 * H, Wo = 2 * W;
  const int ng = ct / 8;
  const long long total = (long long)B * Ho * Wo * ng;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % ng);
    const int ox = (int)((i / ng) % Wo);
    const int oy = (int)((i / ((long long)ng * Wo)) % Ho);
    const int b = (int)(i / ((long long)ng * Wo * Ho));
    const T* s = in + ((size_t)(b * (H + 2) + (oy >> 1) + 1) * (W + 2) + (ox >> 1) + 1) * ct + g * 8;
    T* d = out + ((size_t)(b * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1) * ct + g * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = s[c];
  }
}

__global__ void mish_kernel(const float* x, float* y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = mish_precise(x[i]);
}

__global__ void res2img_kernel(const float* x, const float* up, float* out, long long n, float rescale,
                               int clip, int inverse) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v;
    if (!inverse) {                              // res2img :1118-1125
      v = x[i];
      if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
      v = __fadd_rn(__fdiv_rn(v, rescale), up[i]);
    } else {                                     // img2res :1127-1134
      v = __fmul_rn(__fsub_rn(x[i], up[i]), rescale);
      if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// time table: one CTA per time value.  freqs[dim/2] = exp(-j*ln(1e4)/(dim/2-1)) supplied by
// the host exactly as SinusoidalPosEmb :415-416 computes it.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_table_kernel(
    const float* t_values, int dim, const float* freqs, const float* w0, const float* b0,
    const float* w2, const float* b2, const float* w_blk, const float* b_blk, int S, float* out) {
  extern __shared__ float sm[];
  float* e = sm;                 // [dim]
  float* h1 = e + dim;           // [4*dim]
  float* m = h1 + 4 * dim;       // [dim]
  const float t = t_values[blockIdx.x];
  const int half = dim / 2;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float arg = t * freqs[j];
    e[j] = sinf(arg);
    e[half + j] = cosf(arg);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 4 * dim; j += blockDim.x) {
    float acc = b0[j];
    for (int k = 0; k < dim; ++k) acc = fmaf(w0[(size_t)j * dim + k], e[k], acc);
    h1[j] = mish_precise(acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = b2[j];
    for (int k = 0; k < 4 * dim; ++k) acc = fmaf(w2[(size_t)j * 4 * dim + k], h1[k], acc);
    m[j] = mish_precise(acc);    // Mish at the head of every ResnetBlock.mlp :465
  }
  __syncthreads();
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    float acc = b_blk[j];
    for (int k = 0; k < dim; ++k) acc = fmaf(w_blk[(size_t)j * dim + k], m[k], acc);
    out[(size_t)blockIdx.x * S + j] = acc;
  }
}

// SinusoidalPosEmb.forward :412-419: out[i] = cat(sin(t[i] * f), cos(t[i] * f)), f[dim/2] host-provided
__global__ void sinusoidal_emb_kernel(const float* t, int n, int half, const float* freqs, float* out) {
  const int total = n * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / half, j = i % half;
    const float arg = __fmul_rn(t[r], freqs[j]);
    out[(size_t)r * 2 * half + j] = sinf(arg);
    out[(size_t)r * 2 * half + half + j] = cosf(arg);
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace srd

using namespace srd;

extern "C" {

int srd_time_table(const float* t_values, int32_t n_t, int32_t dim, const float* freqs,
                   const float* w0, const float* b0, const float* w2, const float* b2,
                   const float* w_blk, const float* b_blk, int32_t S, float* out, void* stream) {
  SRD_CHECK(n_t > 0 && dim > 0 && (dim % 2) == 0 && S > 0, "srd_time_table: bad sizes");
  const size_t smem = (size_t)6 * dim * sizeof(float);
  time_table_kernel<<<n_t, 256, smem, (cudaStream_t)stream>>>(t_values, dim, freqs, w0, b0, w2, b2,
                                                              w_blk, b_blk, S, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_sinusoidal_emb(const float* t, int32_t n, int32_t dim, const float* freqs, float* out, void* stream) {
  SRD_CHECK(t && freqs && out && n > 0 && dim >= 4 && (dim % 2) == 0, "srd_sinusoidal_emb: bad arguments");
  sinusoidal_emb_kernel<<<grid_for((long long)n * (dim / 2)), 256, 0, (cudaStream_t)stream>>>(t, n, dim / 2, freqs, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_p_sample_step(const SrdPSampleArgs* a, void* stream) {
  SRD_CHECK(a && a->x && a->eps && a->coef && a->t && a->x_out, "srd_p_sample_step: NULL argument");
  SRD_CHECK(!a->img_out || a->img_lr_up, "srd_p_sample_step: img_out needs img_lr_up");
  SRD_CHECK(!a->nhwc_out || (a->nhwc_ct == 8 && a->C <= 8), "srd_p_sample_step: nhwc_out needs ct == 8, C <= 8");
  const bool v4 = (a->W % 4) == 0;
  const long long total = (long long)a->B * (a->H * a->W / (v4 ? 4 : 1));
  SRD_CHECK(total < (1LL << 31), "srd_p_sample_step: more than 2^31 pixel groups");
  const int grid = grid_for(total);
  cudaStream_t s = (cudaStream_t)stream;
  if (a->act_dtype == SRD_BF16) {
    if (v4 && a->C <= 4) p_sample_kernel<4, __nv_bfloat16, true><<<grid, 256, 0, s>>>(*a);
    else if (v4) p_sample_kernel<4, __nv_bfloat16, false><<<grid, 256, 0, s>>>(*a);
    else p_sample_kernel<1, __nv_bfloat16, false><<<grid, 256, 0, s>>>(*a);
  } else {
    if (v4 && a->C <= 4) p_sample_kernel<4, float, true><<<grid, 256, 0, s>>>(*a);
    else if (v4) p_sample_kernel<4, float, false><<<grid, 256, 0, s>>>(*a);
    else p_sample_kernel<1, float, false><<<grid, 256, 0, s>>>(*a);
  }
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_randn(float* out, int32_t B, int32_t chw, uint64_t seed, uint32_t step,
              uint32_t image_offset, void* stream) {
  SRD_CHECK(out && B > 0 && chw > 0, "srd_randn: bad arguments");
  randn_kernel<<<grid_for((long long)B * ((chw + 3) / 4)), 256, 0, (cudaStream_t)stream>>>(
      out, B, chw, seed, step, image_offset);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_nchw_to_act(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, float scale,
                    float shift, void* out, int32_t ct, int32_t dtype, int32_t border, void* stream) {
  SRD_CHECK(x && out && C <= ct, "srd_nchw_to_act: bad arguments");
  SRD_CHECK(border != SRD_BORDER_REFLECT || (H >= 2 && W >= 2), "srd_nchw_to_act: reflect needs H,W >= 2");
  const int grid = grid_for((long long)B * (H + 2) * (W + 2));
  if (dtype == SRD_BF16)
    nchw_to_act_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        x, B, C, H, W, scale, shift, (__nv_bfloat16*)out, ct, border);
  else
    nchw_to_act_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, C, H, W, scale, shift,
                                                                     (float*)out, ct, border);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_act_to_nchw(const void* act, int32_t ct, int32_t c_off, int32_t dtype, int32_t B, int32_t C,
                    int32_t H, int32_t W, float* out, void* stream) {
  SRD_CHECK(act && out && c_off + C <= ct, "srd_act_to_nchw: bad arguments");
  const int grid = grid_for((long long)B * C * H * W);
  if (dtype == SRD_BF16)
    act_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)act, ct, c_off, B, C, H, W, out);
  else
    act_to_nchw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)act, ct, c_off, B,
                                                                     C, H, W, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_upsample_nearest2x(const void* in, void* out, int32_t dtype, int32_t B, int32_t H, int32_t W,
                           int32_t ct, void* stream) {
  SRD_CHECK(in && out && (ct % 8) == 0, "srd_upsample_nearest2x: ct must be a multiple of 8");
  const int grid = grid_for((long long)B * 4 * H * W * (ct / 8));
  if (dtype == SRD_BF16)
    upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, H, W, ct);
  else
    upsample2x_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, (float*)out, B,
                                                                    H, W, ct);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_mish_f32(const float* x, float* y, int64_t n, void* stream) {
  SRD_CHECK(x && y && n >= 0, "srd_mish_f32: bad arguments");
  if (n == 0) return 0;
  mish_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_res2img(const float* x, const float* img_lr_up, float* out, int64_t n, float res_rescale,
                int32_t clip, int32_t inverse, void* stream) {
  SRD_CHECK(x && img_lr_up && out && n >= 0, "srd_res2img: bad arguments");
  if (n == 0) return 0;
  res2img_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, img_lr_up, out, n, res_rescale, clip,
                                                               inverse);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
