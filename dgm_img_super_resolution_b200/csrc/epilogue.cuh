// Fused convolution epilogue shared by the direct (CUDA-core) and tcgen05 kernels.
//   v = alpha * act(acc + bias) + temb + beta1*res1 + beta2*res2   -> out / out2 / planar
// Replaces, per reference layer (srdiff/model.py): conv bias, Mish :446, time-embedding add
// :476, ResnetBlock residual :480, `x + cond` :830, LeakyReLU :685-688, `x5*0.2+x` :690,
// `out*0.2+x` :706, `fea_first + trunk` :735, clamp/rescale :744-745 -- and the
// nn.ReflectionPad2d(1) of the NEXT layer (:444,498), by writing the mirrored ring here.
#pragma once
#include "common.cuh"

namespace srd {

struct Epi {
  const float* bias;
  const float* temb; int temb_bstride;
  const void* res1; int res1_ct, res1_coff; float beta1;
  const void* res2; int res2_ct, res2_coff; float beta2;
  void* out; int out_ct, out_coff;
  void* out2; int out2_ct, out2_coff;
  float* planar; int planar_c; int planar_clamp01;
  int N, H_out, W_out, act, border;
  float alpha;
};

inline Epi make_epi(const SrdConvDesc& d, const float* temb_override) {
  Epi e;
  e.bias = d.bias;
  e.temb = temb_override ? temb_override : d.temb;
  e.temb_bstride = d.temb_bstride;
  e.res1 = d.res1; e.res1_ct = d.res1_ct; e.res1_coff = d.res1_coff; e.beta1 = d.beta1;
  e.res2 = d.res2; e.res2_ct = d.res2_ct; e.res2_coff = d.res2_coff; e.beta2 = d.beta2;
  e.out = d.out; e.out_ct = d.out_ct; e.out_coff = d.out_coff;
  e.out2 = d.out2; e.out2_ct = d.out2_ct; e.out2_coff = d.out2_coff;
  e.planar = d.planar; e.planar_c = d.planar_c; e.planar_clamp01 = d.planar_clamp01;
  e.N = d.N; e.H_out = d.H_out; e.W_out = d.W_out; e.act = d.act; e.border = d.border;
  e.alpha = d.alpha;
  return e;
}

// Store CH channel values of one pixel into a padded NHWC buffer, replicating into the ring
// rows/cols that mirror this pixel when border == REFLECT (pad row 0 <- row 1, pad row H+1 <-
// row H-2; same for columns; corners follow).
template <typename T, int CH, int VW>
__device__ __forceinline__ void store_nhwc(T* base, int ct, int coff, int b, int oy, int ox,
                                           int H, int W, int c0, int n_valid, int border,
                                           const float (&v)[CH]) {
  const int Wp = W + 2;
  int rows[3], cols[3];
  int nr = 1, nc = 1;
  rows[0] = oy + 1; cols[0] = ox + 1;
  if (border == SRD_BORDER_REFLECT) {
    if (oy == 1) rows[nr++] = 0;
    if (oy == H - 2) rows[nr++] = H + 1;
    if (ox == 1) cols[nc++] = 0;
    if (ox == W - 2) cols[nc++] = W + 1;
  }
  const bool vec_ok = (n_valid >= CH) && ((ct % VW) == 0) && (((coff + c0) % VW) == 0);
  for (int ri = 0; ri < nr; ++ri) {
    for (int ci = 0; ci < nc; ++ci) {
      T* p = base + ((size_t)(b * (H + 2) + rows[ri]) * Wp + cols[ci]) * ct + coff + c0;
      if (vec_ok) {
#pragma unroll
        for (int j = 0; j < CH; j += VW) {
          float t[VW];
#pragma unroll
          for (int k = 0; k < VW; ++k) t[k] = v[j + k];
          st_vec<T, VW>(p + j, t);
        }
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (j < n_valid) Elem<T>::st(p + j, v[j]);
      }
    }
  }
}

template <typename T, int CH, int VW>
__device__ __forceinline__ void add_res(const T* base, int ct, int coff, float beta, int b, int oy,
                                        int ox, int H, int W, int c0, int n_valid, float (&v)[CH]) {
  const T* p = base + ((size_t)(b * (H + 2) + oy + 1) * (W + 2) + ox + 1) * ct + coff + c0;
  const bool vec_ok = (n_valid >= CH) && ((ct % VW) == 0) && (((coff + c0) % VW) == 0);
  if (vec_ok) {
#pragma unroll
    for (int j = 0; j < CH; j += VW) {
      float t[VW];
      ld_vec<T, VW>(p + j, t);
#pragma unroll
      for (int k = 0; k < VW; ++k) v[j + k] = fmaf(beta, t[k], v[j + k]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < n_valid) v[j] = fmaf(beta, Elem<T>::ld(p + j), v[j]);
  }
}

// Pointwise part: v <- alpha*act(v + bias) + temb + residuals.  (b, oy, ox) = output pixel.
template <typename T, int CH, int VW, bool PRECISE>
__device__ __forceinline__ void epilogue_math(const Epi& e, int b, int oy, int ox, int c0,
                                              float (&v)[CH]) {
  const int n_valid = min(CH, e.N - c0);
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    if (j < n_valid) {
      float x = v[j] + __ldg(e.bias + c0 + j);
      x = apply_act<PRECISE>(x, e.act) * e.alpha;
      if (e.temb) x += __ldg(e.temb + (size_t)b * e.temb_bstride + c0 + j);
      v[j] = x;
    } else {
      v[j] = 0.f;
    }
  }
  if (e.res1)
    add_res<T, CH, VW>(static_cast<const T*>(e.res1), e.res1_ct, e.res1_coff, e.beta1, b, oy, ox,
                       e.H_out, e.W_out, c0, n_valid, v);
  if (e.res2)
    add_res<T, CH, VW>(static_cast<const T*>(e.res2), e.res2_ct, e.res2_coff, e.beta2, b, oy, ox,
                       e.H_out, e.W_out, c0, n_valid, v);
}

// Stores: NHWC destinations + optional planar fp32 NCHW (no fused head here).
template <typename T, int CH, int VW>
__device__ __forceinline__ void epilogue_store(const Epi& e, int b, int oy, int ox, int c0,
                                               const float (&v)[CH], bool planar_direct) {
  const int n_valid = min(CH, e.N - c0);
  if (n_valid <= 0) return;
  if (e.out)
    store_nhwc<T, CH, VW>(static_cast<T*>(e.out), e.out_ct, e.out_coff, b, oy, ox, e.H_out,
                          e.W_out, c0, n_valid, e.border, v);
  if (e.out2)
    store_nhwc<T, CH, VW>(static_cast<T*>(e.out2), e.out2_ct, e.out2_coff, b, oy, ox, e.H_out,
                          e.W_out, c0, n_valid, e.border, v);
  if (planar_direct && e.planar) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int c = c0 + j;
      if (j < n_valid && c < e.planar_c) {
        float x = v[j];
        if (e.planar_clamp01) x = fminf(fmaxf(x, 0.f), 1.f) * 2.f - 1.f;
        e.planar[((size_t)(b * e.planar_c + c) * e.H_out + oy) * e.W_out + ox] = x;
      }
    }
  }
}

}  // namespace srd
