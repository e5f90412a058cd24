// Shared device/host helpers for libsrdiff_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/srdiff_b200.h"

namespace srd {

void set_error(const char* fmt, ...);

#define SRD_CHECK(cond, ...)                         \
  do {                                               \
    if (!(cond)) {                                   \
      ::srd::set_error(__VA_ARGS__);                 \
      return 1;                                      \
    }                                                \
  } while (0)

#define SRD_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::srd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                       __FILE__, __LINE__);                                         \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

// ------------------------------------------------------------------------------------------
// element access (fp32 / bf16 activations)
// ------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Load / store VW consecutive elements as fp32, vectorised when aligned.
template <typename T, int VW>
__device__ __forceinline__ void ld_vec(const T* p, float (&v)[VW]);

template <> __device__ __forceinline__ void ld_vec<float, 4>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld_vec<__nv_bfloat16, 4>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <> __device__ __forceinline__ void ld_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(a); v[2 * i + 1] = __high2float(a);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <typename T, int VW>
__device__ __forceinline__ void st_vec(T* p, const float (&v)[VW]);

template <> __device__ __forceinline__ void st_vec<float, 4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st_vec<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t; t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
template <> __device__ __forceinline__ void st_vec<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
  t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = t;
}

// ------------------------------------------------------------------------------------------
// activations
// ------------------------------------------------------------------------------------------
// Mish = x * tanh(softplus(x)) (reference :422-424; torch softplus threshold 20).
// precise: libm-style evaluation for the fp32 validation mode.
__device__ __forceinline__ float mish_precise(float x) {
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
// fast: with n = e^x, tanh(log(1+n)) = (n^2+2n)/(n^2+2n+2): one ex2 + one rcp on the MUFU pipe.
__device__ __forceinline__ float mish_fast(float x) {
  float n, r;
  const float a = fminf(x, 20.f) * 1.4426950408889634f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(a));
  const float m = n * (n + 2.f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m + 2.f));
  return x * (m * r);
}
template <bool PRECISE>
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SRD_ACT_MISH) return PRECISE ? mish_precise(x) : mish_fast(x);
  if (act == SRD_ACT_LRELU02) return x > 0.f ? x : 0.2f * x;
  return x;
}

}  // namespace srd
