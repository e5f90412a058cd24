// SURVEY 8f "next" rows around the sampler: q_sample / interpolate arithmetic, and the image
// pre/post-processing of the reference's driver (st_app1.py:1593-1597, 1682-1684, 1710-1712) and of
// ImgMerger (srdiff/model.py:121-145) -- HBM-bound elementwise / gather kernels on fp32 NCHW images.
#include "common.cuh"

namespace srd {

static inline int img_grid(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// q_sample :1041-1048: (sqrt_ac[t] * x + sqrt_1m_ac[t] * noise) * [t >= 0] + x * (1 - [t >= 0]), t clamped
// at 0 for the table lookup.  Same fp32 operation order as the reference (no FMA contraction).
__global__ void q_sample_kernel(const float* x, const float* noise, const float* sqrt_ac, const float* sqrt_1m_ac,
                                const int* t, int B, long long chw, float* out) {
  const long long total = (long long)B * chw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / chw);
    const int tb = t[b];
    const float cond = tb >= 0 ? 1.f : 0.f;
    const int tc = tb < 0 ? 0 : tb;
    const float v = __fadd_rn(__fmul_rn(sqrt_ac[tc], x[i]), __fmul_rn(sqrt_1m_ac[tc], noise[i]));
    out[i] = __fadd_rn(__fmul_rn(v, cond), __fmul_rn(x[i], 1.f - cond));
  }
}

// out = a * x + b * y (interpolate() :1109: (1 - lam) * xt1 + lam * xt2), two roundings of the products
__global__ void axpby_kernel(const float* x, const float* y, float a, float b, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i]));
}

// F.interpolate(mode='bicubic', align_corners=True) (st_app1.py:1596): cubic convolution, A = -0.75,
// source coordinate dst * (in - 1) / (out - 1), taps clamped to the border.
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  float x = t + 1.f;
  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

__global__ void bicubic_ac_kernel(const float* in, int BC, int H, int W, int Ho, int Wo, float* out) {
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)BC * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long bc = i / ((long long)Wo * Ho);
    const float fy = sy * oy, fx = sx * ox;
    const int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float cy[4], cx[4];
    cubic_coeffs(fy - iy, cy);
    cubic_coeffs(fx - ix, cx);
    const float* p = in + bc * (long long)H * W;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int yy = min(max(iy - 1 + j, 0), H - 1);
      float row = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = min(max(ix - 1 + k, 0), W - 1);
        row += p[(long long)yy * W + xx] * cx[k];
      }
      acc += row * cy[j];
    }
    out[i] = acc;
  }
}

// uint8 HWC [B][H][W][C] -> fp32 NCHW, v = (u / 255 - mean[c]) / std[c]  (ToTensor + Normalize,
// st_app1.py:1585-1589); mean/std NULL -> u / 127.5 - 1 (the [-1, 1] convention of srdiff/model.py:95).
__global__ void u8hwc_to_nchw_kernel(const uint8_t* in, int B, int H, int W, int C, const float* mean,
                                     const float* stdv, float* out) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    const float u = (float)in[(((long long)b * H + y) * W + x) * C + c];
    out[i] = mean ? __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), mean[c]), stdv[c]) : __fsub_rn(__fdiv_rn(u, 127.5f), 1.f);
  }
}

// fp32 NCHW in [-1, 1] -> uint8 HWC: clamp(-1, 1), (x + 1) / 2, clip(x * 255, 0, 255), truncate
// (st_app1.py:1682-1684, 1710-1711: np.clip(img * 255, 0, 255).astype(np.uint8)).
__global__ void nchw_to_u8hwc_kernel(const float* in, int B, int C, int H, int W, uint8_t* out) {
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), x = (int)((i / C) % W);
    const int y = (int)((i / ((long long)C * W)) % H);
    const int b = (int)(i / ((long long)C * W * H));
    float v = in[(((long long)b * C + c) * H + y) * W + x];
    v = fminf(fmaxf(v, -1.f), 1.f);
    v = __fdiv_rn(__fadd_rn(v, 1.f), 2.f);
    v = fminf(fmaxf(__fmul_rn(v, 255.f), 0.f), 255.f);
    out[i] = (uint8_t)v;
  }
}

// Tiles <-> image (ImgMerger.compute :133-145: tile (x, y) = rows [x*th, +th), cols [y*tw, +tw); y tiles are
// concatenated along W, x tiles along H).  tiles: [n][nx*ny][C][th][tw] with tile index x*ny + y;
// image: [n][C][nx*th][ny*tw].  to_tiles = 1: split, 0: merge.
__global__ void tiles_kernel(const float* src, float* dst, int n, int C, int nx, int ny, int th, int tw, int to_tiles) {
  const int H = nx * th, W = ny * tw;
  const long long total = (long long)n * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W), yy = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    const int tx = yy / th, ty = xx / tw;
    const long long ti = ((((long long)b * nx * ny + tx * ny + ty) * C + c) * th + (yy - tx * th)) * tw + (xx - ty * tw);
    if (to_tiles) dst[ti] = src[i];
    else dst[i] = src[ti];
  }
}

}  // namespace srd

using namespace srd;

extern "C" {

int srd_q_sample(const float* x, const float* noise, const float* sqrt_ac, const float* sqrt_1m_ac, const int32_t* t,
                 int32_t B, int64_t chw, float* out, void* stream) {
  SRD_CHECK(x && noise && sqrt_ac && sqrt_1m_ac && t && out && B > 0 && chw > 0, "srd_q_sample: bad arguments");
  q_sample_kernel<<<img_grid((long long)B * chw), 256, 0, (cudaStream_t)stream>>>(x, noise, sqrt_ac, sqrt_1m_ac, t, B,
                                                                                 chw, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_axpby(const float* x, const float* y, float a, float b, int64_t n, float* out, void* stream) {
  SRD_CHECK(x && y && out && n >= 0, "srd_axpby: bad arguments");
  if (n == 0) return 0;
  axpby_kernel<<<img_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, a, b, n, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_bicubic_upsample(const float* in, int32_t BC, int32_t H, int32_t W, int32_t Ho, int32_t Wo, float* out,
                         void* stream) {
  SRD_CHECK(in && out && BC > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "srd_bicubic_upsample: bad arguments");
  bicubic_ac_kernel<<<img_grid((long long)BC * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(in, BC, H, W, Ho, Wo, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_u8hwc_to_nchw(const uint8_t* in, int32_t B, int32_t H, int32_t W, int32_t C, const float* mean,
                      const float* stdv, float* out, void* stream) {
  SRD_CHECK(in && out && B > 0 && H > 0 && W > 0 && C > 0 && ((mean == nullptr) == (stdv == nullptr)),
            "srd_u8hwc_to_nchw: bad arguments");
  u8hwc_to_nchw_kernel<<<img_grid((long long)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(in, B, H, W, C, mean, stdv,
                                                                                            out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_nchw_to_u8hwc(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, uint8_t* out, void* stream) {
  SRD_CHECK(in && out && B > 0 && H > 0 && W > 0 && C > 0, "srd_nchw_to_u8hwc: bad arguments");
  nchw_to_u8hwc_kernel<<<img_grid((long long)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(in, B, C, H, W, out);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

int srd_tiles(const float* src, float* dst, int32_t n, int32_t C, int32_t nx, int32_t ny, int32_t th, int32_t tw,
              int32_t to_tiles, void* stream) {
  SRD_CHECK(src && dst && n > 0 && C > 0 && nx > 0 && ny > 0 && th > 0 && tw > 0, "srd_tiles: bad arguments");
  tiles_kernel<<<img_grid((long long)n * C * nx * th * ny * tw), 256, 0, (cudaStream_t)stream>>>(src, dst, n, C, nx, ny,
                                                                                                th, tw, to_tiles);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
