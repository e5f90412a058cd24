// Generic implicit-GEMM convolution on the CUDA cores (fp32 accumulate).
//
// Role: (1) the fp32 *validation mode* of the whole path (north_star: eps rel-L2 <= 1e-4),
// (2) the layers the tcgen05 kernel does not take (3-channel inputs), (3) the on-GPU
// cross-check for the tcgen05 kernel in tests.  Same descriptor, same fused epilogue.
//
// GEMM view: M = B*H_m*W_m output positions (per sub-pixel phase), N = Cout,
// K = sum_src nky*nkx*cin.  CTA tile 64(M) x 64(N), K step 16, 256 threads, 4x4 per thread.
#include "epilogue.cuh"

namespace srd {

constexpr int DM = 64, DN = 64, DK = 16, DTHREADS = 256;

struct DirectArgs {
  SrdSrc src[SRD_MAX_SRC];
  int n_src;
  int B, H_in, W_in, H_m, W_m;
  int stride, nky, nkx, ky0, kx0, up;
  int K_total;
  const float* w;  // [P][K_total][N]
  Epi ep;
};

template <typename T, bool PRECISE>
__global__ void __launch_bounds__(DTHREADS) conv_direct_kernel(const DirectArgs a) {
  __shared__ float As[DK][DM + 4];
  __shared__ __align__(16) float Ws[DK][DN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int phase = blockIdx.z;
  const int pr = phase / a.up, ps = phase % a.up;
  int ky0 = a.ky0, kx0 = a.kx0;
  if (a.up > 1) { ky0 = (pr >= a.up / 2) ? 1 : 0; kx0 = (ps >= a.up / 2) ? 1 : 0; }
  const int M_total = a.B * a.H_m * a.W_m;
  const int m0 = blockIdx.x * DM, n0 = blockIdx.y * DN;
  const int N = a.ep.N;
  const int Hp = a.H_in + 2, Wp = a.W_in + 2;

  // A-load role: pixel lp = tid/4, channels (tid%4)*4..+3 of the current 16-chunk
  const int lp = tid >> 2, lc = (tid & 3) * 4;
  const int lm = m0 + lp;
  const bool lvalid = lm < M_total;
  int lb = 0, ly = 0, lx = 0;
  if (lvalid) { lb = lm / (a.H_m * a.W_m); int r = lm % (a.H_m * a.W_m); ly = r / a.W_m; lx = r % a.W_m; }
  // W-load role: row kk = tid/16, cols (tid%16)*4
  const int wk = tid >> 4, wc = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* wbase = a.w + (size_t)phase * a.K_total * N;
  int kbase = 0;
  for (int s = 0; s < a.n_src; ++s) {
    const SrdSrc S = a.src[s];
    const T* sp = static_cast<const T*>(S.ptr);
    for (int ky = 0; ky < a.nky; ++ky) {
      for (int kx = 0; kx < a.nkx; ++kx) {
        const int iy = ly * a.stride + ky0 + ky, ix = lx * a.stride + kx0 + kx;
        const T* pix = sp + ((size_t)(lb * Hp + iy) * Wp + ix) * S.ct + S.c_off;
        for (int c0 = 0; c0 < S.cin; c0 += DK) {
          // ---- stage A (64 px x 16 ch) and W (16 x 64)
          float av[4] = {0.f, 0.f, 0.f, 0.f};
          if (lvalid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + lc + j;
              if (c < S.cin) av[j] = Elem<T>::ld(pix + c);
            }
          }
          float wv[4] = {0.f, 0.f, 0.f, 0.f};
          if (c0 + wk < S.cin) {
            const float* wr = wbase + (size_t)(kbase + c0 + wk) * N + n0 + wc;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n0 + wc + j < N) wv[j] = __ldg(wr + j);
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < 4; ++j) As[lc + j][lp] = av[j];
          *reinterpret_cast<float4*>(&Ws[wk][wc]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
          __syncthreads();
#pragma unroll
          for (int kk = 0; kk < DK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
            const float wr4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr4[j], acc[i][j]);
          }
        }
        kbase += S.cin;
      }
    }
  }

  // ---- epilogue: 4 pixels x 4 channels per thread
  const int c0 = n0 + tx * 4;
  if (c0 >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M_total) continue;
    const int b = m / (a.H_m * a.W_m);
    const int r = m % (a.H_m * a.W_m);
    const int y = r / a.W_m, x = r % a.W_m;
    const int oy = y * a.up + pr, ox = x * a.up + ps;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epilogue_math<T, 4, 4, PRECISE>(a.ep, b, oy, ox, c0, v);
    epilogue_store<T, 4, 4>(a.ep, b, oy, ox, c0, v, true);
  }
}

int launch_conv_direct(const SrdConvDesc& d, const float* temb, cudaStream_t stream) {
  DirectArgs a;
  for (int i = 0; i < SRD_MAX_SRC; ++i) a.src[i] = d.src[i];
  a.n_src = d.n_src;
  a.B = d.B; a.H_in = d.H_in; a.W_in = d.W_in;
  a.H_m = d.H_in / d.stride; a.W_m = d.W_in / d.stride;
  a.stride = d.stride; a.nky = d.nky; a.nkx = d.nkx; a.ky0 = d.ky0; a.kx0 = d.kx0; a.up = d.up;
  int K = 0;
  for (int i = 0; i < d.n_src; ++i) K += d.nky * d.nkx * d.src[i].cin;
  a.K_total = K;
  a.w = d.w_direct;
  a.ep = make_epi(d, temb);
  SRD_CHECK(d.w_direct != nullptr, "conv_direct: w_direct is NULL");
  SRD_CHECK(d.head_w == nullptr, "conv_direct: fused 1x1 head is only available on the tc kernel");
  const int M_total = a.B * a.H_m * a.W_m;
  dim3 grid((M_total + DM - 1) / DM, (d.N + DN - 1) / DN, d.up * d.up);
  if (d.dtype == SRD_F32)
    conv_direct_kernel<float, true><<<grid, DTHREADS, 0, stream>>>(a);
  else
    conv_direct_kernel<__nv_bfloat16, false><<<grid, DTHREADS, 0, stream>>>(a);
  SRD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace srd
