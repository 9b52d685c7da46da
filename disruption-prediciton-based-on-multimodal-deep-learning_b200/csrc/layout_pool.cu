// Boundary layout kernels (NCDHW fp32 <-> padded NDHWC, uint8 frames -> NDHWC) and
// AdaptiveAvgPool3d(1) forward/backward on NDHWC.
//
// The reference hands the model NCDHW fp32 clips (/root/reference/src/dataset.py:229-230) made from
// uint8 BGR frames minus a per-channel mean (:104-110,201-205); the pool is R2Plus1D.py:215,224-225.
#include "dp_common.cuh"

namespace dp {

// dst[b][p][cv*8+j] = src[b][cv*8+j][p]   (threads consecutive in p -> coalesced reads)
template <typename T>
__global__ void __launch_bounds__(256)
ncdhw_to_ndhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int Cp, int64_t P) {
  const int b = blockIdx.z, cv = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  f8 v;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    v.v[j] = c < C ? src[((int64_t)b * C + c) * P + p] : 0.f;
  }
  st8(dst + ((int64_t)b * P + p) * Cp + cv * 8, v);
}

template <typename T>
__global__ void __launch_bounds__(256)
ndhwc_to_ncdhw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int Cp, int64_t P) {
  const int b = blockIdx.z, cv = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const f8 v = ld8(src + ((int64_t)b * P + p) * Cp + cv * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    if (c < C) dst[((int64_t)b * C + c) * P + p] = v.v[j];
  }
}

// src: (rows, 3) uint8, dst: (rows, Cp); channel c<3 = src - mean[c], rest 0
template <typename T>
__global__ void __launch_bounds__(256)
u8_to_ndhwc_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, float m0, float m1, float m2,
                   int64_t rows, int Cp) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= rows) return;
  f8 v;
#pragma unroll
  for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
  v.v[0] = (float)src[p * 3 + 0] - m0;
  v.v[1] = (float)src[p * 3 + 1] - m1;
  v.v[2] = (float)src[p * 3 + 2] - m2;
  st8(dst + p * Cp, v);
  f8 z;
#pragma unroll
  for (int j = 0; j < 8; ++j) z.v[j] = 0.f;
  for (int c = 8; c < Cp; c += 8) st8(dst + p * Cp + c, z);
}

// out[b][c] = mean_p x[b][p][c]; one CTA per (b, 8-channel vector group), threads stride over pixels
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t P, int C, int Cp) {
  __shared__ float red[256][8];
  const int b = blockIdx.y, cv = blockIdx.x;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const f8 v = ld8(x + ((int64_t)b * P + p) * Cp + cv * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += v.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = a[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + s][j];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) {
    const int c = cv * 8 + threadIdx.x;
    if (c < C) out[(int64_t)b * C + c] = red[0][threadIdx.x] / (float)P;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dout, T* __restrict__ dx, int64_t P, int C, int Cp, int64_t nvec) {
  const int vpr = Cp >> 3;
  const float inv = 1.f / (float)P;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % vpr);
    const int64_t b = (v / vpr) / P;
    f8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      o.v[j] = c < C ? dout[b * C + c] * inv : 0.f;
    }
    st8(dx + v * 8, o);
  }
}

}  // namespace dp

using namespace dp;

DP_API int dp_ncdhw_f32_to_ndhwc(const float* src, void* dst, int B, int C, int Cp, int T, int H, int W, int dtype,
                                 void* stream) {
  DP_REQUIRE(src && dst, DP_ERR_SHAPE, "dp_ncdhw_f32_to_ndhwc: NULL pointer");
  DP_REQUIRE(B > 0 && C > 0 && Cp >= C && Cp % 8 == 0 && T > 0 && H > 0 && W > 0 && B <= 65535, DP_ERR_SHAPE,
             "dp_ncdhw_f32_to_ndhwc: bad shape");
  const int64_t P = (int64_t)T * H * W;
  dim3 grid(ceil_div(P, 256), Cp / 8, B);
  if (dtype == DP_BF16)
    ncdhw_to_ndhwc_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, C, Cp, P);
  else
    ncdhw_to_ndhwc_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(src, (float*)dst, C, Cp, P);
  return check_launch("dp_ncdhw_f32_to_ndhwc");
}

DP_API int dp_ndhwc_to_ncdhw_f32(const void* src, float* dst, int B, int C, int Cp, int T, int H, int W, int dtype,
                                 void* stream) {
  DP_REQUIRE(src && dst, DP_ERR_SHAPE, "dp_ndhwc_to_ncdhw_f32: NULL pointer");
  DP_REQUIRE(B > 0 && C > 0 && Cp >= C && Cp % 8 == 0 && T > 0 && H > 0 && W > 0 && B <= 65535, DP_ERR_SHAPE,
             "dp_ndhwc_to_ncdhw_f32: bad shape");
  const int64_t P = (int64_t)T * H * W;
  dim3 grid(ceil_div(P, 256), ceil_div(C, 8), B);
  if (dtype == DP_BF16)
    ndhwc_to_ncdhw_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, dst, C, Cp, P);
  else
    ndhwc_to_ncdhw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, dst, C, Cp, P);
  return check_launch("dp_ndhwc_to_ncdhw_f32");
}

DP_API int dp_u8_frames_to_ndhwc(const uint8_t* src, void* dst, const float* mean3, int B, int T, int H, int W, int Cp,
                                 int dtype, void* stream) {
  DP_REQUIRE(src && dst && mean3, DP_ERR_SHAPE, "dp_u8_frames_to_ndhwc: NULL pointer");
  DP_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && Cp >= 8 && Cp % 8 == 0, DP_ERR_SHAPE, "dp_u8_frames_to_ndhwc: bad shape");
  const int64_t rows = (int64_t)B * T * H * W;
  const int grid = ceil_div(rows, 256);
  if (dtype == DP_BF16)
    u8_to_ndhwc_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, mean3[0], mean3[1],
                                                                           mean3[2], rows, Cp);
  else
    u8_to_ndhwc_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(src, (float*)dst, mean3[0], mean3[1], mean3[2], rows,
                                                                   Cp);
  return check_launch("dp_u8_frames_to_ndhwc");
}

DP_API int dp_avgpool_fwd(const void* x, float* out, int B, int64_t pixels, int C, int Cp, int dtype, void* stream) {
  DP_REQUIRE(x && out, DP_ERR_SHAPE, "dp_avgpool_fwd: NULL pointer");
  DP_REQUIRE(B > 0 && B <= 65535 && pixels > 0 && C > 0 && Cp >= C && Cp % 8 == 0, DP_ERR_SHAPE, "dp_avgpool_fwd: bad shape");
  dim3 grid(ceil_div(C, 8), B);
  if (dtype == DP_BF16)
    avgpool_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, out, pixels, C, Cp);
  else
    avgpool_fwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, out, pixels, C, Cp);
  return check_launch("dp_avgpool_fwd");
}

DP_API int dp_avgpool_bwd(const float* dout, void* dx, int B, int64_t pixels, int C, int Cp, int dtype, void* stream) {
  DP_REQUIRE(dout && dx, DP_ERR_SHAPE, "dp_avgpool_bwd: NULL pointer");
  DP_REQUIRE(B > 0 && pixels > 0 && C > 0 && Cp >= C && Cp % 8 == 0, DP_ERR_SHAPE, "dp_avgpool_bwd: bad shape");
  const int64_t nvec = (int64_t)B * pixels * (Cp / 8);
  int64_t g = (nvec + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (dtype == DP_BF16)
    avgpool_bwd_kernel<__nv_bfloat16><<<(int)g, 256, 0, as_stream(stream)>>>(dout, (__nv_bfloat16*)dx, pixels, C, Cp, nvec);
  else
    avgpool_bwd_kernel<float><<<(int)g, 256, 0, as_stream(stream)>>>(dout, (float*)dx, pixels, C, Cp, nvec);
  return check_launch("dp_avgpool_bwd");
}
