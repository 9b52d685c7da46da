// Memory-bound auxiliaries of the SlowFast pathways (BASELINE config 3) on NDHWC activations:
//   squeeze-excite channel scaling fused with Swish   (/root/reference/src/models/resnet.py:63-81 SwishEfficient,
//                                                      :172-200 Bottleneck3D.forward: global_pool -> fc1 -> ReLU -> fc2 ->
//                                                      sigmoid -> scale -> swish)
//   MaxPool3d((1,3,3),(1,2,2),(0,1,1))                (resnet.py:220-225, layer0)
//   channel concatenation of the lateral connections  (/root/reference/src/models/slowfast.py:26-36, torch.cat(dim=1))
// forward and backward, 16-byte vectors along the padded channel dimension where the layout allows.
#include "dp_common.cuh"
#include <float.h>

namespace dp {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---- out = swish(x * gate[b][c]) ----
template <typename T>
__global__ void __launch_bounds__(256)
se_swish_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gate, T* __restrict__ out, int64_t pixels, int C,
                    int Cp, int64_t nvec) {
  const int vpr = Cp >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % vpr);
    const int64_t b = (v / vpr) / pixels;
    f8 a = ld8(x + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      const float g = (gate != nullptr && c < C) ? gate[b * C + c] : 1.f;
      const float u = a.v[j] * g;
      a.v[j] = c < C ? u * sigmoidf_(u) : 0.f;
    }
    st8(out + v * 8, a);
  }
}

// dx = dout * swish'(u) * gate,  dgate[b][c] = sum_pixels dout * swish'(u) * x,   u = x * gate
// one CTA per (8-channel vector, clip): the gate gradient is reduced in the CTA (deterministic order)
template <typename T>
__global__ void __launch_bounds__(256)
se_swish_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gate, const T* __restrict__ dout, T* __restrict__ dx,
                    float* __restrict__ dgate, int64_t pixels, int C, int Cp) {
  const int cv = blockIdx.x, b = blockIdx.y;
  const int64_t base = (int64_t)b * pixels * Cp + cv * 8;
  float g[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    g[j] = (gate != nullptr && c < C) ? gate[(int64_t)b * C + c] : 1.f;
    acc[j] = 0.f;
  }
  for (int64_t p = threadIdx.x; p < pixels; p += blockDim.x) {
    const f8 a = ld8(x + base + p * Cp), d = ld8(dout + base + p * Cp);
    f8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = a.v[j] * g[j];
      const float s = sigmoidf_(u);
      const float ds = d.v[j] * (s * (1.f + u * (1.f - s)));    // dout * d swish / du
      o.v[j] = (cv * 8 + j) < C ? ds * g[j] : 0.f;
      acc[j] = fmaf(ds, a.v[j], acc[j]);
    }
    st8(dx + base + p * Cp, o);
  }
  if (dgate == nullptr) return;
  __shared__ float red[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int c = cv * 8 + threadIdx.x;
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (c < C) dgate[(int64_t)b * C + c] = s;
  }
}

// ---- MaxPool (1,3,3) stride (1,2,2) pad (0,1,1); idx = winning tap 0..8 (first maximum in scan order, as ATen) ----
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, uint8_t* __restrict__ idx, int64_t frames, int H, int W,
                   int Ho, int Wo, int Cp, int64_t nvec) {
  const int vpr = Cp >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % vpr);
    int64_t r = v / vpr;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho); r /= Ho;   // r = frame index (b*T + t)
    f8 best;
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best.v[j] = -FLT_MAX; bi[j] = 0; }
    for (int jh = 0; jh < 3; ++jh) {
      const int h = 2 * ho - 1 + jh;
      if (h < 0 || h >= H) continue;
      for (int jw = 0; jw < 3; ++jw) {
        const int w = 2 * wo - 1 + jw;
        if (w < 0 || w >= W) continue;
        const f8 a = ld8(x + ((r * H + h) * W + w) * Cp + cv * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (a.v[j] > best.v[j]) { best.v[j] = a.v[j]; bi[j] = jh * 3 + jw; }
      }
    }
    st8(out + v * 8, best);
    if (idx != nullptr) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + v * 8) = pk;
    }
  }
}

// gather form (no atomics): an input pixel collects the gradient of the <= 4 windows that contain it and chose it
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ idx, T* __restrict__ dx, int64_t frames, int H,
                   int W, int Ho, int Wo, int Cp, int64_t nvec) {
  const int vpr = Cp >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int cv = (int)(v % vpr);
    int64_t r = v / vpr;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    f8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    // windows ho with 2*ho - 1 <= h <= 2*ho + 1
    for (int ho = (h >> 1); ho <= ((h + 1) >> 1); ++ho) {
      if (ho < 0 || ho >= Ho) continue;
      const int jh = h - (2 * ho - 1);
      for (int wo = (w >> 1); wo <= ((w + 1) >> 1); ++wo) {
        if (wo < 0 || wo >= Wo) continue;
        const int jw = w - (2 * wo - 1);
        const int tap = jh * 3 + jw;
        const int64_t o = ((r * Ho + ho) * Wo + wo) * Cp + cv * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
        const f8 d = ld8(dout + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = (int)(((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xffu);
          if (t == tap) acc.v[j] += d.v[j];
        }
      }
    }
    st8(dx + v * 8, acc);
  }
}

// ---- out[row][0:Ca] = a[row][0:Ca], out[row][Ca:Ca+Cb] = b[row][0:Cb], out[row][Ca+Cb:Cop] = 0 ----
template <typename T>
__global__ void __launch_bounds__(256)
concat_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t rows, int Ca, int Cap, int Cb,
              int Cbp, int Cop) {
  const int64_t total = rows * Cop;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % Cop);
    const int64_t row = i / Cop;
    T v = (T)0.f;
    if (c < Ca) v = a[row * Cap + c];
    else if (c < Ca + Cb) v = b[row * Cbp + (c - Ca)];
    out[i] = v;
  }
}

// backward of the concatenation: da / db from the matching channel ranges of dout, padded channels zero
template <typename T>
__global__ void __launch_bounds__(256)
split_kernel(const T* __restrict__ dout, T* __restrict__ da, T* __restrict__ db, int64_t rows, int Ca, int Cap, int Cb,
             int Cbp, int Cop) {
  const int64_t na = rows * Cap, total = na + rows * Cbp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (i < na) {
      const int c = (int)(i % Cap);
      const int64_t row = i / Cap;
      da[i] = c < Ca ? dout[row * Cop + c] : (T)0.f;
    } else {
      const int64_t k = i - na;
      const int c = (int)(k % Cbp);
      const int64_t row = k / Cbp;
      db[k] = c < Cb ? dout[row * Cop + Ca + c] : (T)0.f;
    }
  }
}

static int ew_blocks(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace dp

using namespace dp;

DP_API int dp_se_swish_fwd(const void* x, const float* gate, void* out, int B, int64_t pixels, int C, int Cp, int dtype,
                           void* stream) {
  DP_REQUIRE(x && out, DP_ERR_SHAPE, "dp_se_swish_fwd: NULL pointer");
  DP_REQUIRE(B > 0 && pixels > 0 && C > 0 && Cp >= C && Cp % 8 == 0, DP_ERR_SHAPE, "dp_se_swish_fwd: bad shape");
  const int64_t nvec = (int64_t)B * pixels * (Cp / 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    se_swish_fwd_kernel<__nv_bfloat16><<<ew_blocks(nvec), 256, 0, st>>>((const __nv_bfloat16*)x, gate, (__nv_bfloat16*)out, pixels, C, Cp, nvec);
  else
    se_swish_fwd_kernel<float><<<ew_blocks(nvec), 256, 0, st>>>((const float*)x, gate, (float*)out, pixels, C, Cp, nvec);
  return check_launch("dp_se_swish_fwd");
}

DP_API int dp_se_swish_bwd(const void* x, const float* gate, const void* dout, void* dx, float* dgate, int B, int64_t pixels,
                           int C, int Cp, int dtype, void* stream) {
  DP_REQUIRE(x && dout && dx, DP_ERR_SHAPE, "dp_se_swish_bwd: NULL pointer");
  DP_REQUIRE((gate == nullptr) == (dgate == nullptr), DP_ERR_SHAPE, "dp_se_swish_bwd: gate and dgate go together");
  DP_REQUIRE(B > 0 && B <= 65535 && pixels > 0 && C > 0 && Cp >= C && Cp % 8 == 0, DP_ERR_SHAPE, "dp_se_swish_bwd: bad shape");
  dim3 grid(Cp / 8, B);
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    se_swish_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, gate, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dx, dgate, pixels, C, Cp);
  else
    se_swish_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, gate, (const float*)dout, (float*)dx, dgate, pixels, C, Cp);
  return check_launch("dp_se_swish_bwd");
}

DP_API int dp_maxpool_hw_fwd(const void* x, void* out, uint8_t* idx, int64_t frames, int H, int W, int Cp, int dtype,
                             void* stream) {
  DP_REQUIRE(x && out, DP_ERR_SHAPE, "dp_maxpool_hw_fwd: NULL pointer");
  DP_REQUIRE(frames > 0 && H > 0 && W > 0 && Cp > 0 && Cp % 8 == 0, DP_ERR_SHAPE, "dp_maxpool_hw_fwd: bad shape");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t nvec = frames * Ho * Wo * (Cp / 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    maxpool_fwd_kernel<__nv_bfloat16><<<ew_blocks(nvec), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, idx, frames, H, W, Ho, Wo, Cp, nvec);
  else
    maxpool_fwd_kernel<float><<<ew_blocks(nvec), 256, 0, st>>>((const float*)x, (float*)out, idx, frames, H, W, Ho, Wo, Cp, nvec);
  return check_launch("dp_maxpool_hw_fwd");
}

DP_API int dp_maxpool_hw_bwd(const void* dout, const uint8_t* idx, void* dx, int64_t frames, int H, int W, int Cp, int dtype,
                             void* stream) {
  DP_REQUIRE(dout && idx && dx, DP_ERR_SHAPE, "dp_maxpool_hw_bwd: NULL pointer");
  DP_REQUIRE(frames > 0 && H > 0 && W > 0 && Cp > 0 && Cp % 8 == 0, DP_ERR_SHAPE, "dp_maxpool_hw_bwd: bad shape");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t nvec = frames * H * W * (Cp / 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    maxpool_bwd_kernel<__nv_bfloat16><<<ew_blocks(nvec), 256, 0, st>>>((const __nv_bfloat16*)dout, idx, (__nv_bfloat16*)dx, frames, H, W, Ho, Wo, Cp, nvec);
  else
    maxpool_bwd_kernel<float><<<ew_blocks(nvec), 256, 0, st>>>((const float*)dout, idx, (float*)dx, frames, H, W, Ho, Wo, Cp, nvec);
  return check_launch("dp_maxpool_hw_bwd");
}

DP_API int dp_concat_channels(const void* a, const void* b, void* out, int64_t rows, int Ca, int Cap, int Cb, int Cbp, int Cop,
                              int dtype, void* stream) {
  DP_REQUIRE(a && b && out, DP_ERR_SHAPE, "dp_concat_channels: NULL pointer");
  DP_REQUIRE(rows > 0 && Ca > 0 && Cb > 0 && Cap >= Ca && Cbp >= Cb && Cop >= Ca + Cb, DP_ERR_SHAPE, "dp_concat_channels: bad shape");
  const int64_t n = rows * Cop;
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    concat_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, rows, Ca, Cap, Cb, Cbp, Cop);
  else
    concat_kernel<float><<<ew_blocks(n), 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, rows, Ca, Cap, Cb, Cbp, Cop);
  return check_launch("dp_concat_channels");
}

DP_API int dp_split_channels(const void* dout, void* da, void* db, int64_t rows, int Ca, int Cap, int Cb, int Cbp, int Cop,
                             int dtype, void* stream) {
  DP_REQUIRE(dout && da && db, DP_ERR_SHAPE, "dp_split_channels: NULL pointer");
  DP_REQUIRE(rows > 0 && Ca > 0 && Cb > 0 && Cap >= Ca && Cbp >= Cb && Cop >= Ca + Cb, DP_ERR_SHAPE, "dp_split_channels: bad shape");
  const int64_t n = rows * (Cap + Cbp);
  cudaStream_t st = as_stream(stream);
  if (dtype == DP_BF16)
    split_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)da, (__nv_bfloat16*)db, rows, Ca, Cap, Cb, Cbp, Cop);
  else
    split_kernel<float><<<ew_blocks(n), 256, 0, st>>>((const float*)dout, (float*)da, (float*)db, rows, Ca, Cap, Cb, Cbp, Cop);
  return check_launch("dp_split_channels");
}
