// Stem fast path: the first convolution of R(2+1)D (3 -> 45 channels, kernel (1,7,7), stride (1,2,2),
// nn.Conv3d at /root/reference/src/models/R2Plus1D.py:137-140 via Conv3dBlock :44-51) on bf16 tensor cores.
//
// With 3 input channels an NDHWC row is 6 bytes: padding it to 16 channels multiplies the largest tensor of the
// network by 5.3 and turns the conv into 49 taps of K=16.  Instead the clip is stored as "packed rows"
//     XP[b][t][h][wp][4]  (bf16; wp = w + pw, zero pixels left and right, channel 3 = 0; WP = 2*Wo + 6)
// in which the kw*C values one output pixel needs from one input row are 32 CONTIGUOUS elements starting at
// pixel 2*wo (16-byte aligned).  A tensor map with an overlapping w stride of 8 elements exposes this as a
// (B,T,H,Wo,32) tensor, and the stem becomes an ordinary convolution over that view with kernel (1,kh,1),
// stride (1,sh,1), 32 "channels" = (jw, c) pairs: K = kh*32 = 224 instead of 784, 7 TMA loads per tile instead
// of 49, and 187 MB instead of 705 MB of input at batch 64.  Forward and weight gradient reuse the generic
// tcgen05 kernels (conv_tc.cu / wgrad_tc.cu) through their view entry points; this file owns the layout
// kernels, the weight packing and the C ABI.  The stem needs no data gradient (clips carry no grad).
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "bn_fin.cuh"

namespace dp {

constexpr int STEM_WIN = 8;   // pixels per window (kw <= 8)
constexpr int STEM_CH = 4;    // channels per packed pixel (C <= 4)

static bool stem_ok(const dp_conv_desc* d) {
  return d != nullptr && d->dtype == DP_BF16 && d->C >= 1 && d->C <= STEM_CH && d->kt == 1 && d->st == 1 && d->pt == 0 &&
         d->kw >= 1 && d->kw <= STEM_WIN && d->sw == 2 && d->pw >= 0 && d->pw <= 8 && d->kh >= 1 && d->kh <= 16 &&
         d->sh >= 1 && d->sh <= 4 && d->Kp % 16 == 0 && d->B > 0 && d->To == d->Ti;
}
static inline int stem_wp(const dp_conv_desc* d) { return 2 * d->Wo + STEM_WIN - 2; }

// the stem as a convolution over the (B,T,H,Wo,32) window view
static dp_conv_desc stem_view(const dp_conv_desc* d, long long* xstr) {
  dp_conv_desc v = *d;
  v.Wi = d->Wo;
  v.C = STEM_WIN * STEM_CH; v.Cp = STEM_WIN * STEM_CH;
  v.kw = 1; v.sw = 1; v.pw = 0;
  const long long WP = stem_wp(d);
  xstr[0] = 2 * STEM_CH;                       // next output pixel = two input pixels further
  xstr[1] = WP * STEM_CH;
  xstr[2] = (long long)d->Hi * WP * STEM_CH;
  xstr[3] = (long long)d->Ti * d->Hi * WP * STEM_CH;
  return v;
}

template <typename SRC>
__global__ void __launch_bounds__(256)
stem_pack_input_kernel(const SRC* __restrict__ src, __nv_bfloat16* __restrict__ xp, int C, int W, int WP, int pw,
                       int64_t rows /* B*T*H */, int64_t plane /* T*H*W */, int TH, float m0, float m1, float m2,
                       int u8_frames) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * WP) return;
  const int wp = (int)(idx % WP);
  const int64_t row = idx / WP;            // (b*T + t)*H + h
  const int w = wp - pw;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (w >= 0 && w < W) {
    if (u8_frames) {                        // frames (B,T,H,W,3) uint8 minus mean
      const int64_t o = (row * W + w) * 3;
      v[0] = (float)src[o] - m0; v[1] = (float)src[o + 1] - m1; v[2] = (float)src[o + 2] - m2;
    } else {                                // NCDHW fp32
      const int64_t b = row / TH, th = row % TH;
      for (int c = 0; c < C; ++c) v[c] = (float)src[(b * C + c) * plane + th * W + w];
    }
  }
  st4(xp + idx * 4, make_float4(v[0], v[1], v[2], v[3]));
}

// wv[k][jh][jw*4 + c] = w[k][c][0][jh][jw]
__global__ void stem_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wv, int K, int Kp, int C,
                                         int kh, int kw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Kp * kh * 32;
  if (idx >= total) return;
  const int e = idx % 32, jh = (idx / 32) % kh, k = idx / (32 * kh);
  const int jw = e / 4, c = e % 4;
  float v = 0.f;
  if (k < K && c < C && jw < kw) v = w[(((int64_t)k * C + c) * kh + jh) * kw + jw];
  wv[idx] = __float2bfloat16(v);
}

// dw[k][c][0][jh][jw] = dwv[k][jw*4 + c][jh]      (dwv in the generic (K, C'=32, taps=kh) order)
__global__ void stem_unpack_dw_kernel(const float* __restrict__ dwv, float* __restrict__ dw, int K, int C, int kh, int kw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = K * C * kh * kw;
  if (idx >= total) return;
  const int jw = idx % kw, jh = (idx / kw) % kh, c = (idx / (kw * kh)) % C, k = idx / (kw * kh * C);
  dw[idx] = dwv[((int64_t)k * 32 + (jw * 4 + c)) * kh + jh];
}

}  // namespace dp

using namespace dp;

DP_API int dp_stem_supported(const dp_conv_desc* d) {
  if (!stem_ok(d)) return 0;
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  return (tc_fwd_supported(&v) && tc_wgrad_supported(&v)) ? 1 : 0;
}

DP_API size_t dp_stem_input_elems(const dp_conv_desc* d) {
  if (!stem_ok(d)) return 0;
  // one extra window of slack so the last overlapping row stays inside the allocation
  return (size_t)d->B * d->Ti * d->Hi * stem_wp(d) * STEM_CH + STEM_WIN * STEM_CH;
}

static int stem_pack_input(const dp_conv_desc* d, const void* src, int u8, const float* mean3, void* xp, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(src && xp, DP_ERR_SHAPE, "stem pack input: NULL pointer");
  DP_REQUIRE(!u8 || d->C == 3, DP_ERR_SHAPE, "stem pack input: uint8 frames have 3 channels");
  const int WP = stem_wp(d);
  const int64_t rows = (int64_t)d->B * d->Ti * d->Hi;
  const int64_t total = rows * WP;
  const float m0 = mean3 ? mean3[0] : 0.f, m1 = mean3 ? mean3[1] : 0.f, m2 = mean3 ? mean3[2] : 0.f;
  const int grid = ceil_div(total, 256);
  if (u8)
    stem_pack_input_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>(
        (const uint8_t*)src, (__nv_bfloat16*)xp, d->C, d->Wi, WP, d->pw, rows, (int64_t)d->Ti * d->Hi * d->Wi,
        d->Ti * d->Hi, m0, m1, m2, 1);
  else
    stem_pack_input_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)src, (__nv_bfloat16*)xp, d->C, d->Wi, WP, d->pw, rows, (int64_t)d->Ti * d->Hi * d->Wi,
        d->Ti * d->Hi, 0.f, 0.f, 0.f, 0);
  return check_launch("stem_pack_input");
}

DP_API int dp_stem_pack_input_f32(const dp_conv_desc* d, const float* ncdhw, void* xp, void* stream) {
  return stem_pack_input(d, ncdhw, 0, nullptr, xp, stream);
}

DP_API int dp_stem_pack_input_u8(const dp_conv_desc* d, const uint8_t* frames, const float* mean3, void* xp,
                                 void* stream) {
  DP_REQUIRE(mean3 != nullptr, DP_ERR_SHAPE, "dp_stem_pack_input_u8: NULL mean");
  return stem_pack_input(d, frames, 1, mean3, xp, stream);
}

DP_API int dp_stem_pack_weights(const dp_conv_desc* d, const float* w, void* wv, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(w && wv, DP_ERR_SHAPE, "dp_stem_pack_weights: NULL pointer");
  const int total = d->Kp * d->kh * 32;
  stem_pack_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wv, d->K, d->Kp, d->C,
                                                                               d->kh, d->kw);
  return check_launch("stem_pack_weights");
}

DP_API int dp_stem_conv_fwd(const dp_conv_desc* d, const void* xp, const void* wv, void* y, float* part, int* nparts,
                            void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(xp && wv && y, DP_ERR_SHAPE, "dp_stem_conv_fwd: NULL pointer");
  DP_REQUIRE(part == nullptr || nparts != nullptr, DP_ERR_SHAPE, "dp_stem_conv_fwd: part given without nparts");
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  return tc_conv_fwd_view(&v, xs, xp, wv, y, part, nparts, as_stream(stream));
}

DP_API int dp_stem_conv_fwd_fin(const dp_conv_desc* d, const void* xp, const void* wv, void* y, float* part,
                                const dp_bn_fin* fin, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(xp && wv && y && part, DP_ERR_SHAPE, "dp_stem_conv_fwd_fin: NULL pointer");
  const int rc = bn_fin_validate(fin, 1, d->Kp, "dp_stem_conv_fwd_fin");
  if (rc != DP_OK) return rc;
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  int nparts = 0;
  return tc_conv_fwd_view(&v, xs, xp, wv, y, part, &nparts, as_stream(stream), fin);
}

DP_API int dp_stem_conv_fwd_bnact(const dp_conv_desc* d, const void* xp, const void* wv, const float* scale_shift, float slope,
                                  void* z, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(xp && wv && z && scale_shift, DP_ERR_SHAPE, "dp_stem_conv_fwd_bnact: NULL pointer");
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  return tc_conv_fwd_bnact(&v, xs, xp, wv, scale_shift, slope, nullptr, 1.f, z, as_stream(stream));
}

DP_API size_t dp_stem_wgrad_workspace(const dp_conv_desc* d) {
  if (!stem_ok(d)) return 0;
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  // split partials + the (K,32,kh) view gradient the unpack kernel reads
  return tc_wgrad_workspace(&v) + (size_t)d->K * 32 * d->kh * sizeof(float) + 256;
}

DP_API int dp_stem_conv_wgrad(const dp_conv_desc* d, const void* xp, const void* dy, float* dw, void* workspace,
                              size_t workspace_bytes, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(xp && dy && dw && workspace, DP_ERR_SHAPE, "dp_stem_conv_wgrad: NULL pointer");
  DP_REQUIRE(workspace_bytes >= dp_stem_wgrad_workspace(d), DP_ERR_SHAPE, "dp_stem_conv_wgrad: workspace too small");
  long long xs[4];
  dp_conv_desc v = stem_view(d, xs);
  const size_t part_bytes = (tc_wgrad_workspace(&v) + 255) / 256 * 256;
  float* dwv = reinterpret_cast<float*>(static_cast<char*>(workspace) + part_bytes);
  int rc = tc_conv_wgrad_view(&v, xs, xp, dy, dwv, workspace, part_bytes, as_stream(stream));
  if (rc != DP_OK) return rc;
  const int total = d->K * d->C * d->kh * d->kw;
  stem_unpack_dw_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(dwv, dw, d->K, d->C, d->kh, d->kw);
  return check_launch("stem_unpack_dw");
}
