// Stem fast path: the first convolution of R(2+1)D (3 -> 45 channels, kernel (1,7,7), stride (1,2,2),
// nn.Conv3d at /root/reference/src/models/R2Plus1D.py:137-140 via Conv3dBlock :44-51) on bf16 tensor cores.
//
// With 3 input channels an NDHWC row is 6 bytes: padding it to 16 channels multiplies the largest tensor of the
// network by 5.3 and turns the conv into 49 taps of K=16.  Instead the clip is stored as "packed row pairs"
//     XP[b][t][P][wp][r][4]   (bf16; padded row h' = h + ph = 2*P + r, wp = w + pw, zeros outside the image,
//                              channel 3 = 0; HP = Ho + ceil(kh/2) - 1 pairs, WP = 2*Wo + 6 pixels)
// in which the 2 x kw x C values one output pixel needs from one PAIR of input rows are 64 CONTIGUOUS elements
// starting at pixel 2*wo (32-byte aligned).  A tensor map with an overlapping w stride of 16 elements exposes this as
// a (B,T,HP,Wo,64) tensor, and -- because output row ho reads the pairs ho .. ho+3 -- the stride-2 7x7 stem becomes
// an ordinary STRIDE-1 convolution over that view with kernel (1,4,1) and 64 "channels" = (jw, r, c) triples:
// K = 4*64 = 256 (the 8th row of the four pairs carries zero weights).  Stride 1 lets the generic tcgen05 kernels
// (conv_tc.cu / wgrad_tc.cu, view entry points) take their halo mode: ONE TMA box of (bh+3) x bw rows of 128 bytes per
// tile feeds all four taps, where the former single-row layout (kernel (1,7,1), stride 2, 32 channels) needed seven
// boxes of 128 rows of 64 bytes -- 896 TMA rows per 128 pixels, the request rate that bounded the stem at 20 % of its
// roofline.  187 -> 193 MB of input at batch 64 (NDHWC with 16 channels would be 705 MB).  This file owns the layout
// kernels, the weight packing and the C ABI.  The stem needs no data gradient (clips carry no grad).
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "bn_fin.cuh"

namespace dp {

constexpr int STEM_WIN = 8;    // pixels per window (kw <= 8)
constexpr int STEM_CH = 4;     // channels per packed pixel (C <= 4)
constexpr int STEM_K = 2 * STEM_WIN * STEM_CH;   // 64 view channels: (jw, r, c)

static bool stem_ok(const dp_conv_desc* d) {
  return d != nullptr && d->dtype == DP_BF16 && d->C >= 1 && d->C <= STEM_CH && d->kt == 1 && d->st == 1 && d->pt == 0 &&
         d->kw >= 1 && d->kw <= STEM_WIN && d->sw == 2 && d->pw >= 0 && d->pw <= 8 && d->kh >= 1 && d->kh <= 16 &&
         d->sh == 2 && d->ph >= 0 && d->Kp % 16 == 0 && d->B > 0 && d->To == d->Ti;
}
static inline int stem_wp(const dp_conv_desc* d) { return 2 * d->Wo + STEM_WIN - 2; }
static inline int stem_npair(const dp_conv_desc* d) { return (d->kh + 1) / 2; }
static inline int stem_hp(const dp_conv_desc* d) { return d->Ho + stem_npair(d) - 1; }

// the stem as a stride-1 convolution over the (B,T,HP,Wo,64) window view
static dp_conv_desc stem_view(const dp_conv_desc* d, long long* xstr) {
  dp_conv_desc v = *d;
  v.Wi = d->Wo;
  v.Hi = stem_hp(d);
  v.C = STEM_K; v.Cp = STEM_K;
  v.kw = 1; v.sw = 1; v.pw = 0;
  v.kh = stem_npair(d); v.sh = 1; v.ph = 0;
  const long long WP = stem_wp(d), HP = stem_hp(d);
  xstr[0] = 2 * 2 * STEM_CH;                   // next output pixel = two input pixels (of two rows each) further
  xstr[1] = WP * 2 * STEM_CH;
  xstr[2] = HP * WP * 2 * STEM_CH;
  xstr[3] = (long long)d->Ti * HP * WP * 2 * STEM_CH;
  return v;
}

// one thread = one packed pixel (both rows of the pair): 8 bf16 = 16 bytes
template <typename SRC>
__global__ void __launch_bounds__(256)
stem_pack_input_kernel(const SRC* __restrict__ src, __nv_bfloat16* __restrict__ xp, int C, int H, int W, int HP, int WP, int ph,
                       int pw, int64_t pairs /* B*T*HP */, int64_t plane /* T*H*W */, int T, float m0, float m1, float m2,
                       int u8_frames) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pairs * WP) return;
  const int wp = (int)(idx % WP);
  const int64_t prow = idx / WP;            // (b*T + t)*HP + P
  const int P = (int)(prow % HP);
  const int64_t bt = prow / HP;
  const int w = wp - pw;
  f8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
  if (w >= 0 && w < W) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int h = 2 * P + r - ph;
      if (h < 0 || h >= H) continue;
      if (u8_frames) {                        // frames (B,T,H,W,3) uint8 minus mean
        const int64_t a = ((bt * H + h) * W + w) * 3;
        o.v[4 * r + 0] = (float)src[a] - m0; o.v[4 * r + 1] = (float)src[a + 1] - m1; o.v[4 * r + 2] = (float)src[a + 2] - m2;
      } else {                                // NCDHW fp32
        const int64_t b = bt / T, t = bt % T;
        for (int c = 0; c < C; ++c) o.v[4 * r + c] = (float)src[(b * C + c) * plane + ((int64_t)t * H + h) * W + w];
      }
    }
  }
  st8(xp + idx * 8, o);
}

// wv[k][j][jw*8 + r*4 + c] = w[k][c][0][2*j + r][jw]   (zero for the row past the kernel)
__global__ void stem_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wv, int K, int Kp, int C,
                                         int kh, int kw, int npair) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Kp * npair * STEM_K;
  if (idx >= total) return;
  const int e = idx % STEM_K, j = (idx / STEM_K) % npair, k = idx / (STEM_K * npair);
  const int jw = e / 8, r = (e / 4) % 2, c = e % 4, jh = 2 * j + r;
  float v = 0.f;
  if (k < K && c < C && jw < kw && jh < kh) v = w[(((int64_t)k * C + c) * kh + jh) * kw + jw];
  wv[idx] = __float2bfloat16(v);
}

// dw[k][c][0][jh][jw] = dwv[k][jw*8 + (jh&1)*4 + c][jh>>1]      (dwv in the generic (K, C'=64, taps=npair) order)
__global__ void stem_unpack_dw_kernel(const float* __restrict__ dwv, float* __restrict__ dw, int K, int C, int kh, int kw,
                                      int npair) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = K * C * kh * kw;
  if (idx >= total) return;
  const int jw = idx % kw, jh = (idx / kw) % kh, c = (idx / (kw * kh)) % C, k = idx / (kw * kh * C);
  dw[idx] = dwv[((int64_t)k * STEM_K + (jw * 8 + (jh & 1) * 4 + c)) * npair + (jh >> 1)];
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
  return (size_t)d->B * d->Ti * stem_hp(d) * stem_wp(d) * 2 * STEM_CH + STEM_K;
}

DP_API size_t dp_stem_weight_elems(const dp_conv_desc* d) {
  if (!stem_ok(d)) return 0;
  return (size_t)d->Kp * stem_npair(d) * STEM_K;
}

static int stem_pack_input(const dp_conv_desc* d, const void* src, int u8, const float* mean3, void* xp, void* stream) {
  DP_REQUIRE(stem_ok(d), DP_ERR_UNSUPPORTED, "stem path: geometry not covered");
  DP_REQUIRE(src && xp, DP_ERR_SHAPE, "stem pack input: NULL pointer");
  DP_REQUIRE(!u8 || d->C == 3, DP_ERR_SHAPE, "stem pack input: uint8 frames have 3 channels");
  const int WP = stem_wp(d), HP = stem_hp(d);
  const int64_t pairs = (int64_t)d->B * d->Ti * HP;
  const int64_t total = pairs * WP;
  const float m0 = mean3 ? mean3[0] : 0.f, m1 = mean3 ? mean3[1] : 0.f, m2 = mean3 ? mean3[2] : 0.f;
  const int grid = ceil_div(total, 256);
  if (u8)
    stem_pack_input_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>(
        (const uint8_t*)src, (__nv_bfloat16*)xp, d->C, d->Hi, d->Wi, HP, WP, d->ph, d->pw, pairs,
        (int64_t)d->Ti * d->Hi * d->Wi, d->Ti, m0, m1, m2, 1);
  else
    stem_pack_input_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)src, (__nv_bfloat16*)xp, d->C, d->Hi, d->Wi, HP, WP, d->ph, d->pw, pairs,
        (int64_t)d->Ti * d->Hi * d->Wi, d->Ti, 0.f, 0.f, 0.f, 0);
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
  const int total = d->Kp * stem_npair(d) * STEM_K;
  stem_pack_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wv, d->K, d->Kp, d->C,
                                                                               d->kh, d->kw, stem_npair(d));
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
  // split partials + the (K,64,npair) view gradient the unpack kernel reads
  return tc_wgrad_workspace(&v) + (size_t)d->K * STEM_K * stem_npair(d) * sizeof(float) + 256;
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
  stem_unpack_dw_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(dwv, dw, d->K, d->C, d->kh, d->kw, stem_npair(d));
  return check_launch("stem_unpack_dw");
}
