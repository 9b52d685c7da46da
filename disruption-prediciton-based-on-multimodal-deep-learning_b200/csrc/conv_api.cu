// C-ABI entry points for the convolution family: validation, weight packing, dispatch between
// the tcgen05 kernels and the CUDA-core kernels.  No cuDNN, no CPU path.
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "bn_fin.cuh"

namespace dp {

static int validate(const dp_conv_desc* d) {
  DP_REQUIRE(d != nullptr, DP_ERR_SHAPE, "conv desc is NULL");
  DP_REQUIRE(d->dtype == DP_F32 || d->dtype == DP_BF16, DP_ERR_UNSUPPORTED, "conv desc: bad dtype %d", d->dtype);
  DP_REQUIRE(d->B > 0 && d->Ti > 0 && d->Hi > 0 && d->Wi > 0 && d->C > 0 && d->K > 0, DP_ERR_SHAPE,
             "conv desc: non-positive dimension");
  DP_REQUIRE(d->Cp >= d->C && d->Cp % 16 == 0 && d->Kp >= d->K && d->Kp % 16 == 0, DP_ERR_ALIGN,
             "conv desc: padded channels must be multiples of 16 (Cp=%d Kp=%d)", d->Cp, d->Kp);
  DP_REQUIRE(d->kt > 0 && d->kh > 0 && d->kw > 0 && d->st > 0 && d->sh > 0 && d->sw > 0, DP_ERR_SHAPE,
             "conv desc: bad kernel/stride");
  const int To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  const int Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  DP_REQUIRE(To == d->To && Ho == d->Ho && Wo == d->Wo && To > 0 && Ho > 0 && Wo > 0, DP_ERR_SHAPE,
             "conv desc: output dims (%d,%d,%d) do not match geometry (%d,%d,%d)", d->To, d->Ho, d->Wo, To, Ho, Wo);
  return DP_OK;
}

template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd,
                                    int K, int C, int Kp, int Cp, int taps) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)Kp * taps * Cp;
  if (idx >= total) return;
  const int c = (int)(idx % Cp), tap = (int)((idx / Cp) % taps), n = (int)(idx / ((int64_t)Cp * taps));
  float v = 0.f;
  if (n < K && c < C) v = w[((int64_t)n * C + c) * taps + tap];
  if (wf != nullptr) wf[idx] = (T)v;
  if (wd != nullptr) wd[((int64_t)c * taps + tap) * Kp + n] = (T)v;
}

unsigned long long g_simt_launches = 0;   // conv launches served by the CUDA-core family (dp_simt_launch_count)
unsigned long long g_simt_fallbacks = 0;  // ... of which DP_IMPL_AUTO fell back from tcgen05 on a bf16 tensor
int g_strict_tc = 0;                      // option "strict_tc": DP_IMPL_AUTO refuses to fall back in bf16

static int resolve_impl(const dp_conv_desc* d, int op, int impl) {
  if (impl == DP_IMPL_SIMT) { ++g_simt_launches; return DP_IMPL_SIMT; }
  bool ok = false;
  if (d->dtype == DP_BF16) {
    if (op == 0) ok = tc_fwd_supported(d);
    else if (op == 1) ok = tc_dgrad_supported(d);
    else ok = tc_wgrad_supported(d);
  }
  if (impl == DP_IMPL_TC) return ok ? DP_IMPL_TC : -1;
  if (ok) return DP_IMPL_TC;
  // DP_IMPL_AUTO and the tcgen05 planner declined: the CUDA-core family serves the call.  That is the fp32 validation
  // mode's normal path; on a bf16 tensor it is a silent slow path, so it is counted (bench.py reports it, must be 0)
  // and refused outright under the "strict_tc" option.
  if (d->dtype == DP_BF16) {
    if (g_strict_tc) return -1;
    ++g_simt_fallbacks;
  }
  ++g_simt_launches;
  return DP_IMPL_SIMT;
}

// queries (dp_conv_supported, workspace sizing) must not disturb the counters
static int resolve_impl_query(const dp_conv_desc* d, int op, int impl) {
  const unsigned long long a = g_simt_launches, b = g_simt_fallbacks;
  const int r = resolve_impl(d, op, impl);
  g_simt_launches = a; g_simt_fallbacks = b;
  return r;
}

}  // namespace dp

using namespace dp;

DP_API int dp_pack_weights(const dp_conv_desc* d, const float* w, void* w_fwd, void* w_dgrad, void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(w != nullptr, DP_ERR_SHAPE, "dp_pack_weights: NULL weight");
  const int taps = d->kt * d->kh * d->kw;
  const int64_t total = (int64_t)d->Kp * taps * d->Cp;
  const int grid = ceil_div(total, 256);
  if (d->dtype == DP_BF16)
    launch_pdl(pack_weights_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), w, (__nv_bfloat16*)w_fwd,
               (__nv_bfloat16*)w_dgrad, d->K, d->C, d->Kp, d->Cp, taps);
  else
    pack_weights_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w, (float*)w_fwd, (float*)w_dgrad, d->K, d->C,
                                                                    d->Kp, d->Cp, taps);
  return check_launch("dp_pack_weights");
}

DP_API int dp_conv_supported(const dp_conv_desc* d, int op, int impl) {
  if (validate(d) != DP_OK) return 0;
  if (op == 3)   // is dp_conv_dgrad_bnstats fused in the tcgen05 epilogue (1) or composed of dgrad + reduction (0)?
    return (resolve_impl_query(d, 1, impl) == DP_IMPL_TC && tc_dgrad_bnstats_supported(d)) ? 1 : 0;
  if (op == 4)   // eval-mode fused conv + BatchNorm + LeakyReLU epilogue on the tcgen05 family
    return (impl != DP_IMPL_SIMT && tc_fwd_bnact_supported(d)) ? 1 : 0;
  if (op < 0 || op > 2) return 0;
  return resolve_impl_query(d, op, impl) > 0 ? 1 : 0;
}

static int conv_fwd_any(const dp_conv_desc* d, const void* x, const void* w_fwd, void* y, float* part, int* nparts,
                        const dp_bn_fin* fin, int impl, void* stream, const char* who) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(x && w_fwd && y, DP_ERR_SHAPE, "%s: NULL pointer", who);
  const int r = resolve_impl(d, 0, impl);
  DP_REQUIRE(r > 0, DP_ERR_UNSUPPORTED, "%s: geometry not covered by the tcgen05 family", who);
  cudaStream_t s = as_stream(stream);
  if (r == DP_IMPL_TC) return tc_conv_fwd(d, x, w_fwd, y, part, nparts, s, fin);
  rc = simt_conv_fwd(d, x, w_fwd, y, s);
  if (rc != DP_OK) return rc;
  if (part != nullptr) {
    const int64_t rows = (int64_t)d->B * d->To * d->Ho * d->Wo;
    return bn_stats_launch(y, rows, d->Kp, d->dtype, part, nparts, s, fin);   // fin: last CTA of the reduction finalises
  }
  return DP_OK;
}

DP_API int dp_conv_fwd(const dp_conv_desc* d, const void* x, const void* w_fwd, void* y, float* part, int* nparts,
                       int impl, void* stream) {
  DP_REQUIRE(part == nullptr || nparts != nullptr, DP_ERR_SHAPE, "dp_conv_fwd: part given without nparts");
  return conv_fwd_any(d, x, w_fwd, y, part, nparts, nullptr, impl, stream, "dp_conv_fwd");
}

DP_API int dp_conv_fwd_fin(const dp_conv_desc* d, const void* x, const void* w_fwd, void* y, float* part,
                           const dp_bn_fin* fin, int impl, void* stream) {
  DP_REQUIRE(d != nullptr && part != nullptr, DP_ERR_SHAPE, "dp_conv_fwd_fin: NULL pointer");
  const int rc = bn_fin_validate(fin, 1, d->Kp, "dp_conv_fwd_fin");
  if (rc != DP_OK) return rc;
  int nparts = 0;
  return conv_fwd_any(d, x, w_fwd, y, part, &nparts, fin, impl, stream, "dp_conv_fwd_fin");
}

DP_API unsigned long long dp_simt_launch_count(void) { return g_simt_launches; }
DP_API unsigned long long dp_simt_fallback_count(void) { return g_simt_fallbacks; }

DP_API int dp_conv_fwd_bnact(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w_fwd,
                             const float* scale_shift, float slope, const void* residual, float slope_res, void* z,
                             int impl, void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(x && w_fwd && z && scale_shift, DP_ERR_SHAPE, "dp_conv_fwd_bnact: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (impl != DP_IMPL_SIMT && tc_fwd_bnact_supported(d))
    return tc_conv_fwd_bnact(d, xstrides, x, w_fwd, scale_shift, slope, residual, slope_res, z, s);
  DP_REQUIRE(impl != DP_IMPL_TC && !(g_strict_tc && d->dtype == DP_BF16), DP_ERR_UNSUPPORTED,
             "dp_conv_fwd_bnact: geometry not covered by the tcgen05 family");
  DP_REQUIRE(xstrides == nullptr, DP_ERR_UNSUPPORTED, "dp_conv_fwd_bnact: input views need the tcgen05 family");
  // CUDA-core composition (fp32 validation mode): conv into z, then the affine + activation pass in place
  ++g_simt_launches;
  if (d->dtype == DP_BF16) ++g_simt_fallbacks;
  rc = simt_conv_fwd(d, x, w_fwd, z, s);
  if (rc != DP_OK) return rc;
  const int64_t rows = (int64_t)d->B * d->To * d->Ho * d->Wo;
  return dp_bn_act_apply(z, scale_shift, scale_shift + d->Kp, slope, residual, slope_res, z, rows, d->Kp, d->dtype, stream);
}

DP_API int dp_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                         int impl, void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(dy && w_dgrad && dx, DP_ERR_SHAPE, "dp_conv_dgrad: NULL pointer");
  const int r = resolve_impl(d, 1, impl);
  DP_REQUIRE(r > 0, DP_ERR_UNSUPPORTED, "dp_conv_dgrad: geometry not covered by the tcgen05 family");
  cudaStream_t s = as_stream(stream);
  if (r == DP_IMPL_TC) return tc_conv_dgrad(d, dy, w_dgrad, addend, dx, s);
  return simt_conv_dgrad(d, dy, w_dgrad, addend, dx, s);
}

DP_API size_t dp_dgrad_classes_weight_elems(const dp_conv_desc* d, int impl) {
  if (validate(d) != DP_OK || impl == DP_IMPL_SIMT) return 0;
  return tc_dgrad_classes_weight_elems(d);
}

DP_API int dp_pack_weights_dgrad_classes(const dp_conv_desc* d, const float* w, void* w_cls, void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(w && w_cls, DP_ERR_SHAPE, "dp_pack_weights_dgrad_classes: NULL pointer");
  return tc_pack_dgrad_classes(d, w, w_cls, as_stream(stream));
}

DP_API int dp_conv_dgrad_classes(const dp_conv_desc* d, const void* dy, const void* w_cls, const void* addend, void* dx,
                                 void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(dy && w_cls && dx, DP_ERR_SHAPE, "dp_conv_dgrad_classes: NULL pointer");
  return tc_conv_dgrad_classes(d, dy, w_cls, addend, dx, as_stream(stream));
}

static int conv_dgrad_bnstats_any(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                                  const void* y_prev, const float* scale_shift, float slope, float* part, int* nparts,
                                  const dp_bn_fin* fin, int impl, void* stream, const char* who) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(dy && w_dgrad && dx && y_prev && scale_shift && part, DP_ERR_SHAPE, "%s: NULL pointer", who);
  const int r = resolve_impl(d, 1, impl);
  DP_REQUIRE(r > 0, DP_ERR_UNSUPPORTED, "%s: geometry not covered by the tcgen05 family", who);
  cudaStream_t s = as_stream(stream);
  if (r == DP_IMPL_TC && tc_dgrad_bnstats_supported(d))
    return tc_conv_dgrad_bnstats(d, dy, w_dgrad, addend, dx, y_prev, scale_shift, slope, part, nparts, s, fin);
  rc = (r == DP_IMPL_TC) ? tc_conv_dgrad(d, dy, w_dgrad, addend, dx, s) : simt_conv_dgrad(d, dy, w_dgrad, addend, dx, s);
  if (rc != DP_OK) return rc;
  const int64_t rows = (int64_t)d->B * d->Ti * d->Hi * d->Wi;
  // mean / rstd are not read by the reduction (the raw-y sums are centred in the finalisation)
  if (fin != nullptr)
    return dp_bn_act_bwd_reduce_fin(dx, y_prev, nullptr, scale_shift, scale_shift + d->Cp, slope, 1.f, part, rows, d->Cp,
                                    d->dtype, fin, stream);
  return dp_bn_act_bwd_reduce(dx, y_prev, nullptr, scale_shift, scale_shift + d->Cp, scale_shift, scale_shift, slope, 1.f, part,
                              nparts, rows, d->Cp, d->dtype, stream);
}

DP_API int dp_conv_dgrad_bnstats(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                                 const void* y_prev, const float* scale_shift, float slope, float* part, int* nparts,
                                 int impl, void* stream) {
  DP_REQUIRE(nparts != nullptr, DP_ERR_SHAPE, "dp_conv_dgrad_bnstats: NULL pointer");
  return conv_dgrad_bnstats_any(d, dy, w_dgrad, addend, dx, y_prev, scale_shift, slope, part, nparts, nullptr, impl, stream,
                                "dp_conv_dgrad_bnstats");
}

DP_API int dp_conv_dgrad_bnstats_fin(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                                     const void* y_prev, const float* scale_shift, float slope, float* part,
                                     const dp_bn_fin* fin, int impl, void* stream) {
  DP_REQUIRE(d != nullptr, DP_ERR_SHAPE, "dp_conv_dgrad_bnstats_fin: NULL pointer");
  const int rc = bn_fin_validate(fin, 2, d->Cp, "dp_conv_dgrad_bnstats_fin");
  if (rc != DP_OK) return rc;
  int nparts = 0;
  return conv_dgrad_bnstats_any(d, dy, w_dgrad, addend, dx, y_prev, scale_shift, slope, part, &nparts, fin, impl, stream,
                                "dp_conv_dgrad_bnstats_fin");
}

DP_API size_t dp_conv_wgrad_workspace(const dp_conv_desc* d, int impl) {
  if (validate(d) != DP_OK) return 0;
  size_t a = simt_wgrad_workspace(d);
  if (impl != DP_IMPL_SIMT && d->dtype == DP_BF16 && tc_wgrad_supported(d)) {  // (sizing query: no counters)
    size_t b = tc_wgrad_workspace(d);
    if (b > a) a = b;
  }
  return a;
}

DP_API int dp_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                         size_t workspace_bytes, int impl, void* stream) {
  int rc = validate(d);
  if (rc != DP_OK) return rc;
  DP_REQUIRE(x && dy && dw && workspace, DP_ERR_SHAPE, "dp_conv_wgrad: NULL pointer");
  const int r = resolve_impl(d, 2, impl);
  DP_REQUIRE(r > 0, DP_ERR_UNSUPPORTED, "dp_conv_wgrad: geometry not covered by the tcgen05 family");
  cudaStream_t s = as_stream(stream);
  if (r == DP_IMPL_TC) {
    DP_REQUIRE(workspace_bytes >= tc_wgrad_workspace(d), DP_ERR_SHAPE, "dp_conv_wgrad: workspace too small");
    return tc_conv_wgrad(d, x, dy, dw, workspace, workspace_bytes, s);
  }
  DP_REQUIRE(workspace_bytes >= simt_wgrad_workspace(d), DP_ERR_SHAPE, "dp_conv_wgrad: workspace too small");
  return simt_conv_wgrad(d, x, dy, dw, workspace, s);
}
