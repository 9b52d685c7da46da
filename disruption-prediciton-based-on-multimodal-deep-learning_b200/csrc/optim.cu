// Fused optimiser tail over ONE flat fp32 parameter/gradient bucket:
// global L2 norm -> clip coefficient (kept on the device) -> AdamW update.
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step over 102 small tensors
// (/root/reference/src/train.py:63-66) with two launches and no host synchronisation.
#include "dp_common.cuh"

namespace dp {

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_GRID = 1184;

// Layout is part of the ABI (include/dp_b200.h): the host reads / restores `step` and `skipped` (checkpoints).
struct OptWs {
  unsigned int counter;
  unsigned int step;      // device-side step count (used when the host passes step == 0: CUDA-graph replay)
  unsigned int skipped;   // steps skipped because the gradient norm was not finite (reference: train.py:55-60 `continue`)
  unsigned int skip_now;  // 1 while the CURRENT step is being skipped (written by sqnorm, read by adamw)
  float partial[OPT_MAX_GRID];
};

__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_kernel(const float* __restrict__ g, int64_t n, float grad_scale, float* __restrict__ norm_out, OptWs* ws,
              int dev_step) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc = fmaf(g[i], g[i], acc);
  __shared__ float sw[OPT_THREADS / 32];
  __shared__ bool is_last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < OPT_THREADS / 32; ++k) a += sw[k];
    ws->partial[blockIdx.x] = a;
    __threadfence();
    is_last = (atomicAdd(&ws->counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) t += (double)*reinterpret_cast<volatile float*>(&ws->partial[k]);
    const float nrm = (float)(sqrt(t) * (double)grad_scale);
    norm_out[0] = nrm;
    ws->counter = 0u;
    // a non-finite loss gives non-finite gradients: the reference skips backward and step (train.py:55-60); here the
    // whole update is skipped on the device (weights, moments and the step count stay untouched) and counted
    const bool bad = !isfinite(nrm);
    ws->skip_now = bad ? 1u : 0u;
    if (bad) ws->skipped += 1u;
    else if (dev_step) ws->step += 1u;
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
             float max_norm, float grad_scale, const float* __restrict__ norm, const OptWs* __restrict__ ws, int dev_step) {
  if (ws->skip_now) return;   // non-finite gradient norm: leave p, m, v as they are
  if (dev_step) {   // bias corrections from the device-side step count
    const float st = (float)ws->step;
    bc1 = 1.f - powf(beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(beta2, st));
  }
  float gs = grad_scale;
  if (max_norm > 0.f) {
    const float coef = max_norm / (norm[0] + 1e-6f);
    gs *= fminf(coef, 1.f);
  }
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * weight_decay;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gs;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
  }
}

}  // namespace dp

using namespace dp;

DP_API size_t dp_optim_workspace(int64_t n) {
  (void)n;
  return sizeof(OptWs);
}

DP_API int dp_grad_sqnorm(const float* g, int64_t n, float grad_scale, int count_step, float* norm_out, void* workspace,
                          void* stream) {
  DP_REQUIRE(g && norm_out && workspace, DP_ERR_SHAPE, "dp_grad_sqnorm: NULL pointer");
  DP_REQUIRE(n > 0, DP_ERR_SHAPE, "dp_grad_sqnorm: n=%lld", (long long)n);
  DP_REQUIRE(((uintptr_t)g & 15) == 0, DP_ERR_ALIGN, "dp_grad_sqnorm: grad bucket must be 16-byte aligned");
  int64_t grid = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  if (grid > OPT_MAX_GRID) grid = OPT_MAX_GRID;
  if (grid < 1) grid = 1;
  sqnorm_kernel<<<(int)grid, OPT_THREADS, 0, as_stream(stream)>>>(g, n, grad_scale, norm_out, (OptWs*)workspace, count_step ? 1 : 0);
  return check_launch("dp_grad_sqnorm");
}

DP_API int dp_adamw_apply(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float max_norm, float grad_scale, const float* norm,
                          const void* workspace, void* stream) {
  DP_REQUIRE(p && g && m && v && norm && workspace, DP_ERR_SHAPE, "dp_adamw_apply: NULL pointer");
  DP_REQUIRE(n > 0 && step >= 0, DP_ERR_SHAPE, "dp_adamw_apply: n=%lld step=%d", (long long)n, step);
  const int dev_step = step == 0 ? 1 : 0;
  const float bc1 = 1.f - powf(beta1, (float)(step > 0 ? step : 1));
  const float bc2 = 1.f - powf(beta2, (float)(step > 0 ? step : 1));
  int64_t g2 = (n + OPT_THREADS - 1) / OPT_THREADS;
  if (g2 > OPT_MAX_GRID) g2 = OPT_MAX_GRID;
  adamw_kernel<<<(int)g2, OPT_THREADS, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2),
                                                max_norm, grad_scale, norm, (const OptWs*)workspace, dev_step);
  return check_launch("dp_adamw_apply");
}

DP_API int dp_clip_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float max_norm, float grad_scale,
                              float* norm_out, void* workspace, void* stream) {
  DP_REQUIRE(p && g && m && v && norm_out && workspace, DP_ERR_SHAPE, "dp_clip_adamw_step: NULL pointer");
  int rc = dp_grad_sqnorm(g, n, grad_scale, step == 0, norm_out, workspace, stream);
  if (rc != DP_OK) return rc;
  return dp_adamw_apply(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, max_norm, grad_scale, norm_out, workspace,
                        stream);
}
