// Fused forward+backward of the three class-imbalance losses of the path, one launch each:
//   CE    : sum_i w[y_i] * CE_i                                 (/root/reference/src/loss.py:71-81)
//   Focal : sum_i w[y_i] * (1 - exp(-CE_i))^gamma * CE_i        (loss.py:14-34, CE unweighted, sum)
//   LDAM  : sum_i w[y_i] * CE(s * (z_i - m[y_i] onehot)) / sum_i w[y_i]     (loss.py:37-69)
// Rows are streamed with coalesced loads, reduced by warp shuffles, block partials go to a
// workspace and the last CTA to finish sums them in fp64 in a fixed order (deterministic).
// The LDAM margins stay on the device (the reference round-trips them through the host).
#include "dp_common.cuh"

namespace dp {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_GRID = 2048;
constexpr int LOSS_MAX_C = 64;

struct LossWs {
  unsigned int counter;
  unsigned int pad[3];
  float2 partial[LOSS_MAX_GRID];
};

template <int KIND, int CT>  // CT: compile-time class count (2) or 0 = runtime C
__global__ void __launch_bounds__(LOSS_THREADS)
loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, const float* __restrict__ weight,
            const float* __restrict__ margins, float gamma, float s, int64_t n, int C,
            float* __restrict__ loss_out, float* __restrict__ dlogits, LossWs* ws) {
  float acc_l = 0.f, acc_w = 0.f;
  const int CC = CT ? CT : C;
  constexpr int ZN = CT ? CT : LOSS_MAX_C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t y = target[i];
    const float* z = logits + i * CC;
    float* d = dlogits + i * CC;
    if (y < 0 || y >= CC) {  // ignore_index-like: no loss, zero gradient
      for (int j = 0; j < CC; ++j) d[j] = 0.f;
      continue;
    }
    const float wy = weight != nullptr ? weight[y] : 1.f;
    const float my = (KIND == DP_LOSS_LDAM) ? margins[y] : 0.f;
    const float sc = (KIND == DP_LOSS_LDAM) ? s : 1.f;
    float zl[ZN];
    float zmax = -INFINITY, zy = 0.f;
    if (CT == 2) {
      const float2 v = *reinterpret_cast<const float2*>(z);
      zl[0] = v.x; zl[1] = v.y;
    } else {
      for (int j = 0; j < CC; ++j) zl[j] = z[j];
    }
#pragma unroll
    for (int j = 0; j < CC; ++j) {
      float v = zl[j];
      if (KIND == DP_LOSS_LDAM) v = sc * (j == y ? v - my : v);
      zl[j] = v;
      if (j == y) zy = v;
      zmax = fmaxf(zmax, v);
    }
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < CC; ++j) se += expf(zl[j] - zmax);
    const float lse = zmax + logf(se);
    const float ce = lse - zy;
    float li, dce;
    if (KIND == DP_LOSS_FOCAL) {
      const float p = expf(-ce);
      const float q = 1.f - p;
      float f, df;  // f = q^gamma, df = d f / d ce = gamma q^(gamma-1) p
      if (gamma == 0.f) { f = 1.f; df = 0.f; }
      else if (gamma == 1.f) { f = q; df = p; }
      else if (gamma == 2.f) { f = q * q; df = 2.f * q * p; }
      else { f = powf(q, gamma); df = gamma * powf(q, gamma - 1.f) * p; }
      li = wy * f * ce;
      dce = wy * (f + df * ce);
    } else {
      li = wy * ce;
      dce = wy * sc;
    }
    acc_l += li;
    acc_w += wy;
    if (CT == 2) {
      float2 o;
      o.x = dce * (expf(zl[0] - lse) - (y == 0 ? 1.f : 0.f));
      o.y = dce * (expf(zl[1] - lse) - (y == 1 ? 1.f : 0.f));
      *reinterpret_cast<float2*>(d) = o;
    } else {
      for (int j = 0; j < CC; ++j) d[j] = dce * (expf(zl[j] - lse) - (j == y ? 1.f : 0.f));
    }
  }
  // block reduction
  __shared__ float sl[LOSS_THREADS / 32], sw[LOSS_THREADS / 32];
  __shared__ bool is_last;
  acc_l = warp_sum(acc_l);
  acc_w = warp_sum(acc_w);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sl[wid] = acc_l; sw[wid] = acc_w; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < LOSS_THREADS / 32; ++k) { a += sl[k]; b += sw[k]; }
    ws->partial[blockIdx.x] = make_float2(a, b);
    __threadfence();
    const unsigned int ticket = atomicAdd(&ws->counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double L = 0.0, Wn = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) {
      const float px = *reinterpret_cast<volatile float*>(&ws->partial[k].x);
      const float py = *reinterpret_cast<volatile float*>(&ws->partial[k].y);
      L += (double)px; Wn += (double)py;
    }
    if (KIND == DP_LOSS_LDAM) {
      loss_out[0] = (float)(L / Wn);
      loss_out[1] = (float)Wn;
    } else {
      loss_out[0] = (float)L;
      loss_out[1] = 1.f;
    }
    ws->counter = 0u;  // self-resetting workspace
  }
}

// Two classes (the disruption / normal problem of this path): a row is 8 B of logits, 8 B of target, 8 B of gradient.
// One thread owns PAIRS of rows -- 16-byte vector loads and stores -- and keeps two pairs in flight per iteration, the
// two class weights / margins live in registers (no dependent gather on weight[y]): the kernel is then bound by HBM,
// not by the latency of one dependent load chain per thread.
template <int KIND>
__device__ __forceinline__ void loss_row_c2(float z0, float z1, int64_t y, float w0, float w1, float m0, float m1, float gamma,
                                            float sc, float& acc_l, float& acc_w, float& o0, float& o1) {
  if (y < 0 || y >= 2) { o0 = 0.f; o1 = 0.f; return; }
  const float wy = y ? w1 : w0;
  if (KIND == DP_LOSS_LDAM) {
    z0 = sc * (y == 0 ? z0 - m0 : z0);
    z1 = sc * (y == 1 ? z1 - m1 : z1);
  }
  const float zy = y ? z1 : z0;
  const float zmax = fmaxf(z0, z1);
  const float se = expf(z0 - zmax) + expf(z1 - zmax);
  const float lse = zmax + logf(se);
  const float ce = lse - zy;
  float li, dce;
  if (KIND == DP_LOSS_FOCAL) {
    const float p = expf(-ce);
    const float q = 1.f - p;
    float f, df;
    if (gamma == 0.f) { f = 1.f; df = 0.f; }
    else if (gamma == 1.f) { f = q; df = p; }
    else if (gamma == 2.f) { f = q * q; df = 2.f * q * p; }
    else { f = powf(q, gamma); df = gamma * powf(q, gamma - 1.f) * p; }
    li = wy * f * ce;
    dce = wy * (f + df * ce);
  } else {
    li = wy * ce;
    dce = wy * sc;
  }
  acc_l += li;
  acc_w += wy;
  o0 = dce * (expf(z0 - lse) - (y == 0 ? 1.f : 0.f));
  o1 = dce * (expf(z1 - lse) - (y == 1 ? 1.f : 0.f));
}

template <int KIND>
__global__ void __launch_bounds__(LOSS_THREADS)
loss_kernel_c2(const float* __restrict__ logits, const int64_t* __restrict__ target, const float* __restrict__ weight,
               const float* __restrict__ margins, float gamma, float s, int64_t n, float* __restrict__ loss_out,
               float* __restrict__ dlogits, LossWs* ws) {
  float acc_l = 0.f, acc_w = 0.f;
  const float w0 = weight != nullptr ? weight[0] : 1.f, w1 = weight != nullptr ? weight[1] : 1.f;
  const float m0 = KIND == DP_LOSS_LDAM ? margins[0] : 0.f, m1 = KIND == DP_LOSS_LDAM ? margins[1] : 0.f;
  const float sc = KIND == DP_LOSS_LDAM ? s : 1.f;
  const int64_t npairs = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4* z4 = reinterpret_cast<const float4*>(logits);
  const longlong2* y2 = reinterpret_cast<const longlong2*>(target);
  float4* d4 = reinterpret_cast<float4*>(dlogits);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < npairs; i += 2 * stride) {   // two pairs (four rows) in flight
    const float4 za = z4[i], zb = z4[i + stride];
    const longlong2 ya = y2[i], yb = y2[i + stride];
    float4 oa, ob;
    loss_row_c2<KIND>(za.x, za.y, ya.x, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, oa.x, oa.y);
    loss_row_c2<KIND>(za.z, za.w, ya.y, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, oa.z, oa.w);
    loss_row_c2<KIND>(zb.x, zb.y, yb.x, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, ob.x, ob.y);
    loss_row_c2<KIND>(zb.z, zb.w, yb.y, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, ob.z, ob.w);
    d4[i] = oa;
    d4[i + stride] = ob;
  }
  if (i < npairs) {
    const float4 za = z4[i];
    const longlong2 ya = y2[i];
    float4 oa;
    loss_row_c2<KIND>(za.x, za.y, ya.x, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, oa.x, oa.y);
    loss_row_c2<KIND>(za.z, za.w, ya.y, w0, w1, m0, m1, gamma, sc, acc_l, acc_w, oa.z, oa.w);
    d4[i] = oa;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd row count: the last row
    float o0, o1;
    loss_row_c2<KIND>(logits[2 * (n - 1)], logits[2 * (n - 1) + 1], target[n - 1], w0, w1, m0, m1, gamma, sc, acc_l, acc_w, o0, o1);
    dlogits[2 * (n - 1)] = o0;
    dlogits[2 * (n - 1) + 1] = o1;
  }
  __shared__ float sl[LOSS_THREADS / 32], sw[LOSS_THREADS / 32];
  __shared__ bool is_last;
  acc_l = warp_sum(acc_l);
  acc_w = warp_sum(acc_w);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sl[wid] = acc_l; sw[wid] = acc_w; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < LOSS_THREADS / 32; ++k) { a += sl[k]; b += sw[k]; }
    ws->partial[blockIdx.x] = make_float2(a, b);
    __threadfence();
    const unsigned int ticket = atomicAdd(&ws->counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {   // fixed-order fp64 combine of the block partials, one warp
    __threadfence();
    double L = 0.0, Wn = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += 32) {
      L += (double)*reinterpret_cast<volatile float*>(&ws->partial[k].x);
      Wn += (double)*reinterpret_cast<volatile float*>(&ws->partial[k].y);
    }
    for (int o = 16; o > 0; o >>= 1) {
      L += __shfl_xor_sync(0xffffffffu, L, o);
      Wn += __shfl_xor_sync(0xffffffffu, Wn, o);
    }
    if (threadIdx.x == 0) {
      if (KIND == DP_LOSS_LDAM) { loss_out[0] = (float)(L / Wn); loss_out[1] = (float)Wn; }
      else { loss_out[0] = (float)L; loss_out[1] = 1.f; }
      ws->counter = 0u;
    }
  }
}

__global__ void loss_bwd_scale_kernel(const float* __restrict__ dlogits, const float* __restrict__ grad_out,
                                      const float* __restrict__ loss_out, float* __restrict__ out, int64_t count) {
  const float f = grad_out[0] / loss_out[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = dlogits[i] * f;
}

}  // namespace dp

using namespace dp;

DP_API size_t dp_loss_workspace(int64_t n) {
  (void)n;
  return sizeof(LossWs);
}

DP_API int dp_loss_fwd_bwd(int kind, const float* logits, const int64_t* target, const float* weight,
                           const float* margins, float gamma, float s, int64_t n, int C, float* loss_out,
                           float* dlogits, void* workspace, void* stream) {
  DP_REQUIRE(logits && target && loss_out && dlogits && workspace, DP_ERR_SHAPE, "dp_loss_fwd_bwd: NULL pointer");
  DP_REQUIRE(n > 0 && C >= 2 && C <= LOSS_MAX_C, DP_ERR_SHAPE, "dp_loss_fwd_bwd: n=%lld C=%d unsupported (2..%d classes)",
             (long long)n, C, LOSS_MAX_C);
  DP_REQUIRE(kind != DP_LOSS_LDAM || margins != nullptr, DP_ERR_SHAPE, "dp_loss_fwd_bwd: LDAM needs margins");
  DP_REQUIRE(kind != DP_LOSS_FOCAL || gamma >= 0.f, DP_ERR_SHAPE, "dp_loss_fwd_bwd: gamma must be >= 0");
  int64_t g = (n + LOSS_THREADS - 1) / LOSS_THREADS;
  if (g > LOSS_MAX_GRID) g = LOSS_MAX_GRID;
  LossWs* ws = (LossWs*)workspace;
  cudaStream_t st = as_stream(stream);
  // two-class fast path: 16-byte vectors need 16-byte aligned tensors
  const bool vec_ok = (((uintptr_t)logits | (uintptr_t)target | (uintptr_t)dlogits) & 15) == 0;
  int64_t g2 = (n / 2 + 2 * LOSS_THREADS - 1) / (2 * LOSS_THREADS);   // two pairs per thread and iteration
  if (g2 < 1) g2 = 1;
  if (g2 > LOSS_MAX_GRID) g2 = LOSS_MAX_GRID;
#define DP_LAUNCH_LOSS(K)                                                                                      \
  do {                                                                                                         \
    if (C == 2 && vec_ok)                                                                                      \
      loss_kernel_c2<K><<<(int)g2, LOSS_THREADS, 0, st>>>(logits, target, weight, margins, gamma, s, n,        \
                                                          loss_out, dlogits, ws);                              \
    else if (C == 2)                                                                                           \
      loss_kernel<K, 2><<<(int)g, LOSS_THREADS, 0, st>>>(logits, target, weight, margins, gamma, s, n, C,      \
                                                         loss_out, dlogits, ws);                               \
    else                                                                                                       \
      loss_kernel<K, 0><<<(int)g, LOSS_THREADS, 0, st>>>(logits, target, weight, margins, gamma, s, n, C,      \
                                                         loss_out, dlogits, ws);                               \
  } while (0)
  switch (kind) {
    case DP_LOSS_CE: DP_LAUNCH_LOSS(DP_LOSS_CE); break;
    case DP_LOSS_FOCAL: DP_LAUNCH_LOSS(DP_LOSS_FOCAL); break;
    case DP_LOSS_LDAM: DP_LAUNCH_LOSS(DP_LOSS_LDAM); break;
    default:
      dp::set_error("dp_loss_fwd_bwd: unknown loss kind %d", kind);
      return DP_ERR_UNSUPPORTED;
  }
#undef DP_LAUNCH_LOSS
  return check_launch("dp_loss_fwd_bwd");
}

DP_API int dp_loss_bwd_scale(const float* dlogits, const float* grad_out, const float* loss_out, float* out,
                             int64_t count, void* stream) {
  DP_REQUIRE(dlogits && grad_out && loss_out && out && count > 0, DP_ERR_SHAPE, "dp_loss_bwd_scale: bad arguments");
  int64_t g = (count + 255) / 256;
  if (g > 1184) g = 1184;
  loss_bwd_scale_kernel<<<(int)g, 256, 0, as_stream(stream)>>>(dlogits, grad_out, loss_out, out, count);
  return check_launch("dp_loss_bwd_scale");
}
