// Memory-bound BatchNorm3d(train) + LeakyReLU kernels over NDHWC activations:
// statistics, finalize (+running stats), apply (+residual), and the two-pass backward.
// 16-byte vector loads along the padded channel dimension, per-thread channel-resident
// accumulators, shared-memory tree, per-CTA partials reduced deterministically in fp64.
//
// Replaces nn.BatchNorm3d / nn.LeakyReLU at /root/reference/src/models/R2Plus1D.py:53-57 and the
// residual add + LeakyReLU at :179-187 (and their autograd).
#include "dp_common.cuh"
#include "conv_internal.cuh"

namespace dp {

constexpr int RED_THREADS = 256;

// ---- generic per-channel column reduction of two quantities over [rows][Cp] ----
// F: __device__ void operator()(int64_t elem_offset, int c0, float* a8, float* b8) accumulates 8 channels
template <typename F>
__global__ void __launch_bounds__(RED_THREADS, 2)
col_reduce2_kernel(F f, int64_t rows, int Cp, float* __restrict__ part) {  // f by value: per-thread register copy
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][RED_THREADS * 8];
  const int vpr = Cp >> 3;                       // 8-channel vectors per row
  const int rpi = RED_THREADS / vpr;             // rows per iteration
  const int active = rpi * vpr;
  const int tid = threadIdx.x;
  const int64_t rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
  if (tid < active) {
    const int cv = tid % vpr, rl = tid / vpr;
    f.init(cv * 8);                       // this thread's 8 channels never change: parameters live in registers
    int64_t r = r0 + rl;
    for (; r + 7 * rpi < r1; r += 8 * rpi) {  // eight independent rows in flight (two CTAs of 256 threads per SM)
#pragma unroll
      for (int u = 0; u < 8; ++u) f((r + u * rpi) * Cp + cv * 8, a, b);
    }
    for (; r + 3 * rpi < r1; r += 4 * rpi) {
      f(r * Cp + cv * 8, a, b);
      f((r + rpi) * Cp + cv * 8, a, b);
      f((r + 2 * rpi) * Cp + cv * 8, a, b);
      f((r + 3 * rpi) * Cp + cv * 8, a, b);
    }
    for (; r < r1; r += rpi) f(r * Cp + cv * 8, a, b);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][tid * 8 + j] = a[j]; red[1][tid * 8 + j] = b[j]; }
  __syncthreads();
  // thread c sums channel c over the rpi row-lanes (layout: [(rl*vpr+cv)*8 + j] == [rl*Cp + c])
  for (int c = tid; c < Cp; c += RED_THREADS) {
    float sa = 0.f, sb = 0.f;
    for (int rl = 0; rl < rpi; ++rl) { sa += red[0][rl * Cp + c]; sb += red[1][rl * Cp + c]; }
    part[((int64_t)blockIdx.x * 2 + 0) * Cp + c] = sa;
    part[((int64_t)blockIdx.x * 2 + 1) * Cp + c] = sb;
  }
}

static int reduce_grid(int64_t rows) {
  int64_t g = (rows + 63) / 64;
  if (g > DP_MAX_PARTS) g = DP_MAX_PARTS;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
struct StatsF {
  const T* y;
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void operator()(int64_t off, float* a, float* b) const {
    const f8 v = ld8(y + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] += v.v[j]; b[j] = fmaf(v.v[j], v.v[j], b[j]); }
  }
};

template <typename T>
struct BwdReduceF {
  const T* dz; const T* y; const T* out;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  float slope, slope_res;
  float r_scale[8], r_shift[8];
  // accumulates sum(g) and sum(g*y) with the RAW conv output y; bn_bwd_finalize turns the second into
  // sum(g*xhat) = (sum(g*y) - mean*sum(g)) * rstd in fp64, so the streaming loop needs two parameters per channel
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { r_scale[j] = scale[c0 + j]; r_shift[j] = shift[c0 + j]; }
  }
  __device__ __forceinline__ void operator()(int64_t off, float* a, float* b) const {
    const f8 g = ld8(dz + off), v = ld8(y + off);
    f8 o;
    if (out != nullptr) o = ld8(out + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gg = g.v[j];
      if (out != nullptr) gg *= (o.v[j] > 0.f ? 1.f : slope_res);
      const float u = fmaf(v.v[j], r_scale[j], r_shift[j]);
      gg *= (u > 0.f ? 1.f : slope);
      a[j] += gg;
      b[j] = fmaf(gg, v.v[j], b[j]);
    }
  }
};

int bn_stats_launch(const void* y, int64_t rows, int Cp, int dtype, float* part, int* nparts, cudaStream_t s) {
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024, DP_ERR_ALIGN, "bn_stats: Cp=%d must be a multiple of 8, <= 1024", Cp);
  DP_REQUIRE(rows > 0, DP_ERR_SHAPE, "bn_stats: no rows");
  const int grid = reduce_grid(rows);
  if (dtype == DP_BF16) {
    StatsF<__nv_bfloat16> f{(const __nv_bfloat16*)y};
    launch_pdl(col_reduce2_kernel<StatsF<__nv_bfloat16>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part);
  } else {
    StatsF<float> f{(const float*)y};
    launch_pdl(col_reduce2_kernel<StatsF<float>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part);
  }
  *nparts = grid;
  return check_launch("bn_stats");
}

// ---- finalize: partials -> mean/rstd/scale/shift, running stats (momentum, unbiased var) ----
// One CTA of 256 threads per 32 channels: 8 part-lanes x 32 channel-lanes, fp64 partial sums combined in a fixed
// order (deterministic), reads coalesced along the channel dimension.
constexpr int FIN_CH = 32, FIN_PL = 8;
__device__ __forceinline__ void fin_reduce(const float* __restrict__ part, int nparts, int Cp, int c, int pl, bool cval,
                                           double& S, double& Q, double (*red)[FIN_PL][FIN_CH]) {
  double s = 0.0, q = 0.0;
  if (cval) {
    // four independent load/accumulate chains per thread: the loop is pure memory latency (<= 592 partials), and this
    // kernel sits on the critical path between every conv and its apply pass; the order stays fixed (deterministic)
    double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
    int p = pl;
    for (; p + 3 * FIN_PL < nparts; p += 4 * FIN_PL) {
      const float a0 = part[((int64_t)p * 2 + 0) * Cp + c], b0 = part[((int64_t)p * 2 + 1) * Cp + c];
      const float a1 = part[((int64_t)(p + FIN_PL) * 2 + 0) * Cp + c], b1 = part[((int64_t)(p + FIN_PL) * 2 + 1) * Cp + c];
      const float a2 = part[((int64_t)(p + 2 * FIN_PL) * 2 + 0) * Cp + c], b2 = part[((int64_t)(p + 2 * FIN_PL) * 2 + 1) * Cp + c];
      const float a3 = part[((int64_t)(p + 3 * FIN_PL) * 2 + 0) * Cp + c], b3 = part[((int64_t)(p + 3 * FIN_PL) * 2 + 1) * Cp + c];
      s += (double)a0; q += (double)b0; s1 += (double)a1; q1 += (double)b1;
      s2 += (double)a2; q2 += (double)b2; s3 += (double)a3; q3 += (double)b3;
    }
    for (; p < nparts; p += FIN_PL) {
      s += (double)part[((int64_t)p * 2 + 0) * Cp + c];
      q += (double)part[((int64_t)p * 2 + 1) * Cp + c];
    }
    s = (s + s1) + (s2 + s3);
    q = (q + q1) + (q2 + q3);
  }
  const int cl = threadIdx.x % FIN_CH;
  red[0][pl][cl] = s;
  red[1][pl][cl] = q;
  __syncthreads();
  S = 0.0; Q = 0.0;
  if (pl == 0) {
#pragma unroll
    for (int i = 0; i < FIN_PL; ++i) { S += red[0][i][cl]; Q += red[1][i][cl]; }
  }
}

__global__ void __launch_bounds__(FIN_CH * FIN_PL)
bn_finalize_kernel(const float* __restrict__ part, int nparts, int C, int Cp, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float momentum, float* running_mean, float* running_var,
                   float* mean, float* rstd, float* scale, float* shift) {
  __shared__ double red[2][FIN_PL][FIN_CH];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * FIN_CH + threadIdx.x % FIN_CH, pl = threadIdx.x / FIN_CH;
  double S, Q;
  fin_reduce(part, nparts, Cp, c, pl, c < C, S, Q, red);
  if (pl != 0 || c >= Cp) return;
  if (c >= C) { mean[c] = 0.f; rstd[c] = 0.f; scale[c] = 0.f; shift[c] = 0.f; return; }
  const double mu = S / count;
  double var = Q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rs;
  mean[c] = (float)mu;
  rstd[c] = rs;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mu * sc;
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* rm, const float* rv, const float* gamma, const float* beta,
                                      float eps, int C, int Cp, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  if (c >= C) { scale[c] = 0.f; shift[c] = 0.f; return; }
  const float rs = rsqrtf(rv[c] + eps);
  const float sc = gamma[c] * rs;
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

__global__ void __launch_bounds__(FIN_CH * FIN_PL)
bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, int Cp, double count,
                       const float* __restrict__ mean, const float* __restrict__ rstd,
                       float* dgamma, float* dbeta, float* coef) {
  __shared__ double red[2][FIN_PL][FIN_CH];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * FIN_CH + threadIdx.x % FIN_CH, pl = threadIdx.x / FIN_CH;
  double S, Q;
  fin_reduce(part, nparts, Cp, c, pl, c < C, S, Q, red);
  if (pl != 0 || c >= Cp) return;
  if (c >= C) { coef[c] = 0.f; coef[Cp + c] = 0.f; return; }
  Q = (Q - (double)mean[c] * S) * (double)rstd[c];   // sum(g*y) -> sum(g*xhat)
  if (dbeta != nullptr) dbeta[c] = (float)S;
  if (dgamma != nullptr) dgamma[c] = (float)Q;
  coef[c] = (float)(S / count);
  coef[Cp + c] = (float)(Q / count);
}

// ---- elementwise passes ----
// Grid-stride loops whose stride is a multiple of the vectors-per-row, so a thread always works on the same
// 8 channels and keeps their parameters in registers (no per-element parameter traffic).
template <typename T>
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                    float slope, const T* __restrict__ residual, float slope_res, T* __restrict__ z,
                    int64_t nvec, int Cp) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = Cp >> 3;
  const int64_t total = (int64_t)gridDim.x * blockDim.x;
  const int64_t stride = (total / vpr) * vpr;
  const int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v0 >= stride) return;
  const int c0 = (int)(v0 % vpr) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
  auto one = [&](const f8& a_in, const f8& r, int64_t v) {
    f8 a = a_in;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float u = lrelu(fmaf(a.v[j], sc[j], sh[j]), slope);
      if (residual != nullptr) u = lrelu(u + r.v[j], slope_res);
      a.v[j] = u;
    }
    st8(z + v * 8, a);
  };
  int64_t v = v0;
  for (; v + stride < nvec; v += 2 * stride) {   // two independent vectors in flight per thread
    const f8 a0 = ld8(y + v * 8), a1 = ld8(y + (v + stride) * 8);
    f8 r0, r1;
    if (residual != nullptr) { r0 = ld8(residual + v * 8); r1 = ld8(residual + (v + stride) * 8); }
    one(a0, r0, v);
    one(a1, r1, v + stride);
  }
  if (v < nvec) {
    const f8 a0 = ld8(y + v * 8);
    f8 r0;
    if (residual != nullptr) r0 = ld8(residual + v * 8);
    one(a0, r0, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ y, const T* __restrict__ out,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ coef, float slope, float slope_res,
                        T* __restrict__ dy, T* __restrict__ dres, int64_t nvec, int Cp) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = Cp >> 3;
  const int64_t total = (int64_t)gridDim.x * blockDim.x;
  const int64_t stride = (total / vpr) * vpr;
  const int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v0 >= stride) return;
  const int c0 = (int)(v0 % vpr) * 8;
  // dy = scale*(g - c0 - xhat*c1),  xhat = (y - mean)*rstd   =>   dy = scale*g + ka + kb*y
  float sc[8], sh[8], ka[8], kb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    sc[j] = scale[c]; sh[j] = shift[c];
    const float c1r = coef[Cp + c] * rstd[c];
    kb[j] = -sc[j] * c1r;
    ka[j] = -sc[j] * (coef[c] - mean[c] * c1r);
  }
  for (int64_t v = v0; v < nvec; v += stride) {
    const f8 g = ld8(dz + v * 8), yy = ld8(y + v * 8);
    f8 o, res, d;
    if (out != nullptr) o = ld8(out + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gg = g.v[j];
      if (out != nullptr) gg *= (o.v[j] > 0.f ? 1.f : slope_res);
      res.v[j] = gg;
      const float u = fmaf(yy.v[j], sc[j], sh[j]);
      gg *= (u > 0.f ? 1.f : slope);
      d.v[j] = fmaf(sc[j], gg, fmaf(kb[j], yy.v[j], ka[j]));
    }
    st8(dy + v * 8, d);
    if (dres != nullptr) st8(dres + v * 8, res);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    f8 x = ld8(a + v * 8);
    const f8 y = ld8(b + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) x.v[j] += y.v[j];
    st8(o + v * 8, x);
  }
}

static int ew_grid(int64_t nvec, int vpr) {
  int64_t g = (nvec + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  const int64_t gmin = (vpr + 255) / 256;   // the loop stride must hold at least one full row of vectors
  if (g < gmin) g = gmin;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace dp

using namespace dp;

DP_API int dp_bn_stats(const void* y, int64_t rows, int Cp, int dtype, float* part, int* nparts, void* stream) {
  DP_REQUIRE(y && part && nparts, DP_ERR_SHAPE, "dp_bn_stats: NULL pointer");
  return bn_stats_launch(y, rows, Cp, dtype, part, nparts, as_stream(stream));
}

DP_API int dp_bn_finalize(const float* part, int nparts, int C, int Cp, double count, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* mean, float* rstd, float* scale, float* shift, void* stream) {
  DP_REQUIRE(part && gamma && beta && mean && rstd && scale && shift, DP_ERR_SHAPE, "dp_bn_finalize: NULL pointer");
  DP_REQUIRE(nparts > 0 && nparts <= DP_MAX_PARTS && C > 0 && Cp >= C && count > 0, DP_ERR_SHAPE,
             "dp_bn_finalize: bad sizes (nparts=%d C=%d Cp=%d)", nparts, C, Cp);
  DP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), DP_ERR_SHAPE,
             "dp_bn_finalize: running_mean/var must both be given or both NULL");
  launch_pdl(bn_finalize_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, C, Cp, count,
             gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift);
  return check_launch("dp_bn_finalize");
}

DP_API int dp_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma,
                             const float* beta, float eps, int C, int Cp, float* scale, float* shift, void* stream) {
  DP_REQUIRE(running_mean && running_var && gamma && beta && scale && shift, DP_ERR_SHAPE,
             "dp_bn_eval_coeffs: NULL pointer");
  bn_eval_coeffs_kernel<<<ceil_div(Cp, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, gamma, beta, eps,
                                                                         C, Cp, scale, shift);
  return check_launch("dp_bn_eval_coeffs");
}

DP_API int dp_bn_act_apply(const void* y, const float* scale, const float* shift, float slope, const void* residual,
                           float slope_res, void* z, int64_t rows, int Cp, int dtype, void* stream) {
  DP_REQUIRE(y && scale && shift && z, DP_ERR_SHAPE, "dp_bn_act_apply: NULL pointer");
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024 && rows > 0, DP_ERR_ALIGN, "dp_bn_act_apply: bad Cp=%d / rows", Cp);
  const int64_t nvec = rows * (Cp / 8);
  const int grid = ew_grid(nvec, Cp / 8);
  const size_t sm = 0;
  if (dtype == DP_BF16)
    launch_pdl(bn_act_apply_kernel<__nv_bfloat16>, dim3(grid), dim3(256), sm, as_stream(stream),
               (const __nv_bfloat16*)y, scale, shift, slope, (const __nv_bfloat16*)residual, slope_res, (__nv_bfloat16*)z,
               nvec, Cp);
  else
    bn_act_apply_kernel<float><<<grid, 256, sm, as_stream(stream)>>>((const float*)y, scale, shift, slope,
                                                                     (const float*)residual, slope_res, (float*)z,
                                                                     nvec, Cp);
  return check_launch("dp_bn_act_apply");
}

DP_API int dp_bn_act_bwd_reduce(const void* dz, const void* y, const void* out, const float* scale,
                                const float* shift, const float* mean, const float* rstd, float slope,
                                float slope_res, float* part, int* nparts, int64_t rows, int Cp, int dtype,
                                void* stream) {
  DP_REQUIRE(dz && y && scale && shift && mean && rstd && part && nparts, DP_ERR_SHAPE,
             "dp_bn_act_bwd_reduce: NULL pointer");
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024 && rows > 0, DP_ERR_ALIGN, "dp_bn_act_bwd_reduce: bad Cp=%d", Cp);
  const int grid = reduce_grid(rows);
  cudaStream_t s = as_stream(stream);
  if (dtype == DP_BF16) {
    BwdReduceF<__nv_bfloat16> f{(const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, (const __nv_bfloat16*)out,
                                scale, shift, mean, rstd, slope, slope_res};
    launch_pdl(col_reduce2_kernel<BwdReduceF<__nv_bfloat16>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part);
  } else {
    BwdReduceF<float> f{(const float*)dz, (const float*)y, (const float*)out, scale, shift, mean, rstd, slope,
                        slope_res};
    launch_pdl(col_reduce2_kernel<BwdReduceF<float>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part);
  }
  *nparts = grid;
  return check_launch("dp_bn_act_bwd_reduce");
}

DP_API int dp_bn_bwd_finalize(const float* part, int nparts, int C, int Cp, double count, const float* mean,
                              const float* rstd, float* dgamma, float* dbeta, float* coef, void* stream) {
  DP_REQUIRE(part && coef && mean && rstd, DP_ERR_SHAPE, "dp_bn_bwd_finalize: NULL pointer");
  DP_REQUIRE(nparts > 0 && nparts <= DP_MAX_PARTS && C > 0 && Cp >= C && count > 0, DP_ERR_SHAPE,
             "dp_bn_bwd_finalize: bad sizes");
  launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, C, Cp,
             count, mean, rstd, dgamma, dbeta, coef);
  return check_launch("dp_bn_bwd_finalize");
}

DP_API int dp_bn_act_bwd_apply(const void* dz, const void* y, const void* out, const float* scale,
                               const float* shift, const float* mean, const float* rstd, const float* coef,
                               float slope, float slope_res, void* dy, void* dres, int64_t rows, int Cp, int dtype,
                               void* stream) {
  DP_REQUIRE(dz && y && scale && shift && mean && rstd && coef && dy, DP_ERR_SHAPE,
             "dp_bn_act_bwd_apply: NULL pointer");
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024 && rows > 0, DP_ERR_ALIGN, "dp_bn_act_bwd_apply: bad Cp=%d", Cp);
  DP_REQUIRE(dres == nullptr || out != nullptr, DP_ERR_SHAPE, "dp_bn_act_bwd_apply: dres needs out");
  const int64_t nvec = rows * (Cp / 8);
  const int grid = ew_grid(nvec, Cp / 8);
  const size_t sm = 0;
  if (dtype == DP_BF16)
    launch_pdl(bn_act_bwd_apply_kernel<__nv_bfloat16>, dim3(grid), dim3(256), sm, as_stream(stream),
               (const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, (const __nv_bfloat16*)out, scale, shift, mean, rstd, coef,
               slope, slope_res, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, nvec, Cp);
  else
    bn_act_bwd_apply_kernel<float><<<grid, 256, sm, as_stream(stream)>>>(
        (const float*)dz, (const float*)y, (const float*)out, scale, shift, mean, rstd, coef, slope, slope_res,
        (float*)dy, (float*)dres, nvec, Cp);
  return check_launch("dp_bn_act_bwd_apply");
}

DP_API int dp_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
  DP_REQUIRE(a && b && out, DP_ERR_SHAPE, "dp_add: NULL pointer");
  DP_REQUIRE(n > 0 && n % 8 == 0, DP_ERR_ALIGN, "dp_add: n=%lld must be a positive multiple of 8", (long long)n);
  const int64_t nvec = n / 8;
  const int grid = ew_grid(nvec, 1);
  if (dtype == DP_BF16)
    add_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                    (__nv_bfloat16*)out, nvec);
  else
    add_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)a, (const float*)b, (float*)out, nvec);
  return check_launch("dp_add");
}
