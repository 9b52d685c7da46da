// Memory-bound BatchNorm3d(train) + LeakyReLU kernels over NDHWC activations:
// statistics, finalize (+running stats), apply (+residual), and the two-pass backward.
// 16-byte vector loads along the padded channel dimension, per-thread channel-resident
// accumulators, shared-memory tree, per-CTA partials reduced deterministically in fp64.
//
// Replaces nn.BatchNorm3d / nn.LeakyReLU at /root/reference/src/models/R2Plus1D.py:53-57 and the
// residual add + LeakyReLU at :179-187 (and their autograd).
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "bn_fin.cuh"

namespace dp {

constexpr int RED_THREADS = 256;

// ---- generic per-channel column reduction of two quantities over [rows][Cp] ----
// F: __device__ void operator()(int64_t elem_offset, int c0, float* a8, float* b8) accumulates 8 channels
template <typename F>
__global__ void __launch_bounds__(RED_THREADS, 2)
col_reduce2_kernel(F f, int64_t rows, int Cp, float* __restrict__ part, const dp_bn_fin fin) {  // f by value: per-thread register copy
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
  if (fin.kind) {   // last CTA done: finalise in this launch (bn_fin.cuh); `red` is free again after the barrier
    __threadfence();
    __syncthreads();
    bn_fin_tail(fin, part, reinterpret_cast<int*>(&red[1][0]), reinterpret_cast<double*>(&red[0][0]));
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

static const dp_bn_fin kNoFin = {};   // kind 0: the kernel writes its partials and stops

int bn_stats_launch(const void* y, int64_t rows, int Cp, int dtype, float* part, int* nparts, cudaStream_t s,
                    const dp_bn_fin* fin) {
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024, DP_ERR_ALIGN, "bn_stats: Cp=%d must be a multiple of 8, <= 1024", Cp);
  DP_REQUIRE(rows > 0, DP_ERR_SHAPE, "bn_stats: no rows");
  const int grid = reduce_grid(rows);
  const dp_bn_fin f_ = fin ? *fin : kNoFin;
  if (dtype == DP_BF16) {
    StatsF<__nv_bfloat16> f{(const __nv_bfloat16*)y};
    launch_pdl(col_reduce2_kernel<StatsF<__nv_bfloat16>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_);
  } else {
    StatsF<float> f{(const float*)y};
    launch_pdl(col_reduce2_kernel<StatsF<float>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_);
  }
  if (nparts != nullptr) *nparts = grid;
  return check_launch("bn_stats");
}

// ---- finalize: partials -> mean/rstd/scale/shift, running stats (momentum, unbiased var) ----
// One CTA of 256 threads per 32 channels (bn_fin.cuh: 8 part-lanes x 32 channel-lanes, fp64 partial sums combined in a
// fixed order, reads coalesced along the channel dimension).  The producing kernels run the same code in their last
// CTA when they are given a dp_bn_fin; these stand-alone launches serve partials that come without one.
__global__ void __launch_bounds__(FIN_CH * FIN_PL)
bn_finalize_group_kernel(const float* __restrict__ part, int nparts, const dp_bn_fin fin) {
  __shared__ double red[2 * FIN_PL * FIN_CH];
  pdl_launch_dependents();
  pdl_wait();
  bn_fin_group(fin, part, nparts, blockIdx.x * FIN_CH, threadIdx.x, red);
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
  return bn_stats_launch(y, rows, Cp, dtype, part, nparts, as_stream(stream), nullptr);
}

DP_API int dp_bn_finalize(const float* part, int nparts, int C, int Cp, double count, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* mean, float* rstd, float* scale, float* shift, void* stream) {
  DP_REQUIRE(part && gamma && beta && mean && rstd && scale && shift, DP_ERR_SHAPE, "dp_bn_finalize: NULL pointer");
  DP_REQUIRE(nparts > 0 && nparts <= DP_MAX_PARTS && C > 0 && Cp >= C && count > 0, DP_ERR_SHAPE,
             "dp_bn_finalize: bad sizes (nparts=%d C=%d Cp=%d)", nparts, C, Cp);
  DP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), DP_ERR_SHAPE,
             "dp_bn_finalize: running_mean/var must both be given or both NULL");
  dp_bn_fin f = {};
  f.kind = 1; f.C = C; f.Cp = Cp; f.count = count; f.gamma = gamma; f.beta = beta; f.eps = eps; f.momentum = momentum;
  f.running_mean = running_mean; f.running_var = running_var; f.mean = mean; f.rstd = rstd; f.scale = scale; f.shift = shift;
  launch_pdl(bn_finalize_group_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, f);
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

static int bwd_reduce_launch(const void* dz, const void* y, const void* out, const float* scale, const float* shift,
                             const float* mean, const float* rstd, float slope, float slope_res, float* part, int* nparts,
                             int64_t rows, int Cp, int dtype, void* stream, const dp_bn_fin* fin, const char* who) {
  DP_REQUIRE(dz && y && scale && shift && mean && rstd && part, DP_ERR_SHAPE, "%s: NULL pointer", who);
  DP_REQUIRE(Cp % 8 == 0 && Cp > 0 && Cp <= 1024 && rows > 0, DP_ERR_ALIGN, "%s: bad Cp=%d", who, Cp);
  const int grid = reduce_grid(rows);
  cudaStream_t s = as_stream(stream);
  const dp_bn_fin f_ = fin ? *fin : kNoFin;
  if (dtype == DP_BF16) {
    BwdReduceF<__nv_bfloat16> f{(const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, (const __nv_bfloat16*)out,
                                scale, shift, mean, rstd, slope, slope_res};
    launch_pdl(col_reduce2_kernel<BwdReduceF<__nv_bfloat16>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_);
  } else {
    BwdReduceF<float> f{(const float*)dz, (const float*)y, (const float*)out, scale, shift, mean, rstd, slope,
                        slope_res};
    launch_pdl(col_reduce2_kernel<BwdReduceF<float>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_);
  }
  if (nparts != nullptr) *nparts = grid;
  return check_launch(who);
}

DP_API int dp_bn_act_bwd_reduce(const void* dz, const void* y, const void* out, const float* scale,
                                const float* shift, const float* mean, const float* rstd, float slope,
                                float slope_res, float* part, int* nparts, int64_t rows, int Cp, int dtype,
                                void* stream) {
  DP_REQUIRE(nparts != nullptr, DP_ERR_SHAPE, "dp_bn_act_bwd_reduce: NULL pointer");
  return bwd_reduce_launch(dz, y, out, scale, shift, mean, rstd, slope, slope_res, part, nparts, rows, Cp, dtype, stream,
                           nullptr, "dp_bn_act_bwd_reduce");
}

DP_API int dp_bn_act_bwd_reduce_fin(const void* dz, const void* y, const void* out, const float* scale,
                                    const float* shift, float slope, float slope_res, float* part, int64_t rows, int Cp,
                                    int dtype, const dp_bn_fin* fin, void* stream) {
  const int rc = bn_fin_validate(fin, 2, Cp, "dp_bn_act_bwd_reduce_fin");
  if (rc != DP_OK) return rc;
  return bwd_reduce_launch(dz, y, out, scale, shift, fin->mean, fin->rstd, slope, slope_res, part, nullptr, rows, Cp, dtype,
                           stream, fin, "dp_bn_act_bwd_reduce_fin");
}

DP_API int dp_bn_bwd_finalize(const float* part, int nparts, int C, int Cp, double count, const float* mean,
                              const float* rstd, float* dgamma, float* dbeta, float* coef, void* stream) {
  DP_REQUIRE(part && coef && mean && rstd, DP_ERR_SHAPE, "dp_bn_bwd_finalize: NULL pointer");
  DP_REQUIRE(nparts > 0 && nparts <= DP_MAX_PARTS && C > 0 && Cp >= C && count > 0, DP_ERR_SHAPE,
             "dp_bn_bwd_finalize: bad sizes");
  dp_bn_fin f = {};
  f.kind = 2; f.C = C; f.Cp = Cp; f.count = count; f.mean = const_cast<float*>(mean); f.rstd = const_cast<float*>(rstd);
  f.dgamma = dgamma; f.dbeta = dbeta; f.coef = coef;
  launch_pdl(bn_finalize_group_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, f);
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
