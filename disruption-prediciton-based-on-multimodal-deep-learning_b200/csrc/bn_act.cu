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

// Traversal direction of the passes (option "bn_sweep", bits) -- an experiment kept as an option, OFF by default.
// Consecutive kernels of a step stream tensors of 0.1-0.9 GB through a 126 MB L2; a consumer that starts where its producer
// STOPPED could find the producer's last ~50-100 MB still cached.  The conv kernels sweep their tiles front to back, so
//   bit 0: bn_act_apply runs back to front (reads the end of y the conv just wrote, leaves the START of z for the next conv);
//   bit 1: the backward reduction sweeps back to front in interleaved row groups (the data gradient that wrote dz stopped
//          at its end) instead of one contiguous row range per CTA;
//   bit 2: bn_act_bwd_apply runs front to back right after a stand-alone reduction over the same dz (which stopped at the
//          front), back to front otherwise (dz comes straight from a data gradient whose epilogue produced the sums).
// Measured on the whole step (profiles/r2/r2z_option_ab.txt, in-process A/B): 16.67-16.74 ms with 0, 16.76-16.84 ms with 7,
// 16.67 with 1 -- nothing carries over in L2 between these kernels that the passes could use; streaming (ld.global.cs)
// loads of the dead operands ("bn_cs") and evict-first activation loads in the conv kernels ("tc_l2hint") change nothing
// either (16.8-17.0 ms).
int g_bn_sweep = 0;
int g_bn_cs = 0;   // option "bn_cs": bit 0 / bit 1: apply / backward-apply read their dead inputs with ld.global.cs
static const void* g_last_reduce_dz = nullptr;   // dz of the most recent stand-alone backward reduction (bit 2)

// 8 consecutive elements exactly as they sit in memory (4 registers for bf16): the reductions below issue the loads of
// several rows BEFORE converting any of them, so that many rows are in flight per thread; a load that returns eight
// converted floats (ld8) makes the compiler serialise rows for lack of registers (measured: 2-4 rows in flight, 81-85 %
// of the copy rate for a read-only kernel)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ Raw8<__nv_bfloat16> ld_raw8(const __nv_bfloat16* p) { Raw8<__nv_bfloat16> r; r.u = *reinterpret_cast<const uint4*>(p); return r; }
__device__ __forceinline__ Raw8<float> ld_raw8(const float* p) {
  Raw8<float> r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
}
// CS: evict-first ("streaming") loads for operands nobody reads again soon, so that what the pass WRITES outlives them in L2
template <bool CS> __device__ __forceinline__ Raw8<__nv_bfloat16> ld_raw8s(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.u = CS ? __ldcs(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
  return r;
}
template <bool CS> __device__ __forceinline__ Raw8<float> ld_raw8s(const float* p) {
  Raw8<float> r;
  if (CS) { r.a = __ldcs(reinterpret_cast<const float4*>(p)); r.b = __ldcs(reinterpret_cast<const float4*>(p + 4)); }
  else { r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); }
  return r;
}
__device__ __forceinline__ f8 cvt8(const Raw8<__nv_bfloat16>& x) {
  f8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}
__device__ __forceinline__ f8 cvt8(const Raw8<float>& x) {
  f8 r;
  r.v[0] = x.a.x; r.v[1] = x.a.y; r.v[2] = x.a.z; r.v[3] = x.a.w; r.v[4] = x.b.x; r.v[5] = x.b.y; r.v[6] = x.b.z; r.v[7] = x.b.w;
  return r;
}

// ---- generic per-channel column reduction of two quantities over [rows][Cp] ----
// F: Raw load(int64_t elem_offset) fetches one 8-channel vector of every input stream; acc(raw, a8, b8) accumulates it
template <typename F>
__global__ void __launch_bounds__(RED_THREADS, 2)
col_reduce2_kernel(F f, int64_t rows, int Cp, float* __restrict__ part, const dp_bn_fin fin, int sweep) {  // f by value: per-thread register copy
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][RED_THREADS * 8];
  const int vpr = Cp >> 3;                       // 8-channel vectors per row
  const int rpi = RED_THREADS / vpr;             // rows per iteration
  const int active = rpi * vpr;
  const int tid = threadIdx.x;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
  if (tid < active) {
    const int cv = tid % vpr, rl = tid / vpr;
    f.init(cv * 8);                       // this thread's 8 channels never change: parameters live in registers
    constexpr int U = F::kRowsInFlight;       // rows in flight per thread (two CTAs of 256 threads per SM)
    auto range = [&](int64_t r0, int64_t r1) {
      int64_t r = r0 + rl;
      for (; r + (U - 1) * rpi < r1; r += U * rpi) {
        typename F::Raw raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = f.load((r + u * rpi) * Cp + cv * 8);
#pragma unroll
        for (int u = 0; u < U; ++u) f.acc(raw[u], a, b);   // same order as row-by-row: the sums do not change
      }
      for (; r < r1; r += rpi) f.acc(f.load(r * Cp + cv * 8), a, b);
    };
    if (sweep == 0) {   // one contiguous row range per CTA
      const int64_t rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
      const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
      const int64_t r1 = r0 + rows_per_cta;
      range(r0, r1 > rows ? rows : r1);
    } else {            // row groups interleaved over the CTAs: the grid sweeps the tensor front to back (1) or back to front (2)
      const int64_t rg = (int64_t)U * rpi;
      const int64_t ngroups = (rows + rg - 1) / rg;
      for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int64_t gg = sweep == 2 ? ngroups - 1 - g : g;
        const int64_t r1 = (gg + 1) * rg;
        range(gg * rg, r1 > rows ? rows : r1);
      }
    }
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

// one wave of two CTAs per SM: with 8 rows x 2 streams in flight per thread that saturates HBM, and the finalize kernel
// behind it reads half the partial rows of the former four CTAs per SM
static int reduce_grid(int64_t rows) {
  int64_t g = (rows + 63) / 64;
  const int64_t cap = 2 * (int64_t)num_sms() < DP_MAX_PARTS ? 2 * (int64_t)num_sms() : DP_MAX_PARTS;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
struct StatsF {
  const T* y;
  struct Raw { Raw8<T> v; };
  static constexpr int kRowsInFlight = sizeof(T) == 2 ? 16 : 8;
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ Raw load(int64_t off) const { Raw r; r.v = ld_raw8(y + off); return r; }
  __device__ __forceinline__ void acc(const Raw& r, float* a, float* b) const {
    const f8 v = cvt8(r.v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] += v.v[j]; b[j] = fmaf(v.v[j], v.v[j], b[j]); }
  }
};

template <typename T, bool HAS_OUT>   // HAS_OUT: a third stream (the block output behind a residual add + activation)
struct BwdReduceF {
  const T* dz; const T* y; const T* out;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  float slope, slope_res;
  float r_scale[8], r_shift[8];
  struct Raw { Raw8<T> g, v, o; };   // (o stays dead without the third stream)
  static constexpr int kRowsInFlight = sizeof(T) == 2 ? (HAS_OUT ? 5 : 8) : (HAS_OUT ? 2 : 4);
  // accumulates sum(g) and sum(g*y) with the RAW conv output y; bn_bwd_finalize turns the second into
  // sum(g*xhat) = (sum(g*y) - mean*sum(g)) * rstd in fp64, so the streaming loop needs two parameters per channel
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { r_scale[j] = scale[c0 + j]; r_shift[j] = shift[c0 + j]; }
  }
  __device__ __forceinline__ Raw load(int64_t off) const {
    Raw r;
    r.g = ld_raw8(dz + off); r.v = ld_raw8(y + off);
    if (HAS_OUT) r.o = ld_raw8(out + off);
    return r;
  }
  __device__ __forceinline__ void acc(const Raw& r, float* a, float* b) const {
    const f8 g = cvt8(r.g), v = cvt8(r.v);
    f8 o;
    if (HAS_OUT) o = cvt8(r.o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gg = g.v[j];
      if (HAS_OUT) gg *= (o.v[j] > 0.f ? 1.f : slope_res);
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
  const int sweep = 0;   // (stand-alone forward statistics: fp32 validation mode and tiles the conv epilogues do not cover)
  if (dtype == DP_BF16) {
    StatsF<__nv_bfloat16> f{(const __nv_bfloat16*)y};
    launch_pdl(col_reduce2_kernel<StatsF<__nv_bfloat16>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
  } else {
    StatsF<float> f{(const float*)y};
    launch_pdl(col_reduce2_kernel<StatsF<float>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
  }
  if (nparts != nullptr) *nparts = grid;
  return check_launch("bn_stats");
}

// ---- finalize: partials -> mean/rstd/scale/shift, running stats (momentum, unbiased var) ----
// One CTA of 256 threads per 16 channels (bn_fin.cuh: 16 part-lanes x 16 channel-lanes, fp64 partial sums combined in a
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
// V vectors in flight per thread: the raw 16-byte loads of all V are issued before any is converted (see Raw8)
template <typename T, bool HAS_RES, int V, bool CS, bool REV>
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
  // REV: the grid sweeps the tensor back to front; nvec and the stride are multiples of vpr, so at(v) stays in one column
  const int64_t last = nvec - 1;
#define at(v) (REV ? last - (v) : (v))
  const int c0 = (int)(at(v0) % vpr) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
  auto one = [&](const Raw8<T>& ya, const Raw8<T>& ra, int64_t v) {
    f8 a = cvt8(ya), r;
    if (HAS_RES) r = cvt8(ra);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float u = lrelu(fmaf(a.v[j], sc[j], sh[j]), slope);
      if (HAS_RES) u = lrelu(u + r.v[j], slope_res);
      a.v[j] = u;
    }
    st8(z + v * 8, a);
  };
  int64_t v = v0;
  for (; v + (V - 1) * stride < nvec; v += V * stride) {
    Raw8<T> ya[V], ra[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
      ya[u] = ld_raw8s<CS>(y + at(v + u * stride) * 8);
      if (HAS_RES) ra[u] = ld_raw8s<CS>(residual + at(v + u * stride) * 8);
    }
#pragma unroll
    for (int u = 0; u < V; ++u) one(ya[u], ra[u], at(v + u * stride));
  }
  for (; v < nvec; v += stride) {
    Raw8<T> ya = ld_raw8s<CS>(y + at(v) * 8), ra;
    if (HAS_RES) ra = ld_raw8s<CS>(residual + at(v) * 8);
    one(ya, ra, at(v));
  }
#undef at
}

template <typename T, bool HAS_OUT, int V, bool CS, bool REV>
__global__ void __launch_bounds__(256, 4)   // 64 registers: four CTAs per SM
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
  const int64_t last = nvec - 1;
#define at(v) (REV ? last - (v) : (v))   // (see bn_act_apply_kernel)
  const int c0 = (int)(at(v0) % vpr) * 8;
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
  auto one = [&](const Raw8<T>& gr, const Raw8<T>& yr, const Raw8<T>& orr, int64_t v) {
    const f8 g = cvt8(gr), yy = cvt8(yr);
    f8 o, res, d;
    if (HAS_OUT) o = cvt8(orr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gg = g.v[j];
      if (HAS_OUT) gg *= (o.v[j] > 0.f ? 1.f : slope_res);
      res.v[j] = gg;
      const float u = fmaf(yy.v[j], sc[j], sh[j]);
      gg *= (u > 0.f ? 1.f : slope);
      d.v[j] = fmaf(sc[j], gg, fmaf(kb[j], yy.v[j], ka[j]));
    }
    st8(dy + v * 8, d);
    if (HAS_OUT && dres != nullptr) st8(dres + v * 8, res);
  };
  int64_t v = v0;
  for (; v + (V - 1) * stride < nvec; v += V * stride) {
    Raw8<T> gr[V], yr[V], orr[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
      gr[u] = ld_raw8s<CS>(dz + at(v + u * stride) * 8);
      yr[u] = ld_raw8s<CS>(y + at(v + u * stride) * 8);
      if (HAS_OUT) orr[u] = ld_raw8s<CS>(out + at(v + u * stride) * 8);
    }
#pragma unroll
    for (int u = 0; u < V; ++u) one(gr[u], yr[u], orr[u], at(v + u * stride));
  }
  for (; v < nvec; v += stride) {
    Raw8<T> gr = ld_raw8s<CS>(dz + at(v) * 8), yr = ld_raw8s<CS>(y + at(v) * 8), orr;
    if (HAS_OUT) orr = ld_raw8s<CS>(out + at(v) * 8);
    one(gr, yr, orr, at(v));
  }
#undef at
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

// grid of a grid-stride elementwise pass: `per_sm` CTAs of 256 threads per SM = exactly the CTAs that are resident at once
// (one wave: a grid of 8 per SM over a kernel that fits 6 ran a second wave at a third of the occupancy)
static int ew_grid(int64_t nvec, int vpr, int per_sm = 8) {
  int64_t g = (nvec + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (g > cap) g = cap;
  const int64_t gmin = (vpr + 255) / 256;   // the loop stride must hold at least one full row of vectors
  if (g < gmin) g = gmin;
  if (g < 1) g = 1;
  return (int)g;
}

// resident CTAs per SM of a 256-thread kernel without dynamic shared memory (cached per kernel and device)
template <typename K>
static int resident_ctas(K kernel) {
  static int cached[DP_MAX_DEVICES] = {};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) n = 4;
    cached[dev] = n;
  }
  return cached[dev];
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
  launch_pdl_small(bn_finalize_group_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, f);
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
  const size_t sm = 0;
  cudaStream_t st_ = as_stream(stream);
#define DP_APPLY_LAUNCH_(T, RES, V, CS, REV)                                                                         \
  launch_pdl(bn_act_apply_kernel<T, RES, V, CS, REV>,                                                                 \
             dim3(ew_grid(nvec, Cp / 8, resident_ctas(bn_act_apply_kernel<T, RES, V, CS, REV>))), dim3(256), sm, st_, \
             (const T*)y, scale, shift, slope, (const T*)residual, slope_res, (T*)z, nvec, Cp)
#define DP_APPLY_LAUNCH(T, RES, V) do {                                                                               \
    if (cs) { if (rev) DP_APPLY_LAUNCH_(T, RES, V, true, true); else DP_APPLY_LAUNCH_(T, RES, V, true, false); }      \
    else { if (rev) DP_APPLY_LAUNCH_(T, RES, V, false, true); else DP_APPLY_LAUNCH_(T, RES, V, false, false); } } while (0)
  const int rev = g_bn_sweep & 1;
  const bool cs = (g_bn_cs & 1) != 0;
  if (dtype == DP_BF16) {
    if (residual != nullptr) DP_APPLY_LAUNCH(__nv_bfloat16, true, 2); else DP_APPLY_LAUNCH(__nv_bfloat16, false, 4);
  } else {
    if (residual != nullptr) DP_APPLY_LAUNCH(float, true, 1); else DP_APPLY_LAUNCH(float, false, 2);
  }
#undef DP_APPLY_LAUNCH_
#undef DP_APPLY_LAUNCH
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
  const int sweep = (g_bn_sweep & 2) ? 2 : 0;
  g_last_reduce_dz = dz;
  if (dtype == DP_BF16) {
    if (out != nullptr) {
      BwdReduceF<__nv_bfloat16, true> f{(const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, (const __nv_bfloat16*)out,
                                        scale, shift, mean, rstd, slope, slope_res};
      launch_pdl(col_reduce2_kernel<BwdReduceF<__nv_bfloat16, true>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
    } else {
      BwdReduceF<__nv_bfloat16, false> f{(const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, nullptr,
                                         scale, shift, mean, rstd, slope, slope_res};
      launch_pdl(col_reduce2_kernel<BwdReduceF<__nv_bfloat16, false>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
    }
  } else {
    if (out != nullptr) {
      BwdReduceF<float, true> f{(const float*)dz, (const float*)y, (const float*)out, scale, shift, mean, rstd, slope, slope_res};
      launch_pdl(col_reduce2_kernel<BwdReduceF<float, true>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
    } else {
      BwdReduceF<float, false> f{(const float*)dz, (const float*)y, nullptr, scale, shift, mean, rstd, slope, slope_res};
      launch_pdl(col_reduce2_kernel<BwdReduceF<float, false>>, dim3(grid), dim3(RED_THREADS), 0, s, f, rows, Cp, part, f_, sweep);
    }
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
  launch_pdl_small(bn_finalize_group_kernel, dim3(ceil_div(Cp, FIN_CH)), dim3(FIN_CH * FIN_PL), 0, as_stream(stream), part, nparts, f);
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
  const size_t sm = 0;
  cudaStream_t st_ = as_stream(stream);
#define DP_BWD_APPLY_LAUNCH_(T, OUT, V, CS, REV)                                                                     \
  launch_pdl(bn_act_bwd_apply_kernel<T, OUT, V, CS, REV>,                                                             \
             dim3(ew_grid(nvec, Cp / 8, resident_ctas(bn_act_bwd_apply_kernel<T, OUT, V, CS, REV>))), dim3(256), sm,  \
             st_, (const T*)dz, (const T*)y, (const T*)out, scale, shift, mean, rstd, coef, slope, slope_res, (T*)dy, \
             (T*)dres, nvec, Cp)
#define DP_BWD_APPLY_LAUNCH(T, OUT, V) do {                                                                           \
    if (cs) { if (rev) DP_BWD_APPLY_LAUNCH_(T, OUT, V, true, true); else DP_BWD_APPLY_LAUNCH_(T, OUT, V, true, false); } \
    else { if (rev) DP_BWD_APPLY_LAUNCH_(T, OUT, V, false, true); else DP_BWD_APPLY_LAUNCH_(T, OUT, V, false, false); } } while (0)
  // after a stand-alone back-to-front reduction over this dz the cache holds the FRONT of dz and y; a dz that comes straight
  // from a data gradient (sums in its epilogue) is cached at its END
  const bool after_reduce = (g_bn_sweep & 2) && g_last_reduce_dz == dz;
  const int rev = (g_bn_sweep & 4) ? (after_reduce ? 0 : 1) : 0;
  g_last_reduce_dz = nullptr;
  const bool cs = (g_bn_cs & 2) != 0;
  if (dtype == DP_BF16) {
    if (out != nullptr) DP_BWD_APPLY_LAUNCH(__nv_bfloat16, true, 1); else DP_BWD_APPLY_LAUNCH(__nv_bfloat16, false, 2);
  } else {
    if (out != nullptr) DP_BWD_APPLY_LAUNCH(float, true, 1); else DP_BWD_APPLY_LAUNCH(float, false, 1);
  }
#undef DP_BWD_APPLY_LAUNCH_
#undef DP_BWD_APPLY_LAUNCH
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
