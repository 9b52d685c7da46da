// BatchNorm finalisation (per-CTA partial sums -> per-channel coefficients), shared by the stand-alone finalize kernels
// (bn_act.cu) and by the optional "last CTA done" tails of the kernels that PRODUCE the partials (tc_gather_gemm_kernel,
// col_reduce2_kernel; entry points dp_*_fin): the producing kernel's last CTA to retire runs the same code over all
// rows of `part`, which removes 64 launches per training step.  Measured on B200 (profiles/README.md, r2h): SLOWER than
// the stand-alone launches (18.98 vs 18.35 ms/step) -- one CTA walks 148-592 partial rows of every channel group while
// the other 147 SMs idle (+15 us per conv, +26 us per reduction), whereas a stand-alone finalize spreads the groups over
// CTAs and costs ~5 us including its launch gap inside the CUDA graph.  The host side keeps it off (DP_FUSE_FIN=1).
//
// kind 1 (forward, nn.BatchNorm3d in train mode at /root/reference/src/models/R2Plus1D.py:53-54):
//   part = (sum y, sum y^2) -> mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, running statistics
// kind 2 (its autograd): part = (sum g', sum g'*y) -> dbeta, dgamma, coef = (sum g'/count, sum g'*xhat/count)
#pragma once
#include "dp_common.cuh"

namespace dp {

// one group = 16 channels x 16 part-lanes = 256 threads: the finalize kernels are pure load latency on the critical path
// between every conv and its apply pass, so the partial rows are spread over as many threads as the CTA has
constexpr int FIN_CH = 16, FIN_PL = 16;

// Threads tid < 256 of a CTA finalise channels [c0, c0 + FIN_CH); every thread of the CTA must call (two CTA barriers).
// red: >= 2 * FIN_PL * FIN_CH doubles of shared memory.  fp64 partial sums combined in a fixed order (deterministic).
__device__ __forceinline__ void bn_fin_group(const dp_bn_fin& f, const float* part, int nparts, int c0, int tid, double* red) {
  const int cl = tid % FIN_CH, pl = tid / FIN_CH, c = c0 + cl, Cp = f.Cp;
  if (tid < FIN_CH * FIN_PL) {
    double s = 0.0, q = 0.0;
    if (c < f.C) {
      // four independent load/accumulate chains per thread: the loop is pure memory latency (<= 592 partials); the
      // partials come from other SMs, so they are read through L2 (ld.global.cg)
      double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
      int p = pl;
      for (; p + 3 * FIN_PL < nparts; p += 4 * FIN_PL) {
        const float a0 = __ldcg(part + ((int64_t)p * 2 + 0) * Cp + c), b0 = __ldcg(part + ((int64_t)p * 2 + 1) * Cp + c);
        const float a1 = __ldcg(part + ((int64_t)(p + FIN_PL) * 2 + 0) * Cp + c), b1 = __ldcg(part + ((int64_t)(p + FIN_PL) * 2 + 1) * Cp + c);
        const float a2 = __ldcg(part + ((int64_t)(p + 2 * FIN_PL) * 2 + 0) * Cp + c), b2 = __ldcg(part + ((int64_t)(p + 2 * FIN_PL) * 2 + 1) * Cp + c);
        const float a3 = __ldcg(part + ((int64_t)(p + 3 * FIN_PL) * 2 + 0) * Cp + c), b3 = __ldcg(part + ((int64_t)(p + 3 * FIN_PL) * 2 + 1) * Cp + c);
        s += (double)a0; q += (double)b0; s1 += (double)a1; q1 += (double)b1;
        s2 += (double)a2; q2 += (double)b2; s3 += (double)a3; q3 += (double)b3;
      }
      for (; p < nparts; p += FIN_PL) {
        s += (double)__ldcg(part + ((int64_t)p * 2 + 0) * Cp + c);
        q += (double)__ldcg(part + ((int64_t)p * 2 + 1) * Cp + c);
      }
      s = (s + s1) + (s2 + s3);
      q = (q + q1) + (q2 + q3);
    }
    red[(0 * FIN_PL + pl) * FIN_CH + cl] = s;
    red[(1 * FIN_PL + pl) * FIN_CH + cl] = q;
  }
  __syncthreads();
  if (tid < FIN_CH && c < Cp) {
    double S = 0.0, Q = 0.0;
#pragma unroll
    for (int i = 0; i < FIN_PL; ++i) { S += red[(0 * FIN_PL + i) * FIN_CH + cl]; Q += red[(1 * FIN_PL + i) * FIN_CH + cl]; }
    if (f.kind == 1) {
      if (c >= f.C) {
        f.mean[c] = 0.f; f.rstd[c] = 0.f; f.scale[c] = 0.f; f.shift[c] = 0.f;
      } else {
        const double mu = S / f.count;
        double var = Q / f.count - mu * mu;
        if (var < 0.0) var = 0.0;
        const float rs = (float)(1.0 / sqrt(var + (double)f.eps));
        const float sc = f.gamma[c] * rs;
        f.mean[c] = (float)mu;
        f.rstd[c] = rs;
        f.scale[c] = sc;
        f.shift[c] = __fmaf_rn(-(float)mu, sc, f.beta[c]);
        if (f.running_mean != nullptr) {
          const double unbiased = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
          // explicit rounding: the same bits whether this code is inlined in a finalize kernel or in a producer's tail
          f.running_mean[c] = __fmaf_rn(f.momentum, (float)mu, __fmul_rn(1.f - f.momentum, f.running_mean[c]));
          f.running_var[c] = __fmaf_rn(f.momentum, (float)unbiased, __fmul_rn(1.f - f.momentum, f.running_var[c]));
        }
      }
    } else {
      if (c >= f.C) {
        f.coef[c] = 0.f; f.coef[Cp + c] = 0.f;
      } else {
        Q = (Q - (double)f.mean[c] * S) * (double)f.rstd[c];   // sum(g*y) -> sum(g*xhat)
        if (f.dbeta != nullptr) f.dbeta[c] = (float)S;
        if (f.dgamma != nullptr) f.dgamma[c] = (float)Q;
        // eval-mode BatchNorm (coef_zero): the statistics are constants, no mean / variance terms in dy
        f.coef[c] = f.coef_zero ? 0.f : (float)(S / f.count);
        f.coef[Cp + c] = f.coef_zero ? 0.f : (float)(Q / f.count);
      }
    }
  }
  __syncthreads();
}

// Tail of a kernel whose CTAs each wrote one row of `part` (row = blockIdx.x, gridDim.x rows): every writer has executed
// __threadfence() after its stores and the CTA has met at a barrier.  The CTA that takes the last ticket finalises all
// channels and re-arms the ticket (the word is zero again when the kernel ends).  All threads of the CTA must call.
// flag: one int of shared memory; red as above (both may alias pipeline buffers that are dead by now).
__device__ __forceinline__ void bn_fin_tail(const dp_bn_fin& f, const float* part, int* flag, double* red) {
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(f.ticket, 1u);
    *flag = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*flag == 0) return;
  __threadfence();   // the other CTAs' partial rows (published before their tickets) are visible from here on
  for (int c0 = 0; c0 < f.Cp; c0 += FIN_CH) bn_fin_group(f, part, (int)gridDim.x, c0, (int)threadIdx.x, red);
  if (threadIdx.x == 0) *f.ticket = 0u;
}

// host side: argument check shared by the entry points that take a dp_bn_fin
inline int bn_fin_validate(const dp_bn_fin* f, int kind, int Cp, const char* who) {
  DP_REQUIRE(f != nullptr && f->kind == kind, DP_ERR_SHAPE, "%s: dp_bn_fin of kind %d expected", who, kind);
  DP_REQUIRE(f->C > 0 && f->Cp == Cp && f->Cp >= f->C && f->count > 0 && f->ticket != nullptr, DP_ERR_SHAPE,
             "%s: dp_bn_fin sizes (C=%d Cp=%d, expected Cp=%d) / ticket", who, f->C, f->Cp, Cp);
  if (kind == 1)
    DP_REQUIRE(f->gamma && f->beta && f->mean && f->rstd && f->scale && f->shift &&
                   ((f->running_mean == nullptr) == (f->running_var == nullptr)),
               DP_ERR_SHAPE, "%s: dp_bn_fin (forward) NULL pointer", who);
  else
    DP_REQUIRE(f->mean && f->rstd && f->coef, DP_ERR_SHAPE, "%s: dp_bn_fin (backward) NULL pointer", who);
  return DP_OK;
}

}  // namespace dp
