// CUDA-core implicit-GEMM convolution: forward, dgrad, wgrad for ANY Conv3d geometry of the
// path, in fp32 (validation mode) or bf16 storage with fp32 accumulation.
//
// This family is (1) the fp32 validation mode north_star asks for and (2) the kernel that
// covers geometries the tcgen05 family does not take yet.  It replaces the cuDNN calls behind
// nn.Conv3d at /root/reference/src/models/R2Plus1D.py:44-51 and its autograd.
#include "dp_common.cuh"
#include "conv_internal.cuh"

namespace dp {

constexpr int BM = 64, BN = 64, BK = 16;

// Geometry of one "gather GEMM": dst[m][n] = sum_{tap,r} src[pix(m,tap)][r] * wgt[n][tap][r]
struct Geom {
  int B;
  int sT, sH, sW, sC;  // source pixels / padded reduction channels
  int dT, dH, dW, dC;  // destination pixels / padded output channels
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
gather_gemm_kernel(Geom g, const T* __restrict__ src, const T* __restrict__ wgt,
                   const T* __restrict__ addend, T* __restrict__ dst, int64_t M) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int64_t m = m0 + lrow;
  const bool mvalid = m < M;
  int w = 0, h = 0, t = 0, b = 0;
  if (mvalid) {
    int64_t r = m;
    w = (int)(r % g.dW); r /= g.dW;
    h = (int)(r % g.dH); r /= g.dH;
    t = (int)(r % g.dT); r /= g.dT;
    b = (int)r;
  }
  const int nld = n0 + lrow;
  const bool nvalid = nld < g.dC;
  const int taps = g.kt * g.kh * g.kw;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    const int dw = tap % g.kw, dh = (tap / g.kw) % g.kh, dt = tap / (g.kw * g.kh);
    bool v = mvalid;
    int ts, hs, ws;
    if (!BWD) {
      ts = t * g.st - g.pt + dt;
      hs = h * g.sh - g.ph + dh;
      ws = w * g.sw - g.pw + dw;
      v = v && ts >= 0 && ts < g.sT && hs >= 0 && hs < g.sH && ws >= 0 && ws < g.sW;
    } else {
      const int tt = t + g.pt - dt, hh = h + g.ph - dh, ww = w + g.pw - dw;
      v = v && tt >= 0 && hh >= 0 && ww >= 0 && (tt % g.st) == 0 && (hh % g.sh) == 0 && (ww % g.sw) == 0;
      ts = tt / g.st; hs = hh / g.sh; ws = ww / g.sw;
      v = v && ts < g.sT && hs < g.sH && ws < g.sW;
    }
    const T* ap = src;
    if (v) ap = src + ((((int64_t)b * g.sT + ts) * g.sH + hs) * g.sW + ws) * g.sC;
    const T* bp = wgt + ((int64_t)(nvalid ? nld : 0) * taps + tap) * g.sC;
    for (int r0 = 0; r0 < g.sC; r0 += BK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bb = a;
      if (v) a = ld4(ap + r0 + lk);
      if (nvalid) bb = ld4(bp + r0 + lk);
      As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      Bs[lk + 0][lrow] = bb.x; Bs[lk + 1][lrow] = bb.y; Bs[lk + 2][lrow] = bb.z; Bs[lk + 3][lrow] = bb.w;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
        const float bq[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bq[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const int n = n0 + tx * 4;
  if (n < g.dC) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t mm = m0 + ty * 4 + i;
      if (mm < M) {
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (addend != nullptr) {
          const float4 ad = ld4(addend + mm * g.dC + n);
          o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
        }
        st4(dst + mm * g.dC + n, o);
      }
    }
  }
}

// partial[split][k][tap][c] = sum over the split's output pixels of dy[m][k] * x[pix(m,tap)][c]
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_kernel(Geom g, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partial,
             int64_t M, int64_t chunk, int ctiles) {
  __shared__ __align__(16) float As[BK][BN + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k0 = (blockIdx.x / ctiles) * BN, c0 = (blockIdx.x % ctiles) * BN;
  const int tap = blockIdx.y;
  const int taps = g.kt * g.kh * g.kw;
  const int dw = tap % g.kw, dh = (tap / g.kw) % g.kh, dt = tap / (g.kw * g.kh);
  const int64_t m_begin = (int64_t)blockIdx.z * chunk;
  const int64_t m_end = (m_begin + chunk < M) ? (m_begin + chunk) : M;
  const int lpix = tid >> 4, lch = (tid & 15) * 4;
  const int Kp = g.dC, Cp = g.sC;
  const bool kvalid = (k0 + lch) < Kp, cvalid = (c0 + lch) < Cp;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t mb = m_begin; mb < m_end; mb += BK) {
    const int64_t m = mb + lpix;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bb = a;
    if (m < m_end) {
      int64_t r = m;
      const int wo = (int)(r % g.dW); r /= g.dW;
      const int ho = (int)(r % g.dH); r /= g.dH;
      const int to = (int)(r % g.dT); r /= g.dT;
      const int b = (int)r;
      const int ts = to * g.st - g.pt + dt, hs = ho * g.sh - g.ph + dh, ws = wo * g.sw - g.pw + dw;
      const bool inb = ts >= 0 && ts < g.sT && hs >= 0 && hs < g.sH && ws >= 0 && ws < g.sW;
      if (kvalid) a = ld4(dy + m * Kp + k0 + lch);
      if (inb && cvalid) bb = ld4(x + ((((int64_t)b * g.sT + ts) * g.sH + hs) * g.sW + ws) * Cp + c0 + lch);
    }
    *reinterpret_cast<float4*>(&As[lpix][lch]) = a;
    *reinterpret_cast<float4*>(&Bs[lpix][lch]) = bb;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bq[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bq[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* out = partial + (int64_t)blockIdx.z * Kp * taps * Cp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= Kp) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < Cp) out[((int64_t)k * taps + tap) * Cp + c] = acc[i][j];
    }
  }
}

// dw[k][c][tap] (PyTorch (K,C,kt,kh,kw) order) = sum_split partial[split][k][tap][c], splits added in a fixed
// order (deterministic).  Threads run along c, so every split's read is a coalesced row.
constexpr int WR_LANES = 16;   // split-lanes per output element (the loop is pure load latency: <= 10 loads per lane at 148 splits)
constexpr int WR_ELEMS = 16;   // consecutive elements (along c) per CTA
__global__ void __launch_bounds__(WR_LANES * WR_ELEMS)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                    int nsplit, int K, int C, int Kp, int Cp, int taps) {
  pdl_launch_dependents();
  pdl_wait();
  // 16 lanes share the <= 148 splits of one element (the loop is pure load latency), each with four independent chains;
  // lanes and chains are combined in a fixed order, so the result is deterministic
  __shared__ float red[WR_LANES][WR_ELEMS];
  const int el = threadIdx.x % WR_ELEMS, ln = threadIdx.x / WR_ELEMS;
  const int idx = blockIdx.x * WR_ELEMS + el;
  const int total = K * taps * Cp;
  float acc = 0.f;
  if (idx < total) {
    const int64_t stride = (int64_t)Kp * taps * Cp;
    const float* p = partial + idx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = ln;
    for (; i + 3 * WR_LANES < nsplit; i += 4 * WR_LANES) {
      s0 += p[(int64_t)i * stride];
      s1 += p[(int64_t)(i + WR_LANES) * stride];
      s2 += p[(int64_t)(i + 2 * WR_LANES) * stride];
      s3 += p[(int64_t)(i + 3 * WR_LANES) * stride];
    }
    for (; i < nsplit; i += WR_LANES) s0 += p[(int64_t)i * stride];
    acc = (s0 + s1) + (s2 + s3);
  }
  red[ln][el] = acc;
  __syncthreads();
  if (ln == 0 && idx < total) {
    const int c = idx % Cp, tap = (idx / Cp) % taps, k = idx / (Cp * taps);
    if (c < C) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < WR_LANES; ++j) t += red[j][el];
      dw[((int64_t)k * C + c) * taps + tap] = t;
    }
  }
}

int wgrad_reduce_launch(const float* partial, float* dw, int nsplit, int K, int C, int Kp, int Cp, int taps,
                        cudaStream_t s) {
  const int total = K * taps * Cp;
  launch_pdl_small(wgrad_reduce_kernel, dim3(ceil_div(total, WR_ELEMS)), dim3(WR_LANES * WR_ELEMS), 0, s, partial, dw, nsplit, K, C, Kp,
             Cp, taps);
  return check_launch("conv_wgrad_reduce");
}

static Geom fwd_geom(const dp_conv_desc* d) {
  Geom g;
  g.B = d->B;
  g.sT = d->Ti; g.sH = d->Hi; g.sW = d->Wi; g.sC = d->Cp;
  g.dT = d->To; g.dH = d->Ho; g.dW = d->Wo; g.dC = d->Kp;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  return g;
}

template <typename T>
static int launch_fwd(const dp_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  Geom g = fwd_geom(d);
  const int64_t M = (int64_t)d->B * d->To * d->Ho * d->Wo;
  dim3 grid(ceil_div(M, BM), ceil_div(d->Kp, BN));
  gather_gemm_kernel<T, false><<<grid, 256, 0, s>>>(g, (const T*)x, (const T*)w, (const T*)nullptr, (T*)y, M);
  return check_launch("conv_fwd_simt");
}

template <typename T>
static int launch_dgrad(const dp_conv_desc* d, const void* dy, const void* w, const void* addend,
                        void* dx, cudaStream_t s) {
  Geom g;
  g.B = d->B;
  g.sT = d->To; g.sH = d->Ho; g.sW = d->Wo; g.sC = d->Kp;
  g.dT = d->Ti; g.dH = d->Hi; g.dW = d->Wi; g.dC = d->Cp;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  const int64_t M = (int64_t)d->B * d->Ti * d->Hi * d->Wi;
  dim3 grid(ceil_div(M, BM), ceil_div(d->Cp, BN));
  gather_gemm_kernel<T, true><<<grid, 256, 0, s>>>(g, (const T*)dy, (const T*)w, (const T*)addend, (T*)dx, M);
  return check_launch("conv_dgrad_simt");
}

static void wgrad_split(const dp_conv_desc* d, int* nsplit, int64_t* chunk) {
  const int64_t M = (int64_t)d->B * d->To * d->Ho * d->Wo;
  const int taps = d->kt * d->kh * d->kw;
  const int base = ceil_div(d->Kp, BN) * ceil_div(d->Cp, BN) * taps;
  int ns = ceil_div(4 * 148, base);
  if (ns > 64) ns = 64;
  const int64_t max_ns = (M + 255) / 256;
  if (ns > max_ns) ns = (int)max_ns;
  if (ns < 1) ns = 1;
  int64_t ch = (M + ns - 1) / ns;
  ch = (ch + BK - 1) / BK * BK;
  ns = (int)((M + ch - 1) / ch);
  *nsplit = ns;
  *chunk = ch;
}

size_t simt_wgrad_workspace(const dp_conv_desc* d) {
  int ns; int64_t ch;
  wgrad_split(d, &ns, &ch);
  return (size_t)ns * d->Kp * d->kt * d->kh * d->kw * d->Cp * sizeof(float);
}

template <typename T>
static int launch_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* ws,
                        cudaStream_t s) {
  Geom g = fwd_geom(d);
  const int64_t M = (int64_t)d->B * d->To * d->Ho * d->Wo;
  const int taps = d->kt * d->kh * d->kw;
  int ns; int64_t ch;
  wgrad_split(d, &ns, &ch);
  const int ctiles = ceil_div(d->Cp, BN);
  dim3 grid(ceil_div(d->Kp, BN) * ctiles, taps, ns);
  wgrad_kernel<T><<<grid, 256, 0, s>>>(g, (const T*)x, (const T*)dy, (float*)ws, M, ch, ctiles);
  int rc = check_launch("conv_wgrad_simt");
  if (rc != DP_OK) return rc;
  return wgrad_reduce_launch((const float*)ws, dw, ns, d->K, d->C, d->Kp, d->Cp, taps, s);
}

int simt_conv_fwd(const dp_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  return d->dtype == DP_BF16 ? launch_fwd<__nv_bfloat16>(d, x, w, y, s) : launch_fwd<float>(d, x, w, y, s);
}
int simt_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                    cudaStream_t s) {
  return d->dtype == DP_BF16 ? launch_dgrad<__nv_bfloat16>(d, dy, w, addend, dx, s)
                             : launch_dgrad<float>(d, dy, w, addend, dx, s);
}
int simt_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* ws,
                    cudaStream_t s) {
  return d->dtype == DP_BF16 ? launch_wgrad<__nv_bfloat16>(d, x, dy, dw, ws, s)
                             : launch_wgrad<float>(d, x, dy, dw, ws, s);
}

}  // namespace dp
