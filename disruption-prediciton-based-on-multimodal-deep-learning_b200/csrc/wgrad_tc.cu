// tcgen05 / TMEM / TMA weight-gradient kernel for the (2+1)D convolutions (bf16 in, fp32 accumulate):
//
//   dw[k][tap][c] = sum over output pixels m of  dy[m][k] * x[pix(m, tap)][c]
//
// i.e. autograd's bwd-filter of the nn.Conv3d calls at /root/reference/src/models/R2Plus1D.py:44-51,57.
//
// GEMM view: the reduction runs over PIXELS, and both operands are stored pixel-major (NDHWC: one row of
// channels per pixel), so both are "MN-major" UMMA operands: a TMA box [128 pixels][64 channels] with the
// 128-byte swizzle is exactly the canonical MN-major SW128 layout (8-pixel x 128-byte atoms, SBO = 1024 B
// between 8-pixel groups, LBO = the byte distance between 64-channel chunks).  No transposition anywhere.
//
// One work item = (pixel split, M tile, tap group).  A CTA streams the split's 128-pixel tiles through two
// TMA rings (dy tiles; x boxes, one per tap or one halo box per kw / for all kt), and for every tap issues
// 8 MMAs (16 pixels each) into that tap's own TMEM columns: all taps of the group accumulate in TMEM at
// once (<= 512 columns), so dy and x are read once per group.  The epilogue drains TMEM into a per-split
// fp32 partial [Kp][taps][Cp]; a second kernel reduces the splits in a fixed order (deterministic).
//
// Roles: whichever of (dy channels, x channels) is larger sits on the M side (TMEM lanes, M = 64 or 128 per
// instruction), the other on the N side (<= 256, columns).  Warp roles as in conv_tc.cu.
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace dp {
using namespace ptx;

constexpr int WG_MAX_LOADS = 49;
constexpr int WG_THREADS = 256;
constexpr int WG_EPI = 128;
constexpr int WG_SMEM_MAX = 232448;
constexpr int WG_P = 128;  // pixels per tile (reduction depth per pipeline step)

struct WgParams {
  int B, ntile_w, ntile_h, ntile_t, num_tiles;
  int bw, bh, bt;
  int mw, mh, mt;
  int nloads, nsub, tap_sub_stride, taps;
  int Kp, Cp;
  int swap;  // 0: M side = dy (rows of D are k), 1: M side = x (rows of D are c)
  // tap stacking (swap == 1, halo loads, Cp <= 64): `stack` sub-taps of one load are stacked along M -- their x windows
  // are the same smem box at row shifts, i.e. MN-major chunks at a constant LBO -- so one MMA covers several taps
  int stack, gpl, cbS, M_last;
  int n_mt, n_tg, lpg, nsplit, tiles_per_split, num_items;
  int cbX, cbD, chunksX, chunksD;
  int x_chunk_bytes, d_chunk_bytes, x_slot_bytes, d_slot_bytes, stage_bytes, num_stages;
  int x_rowbytes, d_rowbytes, x_layout, d_layout;
  int x_shift_bytes, x_box_bytes, d_box_bytes;
  int N, tmem_cols;
  int off_bars;
  signed char off_w[WG_MAX_LOADS], off_h[WG_MAX_LOADS], off_t[WG_MAX_LOADS];
  short tap0[WG_MAX_LOADS];
};

struct WgItem {
  int split, mtile, l0, l1, tile0, tile1, Mi, nchunkM;
};

__device__ __forceinline__ WgItem wg_decode(const WgParams& p, int item) {
  WgItem it;
  const int tg = item % p.n_tg;
  it.mtile = (item / p.n_tg) % p.n_mt;
  it.split = item / (p.n_tg * p.n_mt);
  it.l0 = tg * p.lpg;
  it.l1 = min(p.nloads, it.l0 + p.lpg);
  it.tile0 = it.split * p.tiles_per_split;
  it.tile1 = min(p.num_tiles, it.tile0 + p.tiles_per_split);
  const int CU = p.swap ? p.Cp : p.Kp;
  const int rows = min(128, CU - 128 * it.mtile);
  it.Mi = rows <= 64 ? 64 : 128;
  it.nchunkM = it.Mi / 64;
  return it;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD,
                const __grid_constant__ WgParams p, float* __restrict__ partial, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.num_stages;
  const uint32_t bars = sbase + p.off_bars;
  auto full_bar = [&](int i) { return bars + 8u * i; };
  auto empty_bar = [&](int i) { return bars + 8u * (S + i); };
  const uint32_t tfull = bars + 8u * (2 * S), tempty = tfull + 8u;
  const uint32_t tmem_slot = tempty + 8u;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.off_bars + 8 * (2 * S + 2));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmD);
    for (int i = 0; i < S; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, WG_EPI);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    // one pipeline stage = the dy tile plus every x box of the item's tap group: every mbarrier operation occupies its
    // thread for ~200 cycles on B200 (scripts/ubench/sync_ops.cu), so a tile costs ONE wait / expect_tx / commit round
    int st = 0;
    uint32_t ph = 0;
    long long w_prod = 0;
    const long long t_start = dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const int nD = p.swap ? p.chunksD : it.nchunkM;
      const int nX = p.swap ? it.nchunkM : p.chunksX;
      const int chanD = p.swap ? 0 : 128 * it.mtile;
      const int chanX = p.swap ? 128 * it.mtile : 0;
      const uint32_t tx = (uint32_t)(nD * p.d_box_bytes + (it.l1 - it.l0) * nX * p.x_box_bytes);
      for (int tile = it.tile0; tile < it.tile1; ++tile) {
        int r = tile;
        const int tw = r % p.ntile_w; r /= p.ntile_w;
        const int th = r % p.ntile_h; r /= p.ntile_h;
        const int tt = r % p.ntile_t; r /= p.ntile_t;
        const int b = r;
        const int w0 = tw * p.bw, h0 = th * p.bh, t0 = tt * p.bt;
        long long c0 = dbg ? clock64() : 0;
        mbar_wait(empty_bar(st), ph ^ 1u);
        if (dbg) w_prod += clock64() - c0;
        const uint32_t da = sbase + (uint32_t)(st * p.stage_bytes);
        mbar_expect_tx(full_bar(st), tx);
        for (int j = 0; j < nD; ++j)
          tma_load_5d(&tmD, full_bar(st), da + (uint32_t)(j * p.d_chunk_bytes), chanD + j * p.cbD, w0, h0, t0, b);
        uint32_t xa = da + (uint32_t)p.d_slot_bytes;
        for (int l = it.l0; l < it.l1; ++l, xa += (uint32_t)p.x_slot_bytes)
          for (int j = 0; j < nX; ++j)
            tma_load_5d(&tmX, full_bar(st), xa + (uint32_t)(j * p.x_chunk_bytes), chanX + j * p.cbX,
                        w0 * p.mw + p.off_w[l], h0 * p.mh + p.off_h[l], t0 * p.mt + p.off_t[l], b);
        if (++st == S) { st = 0; ph ^= 1u; }
      }
    }
    if (dbg) { dbg[blockIdx.x * 8 + 0] = w_prod; dbg[blockIdx.x * 8 + 1] = clock64() - t_start; }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    int st = 0;
    uint32_t ph = 0, tph = 0;
    // operand geometry: U = M side, V = N side
    const int u_rowbytes = p.swap ? p.x_rowbytes : p.d_rowbytes;
    const int v_rowbytes = p.swap ? p.d_rowbytes : p.x_rowbytes;
    const uint32_t u_hi = smem_desc_hi((uint32_t)(8 * u_rowbytes), (uint32_t)(p.swap ? p.x_layout : p.d_layout));
    const uint32_t v_hi = smem_desc_hi((uint32_t)(8 * v_rowbytes), (uint32_t)(p.swap ? p.d_layout : p.x_layout));
    const uint32_t u_lbo = (uint32_t)(p.stack > 1 ? p.x_shift_bytes : (p.swap ? p.x_chunk_bytes : p.d_chunk_bytes));
    const uint32_t v_lbo = (uint32_t)(p.swap ? p.d_chunk_bytes : p.x_chunk_bytes);
    const uint32_t u_step = (uint32_t)(16 * u_rowbytes) >> 4, v_step = (uint32_t)(16 * v_rowbytes) >> 4;
    const uint32_t x_shift16 = (uint32_t)p.x_shift_bytes >> 4;
    const uint32_t d_lo0 = smem_desc_lo(sbase, p.swap ? v_lbo : u_lbo);
    const uint32_t x_lo0 = smem_desc_lo(sbase + (uint32_t)p.d_slot_bytes, p.swap ? u_lbo : v_lbo);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, x_slot16 = (uint32_t)p.x_slot_bytes >> 4;
    const int ngrp = p.stack > 1 ? p.gpl : p.nsub;                          // MMA groups (accumulator blocks) per load
    const uint32_t x_adv16 = (uint32_t)(p.stack > 1 ? p.stack : 1) * x_shift16;
    const uint32_t N = (uint32_t)p.N;
    const bool leader = elect_one();
    long long w_full = 0, w_te = 0;
    const long long t_start = dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const uint32_t idesc = make_idesc_bf16(p.stack > 1 ? 128 : it.Mi, p.N, 1, 1);
      const uint32_t idesc_last = p.stack > 1 ? make_idesc_bf16(p.M_last, p.N, 1, 1) : idesc;
      long long c0 = dbg ? clock64() : 0;
      mbar_wait(tempty, tph ^ 1u);
      if (dbg) w_te += clock64() - c0;
      tc_fence_after();
      const int nl = it.l1 - it.l0;
      for (int tile = it.tile0; tile < it.tile1; ++tile) {
        c0 = dbg ? clock64() : 0;
        mbar_wait(full_bar(st), ph);
        if (dbg) w_full += clock64() - c0;
        tc_fence_after();
        if (leader) {
          const uint32_t d_lo = d_lo0 + (uint32_t)st * stage16;
          const uint32_t first = (tile != it.tile0) ? 1u : 0u;
          uint32_t tmem_d = tmem_base;
          uint32_t x_slot = x_lo0 + (uint32_t)st * stage16;
          for (int l = 0; l < nl; ++l, x_slot += x_slot16) {
            uint32_t x_lo = x_slot;
            for (int g = 0; g < ngrp; ++g) {
              const uint32_t u_lo = p.swap ? x_lo : d_lo, v_lo = p.swap ? d_lo : x_lo;
              const uint32_t id = (g == ngrp - 1) ? idesc_last : idesc;
              umma_bf16_lh(tmem_d, u_lo, u_hi, v_lo, v_hi, id, first);
#pragma unroll
              for (int ks = 1; ks < WG_P / 16; ++ks)
                umma_bf16_lh(tmem_d, u_lo + ks * u_step, u_hi, v_lo + ks * v_step, v_hi, id, 1u);
              x_lo += x_adv16;
              tmem_d += N;
            }
          }
          umma_commit(empty_bar(st));
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(tfull);
      __syncwarp();
      tph ^= 1u;
    }
    if (dbg && lane == 0) { dbg[blockIdx.x * 8 + 2] = w_full; dbg[blockIdx.x * 8 + 3] = w_te; dbg[blockIdx.x * 8 + 4] = clock64() - t_start; }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> fp32 partial =====================
    const int q = warp & 3;
    uint32_t tph = 0;
    const int64_t split_stride = (int64_t)p.Kp * p.taps * p.Cp;
    long long w_tf = 0;
    const long long t_start = dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const int row = it.Mi == 128 ? (q * 32 + lane) : (q * 16 + lane);
      const bool lane_ok = it.Mi == 128 || lane < 16;
      const int u = 128 * it.mtile + row;                // channel index on the M side
      const bool valid = lane_ok && u < (p.swap ? p.Cp : p.Kp);
      float* base = partial + (int64_t)it.split * split_stride;
      const long long c0 = dbg ? clock64() : 0;
      mbar_wait(tfull, tph);
      if (dbg) w_tf += clock64() - c0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      if (p.stack > 1) {
        for (int l = it.l0; l < it.l1; ++l) {
          for (int g = 0; g < p.gpl; ++g) {
            const int Mg = (g == p.gpl - 1) ? p.M_last : 128;
            const int r = Mg == 128 ? (q * 32 + lane) : (q * 16 + lane);       // D row held by this thread
            const int sub = r / p.cbS, c = r - sub * p.cbS;
            const int sidx = g * p.stack + sub;
            const bool ok = (Mg == 128 || lane < 16) && sidx < p.nsub && c < p.Cp;
            const int tap = p.tap0[l] + sidx * p.tap_sub_stride;
            const uint32_t col = (uint32_t)(((l - it.l0) * p.gpl + g) * p.N);
            for (int n0 = 0; n0 < p.N; n0 += 16) {
              uint32_t v[16];
              tmem_ld16(taddr + col + (uint32_t)n0, v);
              if (ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  base[((int64_t)(n0 + j) * p.taps + tap) * p.Cp + c] = __uint_as_float(v[j]);
              }
            }
          }
        }
      } else
      for (int l = it.l0; l < it.l1; ++l) {
        for (int s = 0; s < p.nsub; ++s) {
          const int tap = p.tap0[l] + s * p.tap_sub_stride;
          const uint32_t col = (uint32_t)(((l - it.l0) * p.nsub + s) * p.N);
          for (int n0 = 0; n0 < p.N; n0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + col + (uint32_t)n0, v);
            if (valid) {
              if (!p.swap) {   // row = k, columns = c
                float4* dst = reinterpret_cast<float4*>(base + ((int64_t)u * p.taps + tap) * p.Cp + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
              } else {         // row = c, columns = k
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  base[((int64_t)(n0 + j) * p.taps + tap) * p.Cp + u] = __uint_as_float(v[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
      tph ^= 1u;
    }
    if (dbg && threadIdx.x == WG_THREADS - WG_EPI) { dbg[blockIdx.x * 8 + 5] = w_tf; dbg[blockIdx.x * 8 + 6] = clock64() - t_start; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled wg_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  });
  return fn;
}

static int g_wg_enable = 1, g_wg_halo = 1, g_wg_stack = 1, g_wg_items = 1;
int wg_option(const char* name, int value, bool set) {
  int* slot = nullptr;
  if (!strcmp(name, "wg_enable")) slot = &g_wg_enable;
  else if (!strcmp(name, "wg_halo")) slot = &g_wg_halo;
  else if (!strcmp(name, "wg_stack")) slot = &g_wg_stack;
  else if (!strcmp(name, "wg_items")) slot = &g_wg_items;   // work items (pixel splits) per SM
  if (slot == nullptr) return -1;
  if (set) *slot = value;
  return *slot;
}

struct WgPlan {
  WgParams p;
  int grid;
  size_t smem;
  int x_box[5], x_estride[5], d_box[5];
};

static inline int rup(int a, int b) { return (a + b - 1) / b * b; }
static inline int chunk_width(int C) { return C >= 64 ? 64 : (C > 16 ? 32 : 16); }
static inline int layout_of(int cb) { return cb == 64 ? 2 : (cb == 32 ? 4 : 6); }

static bool plan_wgrad(const dp_conv_desc* d, WgPlan* out) {
  if (!g_wg_enable) return false;
  const int taps = d->kt * d->kh * d->kw;
  if (taps > WG_MAX_LOADS) return false;
  if (d->Cp % 16 || d->Kp % 16) return false;
  if (d->st > 8 || d->sh > 8 || d->sw > 8) return false;
  const bool strided = d->st != 1 || d->sh != 1 || d->sw != 1;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.Kp = d->Kp; p.Cp = d->Cp; p.taps = taps;
  p.mw = d->sw; p.mh = d->sh; p.mt = d->st;

  // ---- pixel tile + load mode (0: one x box per tap, 1: one halo box per kw, 2: one halo box for all kt) ----
  double best = 1e30;
  int best_mode = -1, bbw = 0, bbh = 0, bbt = 0;
  for (int relaxed = 0; relaxed < 2 && best_mode < 0; ++relaxed)
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 1 && !(g_wg_halo && !strided && d->kt == 1 && d->kh > 1)) continue;
    if (mode == 2 && !(g_wg_halo && !strided && d->kh == 1 && d->kw == 1 && d->kt > 1)) continue;
    for (int bw = 1; bw <= WG_P; bw <<= 1)
      for (int bh = 1; bh * bw <= WG_P; bh <<= 1) {
        const int bt = WG_P / (bw * bh);
        if (mode == 1 && (bt != 1 || bw % 8)) continue;
        if (mode == 2 && ((bw * bh) % 8)) continue;
        if (bw * d->sw > 256 || bh * d->sh > 256 || bt * d->st > 256) continue;
        if (!relaxed) {
          if (bw > 2 * d->Wo && bw > 8) continue;
          if (bh >= 2 * d->Ho && bh > 1) continue;
          if (bt >= 2 * d->To && bt > 1) continue;
        }
        int rows = WG_P, nloads = taps;
        if (mode == 1) { rows = (bh + d->kh - 1) * bw; nloads = d->kw; if (bh + d->kh - 1 > 256) continue; }
        if (mode == 2) { rows = (bt + d->kt - 1) * bh * bw; nloads = 1; if (bt + d->kt - 1 > 256) continue; }
        const double ntiles = (double)((d->Wo + bw - 1) / bw) * ((d->Ho + bh - 1) / bh) * ((d->To + bt - 1) / bt);
        const double cost = ntiles * ((double)nloads * rows * d->Cp + (double)WG_P * d->Kp) - 1e-3 * bw;
        if (cost < best) { best = cost; best_mode = mode; bbw = bw; bbh = bh; bbt = bt; }
      }
  }
  if (best_mode < 0) return false;
  p.bw = bbw; p.bh = bbh; p.bt = bbt;
  p.ntile_w = (d->Wo + p.bw - 1) / p.bw;
  p.ntile_h = (d->Ho + p.bh - 1) / p.bh;
  p.ntile_t = (d->To + p.bt - 1) / p.bt;
  const int64_t nt = (int64_t)d->B * p.ntile_w * p.ntile_h * p.ntile_t;
  if (nt > 0x7fffffff) return false;
  p.num_tiles = (int)nt;

  int x_rows = WG_P;
  out->x_estride[0] = 1; out->x_estride[1] = d->sw; out->x_estride[2] = d->sh; out->x_estride[3] = d->st; out->x_estride[4] = 1;
  out->x_box[4] = 1;
  auto tap_index = [&](int jt, int jh, int jw) { return (jt * d->kh + jh) * d->kw + jw; };
  int shift_rows = 0;
  if (best_mode == 0) {
    p.nloads = taps; p.nsub = 1; p.tap_sub_stride = 0;
    int l = 0;
    for (int jt = 0; jt < d->kt; ++jt)
      for (int jh = 0; jh < d->kh; ++jh)
        for (int jw = 0; jw < d->kw; ++jw, ++l) {
          p.off_t[l] = (signed char)(jt - d->pt); p.off_h[l] = (signed char)(jh - d->ph); p.off_w[l] = (signed char)(jw - d->pw);
          p.tap0[l] = (short)tap_index(jt, jh, jw);
        }
    out->x_box[1] = p.bw * d->sw; out->x_box[2] = p.bh * d->sh; out->x_box[3] = p.bt * d->st;
  } else if (best_mode == 1) {
    p.nloads = d->kw; p.nsub = d->kh; p.tap_sub_stride = tap_index(0, 1, 0) - tap_index(0, 0, 0);
    for (int jw = 0; jw < d->kw; ++jw) {
      p.off_t[jw] = (signed char)(-d->pt); p.off_h[jw] = (signed char)(-d->ph); p.off_w[jw] = (signed char)(jw - d->pw);
      p.tap0[jw] = (short)tap_index(0, 0, jw);
    }
    x_rows = (p.bh + d->kh - 1) * p.bw;
    shift_rows = p.bw;
    out->x_box[1] = p.bw; out->x_box[2] = p.bh + d->kh - 1; out->x_box[3] = 1;
  } else {
    p.nloads = 1; p.nsub = d->kt; p.tap_sub_stride = tap_index(1, 0, 0) - tap_index(0, 0, 0);
    p.off_t[0] = (signed char)(-d->pt); p.off_h[0] = (signed char)(-d->ph); p.off_w[0] = (signed char)(-d->pw);
    p.tap0[0] = 0;
    x_rows = (p.bt + d->kt - 1) * p.bh * p.bw;
    shift_rows = p.bh * p.bw;
    out->x_box[1] = p.bw; out->x_box[2] = p.bh; out->x_box[3] = p.bt + d->kt - 1;
  }
  out->d_box[1] = p.bw; out->d_box[2] = p.bh; out->d_box[3] = p.bt; out->d_box[4] = 1;

  // ---- roles ----
  p.stack = 1; p.gpl = p.nsub; p.cbS = 0; p.M_last = 0;
  if (g_wg_stack && p.nsub > 1 && d->Cp <= 64 && d->Kp <= 256) {
    const int cbS = d->Cp <= 32 ? 32 : 64, sf = 128 / cbS;
    const int gpl = (p.nsub + sf - 1) / sf;
    if (gpl * d->Kp <= 512) {
      p.stack = sf; p.gpl = gpl; p.cbS = cbS;
      const int cnt_last = p.nsub - (gpl - 1) * sf;
      p.M_last = cnt_last * cbS <= 64 ? 64 : 128;
    }
  }
  int best_swap = -1;
  long best_key = 0;
  for (int swap = 0; swap < 2; ++swap) {
    const int CU = swap ? d->Cp : d->Kp, CV = swap ? d->Kp : d->Cp;
    if (CV > 256) continue;
    const int lpg = 512 / (p.nsub * CV);
    if (lpg < 1) continue;
    const int n_mt = (CU + 127) / 128, n_tg = (p.nloads + lpg - 1) / lpg;
    const long key = (long)n_mt * n_tg * 100000 + (long)n_mt * CV;
    if (best_swap < 0 || key < best_key) { best_swap = swap; best_key = key; }
  }
  if (best_swap < 0) return false;
  p.swap = p.stack > 1 ? 1 : best_swap;
  const int CU = p.swap ? d->Cp : d->Kp, CV = p.swap ? d->Kp : d->Cp;
  p.N = CV;
  p.lpg = 512 / ((p.stack > 1 ? p.gpl : p.nsub) * CV);
  if (p.lpg > p.nloads) p.lpg = p.nloads;
  p.n_mt = (CU + 127) / 128;
  p.n_tg = (p.nloads + p.lpg - 1) / p.lpg;
  int cols = 32;
  while (cols < p.lpg * (p.stack > 1 ? p.gpl : p.nsub) * p.N) cols <<= 1;
  p.tmem_cols = cols;

  // ---- chunking of the two operands ----
  p.cbX = p.stack > 1 ? p.cbS : (p.swap ? 64 : chunk_width(d->Cp));
  p.cbD = p.swap ? chunk_width(d->Kp) : 64;
  p.chunksX = p.swap ? (CU > 64 ? 2 : 1) : (d->Cp + p.cbX - 1) / p.cbX;
  p.chunksD = p.swap ? (d->Kp + p.cbD - 1) / p.cbD : (CU > 64 ? 2 : 1);
  p.x_rowbytes = p.cbX * 2; p.d_rowbytes = p.cbD * 2;
  p.x_layout = layout_of(p.cbX); p.d_layout = layout_of(p.cbD);
  p.x_box_bytes = x_rows * p.x_rowbytes;
  p.d_box_bytes = WG_P * p.d_rowbytes;
  p.x_chunk_bytes = rup(p.x_box_bytes, 1024);
  p.d_chunk_bytes = rup(p.d_box_bytes, 1024);
  p.x_slot_bytes = p.chunksX * p.x_chunk_bytes;
  p.d_slot_bytes = p.chunksD * p.d_chunk_bytes;
  p.x_shift_bytes = shift_rows * p.x_rowbytes;
  out->x_box[0] = p.cbX; out->d_box[0] = p.cbD;

  const int bar_bytes = 1024;
  const int stack_pad = p.stack > 1 ? rup(p.stack * p.x_shift_bytes, 1024) : 0;
  const int avail = WG_SMEM_MAX - 1024 - bar_bytes - stack_pad;
  // one stage = dy tile + the lpg x boxes of a tap group; shrink the tap group until two stages fit
  while (p.lpg > 1 && 2 * (p.d_slot_bytes + p.lpg * p.x_slot_bytes) > avail) --p.lpg;
  p.n_tg = (p.nloads + p.lpg - 1) / p.lpg;
  p.stage_bytes = p.d_slot_bytes + p.lpg * p.x_slot_bytes;
  int ns = avail / p.stage_bytes;
  if (ns > 8) ns = 8;
  if (ns < 2) return false;
  p.num_stages = ns;
  // stacked MMAs read up to `stack` shifted windows: the junk window of a partial stack may run past the last slot
  p.off_bars = p.num_stages * p.stage_bytes + stack_pad;
  out->smem = (size_t)p.off_bars + bar_bytes + 1024;
  if (out->smem > (size_t)WG_SMEM_MAX) return false;

  // ---- pixel splits ----
  const int sms = num_sms();
  const int kinds = p.n_mt * p.n_tg;
  // one item per SM: half the fp32 partials of two items (less drain + reduce traffic); measured 3.45 vs 3.68 ms per step
  // for the 32 weight gradients of the BASELINE model (three items: 3.88 ms)
  int nsplit = ((g_wg_items < 1 ? 1 : g_wg_items) * sms) / kinds;
  if (nsplit < 1) nsplit = 1;
  const int max_split = (p.num_tiles + 3) / 4;   // at least 4 tiles (512 pixels) per split
  if (nsplit > max_split) nsplit = max_split < 1 ? 1 : max_split;
  p.tiles_per_split = (p.num_tiles + nsplit - 1) / nsplit;
  p.nsplit = (p.num_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.num_items = p.nsplit * kinds;
  out->grid = p.num_items < sms ? p.num_items : sms;
  out->p = p;
  return true;
}

static int wg_encode_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, const int* box,
                         const int* estride, int cb, const long long* vstr = nullptr) {
  PFN_encodeTiled enc = wg_get_encode();
  DP_REQUIRE(enc != nullptr, DP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)T * H * W * C * 2};
  if (vstr != nullptr && vstr[0] > 0)
    for (int i = 0; i < 4; ++i) strides[i] = (cuuint64_t)vstr[i] * 2;
  cuuint32_t b[5], es[5];
  for (int i = 0; i < 5; ++i) { b[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)estride[i]; }
  const CUtensorMapSwizzle sw = cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DP_REQUIRE(r == CUDA_SUCCESS, DP_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad) failed: CUresult %d (box %d,%d,%d,%d)",
             (int)r, box[0], box[1], box[2], box[3]);
  return DP_OK;
}

// Output-channel halves: when neither operand order fits one N tile (both channel counts > 256: the 320 -> 512 shortcut
// of SlowFast's last stage), the weight gradient is computed as two problems over halves of K -- dy read through a
// strided view, dw rows [0, K/2) and [K/2, K) are contiguous blocks of the (K, C, taps) master layout.
static bool split_halves(const dp_conv_desc* d, dp_conv_desc* h0, dp_conv_desc* h1) {
  if (d->Kp <= 256 || d->Kp % 32) return false;
  const int half = d->Kp / 2;
  if (d->K <= half) return false;
  *h0 = *d; *h1 = *d;
  h0->Kp = half; h0->K = half;
  h1->Kp = half; h1->K = d->K - half;
  return true;
}

int tc_wgrad_describe(const dp_conv_desc* d, char* out, size_t n) {
  WgPlan plan;
  if (!plan_wgrad(d, &plan)) return DP_ERR_UNSUPPORTED;
  const WgParams& p = plan.p;
  snprintf(out, n, "tile bw=%d bh=%d bt=%d nloads=%d nsub=%d swap=%d stack=%d gpl=%d N=%d n_mt=%d n_tg=%d lpg=%d cbX=%d cbD=%d stages=%d "
           "stage_bytes=%d xslot=%d dslot=%d nsplit=%d tiles_per_split=%d tmem_cols=%d grid=%d smem=%zu",
           p.bw, p.bh, p.bt, p.nloads, p.nsub, p.swap, p.stack, p.gpl, p.N, p.n_mt, p.n_tg, p.lpg, p.cbX, p.cbD, p.num_stages, p.stage_bytes,
           p.x_slot_bytes, p.d_slot_bytes, p.nsplit, p.tiles_per_split, p.tmem_cols, plan.grid, plan.smem);
  return DP_OK;
}

bool tc_wgrad_supported(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16) return false;
  WgPlan plan;
  if (plan_wgrad(d, &plan)) return true;
  dp_conv_desc h0, h1;
  return split_halves(d, &h0, &h1) && plan_wgrad(&h0, &plan);
}

size_t tc_wgrad_workspace(const dp_conv_desc* d) {
  WgPlan plan;
  if (plan_wgrad(d, &plan)) return (size_t)plan.p.nsplit * d->Kp * plan.p.taps * d->Cp * sizeof(float);
  dp_conv_desc h0, h1;
  if (split_halves(d, &h0, &h1) && plan_wgrad(&h0, &plan))
    return (size_t)plan.p.nsplit * h0.Kp * plan.p.taps * d->Cp * sizeof(float);
  return 0;
}

int tc_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  return tc_conv_wgrad_view(d, nullptr, x, dy, dw, ws, ws_bytes, s);
}

static int wgrad_launch(const dp_conv_desc* d, const long long* xstrides, const void* x, const long long* dystrides,
                        const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t s);

int tc_conv_wgrad_view(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* dy, float* dw,
                       void* ws, size_t ws_bytes, cudaStream_t s) {
  WgPlan plan;
  if (plan_wgrad(d, &plan)) return wgrad_launch(d, xstrides, x, nullptr, dy, dw, ws, ws_bytes, s);
  dp_conv_desc h[2];
  DP_REQUIRE(split_halves(d, &h[0], &h[1]) && plan_wgrad(&h[0], &plan), DP_ERR_UNSUPPORTED,
             "tcgen05 wgrad: geometry not supported");
  const long long dys[4] = {(long long)d->Kp, (long long)d->Wo * d->Kp, (long long)d->Ho * d->Wo * d->Kp,
                            (long long)d->To * d->Ho * d->Wo * d->Kp};
  const int taps = d->kt * d->kh * d->kw;
  for (int i = 0; i < 2; ++i) {
    const int rc = wgrad_launch(&h[i], xstrides, x, dys, (const __nv_bfloat16*)dy + i * h[0].Kp,
                                dw + (size_t)i * h[0].Kp * d->C * taps, ws, ws_bytes, s);
    if (rc != DP_OK) return rc;
  }
  return DP_OK;
}

static int wgrad_launch(const dp_conv_desc* d, const long long* xstrides, const void* x, const long long* dystrides,
                        const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t s) {
  WgPlan plan;
  DP_REQUIRE(plan_wgrad(d, &plan), DP_ERR_UNSUPPORTED, "tcgen05 wgrad: geometry not supported");
  DP_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)ws & 15) == 0, DP_ERR_ALIGN,
             "tcgen05 wgrad: tensors must be 16-byte aligned");
  const WgParams& p = plan.p;
  DP_REQUIRE(ws_bytes >= (size_t)p.nsplit * d->Kp * p.taps * d->Cp * sizeof(float), DP_ERR_SHAPE,
             "tcgen05 wgrad: workspace too small");
  CUtensorMap tmX, tmD;
  int rc = wg_encode_map(&tmX, x, d->Cp, d->Wi, d->Hi, d->Ti, d->B, plan.x_box, plan.x_estride, p.cbX, xstrides);
  if (rc != DP_OK) return rc;
  const int ones[5] = {1, 1, 1, 1, 1};
  rc = wg_encode_map(&tmD, dy, d->Kp, d->Wo, d->Ho, d->To, d->B, plan.d_box, ones, p.cbD, dystrides);
  if (rc != DP_OK) return rc;
  static std::mutex attr_mu;
  static bool attr_done[DP_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  cudaError_t attr_err = cudaSuccess;
  {
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(attr_mu);
    if (!attr_done[dev]) {
      attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_MAX);
      attr_done[dev] = attr_err == cudaSuccess;
    }
  }
  DP_REQUIRE(attr_err == cudaSuccess, DP_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s",
             cudaGetErrorString(attr_err));
  launch_pdl(wgrad_tc_kernel, dim3(plan.grid), dim3(WG_THREADS), plan.smem, s, tmX, tmD, p, (float*)ws,
                                                           (g_dbg && g_dbg_slots >= (size_t)plan.grid * 8) ? g_dbg : nullptr);
  if (getenv("DP_DEBUG_PLAN"))
    fprintf(stderr, "[tc_wgrad] out %dx%dx%d Kp=%d Cp=%d taps=%d | tile bw=%d bh=%d bt=%d nloads=%d nsub=%d swap=%d N=%d n_mt=%d n_tg=%d lpg=%d cbX=%d cbD=%d stages=%d stage=%d xslot=%d dslot=%d nsplit=%d tiles/split=%d grid=%d\n",
            d->To, d->Ho, d->Wo, d->Kp, d->Cp, p.taps, p.bw, p.bh, p.bt, p.nloads, p.nsub, p.swap, p.N, p.n_mt, p.n_tg, p.lpg, p.cbX,
            p.cbD, p.num_stages, p.stage_bytes, p.x_slot_bytes, p.d_slot_bytes, p.nsplit, p.tiles_per_split, plan.grid);
  rc = check_launch("wgrad_tc_kernel");
  if (rc != DP_OK) return rc;
  return wgrad_reduce_launch((const float*)ws, dw, p.nsplit, d->K, d->C, d->Kp, d->Cp, p.taps, s);
}

}  // namespace dp
