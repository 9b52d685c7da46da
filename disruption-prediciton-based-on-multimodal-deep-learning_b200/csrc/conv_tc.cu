// tcgen05 / TMEM / TMA implicit-GEMM "gather GEMM" for the (2+1)D convolutions, bf16 in, fp32 accumulate.
//
//   dst[pixel][n] = sum_{tap, r} src[pixel * stride + offset(tap)][r] * wgt[n][tap][r]
//
// used as  forward  (src = x,  wgt = w_fwd  [Kp][taps][Cp], dst = y,  BN statistics in the epilogue)
// and as   dgrad    (src = dy, wgt = w_dgrad[Cp][taps][Kp], dst = dx, taps flipped, optional addend)
// for the nn.Conv3d calls at /root/reference/src/models/R2Plus1D.py:44-51,57 and their autograd.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised; tile = 128 * MT output pixels):
//   warp 0 lane 0 : TMA producer. NDHWC activations are a 5-D tensor map (C,W,H,T,B); one box = one tile of pixels
//                   (bw x bh x bt) x CB channels, zero-filled outside the tensor, so padding costs nothing.  Stride-1
//                   convs load ONE halo box per kw (or one for all kt) and reuse it for the kh (kt) taps by moving
//                   the UMMA descriptor start row.  The last channel block of C % 64 in {16, 32} is a narrow box with
//                   its own swizzle.  A pipeline stage carries `lps` loads (a whole tile where smem allows) and is
//                   armed (expect_tx) after its copies are issued.
//   warps 1, 2    : issue tcgen05.mma (M=128, N=Ntile<=256, K=16) into 2-4 TMEM accumulators.  In dual mode the two
//                   warps take alternate tiles, each on a PRIVATE ring of stages (a parity wait cannot tell phase k+1
//                   from k-1, so two consumers never share an mbarrier); tcgen05.commit releases the stage and
//                   publishes the accumulator.  Warp 2 also allocates / frees TMEM.
//   warp 3        : TMA store of the staged tile (clips the tensor edge); publishes "buffer free" through a counter
//                   in shared memory.
//   warps 4-11    : epilogue, two warps per TMEM lane quadrant, each draining half of the columns:
//                   tcgen05.ld -> (+addend) -> bf16 -> swizzled staging tile.  BatchNorm statistics (forward) and the
//                   producer's BatchNorm-backward sums (dgrad) are accumulated in registers from the fp32 accumulators;
//                   wide tiles fall back to Gram / ones MMAs in TMEM or to the CUDA cores.
// Hand-offs that the hardware does not force onto mbarriers (epilogue <-> store warp <-> MMA warps) use named barriers:
// every mbarrier operation occupies its thread for ~200 cycles on B200 (scripts/ubench/sync_ops.cu).
#include "dp_common.cuh"
#include "conv_internal.cuh"
#include "tc_ptx.cuh"
#include "bn_fin.cuh"
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace dp {
using namespace ptx;

constexpr int TC_MAX_LOADS = 49;
constexpr int TC_THREADS = 384;   // warps: 0 TMA loads, 1-2 MMA (2 also TMEM alloc), 3 TMA stores, 4-11 epilogue
constexpr int TC_EPI = 256;       // two epilogue warps per TMEM lane quadrant, each drains half of the columns
constexpr int TC_SMEM_MAX = 232448;  // 227 KB opt-in limit per CTA

struct TcParams {
  int B, ntile_w, ntile_h, ntile_t, n_ntiles, num_tiles;
  int bw, bh, bt;
  int dW, dH, dT, dC;       // destination dims / padded channels
  int mw, mh, mt;           // source coordinate multipliers (conv stride)
  int nloads, nsub, ncblk, CB, ksteps_last;
  int sub_row_bytes, tap_sub_stride, red_C;
  int Ntile;
  int a_box_bytes, b_box_bytes, b_sub_bytes, stage_bytes, num_stages;
  int lps, slot_bytes;      // loads per pipeline stage; bytes of one full-width load's slot
  // narrow tail block: the last channel block of a source with C % 64 in {16, 32} is loaded as a CBt-wide box with its own
  // swizzle instead of a zero-filled 64-wide one (C = 80: 37 % less smem fill and L2 traffic)
  int CBt, a_box_bytes_t, slot_bytes_t, layout_t, sbo_bytes_t, sub_row_bytes_t, unit_mode;
  int b_box_bytes_t, b_sub_bytes_t, w_row_bytes;   // resident weights: compact tail block; bytes of one (load, sub-tap) row
  int reg_stats;            // BN statistics accumulated in epilogue registers
  int dual_mma;             // two MMA-issuing warps on alternate tiles
  int MT;                   // 128-pixel sub-tiles per tile (1 or 2): one handshake round covers 128 * MT pixels
  int drain_rs;             // unrolled drain: chunks of 16 channels per epilogue warp (0 = generic loop)
  int w_resident, off_wgt;  // all weight blocks loaded once per CTA instead of once per stage
  int mma_stats, stat_M, acc_bufs;  // BN statistics accumulated in TMEM by the tensor core
  // staging tile (epilogue -> TMA store, and B operand of the statistics MMAs): chunks of cw channels, swizzled
  int cw_shift;  // log2(cw) for the chunked layout, -1 for one unswizzled chunk of Ntile channels
  // a staging buffer holds ONE 128-row sub-tile; st_bufs of them form a ring (256-pixel tiles stage and store their two
  // sub-tiles separately: the epilogue only waits for the store issued st_bufs SUB-tiles ago)
  int cw, st_chunks, st_rowbytes, st_chunk_bytes, st_buf_bytes, st_bufs, st_mask, st_layout;
  int sub_ow, sub_oh, sub_ot;   // destination offset of the second sub-tile inside a 256-pixel tile
  // 256-pixel tiles: 1 = every sub-tile is published to the store warp on its own (wide tiles: the epilogue of sub-tile
  // j + st_bufs only waits for the store of sub-tile j), 0 = one hand-off per tile (narrow tiles, where a second
  // fence / barrier / store-completion round per tile costs more than it hides: 164 vs 199 us on the 32-channel stem)
  int pub_sub;
  int off_staging, off_ones, off_stats, off_scratch, off_bars;
  int tmem_cols, layout_type, sbo_bytes;
  int has_stats, has_addend;
  float slope;   // STATS == 2: LeakyReLU slope of the producing layer; epi_bn: slope of this layer's activation
  // eval-mode fused epilogue (BatchNorm with running statistics folded into a per-channel affine, STATS == 0 only):
  //   v = lrelu(acc * scale[c] + shift[c], slope);  with an addend (the residual):  v = lrelu(v + addend, slope_res)
  int epi_bn;
  float slope_res;
  // stride-parity classes of a strided data gradient in ONE launch (STATS == 0): destination column g = n * Ntile + c
  // belongs to class g / cpd, channel g % cpd; each class is written through its own strided view (tmD, tmD1..3) and reads
  // its own addend origin / extent.  ncls <= 1: plain destination
  int ncls, cpd;
  long long a_off_k[8];
  short dWk[8], dHk[8], dTk[8];
  int dbg_skip;  // development: 1 = no epilogue data movement / stores, 2 = no TMA loads, 4 = one MMA per load
  int stats_keep;   // staged-tile statistics: per-thread sums live in registers across tiles (one N tile), combined once per CTA
  int pC, n_base;   // statistics partials: row pitch (channels of the full destination) and first channel of this launch
  int l2_hint;   // 1: activation boxes are loaded with an L2 evict-first policy (what the kernel WRITES outlives them in L2)
  long long a_off, a_sw, a_sh, a_st, a_sb;  // addend view: element offset / strides of (w,h,t,b) in the dst tensor
  signed char off_w[TC_MAX_LOADS], off_h[TC_MAX_LOADS], off_t[TC_MAX_LOADS];
  short tap0[TC_MAX_LOADS];
};

// Tile coordinates (n, w, h, t, b) of tile = blockIdx.x + i * gridDim.x, advanced by mixed-radix addition of the
// constant stride: no division in the per-tile path.
struct TileIter {
  int n, w, h, t, b;       // current coordinates
  int dn, dw, dh, dt, db;  // stride decomposition
  int Rn, Rw, Rh, Rt;
  __device__ __forceinline__ void init(const TcParams& p, int start, int stride) {
    Rn = p.n_ntiles; Rw = p.ntile_w; Rh = p.ntile_h; Rt = p.ntile_t;
    int r = start;
    n = r % Rn; r /= Rn; w = r % Rw; r /= Rw; h = r % Rh; r /= Rh; t = r % Rt; b = r / Rt;
    r = stride;
    dn = r % Rn; r /= Rn; dw = r % Rw; r /= Rw; dh = r % Rh; r /= Rh; dt = r % Rt; db = r / Rt;
  }
  __device__ __forceinline__ void next() {
    n += dn; int c = n >= Rn; n -= c ? Rn : 0;
    w += dw + c; c = w >= Rw; w -= c ? Rw : 0;
    h += dh + c; c = h >= Rh; h -= c ? Rh : 0;
    t += dt + c; c = t >= Rt; t -= c ? Rt : 0;
    b += db + c;
  }
};

// Named barriers (bar.sync / bar.arrive) hand tiles between the epilogue warps, the TMA-store warp and the MMA warp:
// a named-barrier hop costs ~40 cycles on B200 while every mbarrier operation (arrive or try_wait, even on a completed
// phase) occupies its thread for ~180-230 cycles (scripts/ubench/sync_ops.cu).  mbarriers remain only where the
// hardware requires them (TMA completion, tcgen05.commit), and the TMA pipeline moves `lps` loads per stage so that
// one wait / one expect_tx / one commit covers a whole tile's worth of operands where shared memory allows.
constexpr int BAR_SREADY = 3;   // +buf (<= 4 ring slots) : epilogue (arrive) -> store warp (sync): staged sub-tile complete
constexpr int BAR_TEMPTY = 7;   // +acc : epilogue (arrive) -> MMA warp (sync): TMEM accumulator drained
constexpr int BAR_HANDOFF = TC_EPI + 32;

// RS > 0: unrolled accumulator drain, <= RS chunks of 16 channels per epilogue warp.
// STATS (RS <= 3), accumulated in epilogue registers from the fp32 accumulators:
//   1: forward BatchNorm statistics  sum(y), sum(y^2)  of the tile being written;
//   2: (data gradient) the sums the BatchNorm backward of the PRODUCING layer needs,  sum(g'), sum(g' * yp)  with
//      g = this tile (+ addend), yp = that layer's raw conv output, g' = g * lrelu'(scale * yp + shift): the stand-alone
//      reduction pass over (g, yp) disappears
template <int RS, int STATS, int MT>   // MT: 128-row sub-tiles per tile (compile-time: the 128-pixel kernels pay nothing for it)
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD1,
                      const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmD3,
                      const __grid_constant__ TcParams p, const __nv_bfloat16* __restrict__ addend, float* __restrict__ part, long long* __restrict__ dbg,
                      const __nv_bfloat16* __restrict__ yprev, const float* __restrict__ bn_ss,
                      const __grid_constant__ dp_bn_fin fin) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.num_stages;
  const uint32_t bars = sbase + p.off_bars;
  auto full_bar = [&](int i) { return bars + 8u * i; };
  auto empty_bar = [&](int i) { return bars + 8u * (S + i); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  const uint32_t wbar = bars + 8u * (2 * S + 4);
  auto sready_bar = [&](int b) { return bars + 8u * (2 * S + 5 + b); };   // statistics-MMA mode only
  auto sdone_bar = [&](int b) { return bars + 8u * (2 * S + 7 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 11);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.off_bars + 8 * (2 * S + 11));
  const uint32_t sfree_cnt = bars + 8u * (2 * S + 12);   // count of tiles whose TMA store has read its staging buffer
  float* stats_sm = reinterpret_cast<float*>(sm + p.off_stats);
  float* scratch = reinterpret_cast<float*>(sm + p.off_scratch);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (p.CBt) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    for (int i = 0; i < S; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    for (int a = 0; a < 4; ++a) mbar_init(tfull_bar(a), 1);
    for (int a = 0; a < 2; ++a) { mbar_init(sready_bar(a), TC_EPI); mbar_init(sdone_bar(a), 1); }
    mbar_init(wbar, 1);
    st_release_shared(sfree_cnt, 0u);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x >= TC_THREADS - TC_EPI) {
    const int e0 = threadIdx.x - (TC_THREADS - TC_EPI);
    if (p.has_stats && !p.mma_stats && !STATS)
      for (int i = e0; i < 2 * p.dC; i += TC_EPI) stats_sm[i] = 0.f;
    if (p.mma_stats) {   // the all-ones A operand of the column-sum MMA (any canonical layout: every element is 1)
      uint32_t* ones = reinterpret_cast<uint32_t*>(sm + p.off_ones);
      for (int i = e0; i < 512; i += TC_EPI) ones[i] = 0x3F803F80u;
      fence_proxy_async_smem();
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t col_g = (uint32_t)(p.acc_bufs * p.Ntile), col_s = col_g + (uint32_t)p.Ntile;
  // register budget per role: the four control warps (one warpgroup) give registers to the eight epilogue warps, whose
  // statistics accumulators (up to 96 floats) and row prefetch would otherwise spill at the 168 registers of 384 threads
  // (BN-backward variants only -- the tighter budget measurably slows the control warps' issue loops -- setmaxnreg at the top of every role branch: 96 for warps 0-3, 200 for warps 4-11: 4*32*96 + 8*32*200 = 63488 <= the 384*168 registers the CTA is launched with -- an inc beyond the launch-time pool never returns)
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int TL = p.nloads * p.ncblk;   // loads per tile, moved `lps` per pipeline stage

  if (warp == 0) {
    if (STATS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (lane == 0) {
    // ===================== TMA producer =====================
    long long w_prod = 0;
    const long long t_start = dbg ? clock64() : 0;
    const int tailcb = p.CBt ? p.ncblk - 1 : -1;   // index of the narrow tail block (-1: none)
    if (p.w_resident) {
      mbar_expect_tx(wbar, (uint32_t)(p.nloads * p.nsub * ((p.ncblk - (p.CBt ? 1 : 0)) * p.b_box_bytes + (p.CBt ? p.b_box_bytes_t : 0))));
      for (int l = 0; l < p.nloads; ++l)
        for (int s = 0; s < p.nsub; ++s)
          for (int cb = 0; cb < p.ncblk; ++cb)
            tma_load_2d(cb == tailcb ? &tmB2 : &tmB, wbar,
                        sbase + (uint32_t)(p.off_wgt + (l * p.nsub + s) * p.w_row_bytes + cb * p.b_sub_bytes),
                        (p.tap0[l] + s * p.tap_sub_stride) * p.red_C + cb * p.CB, 0);
    }
    const uint32_t tx_w = (uint32_t)(p.w_resident ? 0 : p.nsub * p.b_box_bytes);
    const uint32_t tx = (uint32_t)p.a_box_bytes + tx_w, tx_t = (uint32_t)(p.CBt ? p.a_box_bytes_t : p.a_box_bytes) + tx_w;
    // dual mode: the stages form two private rings, one per MMA warp (tiles alternate), so that no mbarrier is ever
    // waited on by two consumers in different phases (a parity wait cannot tell phase k+1 from phase k-1)
    const int RN = p.dual_mma ? S / 2 : S;
    int rstage0 = 0, rstage1 = 0;
    uint32_t rphase0 = 0, rphase1 = 0;
    int it = 0;
    // loop invariants in registers: this single thread's instruction count per stage is on the critical path of the
    // short-K layers (one 128-pixel tile lasts ~1000 cycles)
    const int lps = p.lps, ncblk = p.ncblk, nsub = p.nsub, CB = p.CB, Ntile = p.Ntile, red_C = p.red_C;
    const int tss = p.tap_sub_stride;
    const bool resident = p.w_resident != 0, dual = p.dual_mma != 0, skip_loads = (p.dbg_skip & 2) != 0;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes, slot_f = (uint32_t)p.slot_bytes;
    const uint32_t slot_t = (uint32_t)(p.unit_mode ? p.slot_bytes_t : p.slot_bytes), b_sub_bytes = (uint32_t)p.b_sub_bytes;
    const int sw = p.bw * p.mw, sh = p.bh * p.mh, st_ = p.bt * p.mt;
    const bool l2_hint = p.l2_hint != 0;
    const uint64_t pol_first = l2_policy_evict_first();
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ti.next(), ++it) {
      const int n_idx = ti.n, b = ti.b;
      const int w0 = ti.w * sw, h0 = ti.h * sh, t0 = ti.t * st_;
      const int r = dual ? (it & 1) : 0;
      int stage = r ? rstage1 : rstage0;
      uint32_t phase = r ? rphase1 : rphase0;
      int l = 0, cb = 0;
      for (int left = TL; left > 0; left -= lps) {
        const int cnt = left < lps ? left : lps;
        const int sidx = r * RN + stage;
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(empty_bar(sidx), phase ^ 1u);
        if (dbg) w_prod += clock64() - c0;
        if (skip_loads) {
          mbar_arrive(full_bar(sidx));
        } else {
          uint32_t sa = sbase + (uint32_t)sidx * stage_bytes;
          uint32_t bytes = 0;
          for (int j = 0; j < cnt; ++j) {
            const bool tail = cb == tailcb;
            if (l2_hint)
              tma_load_5d_hint(tail ? &tmA2 : &tmA, full_bar(sidx), sa, cb * CB, w0 + p.off_w[l], h0 + p.off_h[l], t0 + p.off_t[l], b, pol_first);
            else
              tma_load_5d(tail ? &tmA2 : &tmA, full_bar(sidx), sa, cb * CB, w0 + p.off_w[l], h0 + p.off_h[l], t0 + p.off_t[l], b);
            const uint32_t sz = tail ? slot_t : slot_f;
            bytes += tail ? tx_t : tx;
            if (!resident) {
              for (int s = 0; s < nsub; ++s) {
                const int tap = p.tap0[l] + s * tss;
                tma_load_2d(&tmB, full_bar(sidx), sa + sz - (uint32_t)(nsub - s) * b_sub_bytes, tap * red_C + cb * CB, n_idx * Ntile);
              }
            }
            sa += sz;
            if (++cb == ncblk) { cb = 0; ++l; }
          }
          // armed after the copies are issued: the phase cannot complete before this arrival, and the transaction count
          // may run negative in between
          mbar_expect_tx(full_bar(sidx), bytes);
        }
        if (++stage == RN) { stage = 0; phase ^= 1u; }
      }
      if (r) { rstage1 = stage; rphase1 = phase; } else { rstage0 = stage; rphase0 = phase; }
    }
    if (dbg) { dbg[blockIdx.x * 8 + 0] = w_prod; dbg[blockIdx.x * 8 + 1] = clock64() - t_start; }
    }
  } else if (warp == 1 || (warp == 2 && p.dual_mma)) {
    if (STATS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    // dual mode: warps 1 and 2 take alternate tiles (each owns one TMEM accumulator), so one warp issues MMAs while
    // the other sits in the ~200-cycle mbarrier wait / tcgen05.commit latencies of its own tile
    const int mw = p.dual_mma ? warp - 1 : 0, tstep = p.dual_mma ? 2 : 1;
    const int RN = p.dual_mma ? S / 2 : S, rbase = mw * RN;   // this warp's private ring of stages
    const uint32_t idesc = make_idesc_bf16(128, p.Ntile, 0, 0);
    const uint32_t hi = smem_desc_hi((uint32_t)p.sbo_bytes, (uint32_t)p.layout_type);
    const uint32_t a_sub = (uint32_t)p.sub_row_bytes >> 4, b_sub = (uint32_t)p.b_sub_bytes >> 4;
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, slot16 = (uint32_t)p.slot_bytes >> 4;
    const int tailcb = p.CBt ? p.ncblk - 1 : -1;
    const uint32_t hi_t = smem_desc_hi((uint32_t)p.sbo_bytes_t, (uint32_t)p.layout_t);
    const uint32_t a_sub_t = (uint32_t)p.sub_row_bytes_t >> 4;
    const uint32_t slot16_t = (uint32_t)(p.unit_mode ? p.slot_bytes_t : p.slot_bytes) >> 4;
    const uint32_t nsub_bsub = (uint32_t)p.nsub * ((uint32_t)p.b_sub_bytes >> 4);
    // second 128-row sub-tile of a 256-pixel tile: rows 128.. of the same box, accumulator columns Ntile..
    constexpr bool two = MT == 2;
    const uint32_t a_m1 = (uint32_t)(128 * p.CB * 2) >> 4, a_m1_t = (uint32_t)(128 * p.CBt * 2) >> 4;
    const uint32_t d_m1 = (uint32_t)p.Ntile;
    const uint32_t lo0 = smem_desc_lo(sbase, 16);
    const uint32_t wlo0 = smem_desc_lo(sbase + (uint32_t)p.off_wgt, 16);
    const int ksteps_full = p.CB >> 4, ksteps_last = p.ksteps_last, ncblk = p.ncblk, nsub = p.nsub, lps = p.lps;
    const bool skip_rest = (p.dbg_skip & 4) != 0;
    const bool resident = p.w_resident != 0;
    const uint32_t b_step = resident ? (uint32_t)p.w_row_bytes >> 4 : b_sub;
    const uint32_t w_tap = (uint32_t)nsub * ((uint32_t)p.w_row_bytes >> 4);
    // statistics MMAs: D_g += Y^T Y (diagonal = sum of squares), D_s += 1^T Y (column sums); Y = staged bf16 tile
    const uint32_t st_hi = smem_desc_hi((uint32_t)(8 * p.st_rowbytes), (uint32_t)p.st_layout);
    const uint32_t st_lo0 = smem_desc_lo(sbase + (uint32_t)p.off_staging, (uint32_t)p.st_chunk_bytes);
    const uint32_t st_step = (uint32_t)(16 * p.st_rowbytes) >> 4, st_buf16 = (uint32_t)p.st_buf_bytes >> 4;
    const uint32_t ones_lo = smem_desc_lo(sbase + (uint32_t)p.off_ones, 128), ones_hi = smem_desc_hi(256, 0);
    const uint32_t idesc_g = make_idesc_bf16(p.stat_M, p.Ntile, 1, 1), idesc_s = make_idesc_bf16(64, p.Ntile, 0, 1);
    const bool leader = elect_one();   // the same lane issues every MMA and every commit
    int stage = 0;
    uint32_t phase = 0;
    int it = mw;
    long long w_full = 0, w_te = 0;
    const long long t_start = dbg ? clock64() : 0;
    auto issue_stats = [&](int j) {
      const int buf = j % p.st_bufs;
      mbar_wait(sready_bar(buf), (uint32_t)((j / p.st_bufs) & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t y_lo = st_lo0 + (uint32_t)buf * st_buf16;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t accum = (j > 0 || ks > 0) ? 1u : 0u;
          umma_bf16_lh(tmem_base + col_g, y_lo + ks * st_step, st_hi, y_lo + ks * st_step, st_hi, idesc_g, accum);
          umma_bf16_lh(tmem_base + col_s, ones_lo, ones_hi, y_lo + ks * st_step, st_hi, idesc_s, accum);
        }
        umma_commit(sdone_bar(buf));
      }
      __syncwarp();
    };
    if (resident) { mbar_wait(wbar, 0u); tc_fence_after(); }
    for (; it < my_tiles; it += tstep) {
      const int acc = it & (p.acc_bufs - 1);   // 1, 2 or 4 accumulators
      long long c0 = dbg ? clock64() : 0;
      if (it >= p.acc_bufs) named_bar_sync(BAR_TEMPTY + acc, BAR_HANDOFF);   // epilogue drained this accumulator
      if (dbg) w_te += clock64() - c0;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MT * p.Ntile);
      uint32_t accumulate = 0;
      uint32_t w_lo = wlo0;
      int cb = 0;
      for (int base = 0; base < TL; base += lps) {
        const int cnt = min(lps, TL - base);
        const int sidx = rbase + stage;
        c0 = dbg ? clock64() : 0;
        mbar_wait(full_bar(sidx), phase);
        if (dbg) w_full += clock64() - c0;
        tc_fence_after();
        if (leader) {
          uint32_t a_slot = lo0 + (uint32_t)sidx * stage16;
          uint32_t w_l = w_lo;
          int c = cb;
          bool first = accumulate == 0;
          for (int j = 0; j < cnt; ++j) {
            const bool tail = c == tailcb;
            const int ksteps = (c == ncblk - 1) ? ksteps_last : ksteps_full;
            const uint32_t sz = tail ? slot16_t : slot16;
            const uint32_t a_hi = tail ? hi_t : hi, a_step = tail ? a_sub_t : a_sub;
            const uint32_t b_hi = (tail && resident) ? hi_t : hi;   // streamed weights always arrive as full-width boxes
            uint32_t a_lo = a_slot;
            uint32_t b_lo = resident ? (w_l + (uint32_t)c * b_sub) : (a_slot + sz - nsub_bsub);
            const uint32_t a_m = tail ? a_m1_t : a_m1;
            int s = 0;
            if (first) {   // the tile's first MMA overwrites the accumulator
              umma_bf16_lh(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, 0u);
              if (two) umma_bf16_lh(tmem_d + d_m1, a_lo + a_m, a_hi, b_lo, b_hi, idesc, 0u);
              first = false;
              if (!skip_rest)
                for (int k = 1; k < ksteps; ++k) {
                  umma_bf16_acc(tmem_d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
                  if (two) umma_bf16_acc(tmem_d + d_m1, a_lo + a_m + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
                }
              a_lo += a_step; b_lo += b_step;
              s = 1;
            }
            if (!skip_rest) {
              if (ksteps == 4 && !two) {
                for (; s < nsub; ++s) {
                  umma_bf16_acc(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc);
                  umma_bf16_acc(tmem_d, a_lo + 2u, a_hi, b_lo + 2u, b_hi, idesc);
                  umma_bf16_acc(tmem_d, a_lo + 4u, a_hi, b_lo + 4u, b_hi, idesc);
                  umma_bf16_acc(tmem_d, a_lo + 6u, a_hi, b_lo + 6u, b_hi, idesc);
                  a_lo += a_step; b_lo += b_step;
                }
              } else if (!two) {
                for (; s < nsub; ++s) {
                  for (int k = 0; k < ksteps; ++k) umma_bf16_acc(tmem_d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
                  a_lo += a_step; b_lo += b_step;
                }
              } else {
                for (; s < nsub; ++s) {
                  for (int k = 0; k < ksteps; ++k) {
                    umma_bf16_acc(tmem_d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
                    umma_bf16_acc(tmem_d + d_m1, a_lo + a_m + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
                  }
                  a_lo += a_step; b_lo += b_step;
                }
              }
            }
            a_slot += sz;
            if (++c == ncblk) { c = 0; w_l += w_tap; }
          }
          umma_commit(empty_bar(sidx));  // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        accumulate = 1;
        cb += cnt;
        while (cb >= ncblk) { cb -= ncblk; w_lo += w_tap; }
        if (++stage == RN) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      __syncwarp();
      if (p.mma_stats && it > 0) issue_stats(it - 1);
    }
    if (p.mma_stats && it > 0) issue_stats(it - 1);
    if (dbg && lane == 0 && mw == 0) { dbg[blockIdx.x * 8 + 2] = w_full; dbg[blockIdx.x * 8 + 3] = w_te; dbg[blockIdx.x * 8 + 4] = clock64() - t_start; }
  } else if (warp == 2) {
    if (STATS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");   // idle in single-issuer mode, but part of the warpgroup
  } else if (warp == 3) {
    if (STATS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    // ===================== TMA store warp: staged tile -> global, off the epilogue's critical path ==========
    const int nst = (p.Ntile + p.cw - 1) / p.cw;
    const int R = p.st_bufs;
    int it = 0, buf = 0;
    uint32_t done = 0;   // sub-tiles whose store has read its staging buffer
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it, ti.next()) {
      const bool per_tile = MT == 2 && !p.pub_sub;   // one hand-off for both sub-tiles (slots buf, buf + 1)
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        if (!per_tile || m == 0) named_bar_sync(BAR_SREADY + buf, BAR_HANDOFF);
        if (lane == 0) {
          const uint32_t staging_s = sbase + (uint32_t)(p.off_staging + buf * p.st_buf_bytes);
          const int ow = ti.w * p.bw + (m ? p.sub_ow : 0), oh = ti.h * p.bh + (m ? p.sub_oh : 0), ot = ti.t * p.bt + (m ? p.sub_ot : 0);
          for (int ch = 0; ch < nst && !(p.dbg_skip & 1); ++ch) {
            int col = ti.n * p.Ntile + ch * p.cw;
            const CUtensorMap* dmap = &tmD;
            if (STATS == 0 && p.ncls > 1) {   // stride-parity classes: chunk -> (class view, channel within the class)
              const int k = col / p.cpd;
              col -= k * p.cpd;
              dmap = k == 0 ? &tmD : (k == 1 ? &tmD1 : (k == 2 ? &tmD2 : &tmD3));
            }
            tma_store_5d(dmap, staging_s + (uint32_t)(ch * p.st_chunk_bytes), col, ow, oh, ot, ti.b);
          }
          if (!per_tile || m == MT - 1) {
            tma_store_commit();
            tma_store_wait_read();   // the stores issued so far have finished reading their staging buffers
            // publish "sub-tiles 0..done-1 have left their staging buffers": the epilogue warps poll this counter (a
            // ~30-cycle shared-memory load) instead of meeting at a barrier, so they never wait for each other or for this warp
            done += per_tile ? (uint32_t)MT : 1u;
            st_release_shared(sfree_cnt, done);
          }
        }
        __syncwarp();
        if (++buf == R) buf = 0;
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= 4) {
    if (STATS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    if (STATS == 2 || (STATS == 0 && p.epi_bn)) {   // scale[dC], shift[dC] (global memory: after pdl_wait)
      for (int i = threadIdx.x - (TC_THREADS - TC_EPI); i < 2 * p.dC; i += TC_EPI) stats_sm[i] = bn_ss[i];
      named_bar_sync(1, TC_EPI);
    }
    const bool epi_bn = STATS == 0 && p.epi_bn != 0;
    const float* const ep_sc = stats_sm + 0;          // indexed by absolute destination channel
    const float* const ep_sh = stats_sm + p.dC;
    // ===================== epilogue: TMEM -> bf16 -> swizzled staging tile =====================
    const int et = threadIdx.x - (TC_THREADS - TC_EPI);  // 0..255
    const int q = warp & 3;                              // TMEM lane quadrant this warp may access
    const int e = q * 32 + lane;                         // accumulator row == TMEM lane == pixel of the tile
    const int chalf = (warp - 4) >> 2;                   // which half of the columns this warp drains
    const int csplit = ((p.Ntile >> 4) + 1) / 2 * 16;    // columns [0,csplit) -> half 0, [csplit,Ntile) -> half 1
    const int cbeg = chalf ? csplit : 0, cend = chalf ? p.Ntile : csplit;
    int lw[MT], lh[MT], lt[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int em = e + 128 * m;
      lw[m] = em % p.bw; lh[m] = (em / p.bw) % p.bh; lt[m] = em / (p.bw * p.bh);
    }
    const uint32_t st_mask = (uint32_t)p.st_mask;
    // (statistics outside the epilogue registers exist only in the generic-drain kernel: the planner gives them drain_rs == 0)
    const bool legacy_stats = RS == 0 && p.has_stats && !p.mma_stats && !STATS;
    const bool legacy_keep = legacy_stats && p.n_ntiles == 1 && p.stats_keep != 0;   // statistics of the staged tile kept in registers across tiles
    float lsum[8], lsq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lsum[j] = 0.f; lsq[j] = 0.f; }
    const int st_bufs = p.st_bufs, bw_ = p.bw, bh_ = p.bh, bt_ = p.bt, dW_ = p.dW, dH_ = p.dH, dT_ = p.dT;
    int sbuf = 0;          // ring slot of the next sub-tile to stage
    uint32_t sub_idx = 0;  // sub-tiles staged so far
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    long long* const edbg = STATS == 2 ? nullptr : dbg;   // no role counters in the register-starved BN-backward variant
    long long w_tf = 0, w_a = 0, w_b = 0, w_c = 0, w_d = 0;
    const long long t_start = edbg ? clock64() : 0;
    constexpr int NS = STATS ? RS : 1;
    float ssum[NS][16], ssq[NS][16];   // per-thread (one pixel row) channel sums over all tiles
    if (STATS) {
#pragma unroll
      for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int k = 0; k < 16; ++k) { ssum[j][k] = 0.f; ssq[j][k] = 0.f; }
    }
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int n_idx = ti.n, b = ti.b;
      int buf = sbuf;   // (MT == 1: the tile's only slot; statistics readers below use it)
      uint8_t* staging = sm + p.off_staging + buf * p.st_buf_bytes;
      // per 128-row sub-tile: pixel coordinates, validity, row offset in the destination tensor
      bool valid_m[MT];
      int64_t roff_m[MT];
      const bool cls = STATS == 0 && p.ncls > 1;   // stride-parity classes (MT == 1): the addend origin depends on the column
      int cw0 = 0, ch0 = 0, ct0 = 0;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int w = ti.w * bw_ + lw[m], h = ti.h * bh_ + lh[m], t = ti.t * bt_ + lt[m];
        valid_m[m] = (w < dW_) && (h < dH_) && (t < dT_);
        roff_m[m] = 0;
        if (STATS != 1 && (p.has_addend || STATS == 2) && valid_m[m])
          roff_m[m] = (cls ? 0 : p.a_off + n_idx * p.Ntile) + (int64_t)b * p.a_sb + (int64_t)t * p.a_st + (int64_t)h * p.a_sh + (int64_t)w * p.a_sw;
        if (m == 0) { cw0 = w; ch0 = h; ct0 = t; }
      }
      // addend row of 16 consecutive destination channels starting at tile column c (class mode: nullptr outside the class's extent)
      auto cls_addend = [&](int c) -> const __nv_bfloat16* {
        const int g = n_idx * p.Ntile + c;
        const int k = g / p.cpd;
        if (!(cw0 < p.dWk[k] && ch0 < p.dHk[k] && ct0 < p.dTk[k])) return nullptr;
        return addend + p.a_off_k[k] + roff_m[0] + (g - k * p.cpd);
      };
      const bool valid0 = valid_m[0];
      // STATS == 2: this pixel's row of the producing layer's conv output, fetched before the accumulator wait
      uint4 yq[STATS == 2 ? 2 * RS : 1];
      if (STATS == 2) {
        const int nch = (cend - cbeg) >> 4;
        const __nv_bfloat16* yrow = yprev + roff_m[0] + cbeg;   // (MT == 1 in this mode)
#pragma unroll
        for (int j = 0; j < (STATS == 2 ? RS : 0); ++j) {
          if (j < nch && valid0) {
            yq[2 * j] = *reinterpret_cast<const uint4*>(yrow + 16 * j);
            yq[2 * j + 1] = *reinterpret_cast<const uint4*>(yrow + 16 * j + 8);
          } else {
            yq[2 * j] = make_uint4(0, 0, 0, 0);
            yq[2 * j + 1] = make_uint4(0, 0, 0, 0);
          }
        }
      }

      // addend (shortcut / residual gradient) rows of this pixel, fetched BEFORE the accumulator wait: issued inside the drain
      // their L2 / HBM latency sat on the critical path of every chunk pair (measured: +170 us on the conv3 strided dgrad)
      constexpr bool APF = STATS == 0 && MT == 1 && RS >= 1 && RS <= 4;
      uint4 aq[APF ? 2 * RS : 1];
      const bool apf = APF && p.has_addend != 0;
      if (apf) {
        const int nch = (cend - cbeg) >> 4;
#pragma unroll
        for (int j = 0; j < (APF ? RS : 0); ++j) {
          const __nv_bfloat16* ap = nullptr;
          if (j < nch && valid0) ap = cls ? cls_addend(cbeg + 16 * j) : addend + roff_m[0] + cbeg + 16 * j;
          if (ap != nullptr) {
            aq[2 * j] = *reinterpret_cast<const uint4*>(ap);
            aq[2 * j + 1] = *reinterpret_cast<const uint4*>(ap + 8);
          } else {
            aq[2 * j] = make_uint4(0, 0, 0, 0);
            aq[2 * j + 1] = make_uint4(0, 0, 0, 0);
          }
        }
      }

      const long long c0 = edbg ? clock64() : 0;
      mbar_wait(tfull_bar(acc), acc_phase);
      if (edbg) w_tf += clock64() - c0;
      long long c1 = edbg ? clock64() : 0;
      tc_fence_after();
      if (edbg) { const long long c2 = clock64(); w_a += c2 - c1; c1 = c2; }

      const uint32_t taddr0 = tmem_base + (uint32_t)(acc * MT * p.Ntile) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {   // not unrolled: one copy of the drain code for both sub-tiles
      {
      // this sub-tile's ring slot is free once the TMA store issued st_bufs sub-tiles ago has read it and (statistics
      // MMAs, MT == 1) the MMAs over it have retired
      buf = sbuf;
      staging = sm + p.off_staging + buf * p.st_buf_bytes;
      // (one hand-off per tile: both slots are checked before the first sub-tile and published after the second)
      const bool per_tile = MT == 2 && !p.pub_sub;
      const uint32_t last = per_tile ? sub_idx + (uint32_t)(MT - 1) : sub_idx;   // newest sub-tile about to be overwritten
      if ((!per_tile || m == 0) && last >= (uint32_t)st_bufs) {
        const long long cw0_ = edbg ? clock64() : 0;
        const uint32_t need = last - (uint32_t)st_bufs + 1u;
        while (ld_acquire_shared(sfree_cnt) < need) {}
        if (MT == 1 && p.mma_stats) mbar_wait(sdone_bar(buf), (uint32_t)(((it / st_bufs) - 1) & 1));
        if (edbg) { const long long c2 = clock64(); w_a += c2 - cw0_; c1 += c2 - cw0_; }
      }
      ++sub_idx;
      if (++sbuf == st_bufs) sbuf = 0;
      const bool valid = m ? valid_m[MT - 1] : valid_m[0];
      const __nv_bfloat16* arow = (STATS != 1 && p.has_addend && valid) ? addend + (m ? roff_m[MT - 1] : roff_m[0]) : nullptr;
      const uint32_t row_off = (uint32_t)(e * p.st_rowbytes);
      const uint32_t taddr = taddr0 + (uint32_t)(m * p.Ntile);
      auto store16 = [&](const float* f, int c) {   // 16 consecutive channels starting at channel c of this tile
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h0[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          h1[j] = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
        }
        const int chunk = p.cw_shift >= 0 ? (c >> p.cw_shift) : 0;
        uint32_t off = row_off + (uint32_t)((c - chunk * p.cw) * 2);
        uint8_t* cbase = staging + chunk * p.st_chunk_bytes;
        const uint32_t off0 = off ^ (((off >> 7) & st_mask) << 4);
        off += 16u;
        const uint32_t off1 = off ^ (((off >> 7) & st_mask) << 4);
        *reinterpret_cast<uint4*>(cbase + off0) = o0;
        *reinterpret_cast<uint4*>(cbase + off1) = o1;
      };
      auto emit16 = [&](const uint32_t* v, int c) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.f;
        if (epi_bn) {
          const int cg = n_idx * p.Ntile + c;
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = lrelu(fmaf(f[j], ep_sc[cg + j], ep_sh[cg + j]), p.slope);
        }
        const __nv_bfloat16* ap = arow == nullptr ? nullptr : (cls ? cls_addend(c) : arow + c);
        if (ap != nullptr) {
          const f8 a0 = ld8(ap), a1 = ld8(ap + 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) { f[j] += a0.v[j]; f[8 + j] += a1.v[j]; }
          if (epi_bn) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = lrelu(f[j], p.slope_res);
          }
        }
        store16(f, c);
      };
      if (RS > 0) {
        // <= RS chunks of 16 columns: all TMEM loads in flight at once; statistics from the fp32 accumulators
        if (!(p.dbg_skip & 1)) {
          const int nch = (cend - cbeg) >> 4;
#pragma unroll
          // two chunks in flight per wait (one in the BN-backward mode, whose 96 accumulators + prefetched row leave no
          // registers for a second: a spill in this loop costs far more than the extra TMEM round trip)
          constexpr int STEP = STATS == 2 ? 1 : 2;
#pragma unroll
          for (int j0 = 0; j0 < (RS > 0 ? RS : 1); j0 += STEP) {
            uint32_t v[STEP][16];
#pragma unroll
            for (int jj = 0; jj < STEP; ++jj)
              if (j0 + jj < RS && j0 + jj < nch) tmem_ld16_nowait(taddr + (uint32_t)(cbeg + 16 * (j0 + jj)), v[jj]);
#pragma unroll
            for (int jj = 0; jj < STEP; ++jj) tmem_wait_ld16(v[jj]);
#pragma unroll
            for (int jj = 0; jj < STEP; ++jj) {
              if (j0 + jj < RS && j0 + jj < nch) {
                float f[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) f[k] = valid ? __uint_as_float(v[jj][k]) : 0.f;
                if (STATS == 0 && epi_bn) {
                  const int c = n_idx * p.Ntile + cbeg + 16 * (j0 + jj);   // destination channel
#pragma unroll
                  for (int k = 0; k < 16; ++k) f[k] = lrelu(fmaf(f[k], ep_sc[c + k], ep_sh[c + k]), p.slope);
                }
                if (apf) {   // prefetched rows (zeros where this pixel / class has no addend)
                  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&aq[APF ? 2 * (j0 + jj) : 0]);
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float2 av = __bfloat1622float2(a2[k]);
                    f[2 * k] += av.x; f[2 * k + 1] += av.y;
                  }
                  if (STATS == 0 && epi_bn) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) f[k] = lrelu(f[k], p.slope_res);
                  }
                }
                const __nv_bfloat16* ap = (STATS == 1 || arow == nullptr || apf) ? nullptr
                                          : (cls ? cls_addend(cbeg + 16 * (j0 + jj)) : arow + cbeg + 16 * (j0 + jj));
                if (STATS != 1 && ap != nullptr) {
                  const f8 a0 = ld8(ap), a1 = ld8(ap + 8);
#pragma unroll
                  for (int k = 0; k < 8; ++k) { f[k] += a0.v[k]; f[8 + k] += a1.v[k]; }
                  if (STATS == 0 && epi_bn) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) f[k] = lrelu(f[k], p.slope_res);
                  }
                }
                if (STATS == 1) {
                  const int js = (j0 + jj < NS) ? j0 + jj : 0;
#pragma unroll
                  for (int k = 0; k < 16; ++k) {
                    ssum[js][k] += f[k];
                    ssq[js][k] = fmaf(f[k], f[k], ssq[js][k]);
                  }
                }
                if (STATS == 2) {
                  const int js = (j0 + jj < NS) ? j0 + jj : 0;
                  const int c = cbeg + 16 * (j0 + jj);
                  const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&yq[STATS == 2 ? 2 * js : 0]);
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float2 yv = __bfloat1622float2(y2[k]);
                    float g0 = f[2 * k], g1 = f[2 * k + 1];
                    if (p.slope != 1.f) {
                      const float u0 = fmaf(yv.x, stats_sm[c + 2 * k], stats_sm[p.dC + c + 2 * k]);
                      const float u1 = fmaf(yv.y, stats_sm[c + 2 * k + 1], stats_sm[p.dC + c + 2 * k + 1]);
                      g0 *= u0 > 0.f ? 1.f : p.slope;
                      g1 *= u1 > 0.f ? 1.f : p.slope;
                    }
                    ssum[js][2 * k] += g0;
                    ssum[js][2 * k + 1] += g1;
                    ssq[js][2 * k] = fmaf(g0, yv.x, ssq[js][2 * k]);
                    ssq[js][2 * k + 1] = fmaf(g1, yv.y, ssq[js][2 * k + 1]);
                  }
                }
                store16(f, cbeg + 16 * (j0 + jj));
              }
            }
          }
        }
      } else {
        for (int c0c = cbeg; c0c < cend && !(p.dbg_skip & 1); c0c += 64) {
          // up to 64 columns in flight per wait (column counts are multiples of 16)
          uint32_t va[32], vb[32];
          const int rem = cend - c0c;
          if (rem >= 32) tmem_ld32_nowait(taddr + (uint32_t)c0c, va); else tmem_ld16_nowait(taddr + (uint32_t)c0c, va);
          if (rem >= 64) tmem_ld32_nowait(taddr + (uint32_t)c0c + 32u, vb);
          else if (rem >= 48) tmem_ld16_nowait(taddr + (uint32_t)c0c + 32u, vb);
          tmem_wait_ld16(va); tmem_wait_ld16(va + 16); tmem_wait_ld16(vb); tmem_wait_ld16(vb + 16);
          emit16(va, c0c);
          if (rem >= 32) emit16(va + 16, c0c + 16);
          if (rem >= 48) emit16(vb, c0c + 32);
          if (rem >= 64) emit16(vb + 16, c0c + 48);
        }
      }
      if (MT == 2 && (p.pub_sub || m == MT - 1)) {   // publish this sub-tile, or (narrow tiles) the pair: barrier of the first slot
        fence_proxy_async_smem();
        named_bar_arrive(BAR_SREADY + (p.pub_sub ? buf : buf - 1), BAR_HANDOFF);
      }
      }   // m < MT
      }   // sub-tiles
      if (edbg) { const long long c2 = clock64(); w_b += c2 - c1; c1 = c2; }
      // accumulator drained: hand the TMEM buffer back to the MMA warp; publish the staged tile to the async proxy
      tc_fence_before();
      if (it + p.acc_bufs < my_tiles) named_bar_arrive(BAR_TEMPTY + acc, BAR_HANDOFF);
      if (MT == 1) {
        fence_proxy_async_smem();
        if (legacy_stats) named_bar_sync(1, TC_EPI);
        named_bar_arrive(BAR_SREADY + buf, BAR_HANDOFF);   // the TMA-store warp may read the tile
        if (p.mma_stats) mbar_arrive(sready_bar(buf));     // ... and so may the statistics MMAs
      }
      if (edbg) { const long long c2 = clock64(); w_c += c2 - c1; c1 = c2; }
      if (legacy_stats) {
        // tiles wider than the register / tensor-core statistics cover (N > 128, or several N tiles): per-channel sum / sum of
        // squares of the bf16 tile from the (unswizzled) staging tile.  Thread (row group rg, 8-channel vector cv) keeps its
        // sums in registers ACROSS tiles when every tile covers the same channels (one N tile): the row groups are combined
        // once per CTA after the last tile instead of once per tile (scratch writes with 16-way bank conflicts, a second
        // pass and two more barriers per tile: 4500 of the 7400 cycles a 64 -> 144 tile took, r2g_role_cycles.txt)
        const int row_bytes = p.st_rowbytes;
        const int ncv = p.Ntile >> 3;
        const int nrg = TC_EPI / ncv;
        const int cv = et % ncv, rg = et / ncv;
        if (!legacy_keep) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { lsum[j] = 0.f; lsq[j] = 0.f; }
        }
        if (rg < nrg) {
          for (int row = rg; row < 128; row += nrg) {
            const uint4 u = *reinterpret_cast<const uint4*>(staging + (size_t)row * row_bytes + cv * 16);
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 fv = __bfloat1622float2(hh[j]);
              lsum[2 * j] += fv.x; lsum[2 * j + 1] += fv.y;
              lsq[2 * j] = fmaf(fv.x, fv.x, lsq[2 * j]); lsq[2 * j + 1] = fmaf(fv.y, fv.y, lsq[2 * j + 1]);
            }
          }
        }
        if (!legacy_keep) {
          named_bar_sync(2, TC_EPI);                // previous tile's readers of scratch are done
          if (rg < nrg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              scratch[(rg * p.Ntile + cv * 8 + j) * 2 + 0] = lsum[j];
              scratch[(rg * p.Ntile + cv * 8 + j) * 2 + 1] = lsq[j];
            }
          }
          named_bar_sync(2, TC_EPI);
          for (int c = et; c < p.Ntile; c += TC_EPI) {
            float a0 = 0.f, b0 = 0.f;
            for (int g = 0; g < nrg; ++g) {
              a0 += scratch[(g * p.Ntile + c) * 2 + 0];
              b0 += scratch[(g * p.Ntile + c) * 2 + 1];
            }
            stats_sm[n_idx * p.Ntile + c] += a0;
            stats_sm[p.dC + n_idx * p.Ntile + c] += b0;
          }
        }
        // legacy statistics read the tile in place: the store warp releases the buffer (BAR_SFREE) only after ITS read,
        // and this barrier orders the readers before any thread can start the next drain into it
        named_bar_sync(2, TC_EPI);
      }
      if (edbg) { const long long c2 = clock64(); w_d += c2 - c1; c1 = c2; }
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
    if (edbg && et == 0) {
      edbg[blockIdx.x * 8 + 5] = w_tf; edbg[blockIdx.x * 8 + 6] = clock64() - t_start;
      long long* d2 = edbg + 148 * 8 + blockIdx.x * 8;
      d2[0] = w_a; d2[1] = w_b; d2[2] = w_c; d2[3] = w_d;
    }
    if (STATS) {
      // rows -> channels: butterfly over the 32 pixel rows of the warp, then the 4 row quadrants through smem
      const int nch = (cend - cbeg) >> 4;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float a0 = warp_sum(ssum[j][k]), b0 = warp_sum(ssq[j][k]);
          if (lane == 0 && j < nch) {
            scratch[(q * p.Ntile + cbeg + 16 * j + k) * 2 + 0] = a0;
            scratch[(q * p.Ntile + cbeg + 16 * j + k) * 2 + 1] = b0;
          }
        }
      }
      named_bar_sync(1, TC_EPI);
      float* prow = part + (int64_t)blockIdx.x * 2 * p.pC + p.n_base;
      for (int c = et; c < p.Ntile; c += TC_EPI) {
        float a0 = 0.f, b0 = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) { a0 += scratch[(g * p.Ntile + c) * 2 + 0]; b0 += scratch[(g * p.Ntile + c) * 2 + 1]; }
        prow[c] = a0;
        prow[p.pC + c] = b0;
      }
    }
    if (legacy_keep) {   // combine the row groups once: scratch[rg][channel][2] -> stats_sm (zero so far)
      const int ncv = p.Ntile >> 3;
      const int nrg = TC_EPI / ncv;
      const int cv = et % ncv, rg = et / ncv;
      if (rg < nrg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          scratch[(rg * p.Ntile + cv * 8 + j) * 2 + 0] = lsum[j];
          scratch[(rg * p.Ntile + cv * 8 + j) * 2 + 1] = lsq[j];
        }
      }
      named_bar_sync(2, TC_EPI);
      for (int c = et; c < p.Ntile; c += TC_EPI) {
        float a0 = 0.f, b0 = 0.f;
        for (int g = 0; g < nrg; ++g) {
          a0 += scratch[(g * p.Ntile + c) * 2 + 0];
          b0 += scratch[(g * p.Ntile + c) * 2 + 1];
        }
        stats_sm[c] = a0;
        stats_sm[p.dC + c] = b0;
      }
    }
    if (legacy_stats) {
      named_bar_sync(1, TC_EPI);
      for (int i = et; i < 2 * p.dC; i += TC_EPI)
        part[(int64_t)blockIdx.x * 2 * p.pC + p.n_base + (i < p.dC ? i : p.pC + i - p.dC)] = stats_sm[i];
    }
    if (p.mma_stats && chalf == 0) {
      // all statistics MMAs of this CTA have retired once the last staged tile's are done
      const int last = it - 1;
      mbar_wait(sdone_bar(last % p.st_bufs), (uint32_t)((last / p.st_bufs) & 1));
      tc_fence_after();
      float* prow = part + (int64_t)blockIdx.x * 2 * p.pC + p.n_base;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      if (q == 0) {   // column sums: every row of D_s is the same, row 0 lives in TMEM lane 0
        for (int n0 = 0; n0 < p.Ntile; n0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + col_s + (uint32_t)n0, v);
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) prow[n0 + j] = __uint_as_float(v[j]);
          }
        }
      }
      // sums of squares: diagonal of D_g.  M=128: channel m in lane m; M=64: channel i in lane (i%16) + 32*(i/16)
      const int per_warp = p.stat_M == 128 ? 32 : 16;
      const int cbase = q * per_warp;
      if (cbase < p.Ntile) {
        float val = 0.f;
        for (int n0 = 0; n0 < per_warp; n0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + lane_base + col_g + (uint32_t)(cbase + n0), v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (lane == n0 + j) val = __uint_as_float(v[j]);
        }
        if (lane < per_warp && cbase + lane < p.Ntile) prow[p.pC + cbase + lane] = val;
      }
    }
  }

  if (fin.kind) __threadfence();   // this CTA's row of `part` is visible device-wide before its ticket is taken
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  // last CTA done: BatchNorm finalisation of the partials in this launch (no stand-alone finalize kernel); the pipeline
  // stages are dead by now and lend their shared memory
  if (fin.kind) bn_fin_tail(fin, part, reinterpret_cast<int*>(sm + 4096), reinterpret_cast<double*>(sm));
}

// ------------------------------------------------------------------------------------------------
// host side: planning + tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
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

static int g_opt_halo = 1, g_opt_strided = 1, g_opt_max_stages = 8, g_opt_tc = 1, g_opt_mma_stats = 1, g_opt_resident = 1, g_opt_chunked = 1, g_opt_st_bufs = 2, g_opt_dbg_skip = 0, g_opt_lps_max = 16, g_opt_reg_stats = 1, g_opt_dual = 1, g_opt_tail = 1, g_opt_acc4 = 1, g_opt_bwd_stats_max = 64, g_opt_mt = 0, g_opt_classes = 1, g_opt_pub_sub_min = 64, g_opt_fine_n = 48, g_opt_l2hint = 0, g_opt_nsplit = 0, g_opt_stats_keep = 1;
int tc_option(const char* name, int value, bool set) {
  int* slot = nullptr;
  if (!strcmp(name, "tc_halo")) slot = &g_opt_halo;
  else if (!strcmp(name, "tc_strided")) slot = &g_opt_strided;
  else if (!strcmp(name, "tc_max_stages")) slot = &g_opt_max_stages;
  else if (!strcmp(name, "tc_enable")) slot = &g_opt_tc;
  else if (!strcmp(name, "tc_mma_stats")) slot = &g_opt_mma_stats;
  else if (!strcmp(name, "tc_resident")) slot = &g_opt_resident;
  else if (!strcmp(name, "tc_chunked")) slot = &g_opt_chunked;
  else if (!strcmp(name, "tc_st_bufs")) slot = &g_opt_st_bufs;
  else if (!strcmp(name, "tc_dbg_skip")) slot = &g_opt_dbg_skip;
  else if (!strcmp(name, "tc_nsplit")) slot = &g_opt_nsplit;   // forward: output channels in two launches when that makes the weights resident
  else if (!strcmp(name, "tc_l2hint")) slot = &g_opt_l2hint;
  else if (!strcmp(name, "tc_stats_keep")) slot = &g_opt_stats_keep;   // evict-first loads of the activation operand
  else if (!strcmp(name, "tc_lps_max")) slot = &g_opt_lps_max;
  else if (!strcmp(name, "tc_reg_stats")) slot = &g_opt_reg_stats;
  else if (!strcmp(name, "tc_dual_mma")) slot = &g_opt_dual;
  else if (!strcmp(name, "tc_tail")) slot = &g_opt_tail;
  else if (!strcmp(name, "tc_acc4")) slot = &g_opt_acc4;
  else if (!strcmp(name, "tc_bwd_stats_max")) slot = &g_opt_bwd_stats_max;
  else if (!strcmp(name, "tc_mt")) slot = &g_opt_mt;
  else if (!strcmp(name, "tc_classes")) slot = &g_opt_classes;
  else if (!strcmp(name, "tc_fine_n")) slot = &g_opt_fine_n;   // finest dual-issuer stages up to this many output channels
  else if (!strcmp(name, "tc_pub_sub_min")) slot = &g_opt_pub_sub_min;   // 256-pixel tiles: per-sub-tile hand-off from this many channels
  if (slot == nullptr) return -1;
  if (set) *slot = value;
  return *slot;
}

// Role-agnostic problem: forward or stride-1 dgrad.
struct GatherProblem {
  int B;
  int sT, sH, sW, sC;
  int dT, dH, dW, dC;   // destination VIEW dims (what the kernel tiles over)
  int kt, kh, kw;       // tap positions visited per dim
  int Kt, Kh, Kw;       // full kernel dims (weight layout [N][Kt*Kh*Kw][C])
  int mt, mh, mw;
  int ot, oh, ow;       // source coordinate offset of tap position j = 0 (position j reads offset o + j)
  int ts_t, ts_h, ts_w; // weight tap used by position j is ts + tp * j
  int tp_t, tp_h, tp_w;
  // destination view inside the full tensor (FT,FH,FW): origin and step per dim (strided dgrad parity classes)
  int FT, FH, FW;
  int vo_t, vo_h, vo_w, vs_t, vs_h, vs_w;
  // optional source view: element strides of (w,h,t,b); 0 = dense NDHWC (the packed stem rows overlap in w)
  long long ss_w, ss_h, ss_t, ss_b;
  // output-channel split (forward): this launch produces channels [n0, n0 + dC) of a destination with dCf channels
  // (dCf == 0: the destination has dC channels); weights and statistics partials are offset the same way
  int dCf = 0, n0 = 0;
};

struct TcPlan {
  TcParams p;
  int grid;
  size_t smem;
  int a_box[5];
  int a_estride[5];
};

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

static bool plan_gather_mt(const GatherProblem& g, bool has_stats, TcPlan* out, bool bwd_stats, int MT, int cls_cpd = 0) {
  const int PT = 128 * MT;   // pixels per tile
  if (!g_opt_tc) return false;
  const int taps = g.kt * g.kh * g.kw;
  if (taps > TC_MAX_LOADS || taps < 1) return false;
  if (g.sC % 16 || g.dC % 16) return false;
  const bool strided = (g.mt != 1 || g.mh != 1 || g.mw != 1);
  if (strided && !g_opt_strided) return false;
  if (g.mt > 8 || g.mh > 8 || g.mw > 8) return false;

  TcParams p;
  memset(&p, 0, sizeof(p));
  // N tiling
  p.MT = MT;
  p.n_ntiles = (g.dC + 255) / 256;
  if (g.dC % (16 * p.n_ntiles)) return false;
  p.Ntile = g.dC / p.n_ntiles;
  if (cls_cpd && (MT != 1 || has_stats || p.Ntile % cls_cpd)) return false;   // an N tile holds whole stride-parity classes
  // K blocking
  p.CB = g.sC <= 16 ? 16 : (g.sC <= 32 ? 32 : 64);
  p.ncblk = (g.sC + p.CB - 1) / p.CB;
  p.ksteps_last = (g.sC - (p.ncblk - 1) * p.CB) / 16;
  p.layout_type = p.CB == 64 ? 2 : (p.CB == 32 ? 4 : 6);
  {
    const int rem = g.sC - (p.ncblk - 1) * p.CB;
    p.CBt = (g_opt_tail && p.ncblk > 1 && rem <= 32) ? (rem <= 16 ? 16 : 32) : 0;
    p.layout_t = p.CBt == 32 ? 4 : 6;
    p.sbo_bytes_t = 8 * p.CBt * 2;
  }
  const int rowbytes = p.CB * 2;
  p.sbo_bytes = 8 * rowbytes;
  p.red_C = g.sC;
  p.b_box_bytes = p.Ntile * rowbytes;
  p.b_sub_bytes = round_up(p.b_box_bytes, 1024);
  p.b_box_bytes_t = p.Ntile * p.CBt * 2;
  p.b_sub_bytes_t = round_up(p.b_box_bytes_t, 1024);
  p.w_row_bytes = (p.ncblk - (p.CBt ? 1 : 0)) * p.b_sub_bytes + (p.CBt ? p.b_sub_bytes_t : 0);
  // staging layout / statistics mode / resident weights
  const int used_taps = taps;
  p.has_stats = has_stats ? 1 : 0;
  // statistics: in epilogue registers from the fp32 accumulators (<= 3 chunks of 16 channels per epilogue warp), else
  // on the tensor core (Gram + ones MMAs over the staged bf16 tile), else from the staged tile on the CUDA cores
  p.reg_stats = (has_stats && (g_opt_reg_stats || bwd_stats) && p.n_ntiles == 1 && p.Ntile <= 96) ? 1 : 0;
  // the BN-backward sums exist only in the register mode, and only pay off up to 64 destination channels: each epilogue
  // thread reads its own pixel row of the producer's output (one 32-byte sector per lane and load), which at 80 channels
  // costs more LSU wavefronts than the stand-alone reduction pass saves (measured: 1229 us fused vs 310 + 350 us).
  // Bringing the producer's tile in through TMA instead (ring of 4 tiles in the staging layout, read back from smem) was
  // tried and is slower still at both 32 channels (889 vs 543 us) and 80 (952 vs 603 us composed): the 64-byte-row boxes
  // and the extra per-tile work land on the single producer thread, which these short-K layers cannot spare.
  if (bwd_stats && (!p.reg_stats || p.Ntile > g_opt_bwd_stats_max)) return false;
  p.drain_rs = (!has_stats || p.reg_stats) ? ((p.Ntile >> 4) + 1) / 2 : 0;
  p.mma_stats = (has_stats && !p.reg_stats && g_opt_mma_stats && p.n_ntiles == 1 && p.Ntile <= 128) ? 1 : 0;
  const bool chunked = p.n_ntiles == 1 && (p.mma_stats || ((!has_stats || p.reg_stats) && g_opt_chunked));
  if (cls_cpd && !chunked && p.Ntile != cls_cpd) return false;   // one unswizzled chunk = the whole tile = one class
  if (chunked) {
    p.cw = p.Ntile > 32 ? 64 : (p.Ntile == 32 ? 32 : 16);
    if (cls_cpd) p.cw = cls_cpd % 64 == 0 ? 64 : (cls_cpd % 32 == 0 ? 32 : 16);   // a staging chunk never straddles two classes
    p.st_mask = p.cw == 64 ? 7 : (p.cw == 32 ? 3 : 1);
    p.st_layout = p.cw == 64 ? 2 : (p.cw == 32 ? 4 : 6);
  } else {
    p.cw = p.Ntile; p.st_mask = 0; p.st_layout = 0;
  }
  p.cw_shift = chunked ? (p.cw == 64 ? 6 : (p.cw == 32 ? 5 : 4)) : -1;
  p.stat_M = p.Ntile <= 64 ? 64 : 128;
  p.st_rowbytes = p.cw * 2;
  p.st_chunk_bytes = round_up(128 * p.st_rowbytes, 1024);   // one staging buffer = one 128-row sub-tile
  p.st_chunks = (p.Ntile + p.cw - 1) / p.cw;
  if (p.mma_stats && p.st_chunks < p.stat_M / p.cw) p.st_chunks = p.stat_M / p.cw;   // the Gram A operand spans stat_M channels
  p.st_buf_bytes = p.st_chunks * p.st_chunk_bytes;
  // TMEM accumulators: the commit -> epilogue wake-up -> drain -> release -> next MMA round trip is ~1200 cycles, longer than a
  // short-K tile, so narrow tiles keep four accumulators in flight
  p.acc_bufs = (p.mma_stats && 4 * p.Ntile > 512) ? 1 : ((!p.mma_stats && g_opt_acc4 && 4 * MT * p.Ntile <= 512) ? 4 : 2);
  if (MT == 2 && (p.n_ntiles != 1 || p.mma_stats || (has_stats && !p.reg_stats) || bwd_stats || 2 * MT * p.Ntile > 512)) return false;
  const int wgt_total = used_taps * p.w_row_bytes;
  // class-packed weights of a strided data gradient (4 classes x 4 positions x 128 x 32 channels = 128 KB) stay resident too:
  // streamed per stage they cost more shared-memory fill than the activations
  p.w_resident = (g_opt_resident && p.n_ntiles == 1 && wgt_total <= (cls_cpd ? 131072 : 98304)) ? 1 : 0;
  const int stats_bytes = round_up(2 * g.dC * 4, 16);
  // the all-ones operand exists only for the statistics MMAs, the reduction scratch only with statistics
  // reduction scratch: the register statistics meet once at the end (4 row quadrants x Ntile x 2 floats); the staged-tile
  // statistics keep 256 / (Ntile / 8) row groups per tile (16 KB); the statistics MMAs need none
  const int ones_bytes = p.mma_stats ? 2048 : 0;
  const int scratch_bytes = !has_stats ? 0 : (p.reg_stats ? round_up(32 * p.Ntile, 1024) : (p.mma_stats ? 0 : 16384));
  const int misc = ones_bytes + stats_bytes + scratch_bytes + 256 /*barriers*/ + 1024 /*alignment*/;
  // ring of 128-row staging buffers: at least one per sub-tile of a tile; up to twice that (search below)
  p.pub_sub = (MT == 2 && p.Ntile >= g_opt_pub_sub_min) ? 1 : 0;
  const int st_min = MT, st_max = g_opt_st_bufs == 1 ? MT : 2 * MT;
  p.st_bufs = st_min;
  auto fixed_bytes = [&]() { return p.st_bufs * p.st_buf_bytes + (p.w_resident ? wgt_total : 0) + misc; };
  const int fixed = fixed_bytes();   // tile search budget with the smallest ring (refined below)

  // tile / mode search
  double best = 1e30;
  int best_mode = -1, best_bw = 0, best_bh = 0, best_bt = 0;
  for (int relaxed = 0; relaxed < 2 && best_mode < 0; ++relaxed)
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 1 && !(g_opt_halo && !strided && g.kt == 1 && g.kh > 1)) continue;
    if (mode == 2 && !(g_opt_halo && !strided && g.kh == 1 && g.kw == 1 && g.kt > 1)) continue;
    for (int bw = 1; bw <= PT; bw <<= 1) {
      for (int bh = 1; bh * bw <= PT; bh <<= 1) {
        const int bt = PT / (bw * bh);
        if (mode == 1 && (bt != 1 || bw % 8)) continue;
        if (mode == 2 && ((bw * bh) % 8)) continue;
        if (bw * g.mw > 256 || bh * g.mh > 256 || bt * g.mt > 256) continue;
        // do not take tiles that are mostly outside the tensor
        if (!relaxed) {   // tiny tensors (unit tests) may need tiles that hang over the edges
          if (bw > 2 * g.dW && bw > 8) continue;
          if (bh >= 2 * g.dH && bh > 1) continue;
          if (bt >= 2 * g.dT && bt > 1) continue;
        }
        int rows_l = PT, nloads = taps, nsub = 1;
        if (mode == 1) { rows_l = (bh + g.kh - 1) * bw; nloads = g.kw; nsub = g.kh; if (bh + g.kh - 1 > 256) continue; }
        if (mode == 2) { rows_l = (bt + g.kt - 1) * bh * bw; nloads = 1; nsub = g.kt; if (bt + g.kt - 1 > 256) continue; }
        const int stage = round_up(rows_l * rowbytes, 1024) + (p.w_resident ? 0 : nsub * p.b_sub_bytes);
        if (fixed + 2 * stage > TC_SMEM_MAX) continue;
        const double ntiles = (double)((g.dW + bw - 1) / bw) * ((g.dH + bh - 1) / bh) * ((g.dT + bt - 1) / bt);
        const double cost = ntiles * ((double)nloads * p.ncblk * (rows_l * rowbytes + nsub * p.b_box_bytes) +
                                      0.15 * taps * (double)PT * g.sC * 2.0) - 1e-3 * bw;
        if (cost < best) { best = cost; best_mode = mode; best_bw = bw; best_bh = bh; best_bt = bt; }
      }
    }
  }
  if (best_mode < 0) return false;
  p.bw = best_bw; p.bh = best_bh; p.bt = best_bt;
  p.B = g.B;
  p.ntile_w = (g.dW + p.bw - 1) / p.bw;
  p.ntile_h = (g.dH + p.bh - 1) / p.bh;
  p.ntile_t = (g.dT + p.bt - 1) / p.bt;
  const int64_t nt = (int64_t)g.B * p.ntile_w * p.ntile_h * p.ntile_t * p.n_ntiles;
  if (nt > 0x7fffffff) return false;
  p.num_tiles = (int)nt;
  p.dW = g.dW; p.dH = g.dH; p.dT = g.dT; p.dC = g.dC;
  p.mw = g.mw; p.mh = g.mh; p.mt = g.mt;
  // second 128-row sub-tile of a 256-pixel tile (pixel order: w fastest, then h, then t): whole t slices, else h rows
  p.sub_ow = p.sub_oh = p.sub_ot = 0;
  if (MT == 2) {
    if (p.bt > 1) p.sub_ot = p.bt / 2;
    else if (p.bh > 1) p.sub_oh = p.bh / 2;
    else p.sub_ow = p.bw / 2;
  }

  int rows_l = PT;
  out->a_estride[0] = 1; out->a_estride[1] = g.mw; out->a_estride[2] = g.mh; out->a_estride[3] = g.mt; out->a_estride[4] = 1;
  out->a_box[0] = p.CB; out->a_box[4] = 1;
  auto tap_index = [&](int jt, int jh, int jw) {
    jt = g.ts_t + g.tp_t * jt; jh = g.ts_h + g.tp_h * jh; jw = g.ts_w + g.tp_w * jw;
    return (jt * g.Kh + jh) * g.Kw + jw;
  };
  if (best_mode == 0) {
    p.nloads = taps; p.nsub = 1; p.sub_row_bytes = 0; p.tap_sub_stride = 0;
    int l = 0;
    for (int jt = 0; jt < g.kt; ++jt)
      for (int jh = 0; jh < g.kh; ++jh)
        for (int jw = 0; jw < g.kw; ++jw, ++l) {
          p.off_t[l] = (signed char)(g.ot + jt); p.off_h[l] = (signed char)(g.oh + jh); p.off_w[l] = (signed char)(g.ow + jw);
          p.tap0[l] = (short)tap_index(jt, jh, jw);
        }
    out->a_box[1] = p.bw * g.mw; out->a_box[2] = p.bh * g.mh; out->a_box[3] = p.bt * g.mt;
  } else if (best_mode == 1) {
    p.nloads = g.kw; p.nsub = g.kh; p.sub_row_bytes = p.bw * rowbytes;
    p.tap_sub_stride = tap_index(0, 1, 0) - tap_index(0, 0, 0);
    for (int jw = 0; jw < g.kw; ++jw) {
      p.off_t[jw] = (signed char)g.ot; p.off_h[jw] = (signed char)g.oh; p.off_w[jw] = (signed char)(g.ow + jw);
      p.tap0[jw] = (short)tap_index(0, 0, jw);
    }
    rows_l = (p.bh + g.kh - 1) * p.bw;
    out->a_box[1] = p.bw; out->a_box[2] = p.bh + g.kh - 1; out->a_box[3] = 1;
  } else {
    p.nloads = 1; p.nsub = g.kt; p.sub_row_bytes = p.bw * p.bh * rowbytes;
    p.tap_sub_stride = tap_index(1, 0, 0) - tap_index(0, 0, 0);
    p.off_t[0] = (signed char)g.ot; p.off_h[0] = (signed char)g.oh; p.off_w[0] = (signed char)g.ow;
    p.tap0[0] = (short)tap_index(0, 0, 0);
    rows_l = (p.bt + g.kt - 1) * p.bh * p.bw;
    out->a_box[1] = p.bw; out->a_box[2] = p.bh; out->a_box[3] = p.bt + g.kt - 1;
  }
  p.a_box_bytes = rows_l * rowbytes;
  p.a_box_bytes_t = rows_l * p.CBt * 2;
  p.sub_row_bytes_t = p.sub_row_bytes / rowbytes * p.CBt * 2;
  auto set_slots = [&]() {
    const int wb = p.w_resident ? 0 : p.nsub * p.b_sub_bytes;
    p.slot_bytes = round_up(p.a_box_bytes, 1024) + wb;
    p.slot_bytes_t = p.CBt ? round_up(p.a_box_bytes_t, 1024) + wb : p.slot_bytes;
  };
  set_slots();
  // pipeline shape: `lps` loads per stage (one mbarrier round trip covers all of them).  When lps is a multiple of the
  // block count a stage holds whole load units and the narrow tail block also takes a narrow slot.
  const int total_loads = p.nloads * p.ncblk;
  auto stage_bytes_for = [&](int lps) {
    return (lps % p.ncblk == 0) ? (lps / p.ncblk) * ((p.ncblk - 1) * p.slot_bytes + p.slot_bytes_t) : lps * p.slot_bytes;
  };
  int best_lps = 0, best_stages = 0;
  long best_score = -1;
  auto search = [&]() {
    best_score = -1;
    best_lps = 0;
    const int avail = TC_SMEM_MAX - fixed_bytes();
    for (int lps = 1; lps <= total_loads && lps <= (g_opt_lps_max < 1 ? 1 : g_opt_lps_max); ++lps) {
      int st = avail / stage_bytes_for(lps);
      if (st > g_opt_max_stages) st = g_opt_max_stages;
      if (st > 8) st = 8;
      if (st < 2) continue;
      const int groups = (total_loads + lps - 1) / lps;
      const bool dual_ok = g_opt_dual && p.acc_bufs >= 2 && !p.mma_stats && st / 2 >= groups;
      // within a tier a stage normally carries as many loads as fit (one mbarrier round per tile); the narrow, operand-bound
      // tiles (N <= 48: 32 -> 80 channel dgrad, 72 -> 32 temporal conv) are load-latency bound instead and run 5-8 % faster
      // with the FINEST stages that keep both issuers, because a slice is refilled as soon as its own MMAs retire
      // (profiles/r2/r2p_layer_sweep.txt: 527 -> 487 us, 251 -> 239 us; the 80-channel forward loses 10 % the same way)
      const long gran = (dual_ok && p.Ntile <= g_opt_fine_n) ? 100L * (17 - lps) : 100L * lps;
      const long score = (dual_ok ? 1000000L : 0L) + ((st >= 3 || dual_ok) ? 100000L : 0L) + gran + st;
      if (score > best_score) { best_score = score; best_lps = lps; best_stages = st; }
    }
  };
  // staging ring depth: start from the deepest ring and give slots back while that buys a better tier (two MMA warps >
  // three stages > the rest): a second buffer per sub-tile is worth less than two issuers or pipeline depth, but the ring
  // never goes below one slot per sub-tile, and 256-pixel tiles keep a third slot when it costs no tier
  {
    int keep_lps = 0, keep_st = 0, keep_bufs = 0;
    long keep_score = -1;
    for (int r = st_max; r >= st_min; --r) {
      if (MT == 2 && !p.pub_sub && (r & 1)) continue;   // one hand-off per tile: the two slots of a tile are a pair
      p.st_bufs = r;
      search();
      if (best_lps == 0) continue;
      if (keep_lps == 0 || best_score / 100000L > keep_score / 100000L) {
        keep_lps = best_lps; keep_st = best_stages; keep_bufs = r; keep_score = best_score;
      }
      if (keep_score >= 1000000L) break;   // two issuers already: deeper rings were tried first
    }
    if (keep_lps != 0) { p.st_bufs = keep_bufs; best_lps = keep_lps; best_stages = keep_st; best_score = keep_score; }
    else { p.st_bufs = st_min; best_lps = 0; }
  }
  if (best_lps == 0 && p.w_resident) {
    p.w_resident = 0;
    set_slots();
    search();
  }
  if (best_lps == 0) return false;
  const int stages = best_stages;
  p.lps = best_lps;
  p.unit_mode = (p.lps % p.ncblk == 0) ? 1 : 0;
  p.stage_bytes = stage_bytes_for(p.lps);
  p.num_stages = stages;
  p.off_wgt = stages * p.stage_bytes;
  p.off_staging = p.off_wgt + (p.w_resident ? wgt_total : 0);
  p.off_ones = p.off_staging + p.st_bufs * p.st_buf_bytes;
  p.off_stats = p.off_ones + ones_bytes;
  p.off_scratch = p.off_stats + stats_bytes;
  p.off_bars = p.off_scratch + scratch_bytes;
  int cols = 32;
  const int need_cols = (p.acc_bufs * MT + (p.mma_stats ? 2 : 0)) * p.Ntile;
  while (cols < need_cols) cols <<= 1;
  if (cols > 512) return false;
  p.tmem_cols = cols;
  // two MMA warps only when each private ring can hold a whole tile's stages (else the split rings lose prefetch depth)
  const int groups = (p.nloads * p.ncblk + p.lps - 1) / p.lps;
  p.dual_mma = (g_opt_dual && p.acc_bufs >= 2 && !p.mma_stats && p.num_stages / 2 >= groups) ? 1 : 0;
  out->p = p;
  out->smem = (size_t)p.off_bars + 256 + 1024;
  const int sms = num_sms();
  out->grid = p.num_tiles < sms ? p.num_tiles : sms;
  return out->smem <= (size_t)TC_SMEM_MAX;
}

// 256-pixel tiles (two 128-row MMA sub-tiles per handshake round) when they keep both MMA warps busy and leave enough
// tiles per CTA; else the 128-pixel plan
static bool plan_gather(const GatherProblem& g, bool has_stats, TcPlan* out, bool bwd_stats = false, int cls_cpd = 0) {
  const bool ok1 = plan_gather_mt(g, has_stats, out, bwd_stats, 1, cls_cpd);
  if (g_opt_mt == 1 || cls_cpd) return ok1;
  TcPlan big;
  if (!plan_gather_mt(g, has_stats, &big, bwd_stats, 2)) return ok1;
  // measured rule (scripts/role_variants.py): the larger tile wins unless it costs the second MMA-issuing warp
  // (or keeps a deep pipeline: >= 5 single-load stages)
  const bool take = g_opt_mt == 2 || !ok1 ||
                    (big.p.num_tiles >= 8 * num_sms() && (big.p.dual_mma >= out->p.dual_mma || big.p.num_stages >= 5));
  if (take) *out = big;
  return true;
}

static int encode_act_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, const int* box,
                          const int* estride, CUtensorMapSwizzle sw, const long long* vstr = nullptr) {
  PFN_encodeTiled enc = get_encode();
  DP_REQUIRE(enc != nullptr, DP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)T * H * W * C * 2};
  if (vstr != nullptr && vstr[0] > 0)
    for (int i = 0; i < 4; ++i) strides[i] = (cuuint64_t)vstr[i] * 2;
  cuuint32_t b[5], es[5];
  for (int i = 0; i < 5; ++i) { b[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)estride[i]; }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DP_REQUIRE(r == CUDA_SUCCESS, DP_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: CUresult %d (box %d,%d,%d,%d)",
             (int)r, box[0], box[1], box[2], box[3]);
  return DP_OK;
}

// destination view: explicit element strides (strided-dgrad parity classes write every s-th pixel)
static int encode_view_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, long long sw_, long long sh_,
                           long long st_, long long sb_, const int* box, int layout) {
  PFN_encodeTiled enc = get_encode();
  DP_REQUIRE(enc != nullptr, DP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)sw_ * 2, (cuuint64_t)sh_ * 2, (cuuint64_t)st_ * 2, (cuuint64_t)sb_ * 2};
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) b[i] = (cuuint32_t)box[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   layout == 2 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (layout == 4 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : (layout == 6 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DP_REQUIRE(r == CUDA_SUCCESS, DP_ERR_CUDA, "cuTensorMapEncodeTiled(destination view) failed: CUresult %d", (int)r);
  return DP_OK;
}

static int encode_wgt_map(CUtensorMap* m, const void* ptr, int Ktot, int rows, int boxK, int boxN,
                          CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode();
  DP_REQUIRE(enc != nullptr, DP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t b[2] = {(cuuint32_t)boxK, (cuuint32_t)boxN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DP_REQUIRE(r == CUDA_SUCCESS, DP_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
  return DP_OK;
}

// All stride-parity classes of a strided data gradient as ONE stride-1 gather over dy: position (pt,ph,pw) of the combined
// kernel reads dy at offset omin + p; its weight for class r is w[.., j] with j = r + pad - s * (omin + p) (zero block when j
// falls outside the kernel); destination columns are [class][Cp] (dgrad_class() is the one-class-per-launch form).
struct ClassGeom {
  int ncls;
  int r[8][3];      // (rt, rh, rw) of class k = (rt * sh + rh) * sw + rw
  int vd[8][3];     // extent of the class's view of dx
  int no[3], omin[3], maxd[3];
};

static bool class_geom(const dp_conv_desc* d, ClassGeom* cg) {
  const int in[3] = {d->Ti, d->Hi, d->Wi}, k[3] = {d->kt, d->kh, d->kw}, st[3] = {d->st, d->sh, d->sw}, pad[3] = {d->pt, d->ph, d->pw};
  cg->ncls = st[0] * st[1] * st[2];
  if (cg->ncls < 2 || cg->ncls > 8) return false;
  for (int i = 0; i < 3; ++i) {
    if (in[i] < st[i]) return false;          // every class must own at least one pixel
    int lo = 1 << 20, hi = -(1 << 20);
    for (int r = 0; r < st[i]; ++r) {
      const int j0 = (r + pad[i]) % st[i];
      for (int j = j0; j < k[i]; j += st[i]) {
        const int o = (r + pad[i] - j) / st[i];
        lo = o < lo ? o : lo; hi = o > hi ? o : hi;
      }
    }
    if (hi < lo) { lo = 0; hi = 0; }          // no class has a tap along this dim (cannot happen for k >= 1, s <= k + pad)
    cg->omin[i] = lo; cg->no[i] = hi - lo + 1;
    cg->maxd[i] = (in[i] + st[i] - 1) / st[i];
  }
  for (int rt = 0; rt < st[0]; ++rt)
    for (int rh = 0; rh < st[1]; ++rh)
      for (int rw = 0; rw < st[2]; ++rw) {
        const int kk = (rt * st[1] + rh) * st[2] + rw;
        const int r[3] = {rt, rh, rw};
        for (int i = 0; i < 3; ++i) { cg->r[kk][i] = r[i]; cg->vd[kk][i] = (in[i] - r[i] + st[i] - 1) / st[i]; }
      }
  return true;
}

static GatherProblem class_problem(const dp_conv_desc* d, const ClassGeom& cg) {
  GatherProblem g;
  g.B = d->B;
  g.sT = d->To; g.sH = d->Ho; g.sW = d->Wo; g.sC = d->Kp;
  g.dT = cg.maxd[0]; g.dH = cg.maxd[1]; g.dW = cg.maxd[2]; g.dC = cg.ncls * d->Cp;
  g.kt = cg.no[0]; g.kh = cg.no[1]; g.kw = cg.no[2];
  g.Kt = cg.no[0]; g.Kh = cg.no[1]; g.Kw = cg.no[2];
  g.mt = g.mh = g.mw = 1;
  g.ot = cg.omin[0]; g.oh = cg.omin[1]; g.ow = cg.omin[2];
  g.ts_t = g.ts_h = g.ts_w = 0;
  g.tp_t = g.tp_h = g.tp_w = 1;
  g.FT = d->Ti; g.FH = d->Hi; g.FW = d->Wi;
  g.vo_t = g.vo_h = g.vo_w = 0;
  g.vs_t = d->st; g.vs_h = d->sh; g.vs_w = d->sw;
  g.ss_w = g.ss_h = g.ss_t = g.ss_b = 0;
  return g;
}

static int launch_gather(const GatherProblem& g, const void* src, const void* wgt, void* dst, const void* addend,
                         float* part, int* nparts, cudaStream_t s, const void* yprev = nullptr, const float* bn_ss = nullptr,
                         float slope = 1.f, int epi_bn = 0, float slope_res = 1.f, const ClassGeom* cg = nullptr, int cpd = 0,
                         const dp_bn_fin* fin = nullptr) {
  TcPlan plan;
  const bool bwd_stats = yprev != nullptr;
  DP_REQUIRE(plan_gather(g, part != nullptr, &plan, bwd_stats, cg ? cpd : 0), DP_ERR_UNSUPPORTED, "tcgen05 conv: geometry not supported");
  DP_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)wgt & 15) == 0 && ((uintptr_t)dst & 15) == 0, DP_ERR_ALIGN,
             "tcgen05 conv: tensors must be 16-byte aligned");
  TcParams& p = plan.p;
  p.has_addend = addend != nullptr ? 1 : 0;
  p.slope = slope;
  p.epi_bn = epi_bn;
  p.slope_res = slope_res;
  DP_REQUIRE(!epi_bn || (part == nullptr && !bwd_stats && bn_ss != nullptr), DP_ERR_SHAPE,
             "tcgen05 conv: the fused BatchNorm epilogue excludes statistics");
  p.dbg_skip = g_opt_dbg_skip;
  p.l2_hint = g_opt_l2hint;
  p.stats_keep = g_opt_stats_keep;
  const CUtensorMapSwizzle sw = p.CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (p.CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB, tmD;
  const long long vstr[4] = {g.ss_w, g.ss_h, g.ss_t, g.ss_b};
  int rc = encode_act_map(&tmA, src, g.sC, g.sW, g.sH, g.sT, g.B, plan.a_box, plan.a_estride, sw, vstr);
  if (rc != DP_OK) return rc;
  const int taps = g.Kt * g.Kh * g.Kw;
  const int dCf = g.dCf ? g.dCf : g.dC;
  p.pC = dCf; p.n_base = g.n0;
  wgt = (const __nv_bfloat16*)wgt + (size_t)g.n0 * taps * g.sC;   // rows [n0, n0 + dC) of w[Kp][taps][Cp]
  rc = encode_wgt_map(&tmB, wgt, taps * g.sC, g.dC, p.CB, p.Ntile, sw);
  if (rc != DP_OK) return rc;
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  if (p.CBt) {
    const CUtensorMapSwizzle swt = p.CBt == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    int tbox[5] = {p.CBt, plan.a_box[1], plan.a_box[2], plan.a_box[3], plan.a_box[4]};
    rc = encode_act_map(&tmA2, src, g.sC, g.sW, g.sH, g.sT, g.B, tbox, plan.a_estride, swt, vstr);
    if (rc != DP_OK) return rc;
    rc = encode_wgt_map(&tmB2, wgt, taps * g.sC, g.dC, p.CBt, p.Ntile, swt);
    if (rc != DP_OK) return rc;
  }
  // one TMA store moves one 128-row sub-tile
  const int dbox[5] = {p.cw, p.sub_ow ? p.sub_ow : p.bw, p.sub_oh ? p.sub_oh : p.bh, p.sub_ot ? p.sub_ot : p.bt, 1};
  p.a_sw = (long long)g.vs_w * dCf;
  p.a_sh = (long long)g.vs_h * g.FW * dCf;
  p.a_st = (long long)g.vs_t * g.FH * g.FW * dCf;
  p.a_sb = (long long)g.FT * g.FH * g.FW * dCf;
  p.a_off = (((long long)g.vo_t * g.FH + g.vo_h) * g.FW + g.vo_w) * dCf + g.n0;
  CUtensorMap tmDk[4];
  if (cg == nullptr) {
    rc = encode_view_map(&tmD, (const __nv_bfloat16*)dst + p.a_off, g.dC, g.dW, g.dH, g.dT, g.B, p.a_sw, p.a_sh, p.a_st,
                         p.a_sb, dbox, p.st_layout);
    if (rc != DP_OK) return rc;
    tmDk[1] = tmDk[2] = tmDk[3] = tmD;
  } else {
    DP_REQUIRE(cg->ncls <= 4, DP_ERR_UNSUPPORTED, "tcgen05 dgrad classes: at most four classes per launch");
    p.ncls = cg->ncls; p.cpd = cpd;
    p.a_sw = (long long)g.vs_w * cpd;
    p.a_sh = (long long)g.vs_h * g.FW * cpd;
    p.a_st = (long long)g.vs_t * g.FH * g.FW * cpd;
    p.a_sb = (long long)g.FT * g.FH * g.FW * cpd;
    p.a_off = 0;
    for (int k = 0; k < 4; ++k) {
      const int kk = k < cg->ncls ? k : 0;
      p.a_off_k[k] = (((long long)cg->r[kk][0] * g.FH + cg->r[kk][1]) * g.FW + cg->r[kk][2]) * cpd;
      p.dTk[k] = (short)cg->vd[kk][0]; p.dHk[k] = (short)cg->vd[kk][1]; p.dWk[k] = (short)cg->vd[kk][2];
      rc = encode_view_map(&tmDk[k], (const __nv_bfloat16*)dst + p.a_off_k[k], cpd, cg->vd[kk][2], cg->vd[kk][1], cg->vd[kk][0], g.B,
                           p.a_sw, p.a_sh, p.a_st, p.a_sb, dbox, p.st_layout);
      if (rc != DP_OK) return rc;
    }
    tmD = tmDk[0];
  }

  static std::mutex attr_mu;
  static bool attr_done[DP_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  cudaError_t attr_err = cudaSuccess;
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                         const CUtensorMap, const CUtensorMap, const TcParams, const __nv_bfloat16*, float*,
                         long long*, const __nv_bfloat16*, const float*, const dp_bn_fin);
  static KernFn const kerns[20] = {tc_gather_gemm_kernel<0, 0, 1>, tc_gather_gemm_kernel<1, 1, 1>, tc_gather_gemm_kernel<2, 1, 1>,
                                   tc_gather_gemm_kernel<3, 1, 1>, tc_gather_gemm_kernel<1, 0, 1>, tc_gather_gemm_kernel<2, 0, 1>,
                                   tc_gather_gemm_kernel<3, 0, 1>, tc_gather_gemm_kernel<4, 0, 1>, tc_gather_gemm_kernel<5, 0, 1>,
                                   tc_gather_gemm_kernel<8, 0, 1>, tc_gather_gemm_kernel<1, 2, 1>, tc_gather_gemm_kernel<2, 2, 1>,
                                   tc_gather_gemm_kernel<3, 2, 1>,
                                   // 256-pixel tiles (Ntile <= 128, register statistics or none)
                                   tc_gather_gemm_kernel<1, 1, 2>, tc_gather_gemm_kernel<2, 1, 2>, tc_gather_gemm_kernel<3, 1, 2>,
                                   tc_gather_gemm_kernel<1, 0, 2>, tc_gather_gemm_kernel<2, 0, 2>, tc_gather_gemm_kernel<3, 0, 2>,
                                   tc_gather_gemm_kernel<4, 0, 2>};
  {
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(attr_mu);
    if (!attr_done[dev]) {
      for (int i = 0; i < 20 && attr_err == cudaSuccess; ++i)
        attr_err = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
      attr_done[dev] = attr_err == cudaSuccess;
    }
  }
  DP_REQUIRE(attr_err == cudaSuccess, DP_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s",
             cudaGetErrorString(attr_err));
  long long* dbg = (g_dbg && g_dbg_slots >= (size_t)148 * 16) ? g_dbg : nullptr;
  int ki = 0;
  if (p.reg_stats && bwd_stats) ki = 9 + p.drain_rs;                  // 10..12
  else if (p.reg_stats) ki = p.drain_rs;                              // 1..3
  else if (p.drain_rs >= 1 && p.drain_rs <= 5) ki = 3 + p.drain_rs;   // 4..8
  else if (p.drain_rs >= 6 && p.drain_rs <= 8) ki = 9;
  if (p.MT == 2) {
    DP_REQUIRE(!bwd_stats && p.drain_rs >= 1 && p.drain_rs <= (p.reg_stats ? 3 : 4), DP_ERR_UNSUPPORTED, "tcgen05 conv: no 256-pixel kernel for this shape");
    ki = p.reg_stats ? 12 + p.drain_rs : 15 + p.drain_rs;   // 13..15 / 16..19
  }
  static const dp_bn_fin no_fin = {};
  DP_REQUIRE(fin == nullptr || part != nullptr, DP_ERR_SHAPE, "tcgen05 conv: a fused finalisation needs the partials buffer");
  const size_t smem = (fin && plan.smem < 8192) ? 8192 : plan.smem;   // the finalisation tail borrows 4 KB + a flag word
  launch_pdl(kerns[ki], dim3(plan.grid), dim3(TC_THREADS), smem, s, tmA, tmB, tmD, tmA2, tmB2, tmDk[1], tmDk[2], tmDk[3], p, (const __nv_bfloat16*)addend, part, dbg,
             (const __nv_bfloat16*)yprev, bn_ss, fin ? *fin : no_fin);
  if (getenv("DP_DEBUG_PLAN"))
    fprintf(stderr, "[tc_gather] dst %dx%dx%dx%d src C=%d taps=%d | tile bw=%d bh=%d bt=%d nloads=%d nsub=%d CB=%d ncblk=%d Ntile=%d MT=%d stages=%d lps=%d dual=%d CBt=%d stats=%d/%d stage_bytes=%d a_box=%d tiles=%d grid=%d\n",
            g.dT, g.dH, g.dW, g.dC, g.sC, g.kt * g.kh * g.kw, p.bw, p.bh, p.bt, p.nloads, p.nsub, p.CB, p.ncblk, p.Ntile, p.MT, p.num_stages, p.lps, p.dual_mma, p.CBt, p.reg_stats * 10 + p.drain_rs, p.mma_stats,
            p.stage_bytes, p.a_box_bytes, p.num_tiles, plan.grid);
  if (nparts != nullptr) *nparts = plan.grid;
  return check_launch("tc_gather_gemm_kernel");
}

static GatherProblem fwd_problem(const dp_conv_desc* d) {
  GatherProblem g;
  g.B = d->B;
  g.sT = d->Ti; g.sH = d->Hi; g.sW = d->Wi; g.sC = d->Cp;
  g.dT = d->To; g.dH = d->Ho; g.dW = d->Wo; g.dC = d->Kp;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.mt = d->st; g.mh = d->sh; g.mw = d->sw;
  g.Kt = d->kt; g.Kh = d->kh; g.Kw = d->kw;
  g.ot = -d->pt; g.oh = -d->ph; g.ow = -d->pw;
  g.ts_t = g.ts_h = g.ts_w = 0;
  g.tp_t = g.tp_h = g.tp_w = 1;
  g.FT = g.dT; g.FH = g.dH; g.FW = g.dW;
  g.vo_t = g.vo_h = g.vo_w = 0;
  g.vs_t = g.vs_h = g.vs_w = 1;
  g.ss_w = g.ss_h = g.ss_t = g.ss_b = 0;
  return g;
}

// One parity class (rt,rh,rw) of the data gradient: input pixels p = s*q + r receive
//   dx[p] = sum over taps j with (r + pad - j) % s == 0 of  dy[q + (r + pad - j)/s] * w[j],
// a stride-1 gather over dy with a strided subset of the taps, written to every s-th pixel of dx.
// Returns false when the class has no pixel; *ntaps == 0 when it has pixels but no contributing tap.
static bool dgrad_class(const dp_conv_desc* d, int rt, int rh, int rw, GatherProblem* g, int* ntaps) {
  const int in[3] = {d->Ti, d->Hi, d->Wi}, out[3] = {d->To, d->Ho, d->Wo};
  const int k[3] = {d->kt, d->kh, d->kw}, st[3] = {d->st, d->sh, d->sw}, pad[3] = {d->pt, d->ph, d->pw};
  const int r[3] = {rt, rh, rw};
  int cnt[3], j0[3], off0[3], vd[3];
  for (int i = 0; i < 3; ++i) {
    if (r[i] >= in[i]) return false;
    vd[i] = (in[i] - r[i] + st[i] - 1) / st[i];
    j0[i] = (r[i] + pad[i]) % st[i];
    cnt[i] = j0[i] < k[i] ? (k[i] - 1 - j0[i]) / st[i] + 1 : 0;
    // position q = cnt-1-i  <->  tap j0 + s*i, source offset (r + pad - j0)/s - i
    off0[i] = (r[i] + pad[i] - j0[i]) / st[i] - (cnt[i] - 1);
  }
  (void)out;
  g->B = d->B;
  g->sT = d->To; g->sH = d->Ho; g->sW = d->Wo; g->sC = d->Kp;
  g->dT = vd[0]; g->dH = vd[1]; g->dW = vd[2]; g->dC = d->Cp;
  g->kt = cnt[0]; g->kh = cnt[1]; g->kw = cnt[2];
  g->Kt = d->kt; g->Kh = d->kh; g->Kw = d->kw;
  g->mt = g->mh = g->mw = 1;
  g->ot = off0[0]; g->oh = off0[1]; g->ow = off0[2];
  g->ts_t = j0[0] + st[0] * (cnt[0] - 1); g->ts_h = j0[1] + st[1] * (cnt[1] - 1); g->ts_w = j0[2] + st[2] * (cnt[2] - 1);
  g->tp_t = -st[0]; g->tp_h = -st[1]; g->tp_w = -st[2];
  g->FT = d->Ti; g->FH = d->Hi; g->FW = d->Wi;
  g->vo_t = rt; g->vo_h = rh; g->vo_w = rw;
  g->vs_t = st[0]; g->vs_h = st[1]; g->vs_w = st[2];
  g->ss_w = g->ss_h = g->ss_t = g->ss_b = 0;
  *ntaps = cnt[0] * cnt[1] * cnt[2];
  return true;
}

bool tc_fwd_supported(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16) return false;
  TcPlan plan;
  return plan_gather(fwd_problem(d), true, &plan);
}

bool tc_dgrad_supported(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16) return false;
  const bool strided = d->st != 1 || d->sh != 1 || d->sw != 1;
  if (strided && !g_opt_strided) return false;
  if (d->st > 4 || d->sh > 4 || d->sw > 4) return false;
  for (int rt = 0; rt < d->st; ++rt)
    for (int rh = 0; rh < d->sh; ++rh)
      for (int rw = 0; rw < d->sw; ++rw) {
        GatherProblem g;
        int ntaps = 0;
        if (!dgrad_class(d, rt, rh, rw, &g, &ntaps) || ntaps == 0) continue;
        TcPlan plan;
        if (!plan_gather(g, false, &plan)) return false;
      }
  return true;
}

// Output-channel split of a forward conv whose weights do not fit beside the pipeline (64 -> 144 channels, 9 taps: 166 KB):
// unsplit, every 128-pixel tile streams all of them from L2 again (935 MB of L2 reads per launch at B = 64, the layer's
// actual bound); as two launches over channels [0, na) and [na, dC) each half is resident in shared memory for the whole
// launch, the tiles are narrow enough for register statistics and 256-pixel tiles, and the input is read twice instead.
static bool fwd_split(const GatherProblem& g, bool has_stats, int* na) {
  if (!g_opt_nsplit || g.dC <= 96 || g.dC > 256) return false;
  TcPlan whole, a, b;
  if (!plan_gather(g, has_stats, &whole) || whole.p.w_resident) return false;
  if (g_opt_nsplit == 1 && whole.p.num_tiles < 4 * whole.grid) return false;   // too few tiles per CTA to pay for a second launch (2: always, tests)
  GatherProblem ga = g, gb = g;
  ga.dC = round_up(g.dC / 2, 16);
  gb.dC = g.dC - ga.dC;
  if (gb.dC < 16) return false;
  if (!plan_gather(ga, has_stats, &a) || !plan_gather(gb, has_stats, &b)) return false;
  if (!a.p.w_resident || !b.p.w_resident || a.grid != b.grid) return false;
  *na = ga.dC;
  return true;
}

int tc_conv_fwd(const dp_conv_desc* d, const void* x, const void* w, void* y, float* part, int* nparts,
                cudaStream_t s, const dp_bn_fin* fin) {
  const GatherProblem g = fwd_problem(d);
  int na = 0;
  if (fin == nullptr && fwd_split(g, part != nullptr, &na)) {
    GatherProblem ga = g, gb = g;
    ga.dCf = gb.dCf = g.dC;
    ga.dC = na; ga.n0 = 0;
    gb.dC = g.dC - na; gb.n0 = na;
    int n1 = 0, n2 = 0;
    int rc = launch_gather(ga, x, w, y, nullptr, part, &n1, s);
    if (rc != DP_OK) return rc;
    rc = launch_gather(gb, x, w, y, nullptr, part, &n2, s);
    if (rc != DP_OK) return rc;
    if (nparts != nullptr) *nparts = n1;   // (== n2: fwd_split)
    return DP_OK;
  }
  return launch_gather(g, x, w, y, nullptr, part, nparts, s, nullptr, nullptr, 1.f, 0, 1.f, nullptr, 0, fin);
}

bool tc_fwd_view_supported(const dp_conv_desc* d) { return tc_fwd_supported(d); }

bool tc_fwd_bnact_supported(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16) return false;
  TcPlan plan;
  return plan_gather(fwd_problem(d), false, &plan);
}

// eval-mode Conv3dBlock in one kernel: z = lrelu(conv(x) * scale + shift, slope) [; z = lrelu(z + residual, slope_res)]
int tc_conv_fwd_bnact(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w, const float* scale_shift,
                      float slope, const void* residual, float slope_res, void* z, cudaStream_t s) {
  GatherProblem g = fwd_problem(d);
  if (xstrides != nullptr) { g.ss_w = xstrides[0]; g.ss_h = xstrides[1]; g.ss_t = xstrides[2]; g.ss_b = xstrides[3]; }
  return launch_gather(g, x, w, z, residual, nullptr, nullptr, s, nullptr, scale_shift, slope, 1, slope_res);
}

// stride-1 data gradient whose epilogue also produces the BatchNorm-backward sums of the layer that produced x
bool tc_dgrad_bnstats_supported(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16 || d->st != 1 || d->sh != 1 || d->sw != 1) return false;
  GatherProblem g;
  int ntaps = 0;
  if (!dgrad_class(d, 0, 0, 0, &g, &ntaps) || ntaps == 0) return false;
  TcPlan plan;
  return plan_gather(g, true, &plan, true);
}

int tc_conv_dgrad_bnstats(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                          const void* yprev, const float* bn_scale_shift, float slope, float* part, int* nparts,
                          cudaStream_t s, const dp_bn_fin* fin) {
  GatherProblem g;
  int ntaps = 0;
  DP_REQUIRE(d->st == 1 && d->sh == 1 && d->sw == 1 && dgrad_class(d, 0, 0, 0, &g, &ntaps) && ntaps > 0, DP_ERR_UNSUPPORTED,
             "tcgen05 dgrad+stats: stride-1 convolutions only");
  return launch_gather(g, dy, w, dx, addend, part, nparts, s, yprev, bn_scale_shift, slope, 0, 1.f, nullptr, 0, fin);
}

int tc_conv_fwd_view(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w, void* y,
                     float* part, int* nparts, cudaStream_t s, const dp_bn_fin* fin) {
  GatherProblem g = fwd_problem(d);
  g.ss_w = xstrides[0]; g.ss_h = xstrides[1]; g.ss_t = xstrides[2]; g.ss_b = xstrides[3];
  return launch_gather(g, x, w, y, nullptr, part, nparts, s, nullptr, nullptr, 1.f, 0, 1.f, nullptr, 0, fin);
}

// ---- strided data gradient, all stride-parity classes in one launch ----
size_t tc_dgrad_classes_weight_elems(const dp_conv_desc* d) {
  if (d->dtype != DP_BF16 || (d->st == 1 && d->sh == 1 && d->sw == 1) || !g_opt_strided || !g_opt_classes) return 0;
  ClassGeom cg;
  if (!class_geom(d, &cg) || cg.ncls > 4) return 0;
  // 1x1x1 strided shortcuts have one class with a tap: a single launch plus a copy of the addend beats computing zeros
  if (d->kt * d->kh * d->kw == 1) return 0;
  TcPlan plan;
  if (!plan_gather(class_problem(d, cg), false, &plan, false, d->Cp)) return 0;
  // measured rule (scripts/strided_dgrad_bench.py, B = 64): a temporal stride-2 layer with many tiles per SM is bound by the
  // epilogue of its 2 x Cp wide tile and gains nothing from reading dy once (conv3: 155 vs 120 us); the short ones gain from
  // the saved launch (conv4 / conv5: 28 vs 34, 17 vs 26 us)
  if (d->sh == 1 && d->sw == 1 && plan.p.num_tiles > 20 * num_sms()) return 0;
  return (size_t)cg.ncls * d->Cp * cg.no[0] * cg.no[1] * cg.no[2] * d->Kp;
}

// w_cls[(class k, c)][position (pt,ph,pw)][Kp] = w[kout][c][jt][jh][jw], j = r_k + pad - s * (omin + p); zero outside the kernel
__global__ void pack_class_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, ClassGeom cg,
                                          int K, int C, int Kp, int Cp, int kt, int kh, int kw, int st, int sh, int sw,
                                          int pt, int ph, int pw) {
  pdl_launch_dependents();
  pdl_wait();
  const int npos = cg.no[0] * cg.no[1] * cg.no[2];
  const int64_t total = (int64_t)cg.ncls * Cp * npos * Kp;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ko = (int)(idx % Kp);
  int64_t rest = idx / Kp;
  const int pos = (int)(rest % npos); rest /= npos;
  const int c = (int)(rest % Cp);
  const int k = (int)(rest / Cp);
  const int p2 = pos % cg.no[2], p1 = (pos / cg.no[2]) % cg.no[1], p0 = pos / (cg.no[2] * cg.no[1]);
  const int jt = cg.r[k][0] + pt - st * (cg.omin[0] + p0);
  const int jh = cg.r[k][1] + ph - sh * (cg.omin[1] + p1);
  const int jw = cg.r[k][2] + pw - sw * (cg.omin[2] + p2);
  float v = 0.f;
  if (ko < K && c < C && jt >= 0 && jt < kt && jh >= 0 && jh < kh && jw >= 0 && jw < kw)
    v = w[((((int64_t)ko * C + c) * kt + jt) * kh + jh) * kw + jw];
  out[idx] = __float2bfloat16(v);
}

int tc_pack_dgrad_classes(const dp_conv_desc* d, const float* w, void* out, cudaStream_t s) {
  ClassGeom cg;
  DP_REQUIRE(tc_dgrad_classes_weight_elems(d) > 0 && class_geom(d, &cg), DP_ERR_UNSUPPORTED,
             "dgrad classes: geometry not supported");
  const int64_t total = (int64_t)tc_dgrad_classes_weight_elems(d);
  launch_pdl(pack_class_weights_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, w, (__nv_bfloat16*)out, cg, d->K, d->C,
             d->Kp, d->Cp, d->kt, d->kh, d->kw, d->st, d->sh, d->sw, d->pt, d->ph, d->pw);
  return check_launch("pack_class_weights_kernel");
}

int tc_conv_dgrad_classes(const dp_conv_desc* d, const void* dy, const void* w_cls, const void* addend, void* dx, cudaStream_t s) {
  ClassGeom cg;
  DP_REQUIRE(tc_dgrad_classes_weight_elems(d) > 0 && class_geom(d, &cg), DP_ERR_UNSUPPORTED,
             "dgrad classes: geometry not supported");
  return launch_gather(class_problem(d, cg), dy, w_cls, dx, addend, nullptr, nullptr, s, nullptr, nullptr, 1.f, 0, 1.f, &cg, d->Cp);
}

int tc_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                  cudaStream_t s) {
  // pixels of classes without a contributing tap (1x1x1 strided shortcuts) receive the addend or zero
  bool any_empty = false;
  for (int rt = 0; rt < d->st; ++rt)
    for (int rh = 0; rh < d->sh; ++rh)
      for (int rw = 0; rw < d->sw; ++rw) {
        GatherProblem g;
        int ntaps = 0;
        if (dgrad_class(d, rt, rh, rw, &g, &ntaps) && ntaps == 0) any_empty = true;
      }
  if (any_empty) {
    const size_t bytes = (size_t)d->B * d->Ti * d->Hi * d->Wi * d->Cp * 2;
    cudaError_t e = addend != nullptr ? cudaMemcpyAsync(dx, addend, bytes, cudaMemcpyDeviceToDevice, s)
                                      : cudaMemsetAsync(dx, 0, bytes, s);
    DP_REQUIRE(e == cudaSuccess, DP_ERR_CUDA, "tcgen05 dgrad: clearing dx failed: %s", cudaGetErrorString(e));
  }
  for (int rt = 0; rt < d->st; ++rt)
    for (int rh = 0; rh < d->sh; ++rh)
      for (int rw = 0; rw < d->sw; ++rw) {
        GatherProblem g;
        int ntaps = 0;
        if (!dgrad_class(d, rt, rh, rw, &g, &ntaps) || ntaps == 0) continue;
        const int rc = launch_gather(g, dy, w, dx, addend, nullptr, nullptr, s);
        if (rc != DP_OK) return rc;
      }
  return DP_OK;
}

// development / test aid: the plan the tcgen05 gather kernel would run for a forward (op 0, with or without
// BatchNorm statistics) or a stride-1 data gradient (op 1), as text -- no launch, works without a device
int tc_describe_plan(const dp_conv_desc* d, int op, int has_stats, char* out, size_t n) {
  GatherProblem g;
  if (op == 2) return tc_wgrad_describe(d, out, n);
  if (op == 0) {
    g = fwd_problem(d);
  } else {
    int ntaps = 0;
    if (!dgrad_class(d, 0, 0, 0, &g, &ntaps) || ntaps == 0) return DP_ERR_UNSUPPORTED;
  }
  TcPlan plan;
  if (!plan_gather(g, has_stats != 0, &plan, op == 1 && has_stats != 0)) return DP_ERR_UNSUPPORTED;
  const TcParams& p = plan.p;
  int na = 0;   // forward: channels of the first of two launches when the output channels are split (0: one launch)
  if (op == 0 && !fwd_split(g, has_stats != 0, &na)) na = 0;
  snprintf(out, n,
           "tile bw=%d bh=%d bt=%d MT=%d nloads=%d nsub=%d CB=%d ncblk=%d CBt=%d Ntile=%d n_ntiles=%d stages=%d lps=%d stage_bytes=%d "
           "dual=%d acc_bufs=%d st_bufs=%d pub_sub=%d resident=%d reg_stats=%d mma_stats=%d drain_rs=%d tmem_cols=%d tiles=%d grid=%d smem=%zu "
           "split=%d",
           p.bw, p.bh, p.bt, p.MT, p.nloads, p.nsub, p.CB, p.ncblk, p.CBt, p.Ntile, p.n_ntiles, p.num_stages, p.lps, p.stage_bytes,
           p.dual_mma, p.acc_bufs, p.st_bufs, p.pub_sub, p.w_resident, p.reg_stats, p.mma_stats, p.drain_rs, p.tmem_cols, p.num_tiles, plan.grid,
           plan.smem, na);
  return DP_OK;
}

}  // namespace dp

DP_API int dp_set_option(const char* name, int value) {
  if (name == nullptr) return DP_ERR_SHAPE;
  if (!strcmp(name, "pdl")) { dp::g_pdl = value ? 1 : 0; return DP_OK; }
  if (!strcmp(name, "pdl_small")) { dp::g_pdl_small = value ? 1 : 0; return DP_OK; }
  if (!strcmp(name, "bn_sweep")) { dp::g_bn_sweep = value; return DP_OK; }
  if (!strcmp(name, "bn_cs")) { dp::g_bn_cs = value; return DP_OK; }
  if (!strcmp(name, "strict_tc")) { dp::g_strict_tc = value ? 1 : 0; return DP_OK; }
  if (dp::tc_option(name, value, true) >= 0 || dp::wg_option(name, value, true) >= 0) return DP_OK;
  dp::set_error("dp_set_option: unknown option '%s'", name);
  return DP_ERR_UNSUPPORTED;
}
DP_API int dp_conv_describe_plan(const dp_conv_desc* d, int op, int has_stats, char* out, size_t n) {
  if (d == nullptr || out == nullptr || n == 0) return DP_ERR_SHAPE;
  return dp::tc_describe_plan(d, op, has_stats, out, n);
}
DP_API int dp_get_option(const char* name) {
  if (name == nullptr) return -1;
  if (!strcmp(name, "pdl")) return dp::g_pdl;
  if (!strcmp(name, "pdl_small")) return dp::g_pdl_small;
  if (!strcmp(name, "bn_sweep")) return dp::g_bn_sweep;
  if (!strcmp(name, "bn_cs")) return dp::g_bn_cs;
  if (!strcmp(name, "strict_tc")) return dp::g_strict_tc;
  const int v = dp::tc_option(name, 0, false);
  return v >= 0 ? v : dp::wg_option(name, 0, false);
}
