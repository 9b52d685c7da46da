/*
 * dp_b200.h -- C ABI of the B200-native R(2+1)D hot path.
 *
 * One shared library (libdp_b200.so, built from
 * disruption-prediciton-based-on-multimodal-deep-learning_b200/csrc) exports exactly the symbols declared
 * here.  Every entry point takes plain device pointers and sizes, enqueues
 * work on the cudaStream_t passed as `void* stream`, never allocates device
 * memory, never synchronises, never throws, and returns DP_OK or a negative
 * DP_ERR_* code (dp_last_error() gives the text for the calling thread).
 *
 * The reference (ZINZINBIN/Disruption-Prediciton-based-on-Multimodal-Deep-Learning)
 * has no FFI layer of its own: its boundary for this path is the nn.Module /
 * loss protocol.  Each entry point cites the reference lines whose library
 * calls (ATen / cuDNN) it replaces.  The ctypes binding a maintainer of the
 * reference would add is shown in INTEGRATION.md.
 *
 * Layout contract (internal to the path):
 *   activations : NDHWC, channel dimension padded with zeros to Cp = ceil16(C)
 *   dtype       : DP_BF16 (product path, fp32 accumulate) or DP_F32
 *                 (fp32 validation mode)
 *   weights     : caller keeps fp32 (K,C,kt,kh,kw) master tensors (state_dict
 *                 format); dp_pack_weights makes the private packed copies.
 */
#ifndef DP_B200_H
#define DP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_OK               0
#define DP_ERR_SHAPE       -1
#define DP_ERR_ALIGN       -2
#define DP_ERR_ARCH        -3
#define DP_ERR_CUDA        -4
#define DP_ERR_UNSUPPORTED -5

#define DP_F32  0
#define DP_BF16 1

/* kernel family selector for the conv entry points */
#define DP_IMPL_AUTO 0   /* tcgen05 where the shape is covered, else SIMT */
#define DP_IMPL_SIMT 1   /* CUDA-core implicit GEMM (any shape, both dtypes) */
#define DP_IMPL_TC   2   /* TMA + tcgen05/TMEM implicit GEMM (bf16 only) */

/* loss kinds: src/loss.py:71 (CE), :14 (Focal), :37 (LDAM) */
#define DP_LOSS_CE    0
#define DP_LOSS_FOCAL 1
#define DP_LOSS_LDAM  2

/* rows of per-CTA BatchNorm partial sums any producer may emit */
#define DP_MAX_PARTS 592

/* One Conv3d of the path (nn.Conv3d at src/models/R2Plus1D.py:44-51, bias=False,
 * dilation 1).  Input (B,Ti,Hi,Wi,Cp) -> output (B,To,Ho,Wo,Kp), NDHWC. */
typedef struct dp_conv_desc {
  int32_t B, Ti, Hi, Wi, C, Cp;
  int32_t To, Ho, Wo, K, Kp;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
  int32_t dtype; /* DP_F32 | DP_BF16: storage type of x, y, packed weights */
} dp_conv_desc;

/* "Last CTA done" BatchNorm finalisation, run inside the kernel that produces the per-CTA partial sums (dp_conv_fwd_fin,
 * dp_stem_conv_fwd_fin, dp_conv_dgrad_bnstats_fin, dp_bn_act_bwd_reduce_fin) instead of a stand-alone dp_bn_finalize /
 * dp_bn_bwd_finalize launch: 64 launches less per training step of the BASELINE model.
 *   kind 1 (nn.BatchNorm3d, train mode, R2Plus1D.py:53-54): (sum y, sum y^2) -> mean, rstd, scale, shift written,
 *           running_mean / running_var updated in place when non-NULL (momentum, unbiased variance);
 *   kind 2 (its autograd): (sum g', sum g'*y) -> dbeta, dgamma (may be NULL), coef[2][Cp] = the two means dy needs;
 *           mean / rstd are READ; coef_zero != 0 (eval-mode BatchNorm: constant statistics) writes coef = 0.
 * ticket: one 32-bit word of device memory owned by the caller, zero on entry; the kernel leaves it zero.  A word must
 * not be shared by two launches that can be in flight at the same time. */
typedef struct dp_bn_fin {
  int32_t kind;
  int32_t C, Cp;
  int32_t coef_zero;
  double count;               /* elements per channel (B*T*H*W) */
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* mean;
  float* rstd;
  float* scale;
  float* shift;
  float* dgamma;
  float* dbeta;
  float* coef;
  unsigned int* ticket;
} dp_bn_fin;

int         dp_version(void);
const char* dp_last_error(void);
/* DP_OK iff the current device is compute capability 10.x (sm_100a code). */
int         dp_device_check(void);
int         dp_num_sms(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long dp_launch_count(void);
/* conv calls served by the CUDA-core (SIMT) family, and how many of those were DP_IMPL_AUTO falling back from
 * tcgen05 on a bf16 tensor (a silent slow path: bench.py reports it and it must be 0 on the product path; the
 * option "strict_tc" = 1 turns such a fallback into DP_ERR_UNSUPPORTED) */
unsigned long long dp_simt_launch_count(void);
unsigned long long dp_simt_fallback_count(void);
/* tuning / debug switches (A/B measurements, tests; the defaults are the measured optimum): planner -- "tc_enable",
 * "tc_halo", "tc_strided", "tc_classes", "tc_max_stages", "tc_lps_max", "tc_mt", "tc_dual_mma", "tc_fine_n", "tc_tail",
 * "tc_acc4", "tc_st_bufs", "tc_pub_sub_min", "tc_resident", "tc_chunked", "tc_reg_stats", "tc_mma_stats",
 * "tc_bwd_stats_max", "tc_stats_keep" (staged-tile statistics combined once per CTA, default 1), "tc_nsplit" (forward
 * output channels in two launches with resident weights, default 0), "tc_l2hint" (evict-first activation loads,
 * default 0), "wg_enable", "wg_halo", "wg_stack", "wg_items"; launch -- "pdl" / "pdl_small" (programmatic dependent
 * launch of every kernel / of the small finalize and split-reduce kernels, default 0), "strict_tc"; BatchNorm passes --
 * "bn_sweep" (traversal direction bits, default 0), "bn_cs" (streaming loads of dead operands, default 0) */
int         dp_set_option(const char* name, int value);
int         dp_get_option(const char* name);
/* development / test aid: the tile, pipeline and statistics plan of the tcgen05 gather kernel for a forward (op 0) or
 * stride-1 data gradient (op 1) of this geometry, or of the weight-gradient kernel (op 2), as one line of text (no
 * launch; works without a device) */
int         dp_conv_describe_plan(const dp_conv_desc* d, int op, int has_stats, char* out, size_t n);
/* development aid: device buffer (>= 8 int64 per CTA) receiving per-role wait/busy cycle counters of the tcgen05
 * kernels; NULL switches it off (the default) */
int         dp_set_debug_buffer(void* ptr, size_t bytes);

/* ---- layout (replaces the implicit NCDHW contract of DatasetForVideo, src/dataset.py:229-230) ---- */
int dp_ncdhw_f32_to_ndhwc(const float* src, void* dst, int B, int C, int Cp, int T, int H, int W,
                          int dtype, void* stream);
int dp_ndhwc_to_ncdhw_f32(const void* src, float* dst, int B, int C, int Cp, int T, int H, int W,
                          int dtype, void* stream);
/* uint8 frames (B,T,H,W,3) BGR -> mean-subtracted NDHWC (src/dataset.py:104-110,201-205) */
int dp_u8_frames_to_ndhwc(const uint8_t* src, void* dst, const float* mean3, int B, int T, int H, int W,
                          int Cp, int dtype, void* stream);

/* ---- weights ---- */
/* w: fp32 (K,C,kt,kh,kw).  w_fwd: [Kp][taps][Cp], w_dgrad: [Cp][taps][Kp] in desc->dtype. */
int dp_pack_weights(const dp_conv_desc* d, const float* w, void* w_fwd, void* w_dgrad, void* stream);

/* ---- convolution: forward / dgrad / wgrad (replace cuDNN behind nn.Conv3d, R2Plus1D.py:44-51,57) ---- */
int dp_conv_supported(const dp_conv_desc* d, int op /*0 fwd,1 dgrad,2 wgrad,3 dgrad_bnstats fused in the epilogue,
                                                      4 fwd_bnact fused in the epilogue*/, int impl);
/* y = conv(x, w).  If `part` is non-NULL, per-channel (sum, sumsq) partials of y
 * are written to part[nparts][2][Kp] and *nparts is set (<= DP_MAX_PARTS). */
int dp_conv_fwd(const dp_conv_desc* d, const void* x, const void* w_fwd, void* y,
                float* part, int* nparts, int impl, void* stream);
/* Eval-mode Conv3dBlock in ONE kernel (model.eval() at src/utils/utility.py:936-949, evaluate.py:38-44): BatchNorm
 * with running statistics is the per-channel affine scale_shift = scale[Kp] ++ shift[Kp] (dp_bn_eval_coeffs), applied
 * to the fp32 accumulator in the epilogue:  z = lrelu(conv(x,w)*scale + shift, slope); with `residual` (same shape
 * as z; the tail of SpatioTemporalResBlock, R2Plus1D.py:181-187):  z = lrelu(z + residual, slope_res).  The raw conv
 * output never reaches HBM.  `xstrides` (NULL = dense NDHWC) gives element strides of (w,h,t,b) of an input VIEW;
 * overlapping views are allowed (sliding windows over a per-frame cache share frames: b and t strides both one
 * frame), out-of-range coordinates of the view read as zero (= conv padding). */
int dp_conv_fwd_bnact(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w_fwd,
                      const float* scale_shift, float slope, const void* residual, float slope_res, void* z,
                      int impl, void* stream);
/* dx = conv_transpose(dy, w) (+ addend if non-NULL, same shape as dx). */
int dp_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend,
                  void* dx, int impl, void* stream);
/* Strided data gradient with EVERY stride-parity class in one launch (tcgen05 family, bf16).  The autograd of a strided
 * nn.Conv3d (R2Plus1D.py:44-51 with stride (1,2,2) / (2,1,1): the down-sampling blocks :172-176) scatters dy into
 * s_t*s_h*s_w interleaved pixel classes of dx; each class is a stride-1 gather over dy with a subset of the taps.
 * dp_conv_dgrad runs one launch per class (each re-reading dy); here the classes are the column blocks of ONE stride-1
 * gather whose weights w_cls[(class,c)][position][Kp] hold the class's tap (or zeros) for every dy offset, so dy is read
 * once.  dp_dgrad_classes_weight_elems: elements of w_cls, 0 when the geometry is not covered (use dp_conv_dgrad);
 * dp_pack_weights_dgrad_classes: fp32 master (K,C,kt,kh,kw) -> w_cls. */
size_t dp_dgrad_classes_weight_elems(const dp_conv_desc* d, int impl);
int dp_pack_weights_dgrad_classes(const dp_conv_desc* d, const float* w, void* w_cls, void* stream);
int dp_conv_dgrad_classes(const dp_conv_desc* d, const void* dy, const void* w_cls, const void* addend, void* dx,
                          void* stream);
/* dp_conv_dgrad plus the two per-channel sums the BatchNorm backward of the layer that PRODUCED x needs (the
 * autograd of Conv3dBlock, R2Plus1D.py:44-54): with g = dx, yp = that layer's raw conv output (same shape as dx),
 * g' = g * lrelu'(scale*yp + shift):  part[nparts][2][Cp] = (sum g', sum g'*yp) -- the same partials
 * dp_bn_act_bwd_reduce(dx, yp, NULL, ...) writes.  On stride-1 geometries the tcgen05 kernel accumulates them in its
 * epilogue (no pass over dx / yp); otherwise the data gradient is followed by that reduction.
 * scale_shift = scale[Cp] followed by shift[Cp]. */
int dp_conv_dgrad_bnstats(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend,
                          void* dx, const void* y_prev, const float* scale_shift, float slope,
                          float* part, int* nparts, int impl, void* stream);
/* dp_conv_fwd / dp_conv_dgrad_bnstats with the BatchNorm finalisation of the partials in the same launch (see
 * dp_bn_fin): fin->kind 1 for the forward (the statistics of y = this layer's BatchNorm3d), kind 2 for the data
 * gradient (the sums of the layer that produced x).  `part` is scratch for >= DP_MAX_PARTS * 2 * Kp (Cp) floats.  Where
 * the partials do not come out of the tcgen05 epilogue (CUDA-core family, composed reductions) the finalisation runs in
 * the last CTA of the reduction kernel instead; the result is the same. */
int dp_conv_fwd_fin(const dp_conv_desc* d, const void* x, const void* w_fwd, void* y, float* part,
                    const dp_bn_fin* fin, int impl, void* stream);
int dp_conv_dgrad_bnstats_fin(const dp_conv_desc* d, const void* dy, const void* w_dgrad, const void* addend,
                              void* dx, const void* y_prev, const float* scale_shift, float slope,
                              float* part, const dp_bn_fin* fin, int impl, void* stream);
size_t dp_conv_wgrad_workspace(const dp_conv_desc* d, int impl);
/* dw (fp32, (K,C,kt,kh,kw)) = x^T * dy, deterministic split reduction through `workspace`. */
int dp_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw,
                  void* workspace, size_t workspace_bytes, int impl, void* stream);

/* ---- stem fast path: C <= 4 input channels, kernel (1,kh,kw<=8), stride (1,sh,2), bf16 (R2Plus1D.py:137-140) ----
 * The clip is kept as packed row pairs XP[b][t][P][wp][r][4] (padded row h + ph = 2P + r, wp = w + pw, zero padded;
 * HP = Ho + ceil(kh/2) - 1, WP = 2*Wo + 6) instead of 16-channel NDHWC; forward and weight gradient run on the tcgen05
 * kernels as a stride-1 (1,ceil(kh/2),1) convolution over an overlapping-window view with 64 "channels" (one halo box per
 * tile).  Geometry: kt = 1, sh = sw = 2, kw <= 8, C <= 4.  There is no data gradient on this path (clips carry none). */
int    dp_stem_supported(const dp_conv_desc* d);
size_t dp_stem_input_elems(const dp_conv_desc* d);            /* bf16 elements of XP */
int    dp_stem_pack_input_f32(const dp_conv_desc* d, const float* ncdhw, void* xp, void* stream);
int    dp_stem_pack_input_u8(const dp_conv_desc* d, const uint8_t* frames /* (B,T,H,W,3) */, const float* mean3,
                             void* xp, void* stream);
size_t dp_stem_weight_elems(const dp_conv_desc* d);           /* bf16 elements of wv */
int    dp_stem_pack_weights(const dp_conv_desc* d, const float* w, void* wv /* [Kp][ceil(kh/2)][64] bf16 */, void* stream);
int    dp_stem_conv_fwd(const dp_conv_desc* d, const void* xp, const void* wv, void* y, float* part, int* nparts,
                        void* stream);
/* ... with the BatchNorm finalisation in the same launch (fin->kind 1, see dp_bn_fin / dp_conv_fwd_fin) */
int    dp_stem_conv_fwd_fin(const dp_conv_desc* d, const void* xp, const void* wv, void* y, float* part,
                            const dp_bn_fin* fin, void* stream);
/* eval-mode stem: conv + BatchNorm(running statistics) + LeakyReLU in one kernel (see dp_conv_fwd_bnact) */
int    dp_stem_conv_fwd_bnact(const dp_conv_desc* d, const void* xp, const void* wv, const float* scale_shift, float slope,
                              void* z, void* stream);
size_t dp_stem_wgrad_workspace(const dp_conv_desc* d);
int    dp_stem_conv_wgrad(const dp_conv_desc* d, const void* xp, const void* dy, float* dw, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- BatchNorm3d (train) + LeakyReLU (R2Plus1D.py:53-57,179-187) ---- */
int dp_bn_stats(const void* y, int64_t rows, int Cp, int dtype, float* part, int* nparts, void* stream);
int dp_bn_finalize(const float* part, int nparts, int C, int Cp, double count,
                   const float* gamma, const float* beta, float eps, float momentum,
                   float* running_mean, float* running_var,
                   float* mean, float* rstd, float* scale, float* shift, void* stream);
int dp_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma,
                      const float* beta, float eps, int C, int Cp, float* scale, float* shift,
                      void* stream);
/* z = lrelu(y*scale+shift, slope); if residual: z = lrelu(z + residual, slope_res). */
int dp_bn_act_apply(const void* y, const float* scale, const float* shift, float slope,
                    const void* residual, float slope_res, void* z, int64_t rows, int Cp,
                    int dtype, void* stream);
/* backward of the above, pass 1: per-channel sum(g), sum(g*y) partials (y = raw conv output). */
int dp_bn_act_bwd_reduce(const void* dz, const void* y, const void* out,
                         const float* scale, const float* shift, const float* mean,
                         const float* rstd, float slope, float slope_res,
                         float* part, int* nparts, int64_t rows, int Cp, int dtype, void* stream);
/* pass 1 with dp_bn_bwd_finalize in the same launch (fin->kind 2: mean / rstd are taken from fin) */
int dp_bn_act_bwd_reduce_fin(const void* dz, const void* y, const void* out, const float* scale, const float* shift,
                             float slope, float slope_res, float* part, int64_t rows, int Cp, int dtype,
                             const dp_bn_fin* fin, void* stream);
/* dbeta = sum(g), dgamma = sum(g*xhat) = (sum(g*y) - mean*sum(g))*rstd (fp64), coef = {dbeta, dgamma}/count */
int dp_bn_bwd_finalize(const float* part, int nparts, int C, int Cp, double count, const float* mean,
                       const float* rstd, float* dgamma, float* dbeta, float* coef, void* stream);
/* pass 2: dy = scale*(g - coef0 - xhat*coef1); optionally dres = dz*lrelu'(out). */
int dp_bn_act_bwd_apply(const void* dz, const void* y, const void* out,
                        const float* scale, const float* shift, const float* mean,
                        const float* rstd, const float* coef, float slope, float slope_res,
                        void* dy, void* dres, int64_t rows, int Cp, int dtype, void* stream);
/* plain elementwise sum of two activation tensors (grad fan-in at block inputs) */
int dp_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);

/* ---- AdaptiveAvgPool3d(1) (R2Plus1D.py:215,224-225) ---- */
int dp_avgpool_fwd(const void* x, float* out, int B, int64_t pixels, int C, int Cp, int dtype,
                   void* stream);
int dp_avgpool_bwd(const float* dout, void* dx, int B, int64_t pixels, int C, int Cp, int dtype,
                   void* stream);

/* ---- SlowFast auxiliaries (config 3; src/models/resnet.py:63-81,172-225, src/models/slowfast.py:26-36) ---- */
/* out = swish(x * gate[b][c]) (gate NULL = plain Swish): the squeeze-excite scaling fused with SwishEfficient */
int dp_se_swish_fwd(const void* x, const float* gate /* (B,C) or NULL */, void* out, int B, int64_t pixels, int C,
                    int Cp, int dtype, void* stream);
/* dx = dout * swish'(x*gate) * gate;  dgate[b][c] = sum_pixels dout * swish'(x*gate) * x  (NULL with gate NULL) */
int dp_se_swish_bwd(const void* x, const float* gate, const void* dout, void* dx, float* dgate, int B,
                    int64_t pixels, int C, int Cp, int dtype, void* stream);
/* MaxPool3d(kernel (1,3,3), stride (1,2,2), padding (0,1,1)) over `frames` = B*T images of (H,W,Cp);
 * idx (one byte per output element, may be NULL in inference) records the winning tap for the backward */
int dp_maxpool_hw_fwd(const void* x, void* out, uint8_t* idx, int64_t frames, int H, int W, int Cp, int dtype,
                      void* stream);
int dp_maxpool_hw_bwd(const void* dout, const uint8_t* idx, void* dx, int64_t frames, int H, int W, int Cp,
                      int dtype, void* stream);
/* torch.cat([a, b], dim=channels) on channel-padded NDHWC rows (padding of the result zeroed), and its backward */
int dp_concat_channels(const void* a, const void* b, void* out, int64_t rows, int Ca, int Cap, int Cb, int Cbp,
                       int Cop, int dtype, void* stream);
int dp_split_channels(const void* dout, void* da, void* db, int64_t rows, int Ca, int Cap, int Cb, int Cbp,
                      int Cop, int dtype, void* stream);

/* ---- losses (src/loss.py:14-81) ---- */
size_t dp_loss_workspace(int64_t n);
/* loss_out[0] = loss, loss_out[1] = normaliser (sum w[y] for LDAM, 1 otherwise).
 * dlogits = d(loss * normaliser)/dlogits, i.e. unnormalised. */
int dp_loss_fwd_bwd(int kind, const float* logits, const int64_t* target, const float* weight,
                    const float* margins, float gamma, float s, int64_t n, int C,
                    float* loss_out, float* dlogits, void* workspace, void* stream);
/* out = dlogits * grad_out[0] / loss_out[1]  (device scalars, no host sync) */
int dp_loss_bwd_scale(const float* dlogits, const float* grad_out, const float* loss_out,
                      float* out, int64_t count, void* stream);

/* ---- fused optimiser tail (src/train.py:63-66: clip_grad_norm_ + AdamW.step) ----
 * The workspace starts with four uint32 words the host may read / restore (checkpoints):
 *   [0] internal counter, [1] step = number of updates applied, [2] skipped = steps skipped because the gradient
 *   norm was not finite, [3] skip flag of the current step.
 * A step whose (scaled) gradient norm is NaN/Inf is skipped ON THE DEVICE: weights, moments and the step count stay
 * untouched -- the reference skips backward and step on a non-finite loss (src/train.py:55-60) -- with no host sync,
 * so the guard also works inside a captured CUDA graph. */
size_t dp_optim_workspace(int64_t n);
/* the two halves of dp_clip_adamw_step: norm_out[0] = ||g * grad_scale||_2 (count_step != 0: the device step count
 * advances if the norm is finite); then the AdamW update of one contiguous range (parameters without a gradient are
 * left out of the ranges, as torch.optim.AdamW leaves them untouched). */
int dp_grad_sqnorm(const float* g, int64_t n, float grad_scale, int count_step, float* norm_out, void* workspace,
                   void* stream);
int dp_adamw_apply(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float max_norm, float grad_scale, const float* norm,
                   const void* workspace, void* stream);
/* flat fp32 param/grad/moment buffers of n elements; decoupled weight decay (AdamW).
 * max_norm <= 0 disables clipping.  grad_scale multiplies g first (1/world for DP mean).
 * step >= 1: bias-correction step given by the host; step == 0: the step count lives in `workspace` on the device
 * and is incremented by each call (nothing host-dependent is baked into a captured CUDA graph). */
int dp_clip_adamw_step(float* p, const float* g, float* m, float* v, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay,
                       int step, float max_norm, float grad_scale, float* norm_out,
                       void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DP_B200_H */
