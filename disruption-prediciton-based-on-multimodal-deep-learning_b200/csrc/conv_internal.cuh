// Internal (non-ABI) launchers shared between the conv translation units.
#pragma once
#include "dp_common.cuh"

namespace dp {

// CUDA-core family (conv_simt.cu)
int simt_conv_fwd(const dp_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
int simt_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                    cudaStream_t s);
int simt_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* ws,
                    cudaStream_t s);
size_t simt_wgrad_workspace(const dp_conv_desc* d);
// dw (K,C,taps) = sum over splits of partial[split][Kp][taps][Cp], fixed order
int wgrad_reduce_launch(const float* partial, float* dw, int nsplit, int K, int C, int Kp, int Cp, int taps,
                        cudaStream_t s);
extern int g_strict_tc;   // conv_api.cu: DP_IMPL_AUTO refuses the CUDA-core fallback on bf16 tensors
extern int g_bn_sweep;    // bn_act.cu: traversal direction of the BatchNorm passes (bits, see there)
extern int g_bn_cs;       // bn_act.cu: streaming loads of dead operands in the elementwise BatchNorm passes
int wg_option(const char* name, int value, bool set);
int tc_option(const char* name, int value, bool set);

// tcgen05 family (conv_tc.cu / wgrad_tc.cu)
bool tc_fwd_supported(const dp_conv_desc* d);
bool tc_dgrad_supported(const dp_conv_desc* d);
bool tc_wgrad_supported(const dp_conv_desc* d);
// writes BN partials when part != nullptr; *nparts receives the row count
// fin != nullptr: the last CTA to retire also runs the BatchNorm finalisation over the partials (bn_fin.cuh)
int tc_conv_fwd(const dp_conv_desc* d, const void* x, const void* w, void* y, float* part, int* nparts,
                cudaStream_t s, const dp_bn_fin* fin = nullptr);
int tc_conv_dgrad(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                  cudaStream_t s);
// dgrad whose epilogue also accumulates sum(g'), sum(g' * yprev) per channel of dx (BatchNorm backward of the producer)
bool tc_dgrad_bnstats_supported(const dp_conv_desc* d);
int tc_conv_dgrad_bnstats(const dp_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                          const void* yprev, const float* bn_scale_shift, float slope, float* part, int* nparts,
                          cudaStream_t s, const dp_bn_fin* fin = nullptr);
// strided data gradient with every stride-parity class in one launch (class-packed weights, see conv_tc.cu)
size_t tc_dgrad_classes_weight_elems(const dp_conv_desc* d);   // 0: not supported for this geometry
int tc_pack_dgrad_classes(const dp_conv_desc* d, const float* w, void* out, cudaStream_t s);
int tc_conv_dgrad_classes(const dp_conv_desc* d, const void* dy, const void* w_cls, const void* addend, void* dx, cudaStream_t s);
size_t tc_wgrad_workspace(const dp_conv_desc* d);
int tc_wgrad_describe(const dp_conv_desc* d, char* out, size_t n);   // the weight-gradient plan as text (no launch)
int tc_conv_wgrad(const dp_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
                  cudaStream_t s);

// same kernels over a strided / overlapping VIEW of x (element strides of w,h,t,b): the packed stem rows
int tc_conv_fwd_view(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w, void* y,
                     float* part, int* nparts, cudaStream_t s, const dp_bn_fin* fin = nullptr);
int tc_conv_wgrad_view(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* dy, float* dw,
                       void* ws, size_t ws_bytes, cudaStream_t s);

// eval-mode fused Conv3d -> BatchNorm(running statistics) -> LeakyReLU [-> + residual -> LeakyReLU] (no statistics)
bool tc_fwd_bnact_supported(const dp_conv_desc* d);
int tc_conv_fwd_bnact(const dp_conv_desc* d, const long long* xstrides, const void* x, const void* w, const float* scale_shift,
                      float slope, const void* residual, float slope_res, void* z, cudaStream_t s);

// BN partial statistics over a finished tensor (bn_act.cu)
int bn_stats_launch(const void* y, int64_t rows, int Cp, int dtype, float* part, int* nparts, cudaStream_t s,
                    const dp_bn_fin* fin = nullptr);

}  // namespace dp
