// Shared helpers for the dp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/dp_b200.h"

#define DP_API extern "C" __attribute__((visibility("default")))

namespace dp {

void set_error(const char* fmt, ...);

extern unsigned long long g_launches;  // kernels launched by this library (dp_launch_count)

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DP_ERR_CUDA;
  }
  return DP_OK;
}

#define DP_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      dp::set_error(__VA_ARGS__);        \
      return (code);                     \
    }                                    \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int DP_MAX_DEVICES = 64;
int current_device();   // clamped to [0, DP_MAX_DEVICES)
int num_sms();          // of the current device

// ---- programmatic dependent launch (PDL) ----
// A step is ~370 short kernels; launched with programmatic stream serialization a kernel's CTAs may become resident
// while the previous kernel drains (its prologue overlaps the predecessor's tail).  Every kernel launched this way
// executes pdl_wait() before its first global-memory access and pdl_launch_dependents() at its start.
extern int g_pdl;       // every kernel launched through launch_pdl (measured slower on the whole step: off)
extern int g_pdl_small; // only the small latency-bound kernels that sit between two big ones (launch_pdl_small)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_pdl_if(g_pdl != 0, kernel, grid, block, smem, s, static_cast<Args&&>(args)...);
}
// a few CTAs of 256 threads that fit beside the producer's CTAs: they become resident while the producer drains and sit
// in griddepcontrol.wait, so their launch latency leaves the critical path
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_pdl_if(g_pdl != 0 || g_pdl_small != 0, kernel, grid, block, smem, s, static_cast<Args&&>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// optional device buffer for per-role cycle counters (dp_set_debug_buffer); nullptr in normal operation
extern long long* g_dbg;
extern size_t g_dbg_slots;

// ---- storage-type helpers: 4 consecutive elements <-> float4 ----
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 8 consecutive elements (the 16-byte bf16 vector / two float4 for fp32)
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ f8 ld8(const __nv_bfloat16* p) {
  f8 r;
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const f8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

}  // namespace dp
