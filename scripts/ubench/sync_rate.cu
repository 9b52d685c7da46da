// Microbenchmark: cost of the producer <-> MMA-issuer <-> epilogue handshakes of the conv kernels with no data movement.
//   variant bits: 1 = whole warp polls the full barrier (else lane 0 polls + __syncwarp)
//                 2 = tcgen05.fence::after_thread_sync after the wait
//                 4 = release the stage with tcgen05.commit (else plain mbarrier.arrive by the issuing lane)
//                 8 = issue one MMA per iteration
//                16 = producer issues a real 1-D bulk copy (4 KB) per stage instead of a plain arrive
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../disruption-prediciton-based-on-multimodal-deep-learning_b200/csrc/tc_ptx.cuh"
using namespace dp::ptx;

__device__ int g_wait_mode;   // 0 = library mbar_wait (try_wait loop), 1 = test_wait spin, 2 = try_wait without watchdog
__device__ __forceinline__ void xwait(uint32_t bar, uint32_t parity, int mode) {
  if (mode == 0) { mbar_wait(bar, parity); return; }
  uint32_t done = 0;
  if (mode == 1) {
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
  } else {
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
  }
}

__global__ void __launch_bounds__(128, 1) bench(int variant, int iters, int S, const uint8_t* src, long long* out, int wm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[32];
  __shared__ uint32_t tmem_slot;
  const uint32_t b0 = smem_u32(bars);
  auto full = [&](int i) { return b0 + 8u * i; };
  auto empty = [&](int i) { return b0 + 8u * (8 + i); };
  const uint32_t done = b0 + 8u * 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && (lane == 0 || (variant & 64))) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      xwait(empty(stage), phase ^ 1u, wm);
      if (variant & 16) {
        mbar_expect_tx(full(stage), 4096);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sbase + stage * 4096), "l"(src + (size_t)((blockIdx.x * 64 + (it & 63)) * 4096)), "r"(4096), "r"(full(stage)) : "memory");
      } else {
        if (lane == 0) mbar_arrive(full(stage));
        __syncwarp((variant & 64) ? 0xffffffffu : 1u);
      }
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    const uint32_t hi = smem_desc_hi(1024, 2);
    const uint32_t a_lo = smem_desc_lo(sbase, 16), b_lo = smem_desc_lo(sbase + 65536, 16);
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (variant & 1) xwait(full(stage), phase, wm);
      else { if (lane == 0) xwait(full(stage), phase, wm); __syncwarp(); }
      if (variant & 2) tc_fence_after();
      if (leader) {
        if (variant & 8) umma_bf16_lh(tmem, a_lo, hi, b_lo, hi, idesc, 1u);
        if (variant & 4) umma_commit(empty(stage)); else mbar_arrive(empty(stage));
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (leader) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// epilogue-style handshake: consumer group of NT threads; per tile: wait tfull (1 arrival by tcgen05.commit), all NT threads
// (or one lane per warp) arrive on tempty; the MMA warp waits tempty and commits tfull.
__global__ void __launch_bounds__(384, 1) bench_epi(int per_warp_arrive, int iters, int with_mma, long long* out) {
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b0 = smem_u32(bars);
  auto tfull = [&](int a) { return b0 + 8u * a; };
  auto tempty = [&](int a) { return b0 + 8u * (2 + a); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), per_warp_arrive ? 8 : 256); }
    mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    const uint32_t hi = smem_desc_hi(1024, 2);
    const uint32_t a_lo = smem_desc_lo(sbase, 16), b_lo = smem_desc_lo(sbase + 65536, 16);
    const bool leader = elect_one();
    int acc = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      mbar_wait(tempty(acc), ph ^ 1u);
      tc_fence_after();
      if (leader) {
        if (with_mma) umma_bf16_lh(tmem + acc * 32, a_lo, hi, b_lo, hi, idesc, 0u);
        umma_commit(tfull(acc));
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  } else if (warp >= 4) {
    int acc = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(tfull(acc), ph);
      tc_fence_after();
      tc_fence_before();
      if (per_warp_arrive) { __syncwarp(); if (lane == 0) mbar_arrive(tempty(acc)); }
      else mbar_arrive(tempty(acc));
      if (++acc == 2) { acc = 0; ph ^= 1u; }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 8);
  uint8_t* src; cudaMalloc(&src, (size_t)148 * 64 * 4096);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  printf("producer<->mma handshake: cycles per iteration\n%-8s %-3s %s\n", "variant", "S", "cyc/iter");
  for (int wm : {0, 1, 2})
  for (int S : {2, 8})
    for (int v : {0, 1, 15, 64, 65, 79}) {
      bench<<<148, 128, 200 * 1024>>>(v, iters, S, src, out, wm);
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("wm=%d %-8d %-3d %.1f\n", wm, v, S, (double)c / iters);
    }
  printf("mma<->epilogue handshake (2 accumulators): cycles per tile\n%-16s %-8s %s\n", "per_warp_arrive", "with_mma", "cyc/tile");
  for (int pw : {0, 1})
    for (int wm : {0, 1}) {
      bench_epi<<<148, 384, 200 * 1024>>>(pw, iters, wm, out);
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("%-16d %-8d %.1f\n", pw, wm, (double)c / iters);
    }
  return 0;
}
