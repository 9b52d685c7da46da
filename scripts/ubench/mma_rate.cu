// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) for different N / operand layouts /
// accumulator reuse patterns.  One CTA per SM, one issuing thread, operands = whatever is in smem.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../disruption-prediciton-based-on-multimodal-deep-learning_b200/csrc/tc_ptx.cuh"
using namespace dp::ptx;

// mode 0: K-major SW128 A and B (rows of 128 B), k-steps walk 32 B inside the row, then next 8... (like conv fwd)
// mode 1: MN-major SW128 A and B (like wgrad)
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int mode, int iters, int nacc, int ksteps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {   // whole warp runs the loop (uniform registers), one elected lane issues
    const uint32_t idesc = make_idesc_bf16(M, N, mode, mode);
    const uint32_t hi = smem_desc_hi(1024, 2);
    const uint32_t a_lo = smem_desc_lo(sbase, mode ? 16384 : 16);
    const uint32_t b_lo = smem_desc_lo(sbase + 65536, mode ? 16384 : 16);
    const uint32_t kstep = mode ? (2048 >> 4) : 2;   // MN-major: 16 rows of 128 B; K-major: 32 B inside the row
    const bool leader = elect_one();
    const uint32_t dswap = nacc == 2 ? (uint32_t)N : 0u;
    uint32_t d = tmem;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
        if (ksteps == 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_lh(d, a_lo + k * kstep, hi, b_lo + k * kstep, hi, idesc, 1u);
        } else {
          umma_bf16_lh(d, a_lo, hi, b_lo, hi, idesc, 1u);
        }
      }
      __syncwarp();
      d = (d == tmem) ? tmem + dswap : tmem;
    }
    if (leader) umma_commit(bar_a);
    __syncwarp();
    mbar_wait(bar_a, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  printf("%-6s %-4s %-4s %-5s %-6s %-7s %s\n", "mode", "M", "N", "nacc", "ksteps", "cyc/mma", "math floor (max(M,128)*N/256)");
  for (int mode = 0; mode < 2; ++mode)
    for (int M : {64, 128})
      for (int N : {16, 32, 64, 80, 128, 256})
        for (int nacc : {1, 2})
          for (int ksteps : {1, 4}) {
            if (nacc * N > 512) continue;
            bench<<<148, 128, 200 * 1024>>>(M, N, mode, iters, nacc, ksteps, out);
            long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%-6s %-4d %-4d %-5d %-6d %-7.1f %d\n", mode ? "MN" : "K", M, N, nacc, ksteps, (double)c / (iters * ksteps), (M > 128 ? M : 128) * N / 256);
          }
  return 0;
}
