// Microbenchmark: issue cost of single mbarrier operations by one thread (no contention).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../disruption-prediciton-based-on-multimodal-deep-learning_b200/csrc/tc_ptx.cuh"
using namespace dp::ptx;
__device__ __forceinline__ uint32_t try_once(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
__global__ void k(int mode, int iters, int nthreads, long long* out) {
  __shared__ uint64_t bars[16];
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(b0 + 8 * i, 500000); mbar_fence_init(); }
  __syncthreads();
  uint32_t acc = 0;
  if ((int)threadIdx.x < nthreads) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) mbar_arrive(b0 + 8 * (it & 7));
      else if (mode == 1) acc += try_once(b0 + 8 * (it & 7), 1);            // completed parity: returns true at once
      else if (mode == 2) { mbar_arrive(b0 + 8 * (it & 7)); acc += try_once(b0 + 64 + 8 * (it & 7), 1); }
      else if (mode == 3) { acc += try_once(b0 + 8 * (it & 7), 1); __syncwarp(); }
      else if (mode == 4) { mbar_wait(b0 + 8 * (it & 7), 1); }
      else if (mode == 5) { mbar_arrive(b0 + 8 * (it & 7)); mbar_wait(b0 + 64 + 8 * (it & 7), 1); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
  }
}
int main() {
  long long* out; cudaMalloc(&out, 16);
  const char* names[] = {"arrive", "try_wait(complete)", "arrive+try_wait", "try_wait+syncwarp", "mbar_wait(complete)", "arrive+mbar_wait"};
  for (int nt : {1, 32})
    for (int m = 0; m < 6; ++m) {
      k<<<148, 128>>>(m, 4000, nt, out);
      long long c[2]; cudaError_t e = cudaMemcpy(c, out, 16, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("threads=%-2d %-22s %.1f cyc/iter\n", nt, names[m], (double)c[0] / 4000);
    }
  return 0;
}
