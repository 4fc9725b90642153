// Dev probe: what does a warp do after BAR.SYNC.DEFER_BLOCKING while the barrier is incomplete?
// Warp 1 arrives ~20 000 cycles late; warp 0 stamps the clock after bar.sync, after a run of
// register-only arithmetic, after a shared-memory load and after a global store.
// nvcc -arch=sm_100a -o bar_defer_probe bar_defer_probe.cu && ./bar_defer_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
__global__ void probe(long long* out, double* sink) {
  __shared__ double sh[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sh[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (warp == 1) {
    const long long t0 = clk();
    while (clk() - t0 < 20000) {
    }
    if (lane == 0) out[8] = clk();
    asm volatile("bar.arrive 1, 64;" ::: "memory");
  } else {
    double x = lane * 1e-3 + 1.0;
    const long long t0 = clk();
    asm volatile("bar.sync 1, 64;" ::: "memory");
    const long long t1 = clk();
#pragma unroll
    for (int i = 0; i < 200; ++i) x = fma(x, 1.0000001, 1e-9);
    const long long t2 = clk();
    const double v = sh[lane + 32];
    const long long t3 = clk();
    x += v;
    sink[lane] = x;
    const long long t4 = clk();
    if (lane == 0) {
      out[0] = t0; out[1] = t1; out[2] = t2; out[3] = t3; out[4] = t4;
    }
  }
}
int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 64 * sizeof(double));
  for (int rep = 0; rep < 2; ++rep) probe<<<1, 64>>>(out, sink);
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("after bar.sync: +%lld  after 200 FMAs: +%lld  after LDS: +%lld  after STG: +%lld  (late warp arrives at +%lld)\n",
         h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[8] - h[0]);
  return 0;
}
