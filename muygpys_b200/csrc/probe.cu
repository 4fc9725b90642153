// FP64 roofline probe.  MEASURED_PEAKS.json carries HBM and bf16 peaks only; the
// fused neighbourhood kernel is FP64-bound, so bench.py measures the FP64 DFMA
// and DMMA (mma.sync.m8n8k4.f64) issue rates on the box with these kernels.
#include "common.cuh"

namespace mgp {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__global__ void probe_kernel(int mode, int iters, double* __restrict__ sink) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5;
  double x = 1.0 + 1e-9 * (gid & 7), y = 1e-12;
  int m = mode;
  if (mode == 2) m = (warp & 1);
  if (m == 0) {
    double a0 = 0, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
      a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    sink[gid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  } else if (m == 1) {
    double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
      dmma(c[0], c[1], x, y); dmma(c[2], c[3], x, y);
      dmma(c[4], c[5], x, y); dmma(c[6], c[7], x, y);
    }
    sink[gid] = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
  } else if (m == 5) {  // shuffle issue rate: 8 independent SHFL.IDX per iteration
    int v0 = gid, v1 = gid + 1, v2 = gid + 2, v3 = gid + 3, v4 = gid + 4, v5 = gid + 5,
        v6 = gid + 6, v7 = gid + 7;
    const int src = (threadIdx.x * 5 + 3) & 31;
    for (int i = 0; i < iters; ++i) {
      v0 = __shfl_sync(0xffffffffu, v0, src); v1 = __shfl_sync(0xffffffffu, v1, src);
      v2 = __shfl_sync(0xffffffffu, v2, src); v3 = __shfl_sync(0xffffffffu, v3, src);
      v4 = __shfl_sync(0xffffffffu, v4, src); v5 = __shfl_sync(0xffffffffu, v5, src);
      v6 = __shfl_sync(0xffffffffu, v6, src); v7 = __shfl_sync(0xffffffffu, v7, src);
    }
    sink[gid] = (double)(v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7);
  } else if (m == 6) {  // shuffle latency: one dependent chain
    int v0 = gid;
    for (int i = 0; i < iters; ++i) v0 = __shfl_sync(0xffffffffu, v0, (v0 + 1) & 31);
    sink[gid] = (double)v0;
  } else if (m == 3) {
    double a0 = 0;
    for (int i = 0; i < iters; ++i) a0 = fma(a0, x, y);
    sink[gid] = a0;
  } else {
    double c0 = 0, c1 = 0;
    for (int i = 0; i < iters; ++i) dmma(c0, c1, x, y);
    sink[gid] = c0 + c1;
  }
}

}  // namespace mgp

extern "C" int mgp_fp64_probe(int32_t mode, int32_t blocks, int32_t threads, int32_t iters,
                              double* sink, void* stream) {
  using namespace mgp;
  MGP_REQUIRE(mode >= 0 && mode <= 6, MGP_ERR_BAD_ARG, "unknown probe mode %d", mode);
  MGP_REQUIRE(blocks >= 1 && threads >= 32 && threads <= 1024 && (threads % 32) == 0 &&
                  iters >= 1 && sink,
              MGP_ERR_BAD_ARG, "bad probe launch parameters");
  probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(mode, iters, sink);
  return check_launch("probe_kernel");
}
