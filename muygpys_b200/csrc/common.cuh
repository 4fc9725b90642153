// Shared device/host helpers for the muygpys_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/muygpys_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "muygpys_b200 is written for sm_100a (B200) only"
#endif

namespace mgp {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> MGP_OK / MGP_ERR_CUDA
int sm_count();
int max_smem_optin();

#define MGP_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::mgp::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// ---- model parameters passed by value to kernels ---------------------------
struct Model {
  int kernel_id;
  int metric_id;
  int aniso;      // 0: one length scale, applied after the metric; 1: per-feature
  int d;
  double inv_ls;  // iso: 1/l (L2) or 1/l^2 (F2)
  double inv_ls_vec[MGP_MAX_ANISO_DIM];  // aniso: 1/l_f
};

int make_model(int kernel_id, int metric_id, int d, int length_scale_count,
               const double* length_scale_host, Model* out);

// ---- covariance functions of the scaled distance ---------------------------
// S/_src/gp/kernels/numpy.py:12-31 (reference semantics; x already divided by
// the length scale, RBF takes the squared form).
__device__ __forceinline__ double kernel_eval(int kernel_id, double x) {
  switch (kernel_id) {
    case MGP_KERNEL_RBF:
      return exp(-0.5 * x);
    case MGP_KERNEL_MATERN_05:
      return exp(-x);
    case MGP_KERNEL_MATERN_15: {
      const double s = x * 1.7320508075688772;  // sqrt(3)
      return (1.0 + s) * exp(-s);
    }
    case MGP_KERNEL_MATERN_25: {
      const double s = x * 2.23606797749979;  // sqrt(5)
      return (1.0 + s + s * s / 3.0) * exp(-s);
    }
    default:  // MGP_KERNEL_MATERN_INF
      return exp(-0.5 * x * x);
  }
}

// Turn an accumulated sum of squared (optionally per-feature scaled)
// differences into the kernel's argument.
__device__ __forceinline__ double finish_distance(const Model& m, double sumsq) {
  double x = (m.metric_id == MGP_METRIC_L2) ? sqrt(sumsq) : sumsq;
  if (!m.aniso) x *= m.inv_ls;
  return x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mgp
