// K5 staged single-op kernels: one per reference `_backend_*` hook so every
// stage of the reference pipeline has a 1:1 CUDA replacement for stage-level
// parity tests.  All are streaming, HBM-bound elementwise/gather kernels:
// grid-stride loops, one output element per thread, coalesced stores.
#include "common.cuh"

namespace mgp {

static inline unsigned grid_for(long long work, int threads) {
  long long blocks = (work + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// out[b,j,f] = data[data_idx[b], f] - nn_data[nn_idx[b,j], f]
__global__ void crosswise_diffs_kernel(const double* __restrict__ data,
                                       const double* __restrict__ nn_data,
                                       const int64_t* __restrict__ data_idx,
                                       const int64_t* __restrict__ nn_idx, long long b, int k,
                                       int d, double* __restrict__ out) {
  const long long total = b * k * d;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(e % d);
    const long long bj = e / d;
    const long long row = bj / k;
    const long long q = data_idx ? data_idx[row] : row;
    out[e] = data[q * d + f] - nn_data[nn_idx[bj] * d + f];
  }
}

// out[b,i,j,f] = data[nn[b,i], f] - data[nn[b,j], f]
__global__ void pairwise_diffs_kernel(const double* __restrict__ data,
                                      const int64_t* __restrict__ nn_idx, long long b, int k,
                                      int d, double* __restrict__ out) {
  const long long total = b * k * k * d;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(e % d);
    long long rest = e / d;
    const int j = (int)(rest % k);
    rest /= k;
    const int i = (int)(rest % k);
    const long long row = rest / k;
    const int64_t* nn = nn_idx + row * k;
    out[e] = data[nn[i] * d + f] - data[nn[j] * d + f];
  }
}

struct LsVec {
  int use;
  double inv[MGP_MAX_ANISO_DIM];
};

__global__ void metric_reduce_kernel(int metric_id, const double* __restrict__ diffs,
                                     long long rows, int d, const LsVec ls,
                                     double* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < rows;
       e += (long long)gridDim.x * blockDim.x) {
    const double* v = diffs + e * d;
    double s = 0.0;
    for (int f = 0; f < d; ++f) {
      double x = v[f];
      if (ls.use) x *= ls.inv[f];
      s = fma(x, x, s);
    }
    out[e] = (metric_id == MGP_METRIC_L2) ? sqrt(s) : s;
  }
}

__global__ void crosswise_dists_kernel(int metric_id, const double* __restrict__ data,
                                       const double* __restrict__ nn_data,
                                       const int64_t* __restrict__ data_idx,
                                       const int64_t* __restrict__ nn_idx, long long b, int k,
                                       int d, double* __restrict__ out) {
  const long long total = b * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / k;
    const long long q = data_idx ? data_idx[row] : row;
    const double* x = data + q * d;
    const double* y = nn_data + nn_idx[e] * d;
    double s = 0.0;
    for (int f = 0; f < d; ++f) {
      const double df = x[f] - y[f];
      s = fma(df, df, s);
    }
    out[e] = (metric_id == MGP_METRIC_L2) ? sqrt(s) : s;
  }
}

__global__ void pairwise_dists_kernel(int metric_id, const double* __restrict__ data,
                                      const int64_t* __restrict__ nn_idx, long long b, int k,
                                      int d, double* __restrict__ out) {
  const long long total = b * k * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % k);
    const long long bi = e / k;
    const int i = (int)(bi % k);
    const long long row = bi / k;
    const int64_t* nn = nn_idx + row * k;
    const double* x = data + nn[i] * d;
    const double* y = data + nn[j] * d;
    double s = 0.0;
    for (int f = 0; f < d; ++f) {
      const double df = x[f] - y[f];
      s = fma(df, df, s);
    }
    out[e] = (metric_id == MGP_METRIC_L2) ? sqrt(s) : s;
  }
}

__global__ void kernel_apply_kernel(int kernel_id, const double* __restrict__ in,
                                    double pre_scale, long long count,
                                    double* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = kernel_eval(kernel_id, in[e] * pre_scale);
}

__global__ void perturb_kernel(const double* __restrict__ Kin, long long b, int k, double noise,
                               const double* __restrict__ noise_bk, double* __restrict__ out) {
  const long long total = b * k * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % k);
    const long long bi = e / k;
    const int i = (int)(bi % k);
    double v = Kin[e];
    if (i == j) v += noise_bk ? noise_bk[bi] : noise;
    out[e] = v;
  }
}

// out[b,c] = sum_j Kcross[b,j] * coeffs[b,j,c]; one warp per row
__global__ void rowdot_kernel(const double* __restrict__ Kcross,
                              const double* __restrict__ coeffs, long long b, int k, int r,
                              double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < b; row += nwarps) {
    for (int c = 0; c < r; ++c) {
      double s = 0.0;
      for (int j = lane; j < k; j += 32)
        s = fma(Kcross[row * k + j], coeffs[(row * k + j) * r + c], s);
      s = warp_sum(s);
      if (lane == 0) out[row * r + c] = s;
    }
  }
}

}  // namespace mgp

using namespace mgp;

extern "C" int mgp_crosswise_diffs(const double* data, const double* nn_data,
                                   const int64_t* data_idx, const int64_t* nn_idx, int64_t b,
                                   int32_t k, int32_t d, double* out, void* stream) {
  MGP_REQUIRE(b >= 0 && k >= 1 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(data && nn_data && nn_idx && out, MGP_ERR_BAD_ARG, "null pointer");
  crosswise_diffs_kernel<<<grid_for(b * k * d, 256), 256, 0, (cudaStream_t)stream>>>(
      data, nn_data, data_idx, nn_idx, b, k, d, out);
  return check_launch("crosswise_diffs_kernel");
}

extern "C" int mgp_pairwise_diffs(const double* data, const int64_t* nn_idx, int64_t b,
                                  int32_t k, int32_t d, double* out, void* stream) {
  MGP_REQUIRE(b >= 0 && k >= 1 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(data && nn_idx && out, MGP_ERR_BAD_ARG, "null pointer");
  pairwise_diffs_kernel<<<grid_for(b * k * k * d, 256), 256, 0, (cudaStream_t)stream>>>(
      data, nn_idx, b, k, d, out);
  return check_launch("pairwise_diffs_kernel");
}

extern "C" int mgp_metric_reduce(int32_t metric_id, const double* diffs, int64_t rows, int32_t d,
                                 const double* length_scale, double* out, void* stream) {
  MGP_REQUIRE(metric_id == MGP_METRIC_L2 || metric_id == MGP_METRIC_F2, MGP_ERR_BAD_ARG,
              "unknown metric_id %d", metric_id);
  MGP_REQUIRE(rows >= 0 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (rows == 0) return MGP_OK;
  MGP_REQUIRE(diffs && out, MGP_ERR_BAD_ARG, "null pointer");
  LsVec ls;
  ls.use = 0;
  if (length_scale) {
    MGP_REQUIRE(d <= MGP_MAX_ANISO_DIM, MGP_ERR_UNSUPPORTED,
                "anisotropic deformation supports d <= %d features (got %d)", MGP_MAX_ANISO_DIM,
                d);
    ls.use = 1;
    for (int f = 0; f < d; ++f) {
      MGP_REQUIRE(length_scale[f] > 0.0, MGP_ERR_BAD_ARG, "length scale must be positive");
      ls.inv[f] = 1.0 / length_scale[f];
    }
  }
  metric_reduce_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(metric_id, diffs,
                                                                              rows, d, ls, out);
  return check_launch("metric_reduce_kernel");
}

extern "C" int mgp_crosswise_dists(int32_t metric_id, const double* data, const double* nn_data,
                                   const int64_t* data_idx, const int64_t* nn_idx, int64_t b,
                                   int32_t k, int32_t d, double* out, void* stream) {
  MGP_REQUIRE(metric_id == MGP_METRIC_L2 || metric_id == MGP_METRIC_F2, MGP_ERR_BAD_ARG,
              "unknown metric_id %d", metric_id);
  MGP_REQUIRE(b >= 0 && k >= 1 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(data && nn_data && nn_idx && out, MGP_ERR_BAD_ARG, "null pointer");
  crosswise_dists_kernel<<<grid_for(b * k, 256), 256, 0, (cudaStream_t)stream>>>(
      metric_id, data, nn_data, data_idx, nn_idx, b, k, d, out);
  return check_launch("crosswise_dists_kernel");
}

extern "C" int mgp_pairwise_dists(int32_t metric_id, const double* data, const int64_t* nn_idx,
                                  int64_t b, int32_t k, int32_t d, double* out, void* stream) {
  MGP_REQUIRE(metric_id == MGP_METRIC_L2 || metric_id == MGP_METRIC_F2, MGP_ERR_BAD_ARG,
              "unknown metric_id %d", metric_id);
  MGP_REQUIRE(b >= 0 && k >= 1 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(data && nn_idx && out, MGP_ERR_BAD_ARG, "null pointer");
  pairwise_dists_kernel<<<grid_for(b * k * k, 256), 256, 0, (cudaStream_t)stream>>>(
      metric_id, data, nn_idx, b, k, d, out);
  return check_launch("pairwise_dists_kernel");
}

extern "C" int mgp_kernel_apply(int32_t kernel_id, const double* in, double pre_scale,
                                int64_t count, double* out, void* stream) {
  MGP_REQUIRE(kernel_id >= MGP_KERNEL_RBF && kernel_id <= MGP_KERNEL_MATERN_INF,
              MGP_ERR_BAD_ARG, "unknown kernel_id %d", kernel_id);
  MGP_REQUIRE(count >= 0, MGP_ERR_BAD_ARG, "bad count");
  if (count == 0) return MGP_OK;
  MGP_REQUIRE(in && out, MGP_ERR_BAD_ARG, "null pointer");
  kernel_apply_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      kernel_id, in, pre_scale, count, out);
  return check_launch("kernel_apply_kernel");
}

namespace mgp {
// One warp per neighbourhood row: min / max of labels[nn_idx[row, :]] (column 0 of a strided
// label array) -> 1 when the neighbourhood is NOT constant.  Replaces the (b,k) gather
// `labels[nn_indices]` + two axis reductions of S/optimize/batch.py:58-64,104-110 and
// S/examples/classify.py:577-583 without materialising the gathered labels.
__global__ void nn_label_mask_kernel(const double* __restrict__ labels, long long stride,
                                     const int64_t* __restrict__ nn_idx, long long b, int k,
                                     uint8_t* __restrict__ mask) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= b) return;
  double lo = INFINITY, hi = -INFINITY;
  for (int j = lane; j < k; j += 32) {
    const double v = labels[nn_idx[row * k + j] * stride];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) mask[row] = (hi != lo) ? 1 : 0;
}
}  // namespace mgp

extern "C" int mgp_nn_label_mask(const double* labels, int64_t label_stride,
                                 const int64_t* nn_idx, int64_t b, int32_t k, uint8_t* mask,
                                 void* stream) {
  using namespace mgp;
  MGP_REQUIRE(b >= 0 && k >= 1 && label_stride >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(labels && nn_idx && mask, MGP_ERR_BAD_ARG, "null pointer");
  nn_label_mask_kernel<<<grid_for(b * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      labels, label_stride, nn_idx, b, k, mask);
  return check_launch("nn_label_mask_kernel");
}

extern "C" int mgp_perturb(const double* Kin, int64_t b, int32_t k, double noise,
                           const double* noise_bk, double* out, void* stream) {
  MGP_REQUIRE(b >= 0 && k >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(Kin && out, MGP_ERR_BAD_ARG, "null pointer");
  perturb_kernel<<<grid_for(b * k * k, 256), 256, 0, (cudaStream_t)stream>>>(Kin, b, k, noise,
                                                                             noise_bk, out);
  return check_launch("perturb_kernel");
}

extern "C" int mgp_rowdot(const double* Kcross, const double* coeffs, int64_t b, int32_t k,
                          int32_t r, double* out, void* stream) {
  MGP_REQUIRE(b >= 0 && k >= 1 && r >= 1, MGP_ERR_BAD_ARG, "bad sizes");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(Kcross && coeffs && out, MGP_ERR_BAD_ARG, "null pointer");
  rowdot_kernel<<<grid_for(b * 32, 256), 256, 0, (cudaStream_t)stream>>>(Kcross, coeffs, b, k, r,
                                                                         out);
  return check_launch("rowdot_kernel");
}
