// a14/a15: loss and analytic-scale partial sums over a batch of rows
// (S/_src/optimize/loss/numpy.py:12-112, S/_src/optimize/scale/numpy.py:9-34).
// Two-stage fixed-order reduction -> bitwise reproducible, no fp atomics.
// Every slot is a plain sum so N ranks combine their records with ONE SUM
// all-reduce (replacing S/_src/optimize/loss/mpi.py and scale/mpi.py).
#include "common.cuh"

namespace mgp {

constexpr int LOSS_THREADS = 256;

struct LossArgs {
  int loss_id;
  const double* pred;
  const double* targ;
  const double* var;
  const double* yky;
  const double* scale_dev;
  double delta;
  long long b;
  int r;
};

__device__ __forceinline__ void row_terms(const LossArgs& a, long long row, double* acc) {
  const int r = a.r;
  const double* p = a.pred + row * r;
  const double* t = a.targ ? a.targ + row * r : nullptr;
  bool bad = false;
  if (t) {
    double se = 0.0;
    for (int c = 0; c < r; ++c) {
      const double e = p[c] - t[c];
      se = fma(e, e, se);
      bad |= !(p[c] == p[c]);
    }
    acc[MGP_P_SQERR] += se;
    acc[MGP_P_COUNT] += (double)r;
  }
  acc[MGP_P_ROWS] += 1.0;
  if (a.yky) acc[MGP_P_YKY] += a.yky[row];
  const double sigma2 = a.scale_dev ? *a.scale_dev : 1.0;
  switch (a.loss_id) {
    case MGP_LOSS_LOOL: {  // numpy.py:34-61, variances.ndim == 1 branch
      const double v = a.var[row], e = p[0] - t[0];
      acc[MGP_P_SQERR_V] += e * e / v;
      acc[MGP_P_LOGV] += log(v);
      const double sv = sigma2 * v;
      acc[MGP_P_AUX] += e * e / sv + log(sv);
      break;
    }
    case MGP_LOSS_LOOPH: {  // numpy.py:75-112
      const double v = a.var[row], e = t[0] - p[0];
      acc[MGP_P_SQERR_V] += e * e / v;
      acc[MGP_P_LOGV] += log(v);
      const double sv = sigma2 * v, d2 = a.delta * a.delta;
      acc[MGP_P_AUX] += 2.0 * d2 * (sqrt(1.0 + e * e / (d2 * sv)) - 1.0) + log(sv);
      break;
    }
    case MGP_LOSS_PSEUDO_HUBER: {  // numpy.py:64-72 (delta^2 applied per term)
      double s = 0.0;
      for (int c = 0; c < r; ++c) {
        const double e = (t[c] - p[c]) / a.delta;
        s += sqrt(1.0 + e * e) - 1.0;
      }
      acc[MGP_P_AUX] += a.delta * a.delta * s;
      break;
    }
    case MGP_LOSS_CROSS_ENTROPY: {  // numpy.py:12-19 (sklearn log_loss, eps clip)
      double mx = p[0];
      for (int c = 1; c < r; ++c) mx = fmax(mx, p[c]);
      double z = 0.0;
      for (int c = 0; c < r; ++c) z += exp(p[c] - mx);
      const double eps = 2.220446049250313e-16;
      double s = 0.0;
      for (int c = 0; c < r; ++c) {
        if (t[c] > 0.0) {
          double sm = exp(p[c] - mx) / z;
          sm = fmin(fmax(sm, eps), 1.0 - eps);
          s -= log(sm);
        }
      }
      acc[MGP_P_AUX] += s;
      break;
    }
    default:
      break;
  }
  if (bad) acc[MGP_P_BAD] += 1.0;
}

__device__ __forceinline__ void block_reduce8(double* acc, double* out8) {
  __shared__ double sh[LOSS_THREADS / 32][MGP_PARTIALS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < MGP_PARTIALS; ++s) {
    double v = acc[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < MGP_PARTIALS) {
    double v = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) v += sh[w][threadIdx.x];
    out8[threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_stage1(const LossArgs a,
                                                            double* __restrict__ block_out) {
  double acc[MGP_PARTIALS];
#pragma unroll
  for (int s = 0; s < MGP_PARTIALS; ++s) acc[s] = 0.0;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < a.b;
       row += (long long)gridDim.x * blockDim.x)
    row_terms(a, row, acc);
  block_reduce8(acc, block_out + (size_t)blockIdx.x * MGP_PARTIALS);
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_stage2(const double* __restrict__ block_out,
                                                            int nblocks,
                                                            double* __restrict__ partials) {
  double acc[MGP_PARTIALS];
#pragma unroll
  for (int s = 0; s < MGP_PARTIALS; ++s) acc[s] = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
#pragma unroll
    for (int s = 0; s < MGP_PARTIALS; ++s) acc[s] += block_out[(size_t)i * MGP_PARTIALS + s];
  __shared__ double fin[MGP_PARTIALS];
  block_reduce8(acc, fin);
  __syncthreads();
  if (threadIdx.x < MGP_PARTIALS) partials[threadIdx.x] += fin[threadIdx.x];
}

static int loss_blocks(long long b) {
  long long blocks = (b + LOSS_THREADS - 1) / LOSS_THREADS;
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mgp

using namespace mgp;

extern "C" size_t mgp_loss_workspace_bytes(int64_t b, int32_t r) {
  (void)r;
  return (size_t)loss_blocks(b) * MGP_PARTIALS * sizeof(double);
}

extern "C" int mgp_loss_partials(int32_t loss_id, const double* pred, const double* targets,
                                 const double* var, const double* yky, const double* scale_dev,
                                 double boundary_scale, int64_t b, int32_t r, double* partials,
                                 void* ws, size_t ws_bytes, void* stream) {
  MGP_REQUIRE(loss_id >= MGP_LOSS_NONE && loss_id <= MGP_LOSS_CROSS_ENTROPY, MGP_ERR_BAD_ARG,
              "unknown loss_id %d", loss_id);
  MGP_REQUIRE(b >= 0 && r >= 1, MGP_ERR_BAD_ARG, "bad sizes b=%lld r=%d", (long long)b, r);
  MGP_REQUIRE(partials != nullptr, MGP_ERR_BAD_ARG, "partials is required");
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(pred != nullptr, MGP_ERR_BAD_ARG, "pred is required");
  MGP_REQUIRE(loss_id == MGP_LOSS_NONE || targets != nullptr, MGP_ERR_BAD_ARG,
              "targets are required for loss_id %d", loss_id);
  if (loss_id == MGP_LOSS_LOOL || loss_id == MGP_LOSS_LOOPH) {
    MGP_REQUIRE(var != nullptr, MGP_ERR_BAD_ARG, "variances are required for lool / looph");
    // the reference's multivariate branches are out of scope
    MGP_REQUIRE(r == 1, MGP_ERR_UNSUPPORTED,
                "looph does not yet support multivariate inference (r=%d)", r);
  }
  if (loss_id == MGP_LOSS_LOOPH || loss_id == MGP_LOSS_PSEUDO_HUBER)
    MGP_REQUIRE(boundary_scale > 0.0, MGP_ERR_BAD_ARG, "boundary_scale must be positive");
  const int blocks = loss_blocks(b);
  MGP_REQUIRE(ws != nullptr && ws_bytes >= (size_t)blocks * MGP_PARTIALS * sizeof(double),
              MGP_ERR_WORKSPACE, "loss workspace too small (%zu bytes)", ws_bytes);
  LossArgs a;
  a.loss_id = loss_id;
  a.pred = pred;
  a.targ = targets;
  a.var = var;
  a.yky = yky;
  a.scale_dev = scale_dev;
  a.delta = boundary_scale;
  a.b = b;
  a.r = r;
  cudaStream_t s = (cudaStream_t)stream;
  loss_stage1<<<blocks, LOSS_THREADS, 0, s>>>(a, (double*)ws);
  loss_stage2<<<1, LOSS_THREADS, 0, s>>>((const double*)ws, blocks, partials);
  return check_launch("loss kernels");
}
