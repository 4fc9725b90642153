// K4: fast posterior mean apply (a12) -- crosswise distance + kernel + dot with
// a precomputed coefficient row, one warp per test point.  Replaces
// fast_posterior_mean_from_indices (S/examples/from_indices.py:93-123), i.e.
// crosswise_tensor -> kernel -> einsum('ij,ijk->ik') (S/_src/gp/muygps/numpy.py:70-77)
// without materialising Kcross or gathering coeffs[closest].  HBM/gather-bound:
// per point 8k (ids) + 8kd (rows) + 8kr (coefficients) + 8d + 8 bytes read.
#include "common.cuh"

namespace mgp {

struct FastArgs {
  const double* train_x;
  const double* query_x;
  const int64_t* query_idx;
  const int64_t* nn_idx;
  const int64_t* coeff_row;
  const double* coeffs;
  double* mean;
  long long b;
  int k, d, r;
  Model model;
};

// k <= 64: lane l owns neighbours l and l + 32.  The kernel is bound by dependent gathers
// (index -> feature row), so the indices of the NEXT test point are loaded while the current one
// is evaluated and both neighbours of a lane are in flight together.
__global__ void __launch_bounds__(256) fast_mean_k64_kernel(const FastArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int k = a.k, d = a.d, r = a.r;
  const bool has0 = lane < k, has1 = lane + 32 < k;
  auto load_ids = [&](long long row, long long& i0, long long& i1, long long& q, long long& c) {
    if (row >= a.b) return;
    i0 = has0 ? a.nn_idx[row * k + lane] : 0;
    i1 = has1 ? a.nn_idx[row * k + lane + 32] : 0;
    q = a.query_idx ? a.query_idx[row] : row;
    c = a.coeff_row ? a.coeff_row[row] : row;
  };
  long long i0 = 0, i1 = 0, q = 0, crow = 0;
  load_ids(warp, i0, i1, q, crow);
  for (long long row = warp; row < a.b; row += nwarps) {
    long long n0 = 0, n1 = 0, nq = 0, nc = 0;
    load_ids(row + nwarps, n0, n1, nq, nc);
    const double* xq = a.query_x + q * d;
    const double* y0 = a.train_x + i0 * d;
    const double* y1 = a.train_x + i1 * d;
    double s0 = 0.0, s1 = 0.0;
    for (int f = 0; f < d; ++f) {
      double d0 = xq[f] - y0[f], d1 = xq[f] - y1[f];
      if (a.model.aniso) {
        d0 *= a.model.inv_ls_vec[f];
        d1 *= a.model.inv_ls_vec[f];
      }
      s0 = fma(d0, d0, s0);
      s1 = fma(d1, d1, s1);
    }
    const double k0 = has0 ? kernel_eval(a.model.kernel_id, finish_distance(a.model, s0)) : 0.0;
    const double k1 = has1 ? kernel_eval(a.model.kernel_id, finish_distance(a.model, s1)) : 0.0;
    const double* c0p = a.coeffs + (crow * k + (has0 ? lane : 0)) * r;
    const double* c1p = a.coeffs + (crow * k + (has1 ? lane + 32 : 0)) * r;
    for (int c = 0; c < r; ++c) {
      const double v = warp_sum(fma(k0, c0p[c], k1 * c1p[c]));
      if (lane == 0) a.mean[row * r + c] = v;
    }
    i0 = n0;
    i1 = n1;
    q = nq;
    crow = nc;
  }
}

__global__ void __launch_bounds__(256) fast_mean_kernel(const FastArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int k = a.k, d = a.d, r = a.r;
  for (long long row = warp; row < a.b; row += nwarps) {
    const long long q = a.query_idx ? a.query_idx[row] : row;
    const long long crow = a.coeff_row ? a.coeff_row[row] : row;
    const double* xq = a.query_x + q * d;
    for (int c0 = 0; c0 < r; c0 += 4) {  // up to 4 responses per sweep over neighbours
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int j = lane; j < k; j += 32) {
        const double* y = a.train_x + a.nn_idx[row * k + j] * d;
        double s = 0.0;
        for (int f = 0; f < d; ++f) {
          double df = xq[f] - y[f];
          if (a.model.aniso) df *= a.model.inv_ls_vec[f];
          s = fma(df, df, s);
        }
        const double kv = kernel_eval(a.model.kernel_id, finish_distance(a.model, s));
        const double* cf = a.coeffs + (crow * k + j) * r + c0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c0 + c < r) acc[c] = fma(kv, cf[c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double v = warp_sum(acc[c]);
        if (lane == 0 && c0 + c < r) a.mean[row * r + c0 + c] = v;
      }
    }
  }
}

int validate_problem(const mgp_problem* p);

}  // namespace mgp

extern "C" int mgp_fast_mean(const mgp_problem* p, const int64_t* coeff_row,
                             const double* coeffs, void* stream) {
  using namespace mgp;
  MGP_REQUIRE(p != nullptr, MGP_ERR_BAD_ARG, "null problem");
  MGP_REQUIRE(p->k >= 1 && p->d >= 1 && p->r >= 1 && p->b >= 0, MGP_ERR_BAD_ARG, "bad sizes");
  Model model;
  int rc = make_model(p->kernel_id, p->metric_id, p->d, p->length_scale_count, p->length_scale,
                      &model);
  if (rc != MGP_OK) return rc;
  if (p->b == 0) return MGP_OK;
  MGP_REQUIRE(p->train_x && p->query_x && p->nn_idx && coeffs && p->mean, MGP_ERR_BAD_ARG,
              "train_x, query_x, nn_idx, coeffs and mean are required");
  FastArgs a;
  a.train_x = p->train_x;
  a.query_x = p->query_x;
  a.query_idx = p->query_idx;
  a.nn_idx = p->nn_idx;
  a.coeff_row = coeff_row;
  a.coeffs = coeffs;
  a.mean = p->mean;
  a.b = p->b;
  a.k = p->k;
  a.d = p->d;
  a.r = p->r;
  a.model = model;
  long long blocks = (p->b * 32 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (p->k <= 64)
    fast_mean_k64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  else
    fast_mean_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("fast_mean_kernel");
}
