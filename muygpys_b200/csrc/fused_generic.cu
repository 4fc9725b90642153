// K1 (generic shape) -- fused gather -> distance -> kernel -> nugget ->
// Cholesky elimination -> posterior mean / variance / scale term / coefficients.
//
// One team of WARPS warps per neighbourhood, augmented matrix resident in shared
// memory (see nbhd_smem.cuh).  This variant accepts every (k, d, r); the
// register-resident DMMA variant in fused_tile.cu takes over for the shapes it
// supports.  Nothing of size (b,k,k[,d]) is ever written to HBM: per
// neighbourhood the kernel reads 8k (indices) + 8kd (rows) + 8kr (targets) + 8d
// (query) bytes and writes 8(r+2) bytes.
#include "nbhd_smem.cuh"

namespace mgp {

struct FusedArgs {
  const double* train_x;
  const double* query_x;
  const int64_t* query_idx;
  const int64_t* nn_idx;
  const double* train_y;
  const double* noise_bk;
  double* mean;
  double* var;
  double* yky;
  double* coeffs;
  int32_t* status;
  long long b;
  int k, d, r, m, ld, dchunk;
  double noise, scale;
  Model model;
};

template <int WARPS>
__global__ void __launch_bounds__(256) fused_generic_kernel(const FusedArgs a, int teams_per_block,
                                                            size_t team_doubles) {
  extern __shared__ double smem[];
  const int team_in_block = threadIdx.x / (WARPS * 32);
  const int tid = threadIdx.x % (WARPS * 32);
  const int lane = tid & 31, warp = tid >> 5;
  const int k = a.k, d = a.d, r = a.r, m = a.m, ld = a.ld, dc = a.dchunk;

  double* A = smem + (size_t)team_in_block * team_doubles;
  double* Xs = A + (size_t)m * ld;                    // (k+1) x dc staging, row k = query
  long long* idx = (long long*)(Xs + (size_t)(k + 1) * dc);  // k neighbour ids

  const long long team_global = (long long)blockIdx.x * teams_per_block + team_in_block;
  const long long team_stride = (long long)gridDim.x * teams_per_block;

  for (long long row = team_global; row < a.b; row += team_stride) {
    const long long q = a.query_idx ? a.query_idx[row] : row;
    for (int j = tid; j < k; j += WARPS * 32) idx[j] = a.nn_idx[row * k + j];
    // zero the distance accumulators (lower triangle of K and the kcross row)
    for (int i = warp; i <= k; i += WARPS)
      for (int j = lane; j <= min(i, k - 1); j += 32) A[i * ld + j] = 0.0;
    team_sync<WARPS>(team_in_block);

    // ---- squared (scaled) distances, feature chunk by feature chunk --------
    for (int f0 = 0; f0 < d; f0 += dc) {
      const int fc = min(dc, d - f0);
      for (int e = tid; e < (k + 1) * fc; e += WARPS * 32) {
        const int i = e / fc, f = e - i * fc;
        double v = (i < k) ? a.train_x[idx[i] * d + f0 + f] : a.query_x[q * d + f0 + f];
        if (a.model.aniso) v *= a.model.inv_ls_vec[f0 + f];
        Xs[i * dc + f] = v;
      }
      team_sync<WARPS>(team_in_block);
      for (int i = warp; i <= k; i += WARPS) {
        const double* xi = Xs + i * dc;
        for (int j = lane; j <= min(i, k - 1); j += 32) {
          const double* xj = Xs + j * dc;
          double s = 0.0;
          for (int f = 0; f < fc; ++f) {
            const double df = xi[f] - xj[f];
            s = fma(df, df, s);
          }
          A[i * ld + j] += s;
        }
      }
      team_sync<WARPS>(team_in_block);
    }

    // ---- covariance, nugget, augmented rows --------------------------------
    for (int i = warp; i <= k; i += WARPS) {
      for (int j = lane; j <= min(i, k - 1); j += 32) {
        double v = kernel_eval(a.model.kernel_id, finish_distance(a.model, A[i * ld + j]));
        if (i == j) v += a.noise_bk ? a.noise_bk[row * k + i] : a.noise;
        A[i * ld + j] = v;
      }
    }
    for (int e = tid; e < r * k; e += WARPS * 32) {
      const int j = e / r, c = e - j * r;
      A[(k + 1 + c) * ld + j] = a.train_y ? a.train_y[idx[j] * r + c] : 0.0;
    }
    // trailing block: kout = 1 (RBF/Matern Kout(), S/gp/kernels/rbf.py:113-114), zeros
    for (int e = tid; e < (r + 1) * (r + 1); e += WARPS * 32) {
      const int i = e / (r + 1), j = e - i * (r + 1);
      if (j <= i) A[(k + i) * ld + k + j] = (i == 0) ? 1.0 : 0.0;
    }
    team_sync<WARPS>(team_in_block);

    const bool ok = eliminate<WARPS>(A, ld, k, m, tid, team_in_block);
    team_sync<WARPS>(team_in_block);

    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (tid == 0) {
      if (a.var) a.var[row] = ok ? a.scale * A[k * ld + k] : nan;
      if (a.status) a.status[row] = ok ? 0 : 1;
      if (a.yky) {
        double s = 0.0;
        for (int c = 0; c < r; ++c) s -= A[(k + 1 + c) * ld + k + 1 + c];
        a.yky[row] = ok ? s : nan;
      }
    }
    if (a.mean)
      for (int c = tid; c < r; c += WARPS * 32)
        a.mean[row * r + c] = ok ? -A[(k + 1 + c) * ld + k] : nan;
    if (a.coeffs) {
      if (ok) back_substitute<WARPS>(A, ld, k, r, tid, team_in_block);
      team_sync<WARPS>(team_in_block);
      for (int e = tid; e < k * r; e += WARPS * 32) {
        const int j = e / r, c = e - j * r;
        a.coeffs[(row * k + j) * r + c] = ok ? A[(k + 1 + c) * ld + j] : nan;
      }
    }
    team_sync<WARPS>(team_in_block);
  }
}

static int launch_generic(const mgp_problem* p, const Model& model, cudaStream_t stream) {
  FusedArgs a;
  a.train_x = p->train_x;
  a.query_x = p->query_x;
  a.query_idx = p->query_idx;
  a.nn_idx = p->nn_idx;
  a.train_y = p->train_y;
  a.noise_bk = p->noise_bk;
  a.mean = p->mean;
  a.var = p->var;
  a.yky = p->yky;
  a.coeffs = p->coeffs;
  a.status = p->status;
  a.b = p->b;
  a.k = p->k;
  a.d = p->d;
  a.r = p->r;
  a.m = p->k + 1 + p->r;
  a.ld = a.m | 1;  // odd leading dimension: column walks are bank-conflict free
  a.noise = p->noise;
  a.scale = p->scale;
  a.model = model;
  // stage at most ~1024 doubles of features per chunk
  int dc = 1024 / (p->k + 1);
  if (dc < 1) dc = 1;
  if (dc > p->d) dc = p->d;
  a.dchunk = dc;

  size_t team_doubles = (size_t)a.m * a.ld + (size_t)(p->k + 1) * dc + p->k;
  size_t team_bytes = team_doubles * sizeof(double);
  const size_t smem_max = (size_t)max_smem_optin();
  MGP_REQUIRE(team_bytes <= smem_max, MGP_ERR_UNSUPPORTED,
              "neighbourhood of k=%d, r=%d needs %zu bytes of shared memory (max %zu)", p->k,
              p->r, team_bytes, smem_max);
  const int warps = (a.m <= 64) ? 1 : (a.m <= 128 ? 4 : 8);
  int teams = (int)(smem_max / 2 / team_bytes);  // aim for two resident blocks per SM
  if (teams < 1) teams = 1;
  const int max_teams = 256 / (warps * 32);
  if (teams > max_teams) teams = max_teams;
  if (teams > 15) teams = 15;  // named barriers 1..15
  const size_t smem = team_bytes * teams;
  long long blocks = (p->b + teams - 1) / teams;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;

#define MGP_LAUNCH(W)                                                                       \
  do {                                                                                      \
    cudaFuncSetAttribute(fused_generic_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem_max);                                                    \
    fused_generic_kernel<W><<<(unsigned)blocks, teams * W * 32, smem, stream>>>(a, teams,   \
                                                                                team_doubles); \
  } while (0)
  if (warps == 1)
    MGP_LAUNCH(1);
  else if (warps == 4)
    MGP_LAUNCH(4);
  else
    MGP_LAUNCH(8);
#undef MGP_LAUNCH
  return check_launch("fused_generic_kernel");
}

int fused_variant();
int fused_col_supported(const mgp_problem* p, const Model& model);
int launch_fused_col(const mgp_problem* p, const Model& model, cudaStream_t stream);
int fused_tile_supported(const mgp_problem* p, const Model& model);
int launch_fused_tile(const mgp_problem* p, const Model& model, void* ws, size_t ws_bytes,
                      cudaStream_t stream);

int validate_problem(const mgp_problem* p) {
  MGP_REQUIRE(p != nullptr, MGP_ERR_BAD_ARG, "null problem");
  MGP_REQUIRE(p->b >= 0 && p->n >= 0 && p->t >= 0, MGP_ERR_BAD_ARG, "negative size");
  MGP_REQUIRE(p->k >= 1, MGP_ERR_BAD_ARG, "nn_count k=%d must be >= 1", p->k);
  MGP_REQUIRE(p->r >= 1, MGP_ERR_BAD_ARG, "response count r=%d must be >= 1", p->r);
  MGP_REQUIRE(p->d >= 1, MGP_ERR_BAD_ARG, "feature count d=%d must be >= 1", p->d);
  if (p->b == 0) return MGP_OK;
  MGP_REQUIRE(p->train_x && p->query_x && p->nn_idx, MGP_ERR_BAD_ARG,
              "train_x, query_x and nn_idx are required");
  MGP_REQUIRE(p->train_y || (!p->mean && !p->coeffs && !p->yky), MGP_ERR_BAD_ARG,
              "train_y is required for mean / yky / coeffs outputs");
  return MGP_OK;
}

}  // namespace mgp

extern "C" size_t mgp_fused_workspace_bytes(const mgp_problem* p) {
  (void)p;
  return 0;
}

extern "C" int mgp_fused_posterior(const mgp_problem* p, void* ws, size_t ws_bytes,
                                   void* stream) {
  int rc = mgp::validate_problem(p);
  if (rc != MGP_OK) return rc;
  mgp::Model model;
  rc = mgp::make_model(p->kernel_id, p->metric_id, p->d, p->length_scale_count, p->length_scale,
                       &model);
  if (rc != MGP_OK) return rc;
  if (p->b == 0) return MGP_OK;
  const int variant = mgp::fused_variant();
  if ((variant == 0 || variant == 3 || variant == 4) && mgp::fused_col_supported(p, model))
    return mgp::launch_fused_col(p, model, (cudaStream_t)stream);
  if (variant == 3 || variant == 4) {
    mgp::set_error("the column-direct kernel does not support this shape");
    return MGP_ERR_UNSUPPORTED;
  }
  if (mgp::fused_tile_supported(p, model))
    return mgp::launch_fused_tile(p, model, ws, ws_bytes, (cudaStream_t)stream);
  return mgp::launch_generic(p, model, (cudaStream_t)stream);
}
