// K5 staged batched SPD solve on materialised (b,k,k) tensors: the 1:1
// replacement for the reference's `_backend_mean_fn`, `_backend_var_fn`,
// `_backend_fast_precompute_fn` and AnalyticScale `_backend_fn` hooks
// (S/_src/gp/muygps/numpy.py:17-67,88-95; S/_src/optimize/scale/numpy.py:9-34).
// Same elimination as the fused kernel, but the matrix is streamed from HBM
// (8 k^2 bytes per row), so this path is HBM-bound and only exists for
// stage-level drop-in parity; the fused entry points are the fast path.
#include "nbhd_smem.cuh"

namespace mgp {

struct SolveArgs {
  const double* Kin;
  const double* Kcross;
  const double* Y;
  double* mean;
  double* var;
  double* yky;
  double* coeffs;
  int32_t* status;
  long long b;
  int k, r, m, ld;
  double kout;
};

template <int WARPS>
__global__ void __launch_bounds__(256) solve_kernel(const SolveArgs a, int teams_per_block,
                                                    size_t team_doubles) {
  extern __shared__ double smem[];
  const int team_in_block = threadIdx.x / (WARPS * 32);
  const int tid = threadIdx.x % (WARPS * 32);
  const int k = a.k, r = a.r, m = a.m, ld = a.ld;
  double* A = smem + (size_t)team_in_block * team_doubles;
  const long long team_global = (long long)blockIdx.x * teams_per_block + team_in_block;
  const long long team_stride = (long long)gridDim.x * teams_per_block;

  for (long long row = team_global; row < a.b; row += team_stride) {
    const double* Kr = a.Kin + (size_t)row * k * k;
    for (int e = tid; e < k * k; e += WARPS * 32) {
      const int i = e / k, j = e - i * k;
      if (j <= i) A[i * ld + j] = Kr[e];
    }
    for (int j = tid; j < k; j += WARPS * 32)
      A[k * ld + j] = a.Kcross ? a.Kcross[row * k + j] : 0.0;
    for (int e = tid; e < r * k; e += WARPS * 32) {
      const int j = e / r, c = e - j * r;
      A[(k + 1 + c) * ld + j] = a.Y[((size_t)row * k + j) * r + c];
    }
    for (int e = tid; e < (r + 1) * (r + 1); e += WARPS * 32) {
      const int i = e / (r + 1), j = e - i * (r + 1);
      if (j <= i) A[(k + i) * ld + k + j] = (i == 0) ? a.kout : 0.0;
    }
    team_sync<WARPS>(team_in_block);
    const bool ok = eliminate<WARPS>(A, ld, k, m, tid, team_in_block);
    team_sync<WARPS>(team_in_block);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (tid == 0) {
      if (a.var) a.var[row] = ok ? A[k * ld + k] : nan;
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
        a.coeffs[((size_t)row * k + j) * r + c] = ok ? A[(k + 1 + c) * ld + j] : nan;
      }
    }
    team_sync<WARPS>(team_in_block);
  }
}

}  // namespace mgp

extern "C" size_t mgp_solve_workspace_bytes(int64_t b, int32_t k, int32_t r) {
  (void)b;
  (void)k;
  (void)r;
  return 0;
}

extern "C" int mgp_solve(const double* Kin, const double* Kcross, const double* Y, int64_t b,
                         int32_t k, int32_t r, double kout, double* mean, double* var,
                         double* yky, double* coeffs, int32_t* status, void* ws, size_t ws_bytes,
                         void* stream) {
  using namespace mgp;
  (void)ws;
  (void)ws_bytes;
  MGP_REQUIRE(b >= 0 && k >= 1 && r >= 0, MGP_ERR_BAD_ARG, "bad sizes b=%lld k=%d r=%d",
              (long long)b, k, r);
  if (b == 0) return MGP_OK;
  MGP_REQUIRE(Kin != nullptr, MGP_ERR_BAD_ARG, "Kin is required");
  if (!Y) r = 0;
  MGP_REQUIRE(Y || (!mean && !yky && !coeffs), MGP_ERR_BAD_ARG,
              "Y is required for mean / yky / coeffs");
  MGP_REQUIRE(Kcross || (!mean && !var), MGP_ERR_BAD_ARG, "Kcross is required for mean / var");
  SolveArgs a;
  a.Kin = Kin;
  a.Kcross = Kcross;
  a.Y = Y;
  a.mean = mean;
  a.var = var;
  a.yky = yky;
  a.coeffs = coeffs;
  a.status = status;
  a.b = b;
  a.k = k;
  a.r = r;
  a.m = k + 1 + r;
  a.ld = a.m | 1;
  a.kout = kout;
  const size_t team_doubles = (size_t)a.m * a.ld;
  const size_t team_bytes = team_doubles * sizeof(double);
  const size_t smem_max = (size_t)max_smem_optin();
  MGP_REQUIRE(team_bytes <= smem_max, MGP_ERR_UNSUPPORTED,
              "k=%d, r=%d needs %zu bytes of shared memory (max %zu)", k, r, team_bytes,
              smem_max);
  const int warps = (a.m <= 64) ? 1 : (a.m <= 128 ? 4 : 8);
  int teams = (int)(smem_max / 2 / team_bytes);
  if (teams < 1) teams = 1;
  if (teams > 256 / (warps * 32)) teams = 256 / (warps * 32);
  if (teams > 15) teams = 15;
  const size_t smem = team_bytes * teams;
  long long blocks = (b + teams - 1) / teams;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
#define MGP_LAUNCH(W)                                                                         \
  do {                                                                                        \
    cudaFuncSetAttribute(solve_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                         (int)smem_max);                                                      \
    solve_kernel<W><<<(unsigned)blocks, teams * W * 32, smem, s>>>(a, teams, team_doubles);   \
  } while (0)
  if (warps == 1)
    MGP_LAUNCH(1);
  else if (warps == 4)
    MGP_LAUNCH(4);
  else
    MGP_LAUNCH(8);
#undef MGP_LAUNCH
  return check_launch("solve_kernel");
}
