// dev harness: one instantiation of the thread-per-tile kernel behind a C entry point
#include "fused_tp.cuh"
#include <stdarg.h>
namespace mgp {
static char g_err2[256];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err2, 256, fmt, ap); va_end(ap); }
int check_launch(const char* w) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { set_error("%s: %s", w, cudaGetErrorString(e)); return -3; } return 0; }
int sm_count() { return 148; }
int max_smem_optin() { return 227 * 1024; }
}
extern "C" const char* one_err() { return mgp::g_err2; }

extern "C" int one_run(const double* x, const double* q, const int64_t* nn, const double* y, long long n,
                       long long b, int k, double ls, double noise, double* mean, double* var, void* stream) {
  using namespace mgp;
  mgp_problem p = {};
  p.train_x = x; p.query_x = q; p.nn_idx = nn; p.train_y = y; p.n = n; p.t = b; p.b = b; p.k = k; p.d = 2; p.r = 1;
  p.kernel_id = MGP_KERNEL_MATERN_15; p.metric_id = MGP_METRIC_L2; p.length_scale_count = 1; p.length_scale = &ls;
  p.noise = noise; p.scale = 1.0; p.mean = mean; p.var = var;
  Model m = {}; m.kernel_id = p.kernel_id; m.metric_id = 0; m.d = 2; m.inv_ls = 1.0 / ls;
  TileArgs a; fill_tile_args(&p, m, a);
  ColLoo loo = {}; loo.peers.world = 1;
#ifndef ONE_T
#define ONE_T 7
#endif
  return launch_tp_one<ONE_T, 1, 2>(a, loo, b, nullptr, (cudaStream_t)stream);
}

#ifdef ONE_GRAD
// gradient / coefficient harness: LOO batch (query = train rows), coefficients out
extern "C" int one_run_coeffs(const double* x, const int64_t* bi, const int64_t* nn, const double* y, long long n,
                              long long b, int k, double ls, double noise, double* mean, double* var, double* coeffs,
                              int use_tp, void* stream) {
  using namespace mgp;
  mgp_problem p = {};
  p.train_x = x; p.query_x = x; p.query_idx = bi; p.nn_idx = nn; p.train_y = y; p.n = n; p.t = n; p.b = b; p.k = k; p.d = 2; p.r = 1;
  p.kernel_id = MGP_KERNEL_MATERN_15; p.metric_id = MGP_METRIC_L2; p.length_scale_count = 1; p.length_scale = &ls;
  p.noise = noise; p.scale = 1.0; p.mean = mean; p.var = var; p.coeffs = coeffs;
  Model m = {}; m.kernel_id = p.kernel_id; m.metric_id = 0; m.d = 2; m.inv_ls = 1.0 / ls;
  TileArgs a; fill_tile_args(&p, m, a);
  ColLoo loo = {}; loo.peers.world = 1; loo.backsub = 1;
  if (use_tp) return launch_tp_one<ONE_T, 1, 2, true>(a, loo, b, nullptr, (cudaStream_t)stream);
  return launch_col_one<ONE_T, 1, 2, true>(a, loo, b, nullptr, (cudaStream_t)stream);
}
#endif
#ifdef MGP_TP_TRACE
extern "C" int one_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, mgp::g_tp_trace, sizeof(long long) * 32 * 64);
}
#endif
