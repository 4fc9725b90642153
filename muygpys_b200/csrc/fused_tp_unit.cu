// One translation unit per (covariance formula, tile count) of the thread-per-tile kernel
// (fused_tp.cuh): compiled by the Makefile with -DMGP_COL_F=<0..3> -DMGP_TP_T=<2..13> into
// build/fused_tp_f<F>_t<T>.o: three plain instantiations each (d = 1, 2, 3) and three with the
// back substitution / gradient epilogue (GRAD).  Small units build in parallel.  d = 2 is instantiated LAST on purpose: the code nvcc generates for a kernel depends
// on its position in the unit, and the last one came out measurably better (C2 kernel
// <7, M3/2, 2>: 5 656 instead of 5 672 instructions with a different schedule around the
// barriers, 0.655 instead of 0.689 ms per 100 k neighbourhoods on the same box).
#include "fused_tp.cuh"

namespace mgp {

#define MGP_TP_CAT4_(a, b, c, d) a##b##c##d
#define MGP_TP_CAT4(a, b, c, d) MGP_TP_CAT4_(a, b, c, d)
#define MGP_TP_NAME MGP_TP_CAT4(launch_fused_tp_f, MGP_COL_F, _t, MGP_TP_T)

// back substitution (fast-mean coefficients) and analytic gradient: GRAD instantiations
#define MGP_TPG_NAME MGP_TP_CAT4(launch_fused_tpg_f, MGP_COL_F, _t, MGP_TP_T)
int MGP_TPG_NAME(const mgp_problem* p, const Model& model, const ColLoo& loo, int* grid_out,
                 cudaStream_t stream) {
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  const long long rows = p->b;
  if (a.d == 1)
    return launch_tp_one<MGP_TP_T, MGP_COL_F, 1, true>(a, loo, rows, grid_out, stream);
  if (a.d == 3)
    return launch_tp_one<MGP_TP_T, MGP_COL_F, 3, true>(a, loo, rows, grid_out, stream);
  if (a.d == 2)
    return launch_tp_one<MGP_TP_T, MGP_COL_F, 2, true>(a, loo, rows, grid_out, stream);
  set_error("thread-per-tile kernel: d=%d is not instantiated", a.d);
  return MGP_ERR_UNSUPPORTED;
}

int MGP_TP_NAME(const mgp_problem* p, const Model& model, const ColLoo& loo, int* grid_out,
                cudaStream_t stream) {
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  const long long rows = p->b;
  if (a.d == 1) return launch_tp_one<MGP_TP_T, MGP_COL_F, 1>(a, loo, rows, grid_out, stream);
  if (a.d == 3) return launch_tp_one<MGP_TP_T, MGP_COL_F, 3>(a, loo, rows, grid_out, stream);
  if (a.d == 2) return launch_tp_one<MGP_TP_T, MGP_COL_F, 2>(a, loo, rows, grid_out, stream);
  set_error("thread-per-tile kernel: d=%d is not instantiated", a.d);
  return MGP_ERR_UNSUPPORTED;
}

}  // namespace mgp
