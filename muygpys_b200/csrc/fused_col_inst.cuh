// Instantiates the column-direct kernel for ONE covariance formula (MGP_COL_F), plain or with
// the gradient epilogue (MGP_COL_GRAD), so that the variants compile in parallel translation
// units.  Included by fused_col_m05.cu, ..., fused_colg_m05.cu, ... only.
#include "fused_col.cuh"

namespace mgp {

#define MGP_COL_CAT2(a, b) a##b
#define MGP_COL_CAT(a, b) MGP_COL_CAT2(a, b)
#ifdef MGP_COL_GRAD
#define MGP_COL_NAME MGP_COL_CAT(launch_fused_colg_f, MGP_COL_F)
#define MGP_COL_G true
#else
#define MGP_COL_NAME MGP_COL_CAT(launch_fused_col_f, MGP_COL_F)
#define MGP_COL_G false
#endif

int MGP_COL_NAME(const mgp_problem* p, const Model& model, const ColLoo& loo, int* grid_out,
                 cudaStream_t stream) {
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  const int T = col_tiles(a.k);
  const long long rows = p->b;
#define MGP_COL_CASE(TT, DD)                                                                  \
  if (T == TT && a.d == DD)                                                                   \
    return launch_col_one<TT, MGP_COL_F, DD, MGP_COL_G>(a, loo, rows, grid_out, stream);
  MGP_COL_CASE(2, 1) MGP_COL_CASE(3, 1) MGP_COL_CASE(4, 1) MGP_COL_CASE(5, 1)
  MGP_COL_CASE(6, 1) MGP_COL_CASE(7, 1) MGP_COL_CASE(8, 1)
  MGP_COL_CASE(2, 2) MGP_COL_CASE(3, 2) MGP_COL_CASE(4, 2) MGP_COL_CASE(5, 2)
  MGP_COL_CASE(6, 2) MGP_COL_CASE(7, 2) MGP_COL_CASE(8, 2)
  MGP_COL_CASE(2, 3) MGP_COL_CASE(3, 3) MGP_COL_CASE(4, 3) MGP_COL_CASE(5, 3)
  MGP_COL_CASE(6, 3) MGP_COL_CASE(7, 3) MGP_COL_CASE(8, 3)
#undef MGP_COL_CASE
  set_error("column kernel: T=%d, d=%d is not instantiated", T, a.d);
  return MGP_ERR_UNSUPPORTED;
}

}  // namespace mgp
