// Instantiates the thread-per-tile kernel (fused_tp.cuh) for ONE covariance formula (MGP_COL_F)
// and one range of tile counts (MGP_TP_BIG: T = 9..13, i.e. k = 63..102; otherwise T = 2..8) so
// that they compile in parallel translation units.  Included by fused_tp_m05.cu, _m15.cu,
// _m25.cu, _gauss.cu and fused_tp_big_*.cu only.
#include "fused_tp.cuh"

namespace mgp {

#define MGP_TP_CAT2(a, b) a##b
#define MGP_TP_CAT(a, b) MGP_TP_CAT2(a, b)
#ifdef MGP_TP_BIG
#define MGP_TP_NAME MGP_TP_CAT(launch_fused_tp_big_f, MGP_COL_F)
#else
#define MGP_TP_NAME MGP_TP_CAT(launch_fused_tp_f, MGP_COL_F)
#endif

int MGP_TP_NAME(const mgp_problem* p, const Model& model, const ColLoo& loo, int* grid_out,
                cudaStream_t stream) {
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  const int T = col_tiles(a.k);
  const long long rows = p->b;
#define MGP_TP_CASE(TT, DD)                                                      \
  if (T == TT && a.d == DD)                                                      \
    return launch_tp_one<TT, MGP_COL_F, DD>(a, loo, rows, grid_out, stream);
#ifdef MGP_TP_BIG
  MGP_TP_CASE(9, 1) MGP_TP_CASE(10, 1) MGP_TP_CASE(11, 1) MGP_TP_CASE(12, 1) MGP_TP_CASE(13, 1)
  MGP_TP_CASE(9, 2) MGP_TP_CASE(10, 2) MGP_TP_CASE(11, 2) MGP_TP_CASE(12, 2) MGP_TP_CASE(13, 2)
  MGP_TP_CASE(9, 3) MGP_TP_CASE(10, 3) MGP_TP_CASE(11, 3) MGP_TP_CASE(12, 3) MGP_TP_CASE(13, 3)
#else
  MGP_TP_CASE(2, 1) MGP_TP_CASE(3, 1) MGP_TP_CASE(4, 1) MGP_TP_CASE(5, 1)
  MGP_TP_CASE(6, 1) MGP_TP_CASE(7, 1) MGP_TP_CASE(8, 1)
  MGP_TP_CASE(2, 2) MGP_TP_CASE(3, 2) MGP_TP_CASE(4, 2) MGP_TP_CASE(5, 2)
  MGP_TP_CASE(6, 2) MGP_TP_CASE(7, 2) MGP_TP_CASE(8, 2)
  MGP_TP_CASE(2, 3) MGP_TP_CASE(3, 3) MGP_TP_CASE(4, 3) MGP_TP_CASE(5, 3)
  MGP_TP_CASE(6, 3) MGP_TP_CASE(7, 3) MGP_TP_CASE(8, 3)
#endif
#undef MGP_TP_CASE
  set_error("thread-per-tile kernel: T=%d, d=%d is not instantiated", T, a.d);
  return MGP_ERR_UNSUPPORTED;
}

}  // namespace mgp
