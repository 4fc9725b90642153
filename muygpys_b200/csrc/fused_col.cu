// Dispatch of the column-direct fused kernel (fused_col.cuh; one translation unit per covariance
// formula: fused_col_m05.cu, _m15.cu, _m25.cu, _gauss.cu) and the one-launch leave-one-out
// objective entry point mgp_fused_loo.
#include "fused_col.cuh"
#include "fused_tp.cuh"

namespace mgp {

int launch_fused_tp_big_f0(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_tp_big_f1(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_tp_big_f2(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_tp_big_f3(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
// thread-per-tile kernel: one launcher per (formula, tile count), fused_tp_unit.cu
#define MGP_TP_DECL(F, T)                                                                  \
  int launch_fused_tp_f##F##_t##T(const mgp_problem*, const Model&, const ColLoo&, int*, \
                                  cudaStream_t);
#define MGP_TP_ALL_T(X, F)                                                                   \
  X(F, 2) X(F, 3) X(F, 4) X(F, 5) X(F, 6) X(F, 7) X(F, 8) X(F, 9) X(F, 10) X(F, 11) X(F, 12) \
      X(F, 13)
MGP_TP_ALL_T(MGP_TP_DECL, 0)
MGP_TP_ALL_T(MGP_TP_DECL, 1)
MGP_TP_ALL_T(MGP_TP_DECL, 2)
MGP_TP_ALL_T(MGP_TP_DECL, 3)
#undef MGP_TP_DECL
// ... and with the back substitution / gradient epilogue (GRAD instantiations, every T)
#define MGP_TPG_DECL(F, T)                                                                  \
  int launch_fused_tpg_f##F##_t##T(const mgp_problem*, const Model&, const ColLoo&, int*, \
                                   cudaStream_t);
#define MGP_TPG_ALL_T(X, F) MGP_TP_ALL_T(X, F)
MGP_TPG_ALL_T(MGP_TPG_DECL, 0)
MGP_TPG_ALL_T(MGP_TPG_DECL, 1)
MGP_TPG_ALL_T(MGP_TPG_DECL, 2)
MGP_TPG_ALL_T(MGP_TPG_DECL, 3)
#undef MGP_TPG_DECL
int launch_fused_colg_f0(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_colg_f1(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_colg_f2(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
int launch_fused_colg_f3(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);

int fused_variant();
int validate_problem(const mgp_problem* p);

static int col_formula(const Model& model) {
  if (model.metric_id == MGP_METRIC_L2) {
    switch (model.kernel_id) {
      case MGP_KERNEL_MATERN_05: return F_M05;
      case MGP_KERNEL_MATERN_15: return F_M15;
      case MGP_KERNEL_MATERN_25: return F_M25;
      case MGP_KERNEL_MATERN_INF: return F_GAUSS;
      default: return -1;  // RBF handed l2 distances (reference quirk): tile kernel
    }
  }
  return model.kernel_id == MGP_KERNEL_RBF ? F_GAUSS : -1;
}

// Shapes the column-direct kernels take: r = 1, d <= 3, homoscedastic nugget, the four closed
// covariance formulas; k = 7..102 (T = 2..13 tile rows) for plain prediction and the one-launch
// objective and for the back substitution (fast-mean coefficients, analytic gradient: GRAD
// instantiations of the thread-per-tile kernel); the lane-parallel column kernel (variant 4)
// stops at k = 62 (T <= 8).
static int col_shape_ok(const mgp_problem* p, const Model& model, int max_tiles) {
  if (p->r != 1 || p->d > 3 || p->noise_bk || !p->train_y) return 0;
  const int T = col_tiles(p->k);
  if (T < 2 || T > max_tiles) return 0;
  if (col_formula(model) < 0) return 0;
  // 2-D points are fetched with one 16-byte cp.async each
  if (p->d == 2 && ((((uintptr_t)p->train_x) | ((uintptr_t)p->query_x)) & 15)) return 0;
  return 1;
}

int fused_col_supported(const mgp_problem* p, const Model& model) {
  return col_shape_ok(p, model, TP_MAX_T);
}

static int launch_col(const mgp_problem* p, const Model& model, const ColLoo& loo, int* grid_out,
                      cudaStream_t stream) {
  const int T = col_tiles(p->k);
  typedef int (*launcher)(const mgp_problem*, const Model&, const ColLoo&, int*, cudaStream_t);
  const int f = col_formula(model);
  MGP_REQUIRE(f >= 0 && f < 4 && T >= 2 && T <= TP_MAX_T, MGP_ERR_UNSUPPORTED,
              "column kernels do not support this kernel / metric pair or k = %d", p->k);
  const bool backsub = loo.grad != nullptr || loo.backsub;
  // Variant 4 (cross-check of the thread-per-tile kernels): the lane-parallel column kernel,
  // GRAD instantiation -- its back substitution goes unused for a plain prediction.
  if (fused_variant() == 4 && T <= COL_MAX_T) {
    switch (f) {
      case F_M05: return launch_fused_colg_f0(p, model, loo, grid_out, stream);
      case F_M15: return launch_fused_colg_f1(p, model, loo, grid_out, stream);
      case F_M25: return launch_fused_colg_f2(p, model, loo, grid_out, stream);
      default: return launch_fused_colg_f3(p, model, loo, grid_out, stream);
    }
  }
  if (backsub) {
#define MGP_TPG_ENTRY(F, TT) launch_fused_tpg_f##F##_t##TT,
    static const launcher gtable[4][TP_MAX_T - 1] = {{MGP_TPG_ALL_T(MGP_TPG_ENTRY, 0)},
                                                      {MGP_TPG_ALL_T(MGP_TPG_ENTRY, 1)},
                                                      {MGP_TPG_ALL_T(MGP_TPG_ENTRY, 2)},
                                                      {MGP_TPG_ALL_T(MGP_TPG_ENTRY, 3)}};
#undef MGP_TPG_ENTRY
    return gtable[f][T - 2](p, model, loo, grid_out, stream);
  }
  // Plain prediction and the one-launch objective
#define MGP_TP_ENTRY(F, TT) launch_fused_tp_f##F##_t##TT,
  static const launcher table[4][TP_MAX_T - 1] = {{MGP_TP_ALL_T(MGP_TP_ENTRY, 0)},
                                                  {MGP_TP_ALL_T(MGP_TP_ENTRY, 1)},
                                                  {MGP_TP_ALL_T(MGP_TP_ENTRY, 2)},
                                                  {MGP_TP_ALL_T(MGP_TP_ENTRY, 3)}};
#undef MGP_TP_ENTRY
  return table[f][T - 2](p, model, loo, grid_out, stream);
}

// Neighbourhoods in flight per SM for the kernel a problem takes (the host-buffer pipeline cuts
// its chunks on multiples of one wave): update warps of the thread-per-tile kernel, 12 for the
// tile kernels.
int fused_wave_per_sm(const mgp_problem* p) {
  Model model;
  if (make_model(p->kernel_id, p->metric_id, p->d, p->length_scale_count, p->length_scale,
                 &model) != MGP_OK ||
      !fused_col_supported(p, model) || p->coeffs)
    return 12;
  const int T = col_tiles(p->k), D = p->d;
#define MGP_U_CASE(TT) \
  case TT: return D == 1 ? tp_uwarps<TT, 1>() : D == 2 ? tp_uwarps<TT, 2>() : tp_uwarps<TT, 3>();
  switch (T) {
    MGP_U_CASE(2) MGP_U_CASE(3) MGP_U_CASE(4) MGP_U_CASE(5) MGP_U_CASE(6) MGP_U_CASE(7)
    MGP_U_CASE(8) MGP_U_CASE(9) MGP_U_CASE(10) MGP_U_CASE(11) MGP_U_CASE(12) MGP_U_CASE(13)
    default: return 12;
  }
#undef MGP_U_CASE
}

int launch_fused_col(const mgp_problem* p, const Model& model, cudaStream_t stream) {
  ColLoo loo = {};
  loo.backsub = p->coeffs != nullptr;  // coefficients come from the back substitution
  return launch_col(p, model, loo, nullptr, stream);
}

// workspace of mgp_fused_loo: [counter (256 B)] [per-warp records] -- the record area is sized for
// the largest grid any device of this process can run (4 CTAs per SM)
static size_t loo_ws_bytes() {
  return 256 + (size_t)sm_count() * 4 * COL_NREC_GRAD * sizeof(double);
}

}  // namespace mgp

extern "C" size_t mgp_fused_loo_workspace_bytes(const mgp_problem* p) {
  (void)p;
  return mgp::loo_ws_bytes();
}

namespace mgp {

static int check_peers(const mgp_peer_group* g) {
  MGP_REQUIRE(g->world >= 1 && g->world <= MGP_MAX_PEERS && g->rank >= 0 && g->rank < g->world,
              MGP_ERR_BAD_ARG, "peer group: rank %d of %d (at most %d peers)", g->rank, g->world,
              MGP_MAX_PEERS);
  MGP_REQUIRE(g->epoch >= 1, MGP_ERR_BAD_ARG, "peer group: epochs start at 1");
  for (int p = 0; p < g->world; ++p)
    MGP_REQUIRE(g->peer_buf[p] != nullptr, MGP_ERR_BAD_ARG, "peer group: buffer %d is NULL", p);
  return MGP_OK;
}

__global__ void __launch_bounds__(64) peer_sum8_kernel(double* partials, const mgp_peer_group g) {
  __shared__ double rec[MGP_PARTIALS];
  if (threadIdx.x < MGP_PARTIALS) rec[threadIdx.x] = partials[threadIdx.x];
  __syncthreads();
  peer_sum8_block(g, rec, partials);
}

}  // namespace mgp

extern "C" size_t mgp_peer_buffer_bytes(void) {
  return mgp::PEER_DATA_DOUBLES * sizeof(double) +
         2 * MGP_MAX_PEERS * sizeof(unsigned long long);
}

extern "C" int mgp_peer_sum8(double* partials, const mgp_peer_group* g, void* stream) {
  using namespace mgp;
  MGP_REQUIRE(partials && g, MGP_ERR_BAD_ARG, "partials and peer group are required");
  const int rc = check_peers(g);
  if (rc != MGP_OK) return rc;
  if (g->world == 1) return MGP_OK;
  peer_sum8_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(partials, *g);
  return check_launch("peer_sum8_kernel");
}

extern "C" int mgp_fused_loo(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                             double* partials, void* ws, size_t ws_bytes, void* stream) {
  return mgp_fused_loo_peers(p, loss_id, boundary_scale, partials, ws, ws_bytes, nullptr, stream);
}

extern "C" int mgp_fused_loo_peers(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                                   double* partials, void* ws, size_t ws_bytes,
                                   const mgp_peer_group* g, void* stream) {
  return mgp_fused_loo_grad(p, loss_id, boundary_scale, partials, nullptr, ws, ws_bytes, g,
                            stream);
}

extern "C" int mgp_fused_loo_grad(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                                  double* partials, double* grad, void* ws, size_t ws_bytes,
                                  const mgp_peer_group* g, void* stream) {
  using namespace mgp;
  int rc = validate_problem(p);
  if (rc != MGP_OK) return rc;
  MGP_REQUIRE(partials != nullptr, MGP_ERR_BAD_ARG, "partials is required");
  MGP_REQUIRE(loss_id == MGP_LOSS_NONE || loss_id == MGP_LOSS_MSE || loss_id == MGP_LOSS_LOOL ||
                  loss_id == MGP_LOSS_PSEUDO_HUBER || loss_id == MGP_LOSS_LOOPH,
              MGP_ERR_UNSUPPORTED,
              "mgp_fused_loo finishes mse, lool, pseudo-Huber and (for a known scale) looph in "
              "one launch (loss_id %d needs the two-pass path)", loss_id);
  MGP_REQUIRE(loss_id != MGP_LOSS_LOOPH || p->scale > 0.0, MGP_ERR_BAD_ARG,
              "looph: p->scale must hold the (positive) variance scale sigma^2");
  Model model;
  rc = make_model(p->kernel_id, p->metric_id, p->d, p->length_scale_count, p->length_scale,
                  &model);
  if (rc != MGP_OK) return rc;
  MGP_REQUIRE(col_shape_ok(p, model, TP_MAX_T), MGP_ERR_UNSUPPORTED,
              "mgp_fused_loo: shape not supported by the column kernels (k=%d d=%d r=%d%s)", p->k,
              p->d, p->r, grad ? ", analytic gradient" : "");
  MGP_REQUIRE(p->query_x == p->train_x && p->query_idx != nullptr, MGP_ERR_BAD_ARG,
              "mgp_fused_loo expects a training batch: query_x == train_x and batch indices in "
              "query_idx");
  MGP_REQUIRE(ws != nullptr && ws_bytes >= loo_ws_bytes(), MGP_ERR_WORKSPACE,
              "workspace of %zu bytes required", loo_ws_bytes());
  MGP_REQUIRE(boundary_scale > 0.0 ||
                  (loss_id != MGP_LOSS_PSEUDO_HUBER && loss_id != MGP_LOSS_LOOPH),
              MGP_ERR_BAD_ARG,
              "boundary_scale must be positive");
  ColLoo loo = {};
  if (g != nullptr && g->world > 1) {
    rc = check_peers(g);
    if (rc != MGP_OK) return rc;
    loo.peers = *g;
  } else {
    loo.peers.world = 1;
  }
  loo.counter = (unsigned int*)ws;
  loo.warp_rec = (double*)((char*)ws + 256);
  loo.partials = partials;
  loo.grad = grad;
  if (grad != nullptr) {
    // 1 / l_f as the deformation applies it (isotropic: the same value for every feature)
    for (int f = 0; f < p->d && f < 3; ++f)
      loo.inv_len[f] = 1.0 / p->length_scale[p->length_scale_count == 1 ? 0 : f];
  }
  loo.loss_id = loss_id;
  loo.boundary_scale = boundary_scale > 0.0 ? boundary_scale : 1.0;
  int grid = 0;  // the full persistent grid, whatever the batch size (fixed summation order)
  return launch_col(p, model, loo, &grid, (cudaStream_t)stream);
}
