/*
 * muygpys_b200.h -- C ABI of the B200-native MuyGPyS hot path.
 *
 * One shared library (muygpys_b200/libmuygpys_b200.so, sm_100a) exports exactly
 * the symbols declared here.  Every entry point replaces one reference
 * interface on the per-neighbourhood GP path; the citation next to each is the
 * reference code it stands in for (S/ = /root/reference/src/MuyGPyS/).  A
 * reference maintainer binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - All array arguments are DEVICE pointers to C-contiguous buffers owned by
 *     the caller (float64 / int64, as S/_src/math/numpy.py:92-96), except where
 *     a parameter is documented as HOST.
 *   - `stream` is a cudaStream_t passed as void*; work is stream-ordered, no
 *     entry point synchronises the device or allocates persistent memory.
 *   - Return value: MGP_OK (0) or a negative mgp_status.  mgp_last_error()
 *     returns a thread-local message for the last failing call.
 *   - Numerical failure (non-positive pivot) is reported per row in `status`
 *     (when given) and as NaN outputs, mirroring numpy.linalg.solve which only
 *     raises for exactly singular matrices.
 *   - No CPU fallback exists anywhere in the library.
 */
#ifndef MUYGPYS_B200_H
#define MUYGPYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGP_VERSION 100 /* 0.1.0 */
#define MGP_MAX_ANISO_DIM 32 /* anisotropic length scales live in kernel params */
#define MGP_PARTIALS 8       /* doubles in a loss/scale partials record */
#define MGP_MAX_PEERS 8      /* GPUs of one NVLink domain that can share a partials record */
#define MGP_GRAD_PARAMS 4    /* gradient slots: length scales of features 0..2, then the nugget */
#define MGP_GRAD_DOUBLES 20  /* 5 sums per gradient slot, see mgp_fused_loo_grad */

typedef enum mgp_status {
  MGP_OK = 0,
  MGP_ERR_BAD_ARG = -1,
  MGP_ERR_UNSUPPORTED = -2,
  MGP_ERR_CUDA = -3,
  MGP_ERR_WORKSPACE = -4
} mgp_status;

/* S/_src/gp/kernels/numpy.py:12-31 */
typedef enum mgp_kernel_id {
  MGP_KERNEL_RBF = 0,        /* exp(-x/2), x = F2/l^2          :12-13 */
  MGP_KERNEL_MATERN_05 = 1,  /* exp(-x), x = l2/l              :16-17 */
  MGP_KERNEL_MATERN_15 = 2,  /* (1+sqrt3 x) exp(-sqrt3 x)      :20-22 */
  MGP_KERNEL_MATERN_25 = 3,  /* (1+sqrt5 x+5x^2/3) exp(-sqrt5 x) :25-27 */
  MGP_KERNEL_MATERN_INF = 4  /* exp(-x^2/2)                    :30-31 */
} mgp_kernel_id;

/* S/gp/deformation/metric.py:237-265 */
typedef enum mgp_metric_id {
  MGP_METRIC_L2 = 0, /* sqrt(sum diff^2); length-scale rule x / l   */
  MGP_METRIC_F2 = 1  /* sum diff^2;       length-scale rule x / l^2 */
} mgp_metric_id;

/* S/_src/optimize/loss/numpy.py:12-112 */
typedef enum mgp_loss_id {
  MGP_LOSS_NONE = 0,
  MGP_LOSS_MSE = 1,
  MGP_LOSS_LOOL = 2,
  MGP_LOSS_LOOPH = 3,
  MGP_LOSS_PSEUDO_HUBER = 4,
  MGP_LOSS_CROSS_ENTROPY = 5
} mgp_loss_id;

/* Slots of a partials record (all plain sums, so ranks combine them with one
 * SUM all-reduce -- the replacement for S/_src/optimize/loss/mpi.py:21-104 and
 * S/_src/optimize/scale/mpi.py:19-37). */
enum {
  MGP_P_SQERR = 0,   /* sum (pred-target)^2 over rows and responses          */
  MGP_P_COUNT = 1,   /* number of (row,response) elements                    */
  MGP_P_YKY = 2,     /* sum_rows y^T K^-1 y                                   */
  MGP_P_ROWS = 3,    /* number of rows                                        */
  MGP_P_SQERR_V = 4, /* sum (pred-target)^2 / var          (lool numerator)   */
  MGP_P_LOGV = 5,    /* sum log var                                           */
  MGP_P_AUX = 6,     /* pseudo-huber / cross-entropy / looph sum (loss_id)    */
  MGP_P_BAD = 7      /* number of rows whose factorisation failed             */
};

/*
 * One batch of neighbourhoods for the fused kernel.  Replaces, in one launch:
 *   MuyGPS.make_predict_tensors / make_train_tensors   S/gp/muygps.py:405-551
 *   _crosswise_tensor/_pairwise_tensor, _l2/_F2         S/_src/gp/tensors/numpy.py:47-94
 *   Isotropy/Anisotropy.__call__                        S/gp/deformation/isotropy.py:60-89, anisotropy.py:43-70
 *   _rbf_fn/_matern_*_fn                                S/_src/gp/kernels/numpy.py:12-31
 *   _homoscedastic/_heteroscedastic_perturb             S/_src/gp/noise/numpy.py:9-27,56-67
 *   _muygps_posterior_mean/_muygps_diagonal_variance    S/_src/gp/muygps/numpy.py:17-67
 *   _muygps_fast_posterior_mean_precompute              S/_src/gp/muygps/numpy.py:88-95
 *   _analytic_scale_optim_unnormalized (per-row term)   S/_src/optimize/scale/numpy.py:9-15
 */
typedef struct mgp_problem {
  /* geometry */
  const double* train_x;    /* (n,d) */
  const double* query_x;    /* (t,d); may alias train_x (LOO training batches) */
  const int64_t* query_idx; /* (b) rows of query_x, or NULL for 0..b-1 */
  const int64_t* nn_idx;    /* (b,k) rows of train_x */
  const double* train_y;    /* (n,r); NULL allowed when only var is wanted */
  int64_t n, t, b;
  int32_t k, d, r;
  /* model */
  int32_t kernel_id;          /* mgp_kernel_id */
  int32_t metric_id;          /* mgp_metric_id */
  int32_t length_scale_count; /* 1 = Isotropy, d = Anisotropy (d <= MGP_MAX_ANISO_DIM) */
  const double* length_scale; /* HOST pointer, length_scale_count values */
  double noise;               /* homoscedastic nugget tau^2 (ignored if noise_bk) */
  const double* noise_bk;     /* (b,k) heteroscedastic nugget, or NULL */
  double scale;               /* sigma^2 multiplying var (1.0 = unscaled opt var_fn) */
  /* outputs, each nullable */
  double* mean;    /* (b,r)   K_cross (K+eps)^-1 Y */
  double* var;     /* (b)     scale * (1 - K_cross (K+eps)^-1 K_cross^T) */
  double* yky;     /* (b)     sum_r y_r^T (K+eps)^-1 y_r */
  double* coeffs;  /* (b,k,r) (K+eps)^-1 Y  (fast-mean precompute mode) */
  int32_t* status; /* (b)     0 ok, 1 non-positive pivot */
} mgp_problem;

int mgp_version(void);
const char* mgp_last_error(void);

/* ---- fused path (K1/K3) ------------------------------------------------ */
size_t mgp_fused_workspace_bytes(const mgp_problem* p);
int mgp_fused_posterior(const mgp_problem* p, void* ws, size_t ws_bytes, void* stream);

/* Host-buffer form of the same call -- what `regress_from_indices`
 * (MuyGPyS/examples/from_indices.py:22-63) is handed: neighbour indices (and batch indices) in
 * HOST memory, results wanted in host memory.  `p` is filled as for mgp_fused_posterior, with
 * p->nn_idx (b*k) and, if query_idx_host is given, p->query_idx (b) pointing at DEVICE staging
 * buffers that this call fills; mean_host / var_host (nullable) receive copies of p->mean /
 * p->var.  The batch is processed in chunks on three internal streams (upload, kernel,
 * download, linked by one event per chunk) so that the upload of chunk c+1 and the download of
 * chunk c-1 overlap the kernel of chunk c and the uploads run back to back; the work is ordered
 * after `stream` and joined back into it (synchronise `stream` before reading the host
 * results).  Host buffers should be page-locked for the copies to overlap. */
int mgp_fused_posterior_host(const mgp_problem* p, const int64_t* nn_idx_host,
                             const int64_t* query_idx_host, double* mean_host,
                             double* var_host, void* ws, size_t ws_bytes, void* stream);

/* The same call for a caller that holds its neighbour indices as 32-bit integers (numpy
 * accepts any integer dtype as an index array, S/_src/gp/tensors/numpy.py:47-94): half the
 * bytes cross the host link -- the C2 step uploads 21.6 MB instead of 42.4 MB.  `nn_stage32`
 * is a DEVICE buffer of b*k int32 the chunks land in; a small kernel widens each chunk into
 * p->nn_idx (int64, what every kernel reads) ahead of the chunk's fused launch. */
int mgp_fused_posterior_host32(const mgp_problem* p, const int32_t* nn_idx_host32,
                               int32_t* nn_stage32, const int64_t* query_idx_host,
                               double* mean_host, double* var_host, void* ws, size_t ws_bytes,
                               void* stream);

/* Test/bench hook: 0 = choose automatically (thread-per-tile > tile > generic), 1 = always the
 * generic shared-memory kernel, 2 = the register-tile DMMA kernel where supported, 3 = the
 * thread-per-tile kernel (error if the shape is unsupported: r == 1, d <= 3, homoscedastic
 * nugget, 7 <= k <= 102, with or without coefficients), 4 = its lane-parallel predecessor, the
 * column-direct kernel (k <= 62).  Lets the independently written variants be cross-checked on
 * identical inputs. */
int mgp_set_fused_variant(int32_t variant);

/* One leave-one-out objective evaluation in ONE launch (a14-a16): the fused kernel over a
 * TRAINING batch (p->query_x == p->train_x, p->query_idx = batch indices, p->nn_idx = their
 * neighbours; p->mean / var / yky optional) with the loss and scale partials folded into its
 * epilogue -- per-warp sums in row order, one record per warp, fixed-order tree in the last
 * CTA: bit-reproducible, no floating-point atomics, no extra launches.  `partials`
 * (MGP_PARTIALS doubles, device) is OVERWRITTEN with the record of plain sums
 * (MGP_P_*; the target of batch row i is train_y[query_idx[i]]), ready for one SUM all-reduce
 * across ranks.  Replaces make_loo_crossval_fn's kernel + mean (+ scale + variance) + loss
 * sequence (S/optimize/objective.py:20-105, S/_src/optimize/loss/numpy.py:22-61,
 * S/_src/optimize/scale/numpy.py:9-15) for mse, lool and pseudo-Huber (MGP_LOSS_NONE gives the
 * scale partials only).  looph is nonlinear in the scale: with MGP_LOSS_LOOPH the launch takes
 * sigma^2 from p->scale (a fixed scale, or the analytic scale of a previous MGP_LOSS_NONE
 * launch), MGP_P_AUX receives sum 2 b^2 (sqrt(1 + e^2 / (b^2 sigma^2 v)) - 1) and MGP_P_SQERR_V
 * the Huber-weighted sum e^2 / (v sqrt(1 + u)), so that value and gradient finish like lool.  r == 1, d <= 3, 7 <= k <= 102, homoscedastic nugget
 * (MGP_ERR_UNSUPPORTED otherwise).  `ws` must be ZERO-FILLED before its first use and handed
 * back unchanged afterwards (it carries a self-resetting arrival counter). */
size_t mgp_fused_loo_workspace_bytes(const mgp_problem* p);
int mgp_fused_loo(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                  double* partials, void* ws, size_t ws_bytes, void* stream);

/* ---- cross-GPU SUM of a partials record over NVLink peer memory --------------------------
 * The replacement for the reference's allreduce sites on this path (S/_src/optimize/loss/mpi.py
 * :21-104, S/_src/optimize/scale/mpi.py:19-37) when the ranks are GPUs of one NVLink /
 * NVSwitch domain.  Every rank allocates a peer-mapped buffer of mgp_peer_buffer_bytes() bytes,
 * ZERO-FILLED once (symmetric memory; the Python side uses
 * torch.distributed._symmetric_memory), and fills `peer_buf[p]` with rank p's buffer as mapped
 * into THIS process.  `epoch` must start at 1 and increase by one per call on every rank.
 * The record is pushed into every peer's buffer with 8-byte P2P stores, a release flag
 * follows, and each rank adds the records in rank order once all flags of this epoch have
 * arrived: one block, no separate collective launch, bit-identical sums on every rank.  A peer
 * that does not arrive within ~2 s poisons the result with NaN instead of hanging the GPU. */
typedef struct mgp_peer_group {
  int32_t rank, world;          /* world <= MGP_MAX_PEERS */
  uint64_t epoch;
  void* peer_buf[MGP_MAX_PEERS];
} mgp_peer_group;
size_t mgp_peer_buffer_bytes(void);
/* partials (MGP_PARTIALS doubles, device): local record in, global sum out. */
int mgp_peer_sum8(double* partials, const mgp_peer_group* g, void* stream);
/* mgp_fused_loo with the cross-GPU sum fused into the epilogue of the SAME kernel: the last
 * block to finish reduces the per-warp records and then runs the peer exchange; `partials`
 * receives the sum over all ranks.  g == NULL or g->world == 1 behaves like mgp_fused_loo. */
int mgp_fused_loo_peers(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                        double* partials, void* ws, size_t ws_bytes, const mgp_peer_group* g,
                        void* stream);

/* The same launch with the ANALYTIC GRADIENT of the objective's ingredients (SURVEY.md 8f-2;
 * the reference finite-differences, S/_src/optimize/chassis/numpy.py:68-74).  After the
 * factorisation the kernel back-substitutes w = K^-1 kcross and alpha = K^-1 y on the stored
 * factor and re-evaluates dK/dtheta entry by entry (never stored):
 *   d mean = dc^T alpha - w^T dK alpha,  d var = -2 dc^T w + w^T dK w,  d yky = -alpha^T dK alpha
 * for theta = the length scale of feature 0, 1, 2 (slots 0..2; an isotropic model's derivative
 * is the sum over its features) and the nugget tau^2 (slot 3).  `grad` (MGP_GRAD_DOUBLES doubles,
 * device or pinned host) receives, per slot t, the batch sums
 *   grad[5t+0] = sum 2 e dm        (e = mean - target)      -> d sum e^2           (mse)
 *   grad[5t+1] = sum 2 e dm / v    grad[5t+2] = sum e^2 dv / v^2    grad[5t+3] = sum dv / v
 *   (MGP_LOSS_LOOPH: the terms of grad[5t+1] and grad[5t+2] carry the weight 1 / sqrt(1 + u))
 *   grad[5t+4] = sum d yky                                   -> d sigma^2 (analytic scale)
 * from which the host finishes d mse and d lool (muygpys_b200/objective.py).  grad == NULL is
 * mgp_fused_loo_peers.  The gradient sums are per rank: `partials` goes through the peer
 * exchange, `grad` is summed across ranks by the caller.  grad != NULL takes the same shapes (k <= 102). */
int mgp_fused_loo_grad(const mgp_problem* p, int32_t loss_id, double boundary_scale,
                       double* partials, double* grad, void* ws, size_t ws_bytes,
                       const mgp_peer_group* g, void* stream);

/* ---- losses and scale partials (a14/a15) -------------------------------
 * Accumulates (adds) a partials record over b rows into `partials`
 * (MGP_PARTIALS doubles, zero it first).  `var`/`yky` may be NULL when the loss
 * does not need them.  `scale_dev` is a DEVICE scalar sigma^2 (so a preceding
 * all-reduce can produce it without a host round trip); NULL means 1.0.
 * LOOL/LOOPH use scale*var; LOOPH and PSEUDO_HUBER take `boundary_scale`.
 * Deterministic: fixed-order tree reduction, no floating-point atomics. */
size_t mgp_loss_workspace_bytes(int64_t b, int32_t r);
int mgp_loss_partials(int32_t loss_id, const double* pred, const double* targets,
                      const double* var, const double* yky, const double* scale_dev,
                      double boundary_scale, int64_t b, int32_t r, double* partials,
                      void* ws, size_t ws_bytes, void* stream);

/* ---- exact KNN (a1): NN_Wrapper._get_nns          S/neighbors.py:213-262
 * out_idx (q,k) int64 ascending by distance, out_d2 (q,k) SQUARED l2.
 * exclude_self != 0 drops, for query row i, the train row self_idx[i]
 * (get_batch_nns semantics, S/neighbors.py:169-211).
 * The result is that of a brute-force sweep with separately rounded subtract / multiply / add in
 * feature order (ties to the lower train row), whichever kernel computes it: for n >= 2048,
 * q >= 8, k <= 88 an FP64 tensor-core pre-filter proposes k + 8 candidates per query, they are
 * re-ranked with the exact arithmetic and the answer is certified against a rounding bound;
 * queries that cannot be certified are re-run through the exact sweep inside the same call.
 * `ws` must hold mgp_knn_workspace_bytes(n, q, d, k) bytes. */
size_t mgp_knn_workspace_bytes(int64_t n, int64_t q, int32_t d, int32_t k);
int mgp_knn(const double* train, int64_t n, const double* queries, int64_t q, int32_t d,
            int32_t k, int32_t exclude_self, const int64_t* self_idx, int64_t* out_idx,
            double* out_d2, void* ws, size_t ws_bytes, void* stream);

/* Uniform-grid variant for d <= 3 (same results, bit for bit, as mgp_knn).
 * mgp_knn_grid_cells writes the cell id of every point (cells of edge `cell_size`, `dims[d]`
 * cells per axis starting at `origin[d]`; dims/origin are HOST arrays).  The caller sorts the
 * training points by cell id (any device sort) and passes the sorted points, their original
 * rows and the (ncells+1) cell offsets to mgp_knn_grid_query.  `query_order` (nullable) is the
 * order in which queries are processed (sorted by cell for coherent loads); outputs are always
 * written at the query's own row.  `self_idx` (nullable) as in mgp_knn. */
int mgp_knn_grid_cells(const double* points, int64_t n, int32_t d, const int32_t* dims,
                       const double* origin, double cell_size, int32_t* out_cell, void* stream);
int mgp_knn_grid_query(const double* sorted_points, const int32_t* sorted_ids,
                       const int32_t* cell_start, int64_t n, int32_t d, const int32_t* dims,
                       const double* origin, double cell_size, const double* queries,
                       const int32_t* query_order, int64_t q, int32_t k, const int64_t* self_idx,
                       int64_t* out_idx, double* out_d2, void* stream);

/* ---- fast posterior mean apply (a12, K4): crosswise + kernel + dot ------
 * mean[i,:] = sum_j kernel(dist(query[i], train[nn_idx[i,j]])) * coeffs[coeff_row[i], j, :]
 * Replaces fast_posterior_mean_from_indices        S/examples/from_indices.py:93-123 */
int mgp_fast_mean(const mgp_problem* p, const int64_t* coeff_row, const double* coeffs,
                  void* stream);

/* ---- staged single ops (K5): 1:1 replacements for each _backend_* hook -- */
/* _crosswise_tensor  S/_src/gp/tensors/numpy.py:47-58 ; out (b,k,d) */
int mgp_crosswise_diffs(const double* data, const double* nn_data, const int64_t* data_idx,
                        const int64_t* nn_idx, int64_t b, int32_t k, int32_t d, double* out,
                        void* stream);
/* _pairwise_tensor   S/_src/gp/tensors/numpy.py:61-69 ; out (b,k,k,d) */
int mgp_pairwise_diffs(const double* data, const int64_t* nn_idx, int64_t b, int32_t k,
                       int32_t d, double* out, void* stream);
/* _F2/_l2 over the last axis, optionally dividing each feature by inv-free
 * length scales first (Anisotropy.__call__): in (rows,d) -> out (rows).
 * length_scale: HOST pointer to d values or NULL. */
int mgp_metric_reduce(int32_t metric_id, const double* diffs, int64_t rows, int32_t d,
                      const double* length_scale, double* out, void* stream);
/* fused gather + metric (Isotropy.pairwise_tensor / crosswise_tensor) */
int mgp_crosswise_dists(int32_t metric_id, const double* data, const double* nn_data,
                        const int64_t* data_idx, const int64_t* nn_idx, int64_t b, int32_t k,
                        int32_t d, double* out, void* stream);
int mgp_pairwise_dists(int32_t metric_id, const double* data, const int64_t* nn_idx, int64_t b,
                       int32_t k, int32_t d, double* out, void* stream);
/* out[i] = kernel(in[i] * pre_scale)   (pre_scale = 1/l or 1/l^2; S/gp/deformation/metric.py:241,264) */
int mgp_kernel_apply(int32_t kernel_id, const double* in, double pre_scale, int64_t count,
                     double* out, void* stream);
/* _homoscedastic_perturb / _heteroscedastic_perturb  S/_src/gp/noise/numpy.py:9-27,56-67
 * out may alias Kin. noise_bk NULL -> homoscedastic with `noise`. */
int mgp_perturb(const double* Kin, int64_t b, int32_t k, double noise, const double* noise_bk,
                double* out, void* stream);
/* Batched SPD solve on materialised tensors: for each row, factor Kin (k,k) once and
 *   mean  (b,r) = Kcross^T Kin^-1 Y          _muygps_posterior_mean   muygps/numpy.py:17-41
 *   var   (b)   = kout - Kcross^T Kin^-1 Kcross   _muygps_diagonal_variance :44-67
 *   yky   (b)   = sum_r Y_r^T Kin^-1 Y_r     _analytic_scale_optim_unnormalized scale/numpy.py:9-15
 *   coeffs(b,k,r) = Kin^-1 Y                 _muygps_fast_posterior_mean_precompute :88-95
 * Kin is used as given (already perturbed).  Kcross / Y / outputs nullable. */
size_t mgp_solve_workspace_bytes(int64_t b, int32_t k, int32_t r);
int mgp_solve(const double* Kin, const double* Kcross, const double* Y, int64_t b, int32_t k,
              int32_t r, double kout, double* mean, double* var, double* yky, double* coeffs,
              int32_t* status, void* ws, size_t ws_bytes, void* stream);
/* mask[i] = 1 when the labels of row i's neighbours are not all equal: the "nonconstant
 * neighbourhood" filter of get_balanced_batch / full_filtered_batch
 * (S/optimize/batch.py:58-64,104-110) and classify_any (S/examples/classify.py:577-583).
 * `labels` is read with `label_stride` doubles between consecutive training rows (1 for a flat
 * label array, class_count for column 0 of a one-hot matrix); the (b,k) gathered label tensor
 * is never materialised. */
int mgp_nn_label_mask(const double* labels, int64_t label_stride, const int64_t* nn_idx,
                      int64_t b, int32_t k, uint8_t* mask, void* stream);
/* einsum('ij,ijk->ik')  _muygps_fast_posterior_mean  S/_src/gp/muygps/numpy.py:70-77 */
int mgp_rowdot(const double* Kcross, const double* coeffs, int64_t b, int32_t k, int32_t r,
               double* out, void* stream);

/* ---- measurement helper -------------------------------------------------
 * FP64 roofline probe (MEASURED_PEAKS.json has no FP64 entry).  Launches one
 * kernel of `blocks` x `threads` running `iters` iterations of
 *   mode 0: 8 independent DFMA chains per thread      (2*8*iters flop/thread)
 *   mode 1: 4 independent mma.sync.m8n8k4.f64 chains per warp (2*256*4*iters flop/warp)
 *   mode 2: even warps mode 0, odd warps mode 1 (pipe-sharing test)
 *   mode 3: one dependent DFMA chain per thread (latency)
 *   mode 4: one dependent DMMA chain per warp (latency)
 *   mode 5: 8 independent SHFL.IDX per thread per iteration (shuffle issue rate)
 *   mode 6: one dependent SHFL.IDX chain per thread (shuffle latency)
 * `sink` is a device buffer of blocks*threads doubles.  Timed by the caller. */
int mgp_fp64_probe(int32_t mode, int32_t blocks, int32_t threads, int32_t iters, double* sink,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUYGPYS_B200_H */
