// K2 (low-dimensional variant): exact k-nearest neighbours on a uniform cell grid, d <= 3.
//
// Brute force costs n*q distance evaluations (1e15 at BASELINE config C5).  For the spatial
// configs (d = 2) the training points are bucketed once into a uniform grid (cell ids computed
// here, sorted with a device radix sort by the caller); a query then visits Chebyshev shells of
// cells around its own cell and stops as soon as its current k-th distance is STRICTLY below
// the distance to the nearest unvisited cell, which keeps the search exact including ties.
//
// Distances use the same separately rounded subtract / multiply / add in feature order as the
// brute-force kernel (knn.cu), and the result list is ordered by (distance, train index), so
// both kernels return bit-identical indices and squared distances
// (NN_Wrapper._get_nns semantics, S/neighbors.py:213-262).
//
// One thread per query with a private bounded heap in shared memory; queries are processed in
// cell order (sorted by the caller) so the threads of a warp walk the same cells and their
// loads coalesce.
#include <float.h>
#include <limits.h>

#include "common.cuh"
#include "smem_heap.cuh"

namespace mgp {

struct GridArgs {
  const double* pts;        // (n,D) training points sorted by cell id
  const int32_t* ids;       // (n)   original row of each sorted point
  const int32_t* cell_start;  // (ncells+1)
  const double* queries;    // (q,D)
  const int32_t* order;     // (q) processing order (queries sorted by cell) or NULL
  const int64_t* self_idx;  // (q) train row to skip per query, or NULL
  int64_t* out_idx;         // (q,k)
  double* out_d2;           // (q,k)
  long long n, q;
  int k;
  int dims[3];
  double origin[3];
  double h, inv_h;
};

template <int D>
__device__ __forceinline__ int cell_coord(double x, double origin, double inv_h, int dim) {
  int c = (int)floor((x - origin) * inv_h);
  return min(max(c, 0), dim - 1);
}

template <int D>
__global__ void grid_cell_ids_kernel(const double* __restrict__ pts, long long n, GridArgs g,
                                     int32_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int id = 0;
#pragma unroll
    for (int f = D - 1; f >= 0; --f)
      id = id * g.dims[f] + cell_coord<D>(pts[i * D + f], g.origin[f], g.inv_h, g.dims[f]);
    out[i] = id;
  }
}

template <int D>
__global__ void knn_grid_kernel(const GridArgs g) {
  extern __shared__ double heap_smem[];
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= g.q) return;
  const long long qi = g.order ? g.order[t] : t;
  const int k = g.k;
  SmemHeap top;
  top.hd = heap_smem;
  top.hi = reinterpret_cast<int*>(heap_smem + (size_t)k * blockDim.x);
  top.nt = blockDim.x;
  top.t = threadIdx.x;
  double x[D];
  int c[D];
#pragma unroll
  for (int f = 0; f < D; ++f) {
    x[f] = g.queries[qi * D + f];
    c[f] = cell_coord<D>(x[f], g.origin[f], g.inv_h, g.dims[f]);
  }
  const long long self = g.self_idx ? g.self_idx[qi] : -1;
  int size = 0;                            // heap entries so far (<= k)
  double worst_d = DBL_MAX;                // the root once the heap is full
  int worst_i = INT_MAX;
  int maxr = 0;
#pragma unroll
  for (int f = 0; f < D; ++f) maxr = max(maxr, max(c[f], g.dims[f] - 1 - c[f]));

  for (int r = 0; r <= maxr; ++r) {
    // visit the cells of Chebyshev shell r (clamped to the grid)
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int f = 0; f < D; ++f) {
      lo[f] = max(c[f] - r, 0);
      hi[f] = min(c[f] + r, g.dims[f] - 1);
    }
    for (int cz = lo[2]; cz <= hi[2]; ++cz) {
      const int dz = (D > 2) ? abs(cz - c[2]) : 0;
      for (int cy = lo[1]; cy <= hi[1]; ++cy) {
        const int dy = (D > 1) ? abs(cy - c[1]) : 0;
        const bool edge_row = (dz == r) || (dy == r);
        // inside the shell's slab only the two end cells along x belong to shell r
        const int step = edge_row ? 1 : max(hi[0] - lo[0], 1);
        for (int cx = lo[0]; cx <= hi[0]; cx += step) {
          if (!edge_row && abs(cx - c[0]) != r) continue;
          const long long cell = ((long long)cz * g.dims[1] + cy) * g.dims[0] + cx;
          const int beg = g.cell_start[cell], end = g.cell_start[cell + 1];
          for (int p = beg; p < end; ++p) {
            double s = 0.0;
#pragma unroll
            for (int f = 0; f < D; ++f) {
              const double df = __dsub_rn(x[f], g.pts[(long long)p * D + f]);
              s = __dadd_rn(s, __dmul_rn(df, df));
            }
            if (s <= worst_d) {
              const int id = g.ids[p];
              if (id == self) continue;
              if (size < k) {
                top.push(size, s, id);
                if (++size == k) {
                  worst_d = top.D(0);
                  worst_i = top.I(0);
                }
              } else if (SmemHeap::less(s, id, worst_d, worst_i)) {
                top.replace_root(k, s, id);
                worst_d = top.D(0);
                worst_i = top.I(0);
              }
            }
          }
        }
      }
    }
    // every unvisited point lies outside the block of shells <= r: lower-bound its distance
    // The cell of a point is floor((x - origin) / h) in floating point and the edges below are
    // origin + c h: both carry rounding errors proportional to the COORDINATE magnitude, not
    // to the gap, so each candidate gap is reduced by an absolute slack of a few ulps of the
    // quantities it was formed from before it may justify stopping.
    double gap = DBL_MAX;
#pragma unroll
    for (int f = 0; f < D; ++f) {
      const double span = fabs(x[f]) + fabs(g.origin[f]) + (double)(c[f] + r + 1) * g.h;
      const double slack = 8.0 * DBL_EPSILON * span;
      if (c[f] - r > 0)
        gap = fmin(gap, x[f] - (g.origin[f] + (c[f] - r) * g.h) - slack);
      if (c[f] + r < g.dims[f] - 1)
        gap = fmin(gap, (g.origin[f] + (c[f] + r + 1) * g.h) - x[f] - slack);
    }
    if (gap == DBL_MAX) break;  // the whole grid has been visited
    // ... and stop only when the k-th distance is strictly inside
    gap = gap * (1.0 - 1e-12) - 1e-300;
    if (gap > 0.0 && worst_d < gap * gap) break;
  }
  // heap sort in place: the root (largest key) moves behind the shrinking heap
  for (int m = size - 1; m > 0; --m) {
    const double s = top.D(m);
    const int id = top.I(m);
    top.D(m) = top.D(0);
    top.I(m) = top.I(0);
    top.replace_root(m, s, id);
  }
  for (int i = 0; i < k; ++i) {
    g.out_idx[qi * k + i] = i < size ? top.I(i) : INT_MAX;
    g.out_d2[qi * k + i] = i < size ? top.D(i) : DBL_MAX;
  }
}

template <int D>
static int launch_grid(const GridArgs& g, cudaStream_t s) {
  // threads per CTA: as many as keep several CTAs' heaps (12 bytes per entry) resident
  int nt = 128;
  while (nt > 32 && (size_t)nt * g.k * 12 > 40 * 1024) nt >>= 1;
  const size_t smem = (size_t)nt * g.k * 12 + 8;
  cudaFuncSetAttribute(knn_grid_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)max_smem_optin());
  const unsigned blocks = (unsigned)((g.q + nt - 1) / nt);
  knn_grid_kernel<D><<<blocks, nt, smem, s>>>(g);
  return check_launch("knn_grid_kernel");
}

static int fill_grid(GridArgs& g, int d, const int32_t* dims, const double* origin, double h) {
  MGP_REQUIRE(d >= 1 && d <= 3, MGP_ERR_UNSUPPORTED, "grid KNN supports 1 <= d <= 3 (got %d)", d);
  MGP_REQUIRE(h > 0.0 && dims && origin, MGP_ERR_BAD_ARG, "bad grid description");
  for (int f = 0; f < 3; ++f) {
    g.dims[f] = f < d ? dims[f] : 1;
    g.origin[f] = f < d ? origin[f] : 0.0;
    MGP_REQUIRE(g.dims[f] >= 1, MGP_ERR_BAD_ARG, "grid dimension %d must be >= 1", f);
  }
  g.h = h;
  g.inv_h = 1.0 / h;
  return MGP_OK;
}

}  // namespace mgp

using namespace mgp;

extern "C" int mgp_knn_grid_cells(const double* points, int64_t n, int32_t d,
                                  const int32_t* dims, const double* origin, double cell_size,
                                  int32_t* out_cell, void* stream) {
  GridArgs g = {};
  int rc = fill_grid(g, d, dims, origin, cell_size);
  if (rc != MGP_OK) return rc;
  MGP_REQUIRE(n >= 0, MGP_ERR_BAD_ARG, "bad n");
  if (n == 0) return MGP_OK;
  MGP_REQUIRE(points && out_cell, MGP_ERR_BAD_ARG, "null pointer");
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (d == 1) grid_cell_ids_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(points, n, g, out_cell);
  if (d == 2) grid_cell_ids_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(points, n, g, out_cell);
  if (d == 3) grid_cell_ids_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(points, n, g, out_cell);
  return check_launch("grid_cell_ids_kernel");
}

extern "C" int mgp_knn_grid_query(const double* sorted_points, const int32_t* sorted_ids,
                                  const int32_t* cell_start, int64_t n, int32_t d,
                                  const int32_t* dims, const double* origin, double cell_size,
                                  const double* queries, const int32_t* query_order, int64_t q,
                                  int32_t k, const int64_t* self_idx, int64_t* out_idx,
                                  double* out_d2, void* stream) {
  GridArgs g = {};
  int rc = fill_grid(g, d, dims, origin, cell_size);
  if (rc != MGP_OK) return rc;
  MGP_REQUIRE(n >= 1 && q >= 0, MGP_ERR_BAD_ARG, "bad sizes n=%lld q=%lld", (long long)n,
              (long long)q);
  MGP_REQUIRE(n < (int64_t)INT_MAX, MGP_ERR_UNSUPPORTED, "train_count %lld exceeds 2^31-1",
              (long long)n);
  MGP_REQUIRE(k >= 1 && (int64_t)k + (self_idx ? 1 : 0) <= n, MGP_ERR_BAD_ARG,
              "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld",
              k, (long long)n);
  MGP_REQUIRE(k <= 256, MGP_ERR_UNSUPPORTED, "nn_count %d exceeds the supported maximum 256", k);
  if (q == 0) return MGP_OK;
  MGP_REQUIRE(sorted_points && sorted_ids && cell_start && queries && out_idx && out_d2,
              MGP_ERR_BAD_ARG, "null pointer");
  g.pts = sorted_points;
  g.ids = sorted_ids;
  g.cell_start = cell_start;
  g.queries = queries;
  g.order = query_order;
  g.self_idx = self_idx;
  g.out_idx = out_idx;
  g.out_d2 = out_d2;
  g.n = n;
  g.q = q;
  g.k = k;
  cudaStream_t s = (cudaStream_t)stream;
  if (d == 1) return launch_grid<1>(g, s);
  if (d == 2) return launch_grid<2>(g, s);
  return launch_grid<3>(g, s);
}
