// K2 (high-dimensional variant): exact brute-force KNN for d > 8 with register tiling.
//
// A CTA owns 64 queries and sweeps the training set in tiles of 64 points; each of the 256
// threads accumulates a 4x4 block of squared distances (queries 4 ty + i, points tx + 16 j: with
// the odd row pitch the 16 point rows a half-warp reads land in 16 different banks) over
// feature chunks staged in shared memory, so every staged value feeds four subtractions.  Distances are accumulated feature by
// feature in ascending order with separately rounded subtract / multiply / add -- the same
// arithmetic, in the same order, as a scalar CPU loop and as the small-d kernel -- which makes
// the ranking (ties to the lower train row) bit-identical to the oracle's.  This is why the
// Gram trick (||q||^2 + ||x||^2 - 2 q.x on DMMA) is NOT used for the final ranking: it perturbs
// near-ties; it remains a possible conservative pre-filter ("next").
// FP64-bound: 3 FP64 issue slots per (query, point, feature).
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace mgp {

constexpr int KT_Q = 64;    // queries per CTA
constexpr int KT_X = 64;    // train points per tile
constexpr int KT_F = 32;    // features per staged chunk
constexpr int KT_LD = KT_F + 1;

template <int KMAX>
__global__ void __launch_bounds__(256) knn_tiled_kernel(
    const double* __restrict__ train, long long n, const double* __restrict__ queries,
    long long q, int d, int k, const int64_t* __restrict__ self_idx, long long split_len,
    int nsplit, int32_t* __restrict__ part_idx, double* __restrict__ part_d2,
    int64_t* __restrict__ out_idx, double* __restrict__ out_d2,
    const int32_t* __restrict__ qmap, const int32_t* __restrict__ qcount) {
  __shared__ double stage[2 * KT_Q * KT_LD];
  // optional indirection (re-run of the queries the Gram pre-filter could not certify):
  // local query j is row qmap[j] of `queries` / `self_idx` / the outputs, *qcount of them
  if (qcount) q = *qcount;
  if ((long long)blockIdx.x * KT_Q >= q) return;
  double* Qs = stage;
  double* Xs = stage + KT_Q * KT_LD;
  double* Ds = stage;  // the 64 x 65 tile of squared distances reuses the staging area
  static_assert(KT_Q * (KT_X + 1) <= 2 * KT_Q * KT_LD, "distance tile must fit the staging area");
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const long long q0 = (long long)blockIdx.x * KT_Q;

  // thread t < 64 owns query q0 + t's result list
  double best_d[KMAX];
  int best_i[KMAX];
  const bool owner = tid < KT_Q && q0 + tid < q;
  const long long my_row = owner ? (qmap ? (long long)qmap[q0 + tid] : q0 + tid) : 0;
  long long self = -1;
  if (owner) {
    for (int i = 0; i < k; ++i) {
      best_d[i] = DBL_MAX;
      best_i[i] = INT_MAX;
    }
    if (self_idx) self = self_idx[my_row];
  }
  double worst = DBL_MAX;

  // blockIdx.y sweeps its own slice of the training set (keeps the GPU busy for small q)
  const long long x_begin = (long long)blockIdx.y * split_len;
  const long long x_end = min(n, x_begin + split_len);
  for (long long x0 = x_begin; x0 < x_end; x0 += KT_X) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int f0 = 0; f0 < d; f0 += KT_F) {
      const int fc = min(KT_F, d - f0);
      __syncthreads();
      for (int e = tid; e < KT_Q * KT_F; e += 256) {
        const int r = e / KT_F, f = e - r * KT_F;
        const long long qi = q0 + r, xi = x0 + r;
        const long long qrow = (qmap && qi < q) ? (long long)qmap[qi] : qi;
        Qs[r * KT_LD + f] = (qi < q && f < fc) ? queries[qrow * d + f0 + f] : 0.0;
        Xs[r * KT_LD + f] = (xi < x_end && f < fc) ? train[xi * d + f0 + f] : 0.0;
      }
      __syncthreads();
      for (int f = 0; f < fc; ++f) {
        double qv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qv[i] = Qs[(ty * 4 + i) * KT_LD + f];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = Xs[(tx + 16 * j) * KT_LD + f];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double df = __dsub_rn(qv[i], xv[j]);
            acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(df, df));
          }
      }
    }
    __syncthreads();  // everyone is done with the staged features
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ds[(ty * 4 + i) * (KT_X + 1) + tx + 16 * j] = acc[i][j];
    __syncthreads();
    if (owner) {
      const int cnt = (int)min((long long)KT_X, x_end - x0);
      for (int j = 0; j < cnt; ++j) {
        const double s = Ds[tid * (KT_X + 1) + j];
        if (s < worst && x0 + j != self) {
          int pos = k - 1;
          while (pos > 0 && best_d[pos - 1] > s) {
            best_d[pos] = best_d[pos - 1];
            best_i[pos] = best_i[pos - 1];
            --pos;
          }
          best_d[pos] = s;
          best_i[pos] = (int)(x0 + j);
          worst = best_d[k - 1];
        }
      }
    }
  }
  if (owner) {
    if (nsplit == 1) {
      for (int i = 0; i < k; ++i) {
        out_idx[my_row * k + i] = best_i[i];
        out_d2[my_row * k + i] = best_d[i];
      }
    } else {
      const long long base = ((q0 + tid) * nsplit + blockIdx.y) * k;
      for (int i = 0; i < k; ++i) {
        part_idx[base + i] = best_i[i];
        part_d2[base + i] = best_d[i];
      }
    }
  }
}

// Merge the per-slice lists of one query.  Slices cover ascending index ranges and each list
// is sorted by (distance, index), so taking slices in order with a strict `<` keeps the lower
// train row on ties, exactly like the single-sweep kernel.
template <int KMAX>
__global__ void knn_merge_kernel(const int32_t* __restrict__ part_idx,
                                 const double* __restrict__ part_d2, long long q, int nsplit,
                                 int k, int64_t* __restrict__ out_idx,
                                 double* __restrict__ out_d2, const int32_t* __restrict__ qmap,
                                 const int32_t* __restrict__ qcount) {
  const long long qi = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (qcount) q = *qcount;
  if (qi >= q) return;
  const long long orow = qmap ? (long long)qmap[qi] : qi;
  double best_d[KMAX];
  int best_i[KMAX];
  for (int i = 0; i < k; ++i) {
    best_d[i] = DBL_MAX;
    best_i[i] = INT_MAX;
  }
  for (int s = 0; s < nsplit; ++s) {
    const long long base = (qi * nsplit + s) * k;
    for (int j = 0; j < k; ++j) {
      const double dv = part_d2[base + j];
      if (!(dv < best_d[k - 1])) break;  // the rest of this (sorted) list cannot enter
      int pos = k - 1;
      while (pos > 0 && best_d[pos - 1] > dv) {
        best_d[pos] = best_d[pos - 1];
        best_i[pos] = best_i[pos - 1];
        --pos;
      }
      best_d[pos] = dv;
      best_i[pos] = part_idx[base + j];
    }
  }
  for (int i = 0; i < k; ++i) {
    out_idx[orow * k + i] = best_i[i];
    out_d2[orow * k + i] = best_d[i];
  }
}

static int tiled_splits(long long n, long long q) {
  const long long qblocks = (q + KT_Q - 1) / KT_Q;
  long long want = (4LL * sm_count() + qblocks - 1) / qblocks;  // ~4 CTAs per SM in total
  const long long max_by_n = (n + 4095) / 4096;                  // >= 4096 points per slice
  if (want > max_by_n) want = max_by_n;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

size_t knn_tiled_workspace_bytes(long long n, long long q, int k) {
  const int ns = tiled_splits(n, q);
  return ns == 1 ? 0 : (size_t)q * ns * k * (sizeof(int32_t) + sizeof(double)) + 16;
}

// qmap / qcount (device, optional): run only the *qcount queries listed in qmap (the grid is
// still sized for q; surplus CTAs exit at once)
int launch_knn_tiled(const double* train, long long n, const double* queries, long long q, int d,
                     int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2, void* ws,
                     size_t ws_bytes, cudaStream_t s, const int32_t* qmap,
                     const int32_t* qcount) {
  const int ns = tiled_splits(n, q);
  MGP_REQUIRE(ws_bytes >= knn_tiled_workspace_bytes(n, q, k) && (ns == 1 || ws != nullptr),
              MGP_ERR_WORKSPACE, "KNN workspace too small (%zu bytes)", ws_bytes);
  double* part_d2 = (double*)ws;  // q * ns * k doubles, then the int32 indices
  int32_t* part_idx = ns == 1 ? nullptr : (int32_t*)(part_d2 + (size_t)q * ns * k);
  long long split_len = (n + ns - 1) / ns;
  split_len = (split_len + KT_X - 1) / KT_X * KT_X;
  const dim3 grid((unsigned)((q + KT_Q - 1) / KT_Q), (unsigned)ns);
#define MGP_KT(KM)                                                                            \
  do {                                                                                        \
    knn_tiled_kernel<KM><<<grid, 256, 0, s>>>(train, n, queries, q, d, k, self_idx, split_len, \
                                              ns, part_idx, part_d2, out_idx, out_d2, qmap,   \
                                              qcount);                                        \
    if (ns > 1)                                                                               \
      knn_merge_kernel<KM><<<(unsigned)((q + 127) / 128), 128, 0, s>>>(                       \
          part_idx, part_d2, q, ns, k, out_idx, out_d2, qmap, qcount);                        \
  } while (0)
  if (k <= 64)
    MGP_KT(64);
  else if (k <= 128)
    MGP_KT(128);
  else
    MGP_KT(256);
#undef MGP_KT
  return check_launch("knn_tiled_kernel");
}

}  // namespace mgp
