// K2: exact brute-force k-nearest-neighbour search in squared l2.
// Replaces NN_Wrapper._get_nns for nn_method="exact" (S/neighbors.py:213-262):
// indices int64 ascending by distance, distances SQUARED.
//
// Distances are accumulated feature by feature with separately rounded multiply
// and add (no FMA contraction, same order as a scalar CPU loop) so that the
// ranking -- including near-ties -- is the one a direct-difference CPU search
// produces.  Exact ties resolve to the lower train index.
//
//  * d <= 8  : one thread per query, query in registers, train rows staged in
//              shared memory tiles that every thread of the block scans
//              (broadcast reads), private sorted top-k list per thread.
//  * d  > 8  : one warp per query, features split across lanes.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace mgp {

constexpr int KNN_TILE_DOUBLES = 4096;  // 32 KB of staged train rows (small-d path)

template <int KMAX>
struct TopK {
  double dist[KMAX];
  int idx[KMAX];
  __device__ __forceinline__ void init(int k) {
    for (int i = 0; i < k; ++i) {
      dist[i] = DBL_MAX;
      idx[i] = INT_MAX;
    }
  }
  // candidates arrive in increasing index order, so `<` keeps the lower index on ties
  __device__ __forceinline__ void push(int k, double dv, int iv) {
    int pos = k - 1;
    while (pos > 0 && dist[pos - 1] > dv) {
      dist[pos] = dist[pos - 1];
      idx[pos] = idx[pos - 1];
      --pos;
    }
    dist[pos] = dv;
    idx[pos] = iv;
  }
};

template <int D, int KMAX>
__global__ void __launch_bounds__(128) knn_small_d_kernel(
    const double* __restrict__ train, long long n, const double* __restrict__ queries,
    long long q, int k, const int64_t* __restrict__ self_idx, int64_t* __restrict__ out_idx,
    double* __restrict__ out_d2) {
  constexpr int KNN_TILE = KNN_TILE_DOUBLES / D;
  __shared__ double tile[KNN_TILE * D];
  const long long qi = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool active = qi < q;
  double x[D];
#pragma unroll
  for (int f = 0; f < D; ++f) x[f] = active ? queries[qi * D + f] : 0.0;
  const long long self = (active && self_idx) ? self_idx[qi] : -1;
  TopK<KMAX> top;
  top.init(k);
  double worst = DBL_MAX;
  for (long long base = 0; base < n; base += KNN_TILE) {
    const int cnt = (int)min((long long)KNN_TILE, n - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * D; e += blockDim.x) tile[e] = train[base * D + e];
    __syncthreads();
    if (!active) continue;
    for (int j = 0; j < cnt; ++j) {
      double s = 0.0;
#pragma unroll
      for (int f = 0; f < D; ++f) {
        const double df = __dsub_rn(x[f], tile[j * D + f]);
        s = __dadd_rn(s, __dmul_rn(df, df));
      }
      if (s < worst && base + j != self) {
        top.push(k, s, (int)(base + j));
        worst = top.dist[k - 1];
      }
    }
  }
  if (active) {
    for (int i = 0; i < k; ++i) {
      out_idx[qi * k + i] = top.idx[i];
      out_d2[qi * k + i] = top.dist[i];
    }
  }
}

// general d: warp per query; lanes own features f = lane, lane+32, ...
template <int KMAX>
__global__ void __launch_bounds__(128) knn_warp_kernel(
    const double* __restrict__ train, long long n, const double* __restrict__ queries,
    long long q, int d, int k, const int64_t* __restrict__ self_idx,
    int64_t* __restrict__ out_idx, double* __restrict__ out_d2) {
  extern __shared__ double qs[];  // warps_per_block x d query features
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const long long qi = blockIdx.x * (long long)wpb + warp;
  if (qi >= q) return;
  double* xq = qs + (size_t)warp * d;
  for (int f = lane; f < d; f += 32) xq[f] = queries[qi * d + f];
  __syncwarp();
  const long long self = self_idx ? self_idx[qi] : -1;
  TopK<KMAX> top;  // maintained redundantly by every lane (uniform control flow)
  top.init(k);
  double worst = DBL_MAX;
  for (long long j = 0; j < n; ++j) {
    const double* y = train + j * d;
    double s = 0.0;
    for (int f = lane; f < d; f += 32) {
      const double df = __dsub_rn(xq[f], y[f]);
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    s = warp_sum(s);
    if (s < worst && j != self) {
      top.push(k, s, (int)j);
      worst = top.dist[k - 1];
    }
  }
  for (int i = lane; i < k; i += 32) {
    out_idx[qi * k + i] = top.idx[i];
    out_d2[qi * k + i] = top.dist[i];
  }
}

size_t knn_tiled_workspace_bytes(long long n, long long q, int k);
int launch_knn_tiled(const double* train, long long n, const double* queries, long long q, int d,
                     int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2, void* ws,
                     size_t ws_bytes, cudaStream_t s, const int32_t* qmap, const int32_t* qcount);
bool knn_gram_supported(long long n, long long q, int d, int k);
size_t knn_gram_workspace_bytes(long long n, long long q, int k);
int launch_knn_gram(const double* train, long long n, const double* queries, long long q, int d,
                    int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2, void* ws,
                    size_t ws_bytes, cudaStream_t s);

template <int KMAX>
static int launch_knn(const double* train, long long n, const double* queries, long long q,
                      int d, int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2,
                      cudaStream_t s) {
  if (d <= 8) {
    int threads = 128;
    if (q < (long long)sm_count() * 128) threads = 64;
    if (q < (long long)sm_count() * 64) threads = 32;
    const unsigned blocks = (unsigned)((q + threads - 1) / threads);
#define MGP_KNN_D(D)                                                                           \
  case D:                                                                                      \
    knn_small_d_kernel<D, KMAX><<<blocks, threads, 0, s>>>(train, n, queries, q, k, self_idx,  \
                                                           out_idx, out_d2);                   \
    break;
    switch (d) {
      MGP_KNN_D(1)
      MGP_KNN_D(2)
      MGP_KNN_D(3)
      MGP_KNN_D(4)
      MGP_KNN_D(5)
      MGP_KNN_D(6)
      MGP_KNN_D(7)
      MGP_KNN_D(8)
    }
#undef MGP_KNN_D
    return check_launch("knn_small_d_kernel");
  }
  const int wpb = 4;
  const unsigned blocks = (unsigned)((q + wpb - 1) / wpb);
  const size_t smem = (size_t)wpb * d * sizeof(double);
  MGP_REQUIRE(smem <= 48 * 1024, MGP_ERR_UNSUPPORTED, "feature count d=%d too large for KNN", d);
  knn_warp_kernel<KMAX><<<blocks, wpb * 32, smem, s>>>(train, n, queries, q, d, k, self_idx,
                                                       out_idx, out_d2);
  return check_launch("knn_warp_kernel");
}

}  // namespace mgp

using namespace mgp;

extern "C" size_t mgp_knn_workspace_bytes(int64_t n, int64_t q, int32_t d, int32_t k) {
  if (n < 1 || q < 8 || k < 1) return 0;
  if (mgp::knn_gram_supported(n, q, d, k)) return mgp::knn_gram_workspace_bytes(n, q, k);
  if (d <= 8) return 0;
  return mgp::knn_tiled_workspace_bytes(n, q, k);
}

extern "C" int mgp_knn(const double* train, int64_t n, const double* queries, int64_t q,
                       int32_t d, int32_t k, int32_t exclude_self, const int64_t* self_idx,
                       int64_t* out_idx, double* out_d2, void* ws, size_t ws_bytes,
                       void* stream) {
  MGP_REQUIRE(n >= 1 && q >= 0 && d >= 1, MGP_ERR_BAD_ARG, "bad sizes n=%lld q=%lld d=%d",
              (long long)n, (long long)q, d);
  MGP_REQUIRE(n < (int64_t)INT_MAX, MGP_ERR_UNSUPPORTED, "train_count %lld exceeds 2^31-1",
              (long long)n);
  MGP_REQUIRE(k >= 1 && (int64_t)k + (exclude_self ? 1 : 0) <= n, MGP_ERR_BAD_ARG,
              "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld",
              k, (long long)n);
  MGP_REQUIRE(k <= 256, MGP_ERR_UNSUPPORTED, "nn_count %d exceeds the supported maximum 256", k);
  if (q == 0) return MGP_OK;
  MGP_REQUIRE(train && queries && out_idx && out_d2, MGP_ERR_BAD_ARG, "null pointer");
  MGP_REQUIRE(!exclude_self || self_idx, MGP_ERR_BAD_ARG, "exclude_self needs self_idx");
  const int64_t* self = exclude_self ? self_idx : nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  // DMMA pre-filter + certified exact re-rank where it pays
  if (q >= 8 && knn_gram_supported(n, q, d, k))
    return launch_knn_gram(train, n, queries, q, d, k, self, out_idx, out_d2, ws, ws_bytes, s);
  if (d > 8 && q >= 8)  // register-tiled exact sweep; a handful of queries take the warp kernel
    return launch_knn_tiled(train, n, queries, q, d, k, self, out_idx, out_d2, ws, ws_bytes, s,
                            nullptr, nullptr);
  if (k <= 64) return launch_knn<64>(train, n, queries, q, d, k, self, out_idx, out_d2, s);
  if (k <= 128) return launch_knn<128>(train, n, queries, q, d, k, self, out_idx, out_d2, s);
  return launch_knn<256>(train, n, queries, q, d, k, self, out_idx, out_d2, s);
}
