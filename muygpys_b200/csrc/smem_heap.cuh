// Bounded max-heap in shared memory shared by the KNN kernels (knn_grid.cu, knn_gram.cu).
#pragma once

namespace mgp {

// Per-query bounded MAX-heap on the key (distance, train row), resident in SHARED memory
// (element i of thread t at [i * NT + t]: the lanes of a warp sit at different heap positions
// but always in different banks).  Per-thread arrays in local memory spilled to L2 -- 1.5 MB of
// lists per SM -- and every shift of the sorted insertion was an L2 round trip; the heap needs
// log2(k) shared-memory steps per accepted candidate.  The root is the current worst entry.
// IdxT is the stored row type (int, or unsigned short for in-slice offsets below 65536).
template <typename IdxT>
struct SmemHeapT {
  double* hd;
  IdxT* hi;
  int nt, t;
  __device__ __forceinline__ double& D(int i) { return hd[i * nt + t]; }
  __device__ __forceinline__ IdxT& I(int i) { return hi[i * nt + t]; }
  static __device__ __forceinline__ bool less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
  }
  // heap of `size` < k entries: add (s, id)
  __device__ __forceinline__ void push(int size, double s, int id) {
    int i = size;
    while (i > 0) {
      const int p = (i - 1) >> 1;
      const double dp = D(p);
      const int ip = I(p);
      if (!less(dp, ip, s, id)) break;
      D(i) = dp;
      I(i) = (IdxT)ip;
      i = p;
    }
    D(i) = s;
    I(i) = (IdxT)id;
  }
  // replace the root of a heap of `size` entries by (s, id) and restore the heap
  __device__ __forceinline__ void replace_root(int size, double s, int id) {
    int i = 0;
    for (;;) {
      int c = 2 * i + 1;
      if (c >= size) break;
      double dc = D(c);
      int ic = I(c);
      if (c + 1 < size) {
        const double d2 = D(c + 1);
        const int i2 = I(c + 1);
        if (less(dc, ic, d2, i2)) {
          dc = d2;
          ic = i2;
          ++c;
        }
      }
      if (!less(s, id, dc, ic)) break;
      D(i) = dc;
      I(i) = (IdxT)ic;
      i = c;
    }
    D(i) = s;
    I(i) = (IdxT)id;
  }
};

using SmemHeap = SmemHeapT<int>;

}  // namespace mgp
