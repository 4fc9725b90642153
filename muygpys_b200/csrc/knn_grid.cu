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

// One query by ONE thread: Chebyshev shells around the query's cell, a bounded max-heap of the k
// best (distance, row) keys.  The thread-per-query kernel below and the overflow path of the
// warp-per-query kernel run it.
template <int D>
__device__ void grid_query_serial(const GridArgs& g, long long qi, SmemHeap top) {
  const int k = g.k;
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
__global__ void knn_grid_kernel(const GridArgs g) {
  extern __shared__ double heap_smem[];
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= g.q) return;
  SmemHeap top;
  top.hd = heap_smem;
  top.hi = reinterpret_cast<int*>(heap_smem + (size_t)g.k * blockDim.x);
  top.nt = blockDim.x;
  top.t = threadIdx.x;
  grid_query_serial<D>(g, g.order ? g.order[t] : t, top);
}

// ---- warp per query ----------------------------------------------------------------------------
// The thread-per-query kernel is bound by latency: 100 k queries are two thirds of ONE wave of
// threads, every thread walks its own cells and sifts its own heap.  Here a WARP owns a query:
//   * a row of cells along x is one contiguous range of the cell-sorted point array, so the
//     lanes stream a whole row of the block of shells <= r with coalesced loads and append the
//     candidates (distance, row) to a list in shared memory with a ballot / prefix count;
//   * the search starts at the smallest r whose block holds k points at all (cell_start gives
//     the count without touching a point) -- no earlier shell can end the search;
//   * the list is sorted with a bitonic network over the warp, cut to the k best, and the k-th
//     key is the acceptance threshold for the next shell (if one is needed: same stopping rule,
//     same slack as above).
// Same arithmetic, same key order, same stopping rule as grid_query_serial: bit-identical
// results.  A list that would exceed its capacity (heavily clustered or duplicated data)
// sends the query through grid_query_serial on lane 0, with the list's memory as its heap.
// candidate list entries per warp: 256 up to k = 64 (two thirds more warps per SM), 512 above
constexpr int GW_WARPS = 4;    // warps (queries) per CTA
constexpr int GW_BINS = 256;   // histogram bins of the selection
constexpr int GW_MAX_K = 128;  // larger k: thread-per-query kernel

__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

// ascending bitonic sort of the first P (power of two, <= list capacity) entries by (distance, row)
__device__ __forceinline__ void warp_bitonic_sort(double* sd, int* si, int P, int lane) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int a = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int b = a | stride;
        const bool up = (a & size) == 0;
        const double da = sd[a], db = sd[b];
        const int ia = si[a], ib = si[b];
        if (key_less(db, ib, da, ia) == up) {
          sd[a] = db;
          si[a] = ib;
          sd[b] = da;
          si[b] = ia;
        }
      }
      __syncwarp();
    }
  }
}

// Exact selection of the k smallest keys of the list's first `count` (> k) entries WITHOUT sorting
// them: a most-significant-bits-first radix selection on the bit pattern of the distance (a
// non-negative double orders like its 64-bit pattern).  Each pass maps the keys of the current
// range [lo, hi] onto 256 buckets by a shift, histograms them with shared-memory atomics and
// narrows the range to the bucket that holds rank k - 1; as soon as that bucket has at most 32
// entries they are sorted by (distance, row) in registers and the k-th key read off.  The list
// is then compacted in place to the k entries <= that key (order not preserved) and their
// number is returned.  Returns -1, list untouched, if the range collapses to one distance with
// more than 32 rows (massive duplicates): the caller sorts instead.
__device__ __forceinline__ int warp_select_k(double* sd, int* si, int count, int k, int lane,
                                              unsigned* hist, double& tau_d, int& tau_i) {
  typedef unsigned long long u64;
  u64 lo = ~0ull, hi = 0ull;
  for (int i = lane; i < count; i += 32) {
    const u64 key = (u64)__double_as_longlong(sd[i]);
    lo = key < lo ? key : lo;
    hi = key > hi ? key : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  int want = k - 1;  // rank (0-based) of the wanted key among the entries in [lo, hi]
  int m = count;     // entries in [lo, hi]
  for (int pass = 0; pass < 12 && m > 32; ++pass) {
    const u64 range = hi - lo;
    if (range == 0) return -1;
    const int bits = 64 - __clzll((long long)range);  // range < 2^bits
    const int shift = bits > 8 ? bits - 8 : 0;
    for (int b = lane; b < GW_BINS; b += 32) hist[b] = 0;
    __syncwarp();
    for (int i = lane; i < count; i += 32) {
      const u64 key = (u64)__double_as_longlong(sd[i]);
      if (key >= lo && key <= hi) atomicAdd(&hist[(unsigned)((key - lo) >> shift)], 1u);
    }
    __syncwarp();
    // lane l owns bins 8 l .. 8 l + 7
    unsigned mine[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[j] = hist[8 * lane + j];
      tot += mine[j];
    }
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned excl = incl - tot;
    const bool here = (unsigned)want >= excl && (unsigned)want < incl;
    int bin = -1;
    unsigned before = 0, inbin = 0;
    if (here) {
      unsigned acc = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (bin < 0 && (unsigned)want < acc + mine[j]) {
          bin = 8 * lane + j;
          before = acc;
          inbin = mine[j];
        }
        acc += mine[j];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
    bin = __shfl_sync(0xffffffffu, bin, src);
    before = __shfl_sync(0xffffffffu, before, src);
    inbin = __shfl_sync(0xffffffffu, inbin, src);
    want -= (int)before;
    m = (int)inbin;
    const u64 nlo = lo + ((u64)bin << shift);
    const u64 nhi = nlo + (((u64)1 << shift) - 1);
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
    __syncwarp();
  }
  if (m > 32) return -1;
  // the (at most 32) entries of the final range, one per lane (through the histogram's memory),
  // sorted by (distance, row)
  double* scr_d = reinterpret_cast<double*>(hist);       // 32 doubles
  int* scr_i = reinterpret_cast<int*>(hist + 64);         // 32 ints behind them
  int got = 0;
  for (int i0 = 0; i0 < count; i0 += 32) {
    const int i = i0 + lane;
    bool in = false;
    double s = 0.0;
    if (i < count) {
      s = sd[i];
      const u64 key = (u64)__double_as_longlong(s);
      in = key >= lo && key <= hi;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int pos = got + __popc(mask & ((1u << lane) - 1));
      scr_d[pos] = s;
      scr_i[pos] = si[i];
    }
    got += __popc(mask);
  }
  __syncwarp();
  double ms = lane < got ? scr_d[lane] : DBL_MAX;
  int mi = lane < got ? scr_i[lane] : INT_MAX;
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, ms, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, stride);
      const bool lower = (lane & stride) == 0;          // this lane keeps the smaller key ...
      const bool up = (lane & size) == 0;               // ... in ascending runs
      const bool other_less = key_less(os, oi, ms, mi);
      if (other_less == (lower == up)) {
        ms = os;
        mi = oi;
      }
    }
  }
  tau_d = __shfl_sync(0xffffffffu, ms, want);
  tau_i = __shfl_sync(0xffffffffu, mi, want);
  // compact the list in place to the entries <= tau (exactly k: the keys are distinct)
  int out = 0;
  for (int i0 = 0; i0 < count; i0 += 32) {
    const int i = i0 + lane;
    double s = 0.0;
    int id = 0;
    bool keep = false;
    if (i < count) {
      s = sd[i];
      id = si[i];
      keep = !key_less(tau_d, tau_i, s, id);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const int pos = out + __popc(mask & ((1u << lane) - 1));
      sd[pos] = s;
      si[pos] = id;
    }
    out += __popc(mask);
    __syncwarp();
  }
  return out;
}

template <int D, int GW_CAP>
__global__ void __launch_bounds__(GW_WARPS * 32) knn_grid_warp_kernel(const GridArgs g) {
  __shared__ double s_d[GW_WARPS][GW_CAP];
  __shared__ int s_i[GW_WARPS][GW_CAP];
  __shared__ __align__(16) unsigned s_hist[GW_WARPS][GW_BINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t = blockIdx.x * (long long)GW_WARPS + warp;
  if (t >= g.q) return;
  const long long qi = g.order ? g.order[t] : t;
  const int k = g.k;
  double* sd = s_d[warp];
  int* si = s_i[warp];
  double x[D];
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int f = 0; f < D; ++f) {
    x[f] = g.queries[qi * D + f];
    c[f] = cell_coord<D>(x[f], g.origin[f], g.inv_h, g.dims[f]);
  }
  const long long self = g.self_idx ? g.self_idx[qi] : -1;
  const int need = k + (self >= 0 ? 1 : 0);
  int maxr = 0;
#pragma unroll
  for (int f = 0; f < D; ++f) maxr = max(maxr, max(c[f], g.dims[f] - 1 - c[f]));

  // rows (fixed cy, cz) of the block of shells <= r, clamped to the grid: lane-parallel count
  auto block_count = [&](int r) -> long long {
    const int lo0 = max(c[0] - r, 0), hi0 = min(c[0] + r, g.dims[0] - 1);
    const int lo1 = (D > 1) ? max(c[1] - r, 0) : 0, hi1 = (D > 1) ? min(c[1] + r, g.dims[1] - 1) : 0;
    const int lo2 = (D > 2) ? max(c[2] - r, 0) : 0, hi2 = (D > 2) ? min(c[2] + r, g.dims[2] - 1) : 0;
    const int n1 = hi1 - lo1 + 1, rows = n1 * (hi2 - lo2 + 1);
    long long cnt = 0;
    for (int i = lane; i < rows; i += 32) {
      const int cz = (D > 2) ? lo2 + i / n1 : 0, cy = (D > 2) ? lo1 + i % n1 : lo1 + i;
      const long long base = ((long long)cz * g.dims[1] + cy) * g.dims[0];
      cnt += g.cell_start[base + hi0 + 1] - g.cell_start[base + lo0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
  };
  int r = 0;
  while (r < maxr && block_count(r) < need) ++r;

  int count = 0;               // entries in the list
  bool full = false;           // the list holds the k best of everything scanned: tau is valid
  double tau_d = DBL_MAX;
  int tau_i = INT_MAX;
  bool overflow = false;
  // candidates of the contiguous point range [beg, end): keys below tau are appended
  auto scan_range = [&](int beg, int end) {
    for (int p0 = beg; p0 < end && !overflow; p0 += 32) {
      const int p = p0 + lane;
      bool take = false;
      double s = 0.0;
      int id = 0;
      if (p < end) {
#pragma unroll
        for (int f = 0; f < D; ++f) {
          const double df = __dsub_rn(x[f], g.pts[(long long)p * D + f]);
          s = __dadd_rn(s, __dmul_rn(df, df));
        }
        if (s <= tau_d) {
          id = g.ids[p];
          take = id != self && (!full || key_less(s, id, tau_d, tau_i));
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, take);
      const int add = __popc(m);
      if (count + add > GW_CAP) {
        overflow = true;
        break;
      }
      if (take) {
        const int pos = count + __popc(m & ((1u << lane) - 1));
        sd[pos] = s;
        si[pos] = id;
      }
      count += add;
    }
  };
  // cells of shells first..last (Chebyshev rings) as contiguous point ranges
  auto scan_shells = [&](int first, int last) {
    const int lo1 = (D > 1) ? max(c[1] - last, 0) : 0, hi1 = (D > 1) ? min(c[1] + last, g.dims[1] - 1) : 0;
    const int lo2 = (D > 2) ? max(c[2] - last, 0) : 0, hi2 = (D > 2) ? min(c[2] + last, g.dims[2] - 1) : 0;
    for (int cz = lo2; cz <= hi2 && !overflow; ++cz) {
      const int dz = (D > 2) ? abs(cz - c[2]) : 0;
      for (int cy = lo1; cy <= hi1 && !overflow; ++cy) {
        const int dy = (D > 1) ? abs(cy - c[1]) : 0;
        const long long base = ((long long)cz * g.dims[1] + cy) * g.dims[0];
        const int ring = max(dy, dz);  // every cell of this row lies in shell >= ring
        if (ring >= first) {
          // the whole row out to +-last belongs to shells first..last
          const int a = max(c[0] - last, 0), b = min(c[0] + last, g.dims[0] - 1);
          scan_range(g.cell_start[base + a], g.cell_start[base + b + 1]);
        } else {
          // only |cx - c0| in [first, last] on either side
          const int a0 = max(c[0] - last, 0), b0 = min(c[0] - first, g.dims[0] - 1);
          if (b0 >= a0) scan_range(g.cell_start[base + a0], g.cell_start[base + b0 + 1]);
          const int a1 = max(c[0] + first, 0), b1 = min(c[0] + last, g.dims[0] - 1);
          if (b1 >= a1) scan_range(g.cell_start[base + a1], g.cell_start[base + b1 + 1]);
        }
      }
    }
  };

  int scanned = -1;  // shells 0..scanned are in the list (or were rejected by tau)
  for (; r <= maxr; ++r) {
    scan_shells(scanned + 1, r);
    scanned = r;
    if (overflow) break;
    __syncwarp();
    bool selected = false;
    if (count > k) {
      // the k best by exact selection; the k-th key is the new threshold
      const int kept = warp_select_k(sd, si, count, k, lane, s_hist[warp], tau_d, tau_i);
      if (kept >= 0) {
        selected = true;
        count = kept;  // (== k: the keys are distinct)
        full = true;
      }
      __syncwarp();
    }
    if (!selected && (count > k || !full)) {
      // (few candidates, or a distance shared by many rows): sort, keep the k best
      int P = 2;
      while (P < count) P <<= 1;
      for (int i = count + lane; i < P; i += 32) {
        sd[i] = DBL_MAX;
        si[i] = INT_MAX;
      }
      __syncwarp();
      warp_bitonic_sort(sd, si, P, lane);
      if (count >= k) {
        count = k;
        full = true;
        tau_d = sd[k - 1];
        tau_i = si[k - 1];
      }
    }
    // every unvisited point lies outside the block of shells <= r (same bound and slack as
    // grid_query_serial)
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
    gap = gap * (1.0 - 1e-12) - 1e-300;
    if (full && gap > 0.0 && tau_d < gap * gap) break;
  }
  if (overflow) {
    __syncwarp();
    if (lane == 0) {
      SmemHeap top;
      top.hd = sd;
      top.hi = si;
      top.nt = 1;
      top.t = 0;
      grid_query_serial<D>(g, qi, top);
    }
    return;
  }
  __syncwarp();
  if (count <= 64) {
    // The list holds the answer: order it by (distance, row) in REGISTERS -- entries e = lane
    // and e = lane + 32, a 64-element bitonic network whose only intra-lane step is stride 32
    // (half the instructions of the shared-memory network below, no barriers).
    double d0 = lane < count ? sd[lane] : DBL_MAX, d1 = lane + 32 < count ? sd[lane + 32] : DBL_MAX;
    int i0 = lane < count ? si[lane] : INT_MAX, i1 = lane + 32 < count ? si[lane + 32] : INT_MAX;
    auto cross = [&](double& d, int& i, int stride, bool up) {
      const double od = __shfl_xor_sync(0xffffffffu, d, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
      const bool lower = (lane & stride) == 0;
      if (key_less(od, oi, d, i) == (lower == up)) {
        d = od;
        i = oi;
      }
    };
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        // row 0 (e < 32) sorts ascending, row 1 descending at size 32: a bitonic sequence of 64
        const bool up0 = (lane & size) == 0;
        cross(d0, i0, stride, size == 32 ? true : up0);
        cross(d1, i1, stride, size == 32 ? false : up0);
      }
    }
    if (key_less(d1, i1, d0, i0)) {  // stride 32
      const double td = d0;
      const int ti = i0;
      d0 = d1;
      i0 = i1;
      d1 = td;
      i1 = ti;
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
      cross(d0, i0, stride, true);
      cross(d1, i1, stride, true);
    }
    if (lane < k) {
      g.out_idx[qi * k + lane] = lane < count ? i0 : INT_MAX;
      g.out_d2[qi * k + lane] = lane < count ? d0 : DBL_MAX;
    }
    if (lane + 32 < k) {
      g.out_idx[qi * k + lane + 32] = lane + 32 < count ? i1 : INT_MAX;
      g.out_d2[qi * k + lane + 32] = lane + 32 < count ? d1 : DBL_MAX;
    }
    return;
  }
  {
    // (k > 64) the same through shared memory
    int P = 2;
    while (P < count) P <<= 1;
    for (int i = count + lane; i < P; i += 32) {
      sd[i] = DBL_MAX;
      si[i] = INT_MAX;
    }
    __syncwarp();
    warp_bitonic_sort(sd, si, P, lane);
  }
  for (int i = lane; i < k; i += 32) {
    g.out_idx[qi * k + i] = i < count ? si[i] : INT_MAX;
    g.out_d2[qi * k + i] = i < count ? sd[i] : DBL_MAX;
  }
}

template <int D>
static int launch_grid(const GridArgs& g, cudaStream_t s) {
  // A warp per query is 3.6x faster than a thread per query at 10 k queries (k = 101: the
  // thread-per-query kernel is latency-bound there), 1.45x at 100 k and 1.25x at 1 M (k = 50).
  // MGP_KNN_GRID = thread | warp overrides (dev switch).
  static const char* force = getenv("MGP_KNN_GRID");
  const bool warp_per_query = force ? force[0] == 'w' : true;
  if (g.k <= GW_MAX_K && warp_per_query) {
    const unsigned blocks = (unsigned)((g.q + GW_WARPS - 1) / GW_WARPS);
    if (g.k <= 64)
      knn_grid_warp_kernel<D, 256><<<blocks, GW_WARPS * 32, 0, s>>>(g);
    else
      knn_grid_warp_kernel<D, 512><<<blocks, GW_WARPS * 32, 0, s>>>(g);
    return check_launch("knn_grid_warp_kernel");
  }
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
