// K1, feature counts above TILE_MAX_D: squared distances of one neighbourhood on the FP64
// tensor path.
//
// With d in the hundreds (BASELINE config C3: d = 784) evaluating k(k+1)/2 pairwise
// differences feature by feature is ~60k FP64 + load instructions per neighbourhood and the
// warp that owns the neighbourhood is instruction-bound.  Here the rows are CENTRED ON THE
// QUERY POINT, u_i = (x_i - q) * s, streamed once from global memory straight into
// mma.sync.m8n8k4.f64 fragments (the A fragment of row tile I and the B fragment of row tile J
// are the same registers, so a row costs one 16-byte load per lane per 8 features), and
//
//     |u_i - u_j|^2 = |u_i|^2 + |u_j|^2 - 2 u_i . u_j
//
// is taken from the Gram tiles: 1 DMMA per 8x8 tile per 4 features instead of 64 lane-FMAs plus
// their operand loads.  Centring keeps |u_i| of the order of the neighbourhood radius, so the
// cancellation in the identity is bounded by (|u_i|^2 + |u_j|^2) / |u_i - u_j|^2; pairs where
// that ratio exceeds 2^8 (near-duplicate neighbours, or a query far outside its neighbourhood)
// are recomputed by direct differences in gram_fixup(), which keeps every distance within a
// few ulp of the oracle's.  The query row itself is u = 0, so the cross-covariance distances
// are the row norms, which are accumulated lane-locally next to the DMMAs.
//
// Output: RAW (length-scaled) squared distances written into the tile image at the cells the
// covariance pass (assemble<F, -1> in fused_tile.cu) then transforms in place.
#pragma once

#include "tile_common.cuh"

namespace mgp {
namespace {

constexpr double GRAM_FIXUP_RATIO = 1.0 / 256.0;
#ifndef MGP_GRAM_CHUNKS
#define MGP_GRAM_CHUNKS 1
#endif

struct GramCtx {
  const double* train_x;
  const double* qrow;       // the query point's row
  const int64_t* nn_row;    // this neighbourhood's k train rows
  const double* scale;      // per-feature multipliers (shared memory) or nullptr
  double cs2;               // isotropic: squared multiplier applied to the finished distances
  int k, kp, d;
};

// Tiles (I0 + i, J0 + j), i < NI, j < NJ of the Gram matrix; DIAG: I0 == J0, lower half only.
// VEC: d is even and the arrays are 16-byte aligned, so a lane's two features are one LDG.128.
// SCALED: per-feature multipliers g.scale (anisotropic deformation).
template <int NI, int NJ, bool DIAG, bool VEC, bool SCALED>
__device__ __noinline__ void gram_block(double* __restrict__ tiles, const GramCtx g, int I0,
                                        int J0, int lane) {
  constexpr int NJR = DIAG ? 1 : NJ;  // separately streamed column-side row tiles
  const int rho = lane >> 2, q = lane & 3;
  const int fo0 = VEC ? 2 * q : q, fo1 = VEC ? 2 * q + 1 : q + 4;
  const int d = g.d, k = g.k;
  const double* rI[NI];
  const double* rJ[NJR];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int p = 8 * (I0 + i) + rho;
    rI[i] = p < k ? g.train_x + g.nn_row[p] * (long long)d : g.qrow;  // p >= k: u = 0
  }
#pragma unroll
  for (int j = 0; j < NJR; ++j) {
    const int p = 8 * (J0 + j) + rho;
    rJ[j] = (!DIAG && p < k) ? g.train_x + g.nn_row[p] * (long long)d : g.qrow;
  }
  auto ld = [&](const double* p, int f0) -> double2 {
    if (VEC) {
      return (f0 + fo0 < d) ? *reinterpret_cast<const double2*>(p + f0 + fo0)
                            : make_double2(0.0, 0.0);
    }
    double2 v;
    v.x = (f0 + fo0 < d) ? p[f0 + fo0] : 0.0;
    v.y = (f0 + fo1 < d) ? p[f0 + fo1] : 0.0;
    return v;
  };

  double acc[NI][NJ][2];
  double nI[NI], nJ[NJR];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    nI[i] = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < NJR; ++j) nJ[j] = 0.0;

  // One step = CH chunks of 8 features, requested back to back so that every row is read in
  // runs of 64 * CH contiguous bytes.  The loaded values are consumed (centred, scaled) into
  // the fragment registers FIRST and only then is the next step requested into the same
  // buffer registers: hardware scoreboards are shared between loads, so a wait that is issued
  // after newer loads would also wait for those (measured: 50 % of the kernel's stall samples
  // sat on that one wait with a register ring).  The DMMAs of this step cover the latency of
  // the next one.  Loads past d return 0.
  constexpr int CH = (!DIAG && NI + NJ > 6) ? 1 : MGP_GRAM_CHUNKS;
  double2 bI[CH][NI], bJ[CH][NJR], bq[CH];
  auto request = [&](int f0) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
#pragma unroll
      for (int i = 0; i < NI; ++i) bI[c][i] = ld(rI[i], f0 + 8 * c);
#pragma unroll
      for (int j = 0; j < NJR; ++j)
        bJ[c][j] = DIAG ? make_double2(0.0, 0.0) : ld(rJ[j], f0 + 8 * c);
      bq[c] = ld(g.qrow, f0 + 8 * c);
    }
  };
  request(0);
  for (int fb = 0; fb < d; fb += 8 * CH) {
    double2 uI[CH][NI], uJ[CH][NJR];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int f0 = fb + 8 * c;
      double sx = 1.0, sy = 1.0;
      if (SCALED) {
        sx = (f0 + fo0 < d) ? g.scale[f0 + fo0] : 0.0;
        sy = (f0 + fo1 < d) ? g.scale[f0 + fo1] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        uI[c][i].x = bI[c][i].x - bq[c].x;
        uI[c][i].y = bI[c][i].y - bq[c].y;
        if (SCALED) {
          uI[c][i].x *= sx;
          uI[c][i].y *= sy;
        }
      }
#pragma unroll
      for (int j = 0; j < NJR; ++j) {
        uJ[c][j].x = bJ[c][j].x - bq[c].x;
        uJ[c][j].y = bJ[c][j].y - bq[c].y;
        if (SCALED) {
          uJ[c][j].x *= sx;
          uJ[c][j].y *= sy;
        }
      }
    }
    request(fb + 8 * CH);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        nI[i] = fma(uI[c][i].x, uI[c][i].x, nI[i]);
        nI[i] = fma(uI[c][i].y, uI[c][i].y, nI[i]);
      }
      if (!DIAG) {
#pragma unroll
        for (int j = 0; j < NJR; ++j) {
          nJ[j] = fma(uJ[c][j].x, uJ[c][j].x, nJ[j]);
          nJ[j] = fma(uJ[c][j].y, uJ[c][j].y, nJ[j]);
        }
      }
      // all tiles for the even features, then all tiles for the odd ones: consecutive
      // DMMAs never touch the same accumulator
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (!DIAG || j <= i)
            dmma_free(acc[i][j][0], acc[i][j][1], uI[c][i].x, DIAG ? uI[c][j].x : uJ[c][j].x);
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (!DIAG || j <= i)
            dmma_free(acc[i][j][0], acc[i][j][1], uI[c][i].y, DIAG ? uI[c][j].y : uJ[c][j].y);
    }
  }

  // row norms: the four lanes of a quad hold partial sums of the same row
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    nI[i] += __shfl_xor_sync(0xffffffffu, nI[i], 1);
    nI[i] += __shfl_xor_sync(0xffffffffu, nI[i], 2);
  }
#pragma unroll
  for (int j = 0; j < NJR; ++j) {
    nJ[j] += __shfl_xor_sync(0xffffffffu, nJ[j], 1);
    nJ[j] += __shfl_xor_sync(0xffffffffu, nJ[j], 2);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const double nj = DIAG ? nI[j] : nJ[j];
    const double c0 = shfl_d(nj, (2 * q) * 4);      // |u|^2 of column point 8(J0+j) + 2q
    const double c1 = shfl_d(nj, (2 * q + 1) * 4);  //                     ... + 2q + 1
    const int col = 8 * (J0 + j) + 2 * q;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (DIAG && j > i) continue;
      const int row = 8 * (I0 + i) + rho;
      double u0 = fmax(0.0, fma(-2.0, acc[i][j][0], nI[i] + c0)) * g.cs2;
      double u1 = fmax(0.0, fma(-2.0, acc[i][j][1], nI[i] + c1)) * g.cs2;
      if (row == col) u0 = 0.0;
      if (row == col + 1) u1 = 0.0;
      // point k is the query: its distances are the cross row kp of the image
      const int irow = (row == k) ? g.kp : row;
      const int jmax = (row < k) ? row : (row == k ? k - 1 : -1);
      double* dst = tiles + elem_off(irow, col);
      if (col + 1 <= jmax)
        *reinterpret_cast<double2*>(dst) = make_double2(u0, u1);
      else if (col <= jmax)
        *dst = u0;
    }
  }
}

template <bool VEC, bool SCALED>
__device__ __forceinline__ void gram_distances(double* tiles, const GramCtx& g, int lane) {
  const int TR = (g.k + 1 + 7) >> 3;  // row tiles that hold the k neighbours and the query
  for (int A = 0; 4 * A < TR; ++A) {
    const int na = min(4, TR - 4 * A);
    for (int B = 0; B < A; ++B) {
      switch (na) {
        case 1: gram_block<1, 4, false, VEC, SCALED>(tiles, g, 4 * A, 4 * B, lane); break;
        case 2: gram_block<2, 4, false, VEC, SCALED>(tiles, g, 4 * A, 4 * B, lane); break;
        case 3: gram_block<3, 4, false, VEC, SCALED>(tiles, g, 4 * A, 4 * B, lane); break;
        default: gram_block<4, 4, false, VEC, SCALED>(tiles, g, 4 * A, 4 * B, lane);
      }
    }
    switch (na) {
      case 1: gram_block<1, 1, true, VEC, SCALED>(tiles, g, 4 * A, 4 * A, lane); break;
      case 2: gram_block<2, 2, true, VEC, SCALED>(tiles, g, 4 * A, 4 * A, lane); break;
      case 3: gram_block<3, 3, true, VEC, SCALED>(tiles, g, 4 * A, 4 * A, lane); break;
      default: gram_block<4, 4, true, VEC, SCALED>(tiles, g, 4 * A, 4 * A, lane);
    }
  }
}

// Recompute, by direct differences, every pair whose Gram-identity distance lost more than
// ~8 bits to cancellation.  The row norms |u_i|^2 are the cross row (kp) of the image.
__device__ __noinline__ void gram_fixup(double* __restrict__ tiles,
                                        const unsigned* __restrict__ etab, const GramCtx g,
                                        int lane) {
  const int tri = g.k * (g.k + 1) / 2;
  for (int base = 0; base < tri; base += 32) {
    const int e = base + lane;
    const unsigned p = etab[e < tri ? e : 0];
    const int pi = (p >> 8) & 255, pj = p & 255;
    const double u2 = tiles[p >> 16];
    const double lim = GRAM_FIXUP_RATIO * (tiles[elem_off(g.kp, pi)] + tiles[elem_off(g.kp, pj)]);
    unsigned m = __ballot_sync(0xffffffffu, e < tri && pi != pj && u2 < lim);
    while (m) {
      const int s = __ffs(m) - 1;
      m &= m - 1;
      const int ii = __shfl_sync(0xffffffffu, pi, s), jj = __shfl_sync(0xffffffffu, pj, s);
      const double* a = g.train_x + g.nn_row[ii] * (long long)g.d;
      const double* b = g.train_x + g.nn_row[jj] * (long long)g.d;
      double sum = 0.0;
      for (int f = lane; f < g.d; f += 32) {
        double df = a[f] - b[f];
        if (g.scale) df *= g.scale[f];
        sum = fma(df, df, sum);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == s) tiles[p >> 16] = sum * g.cs2;
    }
  }
}

}  // namespace
}  // namespace mgp
