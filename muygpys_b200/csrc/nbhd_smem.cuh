// Shared-memory neighbourhood solver used by the generic fused kernel and by the
// staged batched solve.  A "team" (1, 2, 4 or 8 warps) owns one augmented matrix
//
//        [ K + eps   .      .  ]   rows 0..k-1      (lower triangle stored)
//   A =  [ kcross^T  kout   .  ]   row  k
//        [ Y^T       0      0  ]   rows k+1..k+r
//
// and runs k steps of right-looking Cholesky elimination over all m = k+1+r
// rows.  Afterwards the trailing (1+r)x(1+r) Schur complement holds
//   A[k][k]         = kout - kcross^T K^-1 kcross        (posterior variance)
//   A[k+1+c][k]     = - y_c^T K^-1 kcross                (minus posterior mean)
//   A[k+1+c][k+1+c] = - y_c^T K^-1 y_c                   (minus the scale term)
// so one factorisation yields mean, variance and the analytic-scale partial;
// the reference does three separate LU solves for the same quantities
// (S/_src/gp/muygps/numpy.py:37,63; S/_src/optimize/scale/numpy.py:14).
// Rows k+1.. hold (L^-1 Y)^T in their first k columns, which a back
// substitution turns into the fast-mean coefficients K^-1 Y (numpy.py:88-95).
#pragma once

#include "common.cuh"

namespace mgp {

template <int WARPS>
__device__ __forceinline__ void team_sync(int team_in_block) {
  if (WARPS == 1) {
    __syncwarp();
  } else {
    // named barrier 1..15, one per team in the block
    asm volatile("bar.sync %0, %1;" ::"r"(team_in_block + 1), "r"(WARPS * 32) : "memory");
  }
}

// Runs the elimination. `tid` in [0, WARPS*32). Returns false if a pivot was
// non-positive (matrix left partially factorised).
template <int WARPS>
__device__ __forceinline__ bool eliminate(double* __restrict__ A, int ld, int k, int m, int tid,
                                          int team_in_block) {
  const int lane = tid & 31;
  const int warp = tid >> 5;
  bool ok = true;
  for (int j = 0; j < k; ++j) {
    const double pivot = A[j * ld + j];
    if (!(pivot > 0.0)) {
      ok = false;
      break;  // uniform across the team: everyone reads the same value
    }
    const double inv = 1.0 / sqrt(pivot);
    team_sync<WARPS>(team_in_block);  // everyone has read the pivot
    for (int i = j + tid; i < m; i += WARPS * 32) A[i * ld + j] *= inv;
    team_sync<WARPS>(team_in_block);
    // trailing update of the lower triangle: rows over warps, columns over lanes
    for (int i = j + 1 + warp; i < m; i += WARPS) {
      const double lij = A[i * ld + j];
      for (int l = j + 1 + lane; l <= i; l += 32) {
        A[i * ld + l] = fma(-lij, A[l * ld + j], A[i * ld + l]);
      }
    }
    team_sync<WARPS>(team_in_block);
  }
  return ok;
}

// Back substitution L^T C = U for the r right-hand sides stored in rows
// k+1..k+r (U^T in columns 0..k-1).  Overwrites those rows with C^T.
template <int WARPS>
__device__ __forceinline__ void back_substitute(double* __restrict__ A, int ld, int k, int r,
                                                int tid, int team_in_block) {
  for (int i = k - 1; i >= 0; --i) {
    const double inv = 1.0 / A[i * ld + i];
    team_sync<WARPS>(team_in_block);
    // c_i for every rhs, then eliminate it from the remaining unknowns j < i
    for (int c = tid; c < r; c += WARPS * 32) A[(k + 1 + c) * ld + i] *= inv;
    team_sync<WARPS>(team_in_block);
    for (int e = tid; e < r * i; e += WARPS * 32) {
      const int c = e / i, j = e - c * i;
      double* row = A + (size_t)(k + 1 + c) * ld;
      row[j] = fma(-row[i], A[i * ld + j], row[j]);
    }
    team_sync<WARPS>(team_in_block);
  }
}

}  // namespace mgp
