// The register-tile fused kernel template and its launcher, shared by two translation units
// (fused_tile.cu: d <= 8, fused_tile_gram.cu: d > 8) so that the ~26 instantiations compile in
// parallel.  See fused_tile.cu for the description of the algorithm.
#pragma once

#include <cstdint>
#include <cstdlib>

#include "gram.cuh"
#include "tile_common.cuh"

namespace mgp {

namespace {

// Flat, perfectly balanced evaluation of the k(k+1)/2 + k covariances; each table
// entry is (tile-image offset << 16) | (row point << 8) | column point.  Writes
// N = -K.  Compiled once per (formula, D) and called from every tile kernel.
template <int F, int D>
__device__ __noinline__ void assemble(double* __restrict__ tiles, const double* __restrict__ pts,
                                      const unsigned* __restrict__ etab,
                                      double tab64, int n_elem, int lane,
                                      int d, double post_scale, int kernel_id) {
#ifndef MGP_ASM_WAYS
#define MGP_ASM_WAYS 2
#endif
  constexpr int W = MGP_ASM_WAYS;  // independent elements in flight per lane
  // uniform trip count (the exp table is read with warp shuffles): lanes past the end
  // evaluate a dummy entry that lands in the scratch cell behind the image
  for (int base = 0; base < n_elem; base += 32 * W) {
    unsigned p[W];
    double u[W], v[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int e = base + 32 * w + lane;
      p[w] = etab[e < n_elem ? e : n_elem];
    }
#pragma unroll
    for (int w = 0; w < W; ++w)
      u[w] = (D < 0) ? tiles[p[w] >> 16]  // gram mode: the raw distance is already in place
                     : sqdist<(D < 0 ? 0 : D)>(pts, (p[w] >> 8) & 255, p[w] & 255, d);
#pragma unroll
    for (int w = 0; w < W; ++w) v[w] = neg_cov<F>(u[w], tab64, post_scale, kernel_id);
    // lanes past the end evaluated the dummy entry: do not store it (every such lane would
    // write the same scratch cell -- harmless, but a write-write hazard for racecheck)
#pragma unroll
    for (int w = 0; w < W; ++w)
      if (base + 32 * w + lane < n_elem) tiles[p[w] >> 16] = v[w];
  }
}

template <int F>
__device__ __forceinline__ void assemble_d(double* tiles, const double* pts, const unsigned* etab,
                                           double tab64, int n_elem, int lane, int d,
                                           double post_scale, int kernel_id) {
  switch (d) {
    case 1: assemble<F, 1>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case 2: assemble<F, 2>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case 3: assemble<F, 3>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case -1: assemble<F, -1>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    default: assemble<F, 0>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id);
  }
}

__device__ __forceinline__ void assemble_any(int formula, double* tiles, const double* pts,
                                             const unsigned* etab, double tab64,
                                             int n_elem, int lane, int d, double post_scale,
                                             int kernel_id) {
  switch (formula) {
    case F_M05: assemble_d<F_M05>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case F_M15: assemble_d<F_M15>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case F_M25: assemble_d<F_M25>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case F_GAUSS: assemble_d<F_GAUSS>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    case F_RBF_L2: assemble_d<F_RBF_L2>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id); break;
    default: assemble_d<F_F2_ANY>(tiles, pts, etab, tab64, n_elem, lane, d, post_scale, kernel_id);
  }
}


// SMEM_L = false: finished tiles live in registers (T <= 7, three CTAs per SM).
// SMEM_L = true : finished tiles are written back over their own cells of the shared-memory
//                 image and re-read as DMMA fragments (one LDS.128 per tile per use), which
//                 takes T up to 16 (k <= 124; k ~ 100 is BASELINE config C4).
template <int T, bool SMEM_L, bool GRAM>
__global__ void __launch_bounds__(TILE_WARPS * 32, SMEM_L ? 1 : 3)
    fused_tile_kernel(const TileArgs a, size_t warp_doubles) {
  extern __shared__ double smem[];
  constexpr int NT = T * (T + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rho = lane >> 2, q = lane & 3;
  const int k = a.k, kp = a.kp, d = a.d, r = a.r;

  // CTA-shared: exp table and the table of the flat element list
  const double tab64 = a.exp_tab[lane];  // this lane's entry of the 2^(j/32) table
  unsigned* etab = (unsigned*)smem;       // n_elem entries + 1 dummy
  const int etab_doubles = (((a.n_elem + 2) / 2) + 1) & ~1;
  double* sscale = smem + etab_doubles;  // gram mode: per-feature multipliers (anisotropic)
  double* wbase = sscale + MGP_MAX_ANISO_DIM + (size_t)warp * warp_doubles;
  double* tiles = wbase;  // NT * 64 doubles (+2 scratch), tile-major, see elem_off()
  const int ds = GRAM ? 0 : d;  // staged coordinates per point
  const int pts_doubles = ((k + 1) * ds + 1) & ~1;
  const int ys_doubles = (k * r + 1) & ~1;
  double* pts_buf = tiles + NT * 64 + 2;       // 2 x (k+1) x d coordinates, row k = query
  double* ys_buf = pts_buf + 2 * pts_doubles;  // 2 x k x r targets

  if (threadIdx.x == 0) etab[a.n_elem] = (unsigned)(NT * 64) << 16;  // dummy -> scratch cell
  if (GRAM && threadIdx.x < MGP_MAX_ANISO_DIM) sscale[threadIdx.x] = a.coord_scale[threadIdx.x];
  for (int e = threadIdx.x; e < a.n_elem; e += blockDim.x) {
    // e < k(k+1)/2: lower triangle in row-major order; then the k cross entries
    const int tri = k * (k + 1) / 2;
    int ti, pi, j;
    if (e < tri) {
      int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      while (i * (i + 1) / 2 > e) --i;
      ti = pi = i;
      j = e - i * (i + 1) / 2;
    } else {
      ti = kp;
      pi = k;
      j = e - tri;
    }
    etab[e] = ((unsigned)elem_off(ti, j) << 16) | ((unsigned)pi << 8) | (unsigned)j;
  }
  __syncthreads();

  const int nwarps = blockDim.x >> 5;
  const long long wglobal = (long long)blockIdx.x * nwarps + warp;
  const long long wstride = (long long)gridDim.x * nwarps;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const int qb = lane & ~3;
  const int zero_from = (k >> 3);  // first tile row that holds padding / augmented rows

  // ---- software pipeline over neighbourhoods: indices 2 ahead, rows 1 ahead (cp.async) ----
  // lane l stages points l and l+32 (point k is the query)
  auto load_src = [&](long long row, int i) -> long long {
    if (row >= a.b || i > k) return -1;
    if (i == k) return a.query_idx ? a.query_idx[row] : row;
    return a.nn_idx[row * k + i];
  };
  auto issue_rows = [&](int buf, int i, long long src) {
    if (src < 0) return;
    const double* px = ((i == k) ? a.query_x : a.train_x) + src * d;
    double* dst = pts_buf + buf * pts_doubles + i * ds;
    for (int f = 0; f < ds; ++f) cp_async8(dst + f, px + f);
    if (i < k && a.train_y) {
      const double* py = a.train_y + src * r;
      double* dy = ys_buf + buf * ys_doubles + i * r;
      for (int c = 0; c < r; ++c) cp_async8(dy + c, py + c);
    }
  };
  // lane l stages points l, l+32, ... (NS = 2 covers k <= 63, NS = 4 covers k <= 127)
  constexpr int NS = (T <= 7) ? 2 : 4;
  long long src[NS];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    src[m] = load_src(wglobal, lane + 32 * m);
    issue_rows(0, lane + 32 * m, src[m]);
  }
  cp_async_commit();
#pragma unroll
  for (int m = 0; m < NS; ++m) src[m] = load_src(wglobal + wstride, lane + 32 * m);
  int buf = 0;

  for (long long row = wglobal; row < a.b; row += wstride, buf ^= 1) {
    // ---- zero the tile rows that keep padding (previous outputs were read already) ----
    {
      double2* z = reinterpret_cast<double2*>(tiles + tile_base(zero_from, 0));
      const int cnt = (NT * 64 - tile_base(zero_from, 0)) / 2;
      for (int e = lane; e < cnt; e += 32) z[e] = make_double2(0.0, 0.0);
      // The strictly upper halves of the diagonal tiles are never assembled; they only ever
      // feed other upper cells, but a stale NaN there would leak into finished entries
      // through 0 * NaN in the masked column updates, so keep them finite.
#pragma unroll
      for (int J = 0; J < T; ++J)
        reinterpret_cast<double2*>(tiles + tile_base(J, J))[lane] = make_double2(0.0, 0.0);
    }
    cp_async_wait_all();
    __syncwarp();
    // rows of the next neighbourhood start flowing in; its successor's indices follow
#pragma unroll
    for (int m = 0; m < NS; ++m) issue_rows(buf ^ 1, lane + 32 * m, src[m]);
    cp_async_commit();
#pragma unroll
    for (int m = 0; m < NS; ++m) src[m] = load_src(row + 2 * wstride, lane + 32 * m);
    double* pts = pts_buf + buf * pts_doubles;
    const double* ys = ys_buf + buf * ys_doubles;
    // fold the length scale(s) into the staged coordinates; scatter -y into rows kp+1+c
    for (int e = lane; e < (k + 1) * ds; e += 32) pts[e] *= a.coord_scale[GRAM ? 0 : e % ds];
    if (a.train_y)
      for (int e = lane; e < k * r; e += 32) {
        const int i = e / r, c = e - i * r;
        tiles[elem_off(kp + 1 + c, i)] = -ys[e];
      }
    // identity padding rows k..kp-1 and Kout = 1 at (kp,kp)
    if (lane <= kp - k) tiles[elem_off(k + lane, k + lane)] = -1.0;
    __syncwarp();

    // ---- covariance assembly -------------------------------------------------
    if (GRAM) {
      GramCtx g;
      g.train_x = a.train_x;
      g.qrow = a.query_x + (a.query_idx ? a.query_idx[row] : row) * d;
      g.nn_row = a.nn_idx + row * k;
      g.scale = a.aniso ? sscale : nullptr;
      g.cs2 = a.aniso ? 1.0 : a.coord_scale[0] * a.coord_scale[0];
      g.k = k;
      g.kp = kp;
      g.d = d;
      if (a.gram == 2) {
        if (a.aniso) gram_distances<true, true>(tiles, g, lane);
        else gram_distances<true, false>(tiles, g, lane);
      } else {
        if (a.aniso) gram_distances<false, true>(tiles, g, lane);
        else gram_distances<false, false>(tiles, g, lane);
      }
      __syncwarp();
      gram_fixup(tiles, etab, g, lane);
      __syncwarp();
    }
    assemble_any(a.formula, tiles, pts, etab, tab64, a.n_elem, lane, GRAM ? -1 : d,
                 a.post_scale, a.kernel_id);
    __syncwarp();
    for (int i = lane; i < k; i += 32)  // nugget on the diagonal (N = -(K + eps))
      tiles[elem_off(i, i)] -= a.noise_bk ? a.noise_bk[row * k + i] : a.noise;
    __syncwarp();

#ifdef MGP_DEBUG_ASM_ONLY
    if (lane == 0 && a.var) a.var[row] = tiles[elem_off(kp, 0)];
    continue;
#endif
    // ---- left-looking tiled LDL^T in registers ---------------------------------
    // Finished tiles hold U = L D (unscaled columns) in ACCUMULATOR layout (lane
    // (rho,q): columns 2q, 2q+1).  Register 0 of every lane read as an A (or B)
    // fragment is the 8x4 slice of the EVEN columns {0,2,4,6}, register 1 the slice
    // of the odd columns; the contraction index of U_I D^-1 U_J^T may be visited in
    // any order, so two DMMAs (even, odd) update a tile with no re-layout at all.
    double l0[SMEM_L ? 1 : T][SMEM_L ? 1 : T], l1[SMEM_L ? 1 : T][SMEM_L ? 1 : T];  // [I][P]
    double dinv0[T], dinv1[T];  // 1/d for this lane's two columns of tile column P
    bool ok = true;
#pragma unroll
    for (int J = 0; J < T; ++J) {
      double c[T][2];
#pragma unroll
      for (int I = J; I < T; ++I) {
        const double2 v =
            *reinterpret_cast<const double2*>(tiles + frag_off(I, J, rho, q));
        c[I][0] = v.x;
        c[I][1] = v.y;
      }
#pragma unroll
      for (int P = 0; P < J; ++P) {
        if (8 * P < kp) {  // tile column P carries eliminated columns
          if (SMEM_L) {
            const double2 lj = *reinterpret_cast<const double2*>(tiles + frag_off(J, P, rho, q));
            const double b0 = lj.x * dinv0[P], b1 = lj.y * dinv1[P];
#pragma unroll
            for (int I = J; I < T; ++I) {
              const double2 li =
                  *reinterpret_cast<const double2*>(tiles + frag_off(I, P, rho, q));
              dmma_acc(c[I][0], c[I][1], li.x, b0);
              dmma_acc(c[I][0], c[I][1], li.y, b1);
            }
          } else {
            const double b0 = l0[J][P] * dinv0[P], b1 = l1[J][P] * dinv1[P];
            // even slices of every tile, then the odd ones: no back-to-back DMMAs on one
            // accumulator (the register variant has the operands at hand, so this is free)
#pragma unroll
            for (int I = J; I < T; ++I) dmma_acc(c[I][0], c[I][1], l0[I][P], b0);
#pragma unroll
            for (int I = J; I < T; ++I) dmma_acc(c[I][0], c[I][1], l1[I][P], b1);
          }
        }
      }
#pragma unroll
      for (int I = J; I < T; ++I) {
        c[I][0] = -c[I][0];
        c[I][1] = -c[I][1];
      }
      const int ncols = min(8, kp - 8 * J);  // eliminated columns in this tile column
      if (ncols > 0) {
        // Column operations are right-multiplications by a matrix M that depends on the
        // diagonal tile only.  Run them on the diagonal tile and on an identity tile V
        // (-> V = M); every tile below the diagonal then becomes S*M with two DMMAs.
        double v0 = (rho == 2 * q) ? 1.0 : 0.0, v1 = (rho == 2 * q + 1) ? 1.0 : 0.0;
        double di0 = 0.0, di1 = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < ncols) {
            const int qj = j >> 1, bj = j & 1;
            // everything below reads the (final, unscaled) column j at once
            const double p = shfl_d(c[J][bj], j * 4 + qj);
            const double uc0 = shfl_d(c[J][bj], (2 * q) * 4 + qj);      // U[2q][j]
            const double uc1 = shfl_d(c[J][bj], (2 * q + 1) * 4 + qj);  // U[2q+1][j]
            const double lr = shfl_d(c[J][bj], qb | qj);                // U[row][j]
            const double vr = shfl_d(bj == 0 ? v0 : v1, qb | qj);       // V[row][j]
            ok = ok && (p > 0.0);
            const double pinv = rcp_fast(p);
            if (bj == 0) di0 = sel_d(q == qj, pinv, di0); else di1 = sel_d(q == qj, pinv, di1);
            // U[c][j] / d_j for this lane's columns c = 2q, 2q+1 (zero for finished columns)
            const double t0 = sel_d(2 * q > j, uc0, 0.0) * pinv;
            const double t1 = sel_d(2 * q + 1 > j, uc1, 0.0) * pinv;
            if (j < 6) {
              c[J][0] = fma(-lr, t0, c[J][0]);
              v0 = fma(-vr, t0, v0);
            }
            if (j < 7) {
              c[J][1] = fma(-lr, t1, c[J][1]);
              v1 = fma(-vr, t1, v1);
            }
          }
        }
        // a half-eliminated tile column only contributes its first four columns later
        const bool keep = (ncols == 8) || (q < 2);
        dinv0[J] = sel_d(keep, di0, 0.0);
        dinv1[J] = sel_d(keep, di1, 0.0);
        if (J + 1 < T) {
          // B fragments of M: even rows {0,2,4,6} and odd rows, lane l = (kk = l&3, n = l>>2)
          const int srcE = 8 * q + (lane >> 3), par = (lane >> 2) & 1;
          const double e0 = shfl_d(v0, srcE), e1 = shfl_d(v1, srcE);
          const double o0 = shfl_d(v0, srcE + 4), o1 = shfl_d(v1, srcE + 4);
          const double bm0 = sel_d(par, e1, e0), bm1 = sel_d(par, o1, o0);
#pragma unroll
          for (int I = J + 1; I < T; ++I) {
            double n0 = 0.0, n1 = 0.0;
            dmma_acc(n0, n1, c[I][0], bm0);
            dmma_acc(n0, n1, c[I][1], bm1);
            c[I][0] = n0;
            c[I][1] = n1;
            if (!SMEM_L) {
              l0[I][J] = n0;
              l1[I][J] = n1;
            }
          }
        }
      }
      // write back what later phases read from shared memory
      if (SMEM_L || a.coeffs || 8 * J + 8 > kp) {
#pragma unroll
        for (int I = J; I < T; ++I)
          *reinterpret_cast<double2*>(tiles + frag_off(I, J, rho, q)) =
              make_double2(c[I][0], c[I][1]);
      }
    }
    __syncwarp();

    // ---- outputs from the Schur complement: lane 0 var/yky/status, lanes 1..r mean ----
    if (lane == 0) {
      if (a.var) a.var[row] = ok ? a.scale * tiles[elem_off(kp, kp)] : nan;
      if (a.status) a.status[row] = ok ? 0 : 1;
      if (a.yky) {
        double s = 0.0;
        for (int c2 = 0; c2 < r; ++c2) s -= tiles[elem_off(kp + 1 + c2, kp + 1 + c2)];
        a.yky[row] = ok ? s : nan;
      }
    } else if (a.mean) {
      for (int c2 = lane - 1; c2 < r; c2 += 31)
        a.mean[row * r + c2] = ok ? -tiles[elem_off(kp + 1 + c2, kp)] : nan;
    }
    if (a.coeffs) {
      // The image holds U = L D (L unit lower) and, in rows kp+1+c, u = L^-1 y.  With
      // K^-1 y = L^-T D^-1 u:  C_i = (u_i - sum_{j>i} U[j][i] C_j) / d_i, column by column.
      __syncwarp();
      for (int i = k - 1; i >= 0; --i) {
        const double inv = 1.0 / tiles[elem_off(i, i)];
        __syncwarp();
        for (int c2 = lane; c2 < r; c2 += 32) tiles[elem_off(kp + 1 + c2, i)] *= inv;
        __syncwarp();
        for (int e = lane; e < r * i; e += 32) {
          const int c2 = e / i, j = e - c2 * i;
          const int ti = kp + 1 + c2;
          double* dst = &tiles[elem_off(ti, j)];
          *dst = fma(-tiles[elem_off(ti, i)], tiles[elem_off(i, j)], *dst);
        }
        __syncwarp();
      }
      for (int e = lane; e < k * r; e += 32) {
        const int j = e / r, c2 = e - j * r;
        a.coeffs[(row * k + j) * r + c2] = ok ? tiles[elem_off(kp + 1 + c2, j)] : nan;
      }
    }
    __syncwarp();
  }
  cp_async_wait_all();
}

}  // namespace

namespace {

// Launch the instantiation for T tile rows; GRAM selects the d > 8 assembly (gram.cuh).
// Persistent grid: warps per CTA (<= TILE_WARPS) and CTAs per SM are chosen to maximise the
// neighbourhoods in flight per SM under the register and shared-memory limits -- e.g. the
// shared-memory-factor variants fit 2 x 4 warps up to T = 9 and 2 x 3 warps at T = 10.
// Only the instantiations with TLO <= T <= THI are compiled into the calling translation unit
// (the large-T ones take tens of seconds each; four units build in parallel).
constexpr int TILE_T_SPLIT = 10;  // fused_tile*.cu: T <= 10, fused_tile*_big.cu: T >= 11

template <bool GRAM, int TLO, int THI>
int launch_tile_instance(const TileArgs& a, int T, long long rows, size_t shared_doubles,
                         size_t warp_doubles, cudaStream_t stream) {
  const size_t smem_max = (size_t)max_smem_optin();
#define MGP_TILE(TT, SL)                                                                      \
  case TT:                                                                                    \
    if constexpr (TT >= TLO && TT <= THI) {                                                   \
    cudaFuncSetAttribute(fused_tile_kernel<TT, SL, GRAM>,                                     \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);         \
    int best_w = 0, best_per_sm = 0;                                                          \
    for (int w = TILE_WARPS; w >= 1; --w) {                                                   \
      const size_t smem_w = (shared_doubles + warp_doubles * w) * sizeof(double);             \
      if (smem_w > smem_max) continue;                                                        \
      int per_sm = 0;                                                                         \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                      \
              &per_sm, fused_tile_kernel<TT, SL, GRAM>, w * 32, smem_w) != cudaSuccess)       \
        per_sm = 0;                                                                           \
      if (w * per_sm > best_w * best_per_sm) {                                                \
        best_w = w;                                                                           \
        best_per_sm = per_sm;                                                                 \
      }                                                                                       \
    }                                                                                         \
    MGP_REQUIRE(best_w > 0, MGP_ERR_UNSUPPORTED,                                              \
                "tile kernel: no launch configuration fits (T=%d, %zu doubles per warp)", T,  \
                warp_doubles);                                                                \
    const size_t smem = (shared_doubles + warp_doubles * best_w) * sizeof(double);            \
    long long blocks = (rows + best_w - 1) / best_w;                                          \
    const long long cap = (long long)sm_count() * best_per_sm;                                \
    if (blocks > cap) blocks = cap;                                                           \
    fused_tile_kernel<TT, SL, GRAM><<<(unsigned)blocks, best_w * 32, smem, stream>>>(         \
        a, warp_doubles);                                                                     \
    break;                                                                                    \
    } else {                                                                                  \
      set_error("tile variant: T=%d is not compiled into this translation unit", T);          \
      return MGP_ERR_UNSUPPORTED;                                                             \
    }
  switch (T) {
    MGP_TILE(1, false)
    MGP_TILE(2, false)
    MGP_TILE(3, false)
    MGP_TILE(4, false)
    MGP_TILE(5, false)
    MGP_TILE(6, false)
    MGP_TILE(7, false)
    MGP_TILE(8, true)
    MGP_TILE(9, true)
    MGP_TILE(10, true)
    MGP_TILE(11, true)
    MGP_TILE(12, true)
    MGP_TILE(13, true)
    MGP_TILE(14, true)
    MGP_TILE(15, true)
    MGP_TILE(16, true)
    default:
      set_error("tile variant does not support %d tile rows", T);
      return MGP_ERR_UNSUPPORTED;
  }
#undef MGP_TILE
  return check_launch("fused_tile_kernel");
}

}  // namespace
}  // namespace mgp
