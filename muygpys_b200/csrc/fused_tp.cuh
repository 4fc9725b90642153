// K1 (thread-per-tile variant of the column-direct kernel) -- the kernel bench.py times.
//
// fused_col_kernel spends 39 % of its instructions in the 52 in-tile LDL^T column steps of a
// neighbourhood: lane-parallel steps (five 64-bit shuffles, a reciprocal, selects, six FP64
// instructions: ~35 instructions per step) that use ~10 % of their FP64 lanes.  Here a CTA (one
// per SM) holds U update warps (one neighbourhood each) and ONE factor warp:
//
//   * update warp: gather, evaluate the covariance entries of a tile column straight into
//     accumulator fragments, DMMA-update them with the finished tile columns, park the updated
//     DIAGONAL tile in 512 bytes of shared memory (hand-off), later pick up M_J = L_JJ^-T
//     (B-fragment order) and -1/d and finish the tiles below the diagonal with two DMMAs each;
//   * factor warp: lane g factorises the diagonal tile of update warp g ENTIRELY IN ITS OWN
//     REGISTERS -- 8x8 LDL^T (84 FMAs, 28 multiplies, 8 reciprocals) and M = L^-T (56 FMAs) on
//     compile-time register indices, no shuffles, no selects: ~280 instructions for U tiles
//     instead of 8 x 35 per tile.  Same operations in the same order as the lane-parallel steps
//     of fused_col_kernel: results are bit-identical.
//
// Hand-offs use two named barriers over the CTA (1: tiles parked, 2: factors ready; bar.arrive
// by the producer, bar.sync by the consumer).  The update warps run a software pipeline over
// the tile columns: while the factor warp holds the diagonal tile of column J they build column
// J + 1 without column J's term; when M_J arrives they finish tile (J+1, J) first, complete the
// next diagonal tile and hand it off before touching the other rows.  The last tile column
// (at most six columns to eliminate, nothing below it) is factorised by the update warp itself,
// and the next neighbourhood is prepared (coordinate scaling, staging of the one after it,
// first tile pair of column 0) in the last two stages, where the update warps would otherwise
// wait for the factor warp.  Finished tiles live in a shared-memory store whose slots are
// REUSED once a tile row is complete (15 instead of 21 slots at T = 7, 48 instead of 78 at
// T = 13).
//
// Same numerics and layout as fused_col_kernel; r == 1, d <= 3, homoscedastic nugget, 7 <= k <=
// 102 (T = 2..13 tile rows); prediction and the one-launch objective (no coefficient output, no
// gradient: those need the whole factor at the end and stay with fused_col_kernel).
// DESIGN.md section 4 has the measurements and the variants that were tried and dropped.
#pragma once

#include "fused_col.cuh"

namespace mgp {
namespace {

// Update warps per CTA (one CTA per SM).  16 warps is what 128 registers per thread allow: four
// warps per scheduler (the register file is split per scheduler, so a 17th warp would cap
// everybody at 96).  Shared memory allows fewer from T = 8 on (tp_uwarps).
#ifndef MGP_TP_MAX_UWARPS
#define MGP_TP_MAX_UWARPS 15
#endif
constexpr int TP_MAX_UWARPS = MGP_TP_MAX_UWARPS;

// ---- shared-memory slots of the finished tiles, reused over the factorisation --------------
// Tile (I,P), I > P, is written at the end of tile column P and last read during the update of
// tile column I.  The tiles of the LAST tile row also park the compactly evaluated raw entries
// from the start of the neighbourhood: they keep fixed slots 0 .. T-2.  Every other tile takes
// the lowest slot whose previous occupant (I', P') has I' <= P.
template <int T>
struct TpSlots {
  int s[T > 0 ? T : 1][T > 0 ? T : 1];
  int count;
};

// REUSE = false (back substitution wanted: the whole factor must survive): one slot per tile.
template <int T, bool REUSE = true>
constexpr TpSlots<T> tp_make_slots() {
  TpSlots<T> m{};
  if (!REUSE) {
    int n = 0;
    for (int P = 0; P + 1 < T; ++P) m.s[T - 1][P] = n++;
    for (int I = 1; I + 1 < T; ++I)
      for (int P = 0; P < I; ++P) m.s[I][P] = n++;
    m.count = n;
    return m;
  }
  int busy_until[T * T + 1] = {};  // slot -> tile row of its occupant (free again when <= P)
  int count = T - 1;
  for (int P = 0; P + 1 < T; ++P) m.s[T - 1][P] = P;
  for (int i = 0; i < T - 1; ++i) busy_until[i] = T;  // never free
  for (int P = 0; P + 2 < T; ++P) {
    for (int I = P + 1; I <= T - 2; ++I) {
      int slot = T - 1;
      while (slot < count && busy_until[slot] > P) ++slot;
      if (slot == count) ++count;
      busy_until[slot] = I;
      m.s[I][P] = slot;
    }
  }
  m.count = count;
  return m;
}

#ifdef MGP_TP_TRACE
__device__ long long g_tp_trace[32][64];
__device__ __forceinline__ long long tp_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define TP_TRACE(w, e) \
  do { if (blockIdx.x == 3 && it == 20 && lane == 0) g_tp_trace[w][e] = tp_clock(); } while (0)
#else
#define TP_TRACE(w, e) do { } while (0)
#endif

__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// per update warp, in doubles: finished tiles | -1/d | diagonal tile in | M out | outputs
// (three Schur entries, ok) | 2 x points | 2 x targets; the stride is 2 (mod 16) doubles so that
// the factor warp's lanes, which read one tile each, start in different banks
// GRAD adds: M_J of every tile column (accumulator order), the last two diagonal tiles after
// their column steps, and the solution vectors w = K^-1 kcross, alpha = K^-1 y.
template <int T, bool GRAD = false>
static inline size_t tp_warp_doubles(int k, int d) {
  constexpr TpSlots<T> SL = tp_make_slots<T, !GRAD>();
  const size_t pts = (size_t)((((k + 1) * d) + 1) & ~1);
  const size_t ys = (size_t)((k + 2) & ~1);
  size_t n = (size_t)SL.count * 64 + 8 * (size_t)T + 64 + 64 + 8 + 2 * pts + 2 * ys +
             (GRAD ? (size_t)(T + 2) * 64 + 16 * (size_t)T : 0);
  while ((n & 15) != 2) n += 2;
  return n;
}

// -1/p: MUFU.RCP64H seed (~20 bits, sign flipped on the integer pipe) + one third-order step;
// bit-identical to -rcp_fast(p)
__device__ __forceinline__ double neg_rcp_fast(double p) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  const double nr = __hiloint2double(__double2hiint(r) ^ 0x80000000, __double2loint(r));
  const double e = fma(-p, r, 1.0);
  const double q1 = nr * e;
  const double w = 1.0 + e;
  return fma(q1, w, nr);
}

// One lane: LDL^T of the first `ncols` columns of an 8x8 symmetric tile (row-major at xin,
// lower triangle significant) and M = L^-T, the product of the column operations.
//   dinv[c]      = -1/d_c (eliminated columns), -0 otherwise
//   xout[8 c + r] = M[r][c] for r < c  (M is unit upper triangular; the diagonal and the zeros
//                   below it are preset once per kernel) -- the B-fragment order of the DMMAs
//   xo[0..2]     = entries (n,n), (n+1,n), (n+1,n+1) of the partially eliminated tile, n = ncols
//   xo[3]        = 1 if every pivot was positive, finite and normal
// FULL: all eight columns are eliminated (every tile column but the last two): no predicates, no
// zero fill, nothing captured (xo[3] only).
// xdiag (not FULL, may be null): the tile after its column steps, row-major lower triangle
// (entries (i,j), j < ncols, are U = L D; the rest is the Schur complement).
template <bool FULL>
__device__ __forceinline__ void tp_factor_tile(const double* __restrict__ xin,
                                               double* __restrict__ xout,
                                               double* __restrict__ dinv,
                                               double* __restrict__ xo, int ncols,
                                               double* __restrict__ xdiag = nullptr) {
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; j += 2) {
      const double2 t = *reinterpret_cast<const double2*>(xin + 8 * i + j);
      a[i][j] = t.x;
      if (j + 1 <= i) a[i][j + 1] = t.y;
    }
  }
  double v[8][8];
  if (!FULL) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[r][c] = 0.0;
  }
  double nd[8];
  double cap0 = 0.0, cap1 = 0.0, cap2 = 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (!FULL) nd[j] = -0.0;
    if (FULL || j < ncols) {
      const double p = a[j][j];
      ok = ok && ((unsigned)(__double2hiint(p) - 1) < 0x7fefffffu);
      const double npinv = neg_rcp_fast(p);
      nd[j] = npinv;
      double nl[8];  // negated multipliers -a[c][j] / p
#pragma unroll
      for (int c = j + 1; c < 8; ++c) nl[c] = a[c][j] * npinv;
#pragma unroll
      for (int c = j + 1; c < 8; ++c) {
#pragma unroll
        for (int i = c; i < 8; ++i) a[i][c] = fma(a[i][j], nl[c], a[i][c]);
      }
#pragma unroll
      for (int c = j + 1; c < 8; ++c) {
#pragma unroll
        for (int r = 0; r < j; ++r) v[r][c] = fma(v[r][j], nl[c], v[r][c]);
        v[j][c] = nl[c];
      }
    } else if (j == ncols) {
      cap0 = a[j][j];
      if (j < 7) {
        cap1 = a[j + 1][j];
        cap2 = a[j + 1][j + 1];
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; c += 2)
    *reinterpret_cast<double2*>(dinv + c) = make_double2(nd[c], nd[c + 1]);
#pragma unroll
  for (int c = 1; c < 8; ++c) {
#pragma unroll
    for (int r = 0; r < c; r += 2) {
      const double lo = v[r][c];
      const double hi = (r + 1 == c) ? 1.0 : ((r + 1 > c) ? 0.0 : v[r + 1][c]);
      *reinterpret_cast<double2*>(xout + 8 * c + r) = make_double2(lo, hi);
    }
  }
  if (!FULL && xdiag != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) xdiag[8 * i + j] = a[i][j];
  }
  if (FULL) {
    xo[3] = ok ? 1.0 : 0.0;
  } else {
    *reinterpret_cast<double2*>(xo) = make_double2(cap0, cap1);
    *reinterpret_cast<double2*>(xo + 2) = make_double2(cap2, ok ? 1.0 : 0.0);
  }
}

template <int T, int F, int D, int TP_UWARPS, bool GRAD = false>
__global__ void __launch_bounds__((TP_UWARPS + 1) * 32, 1)
    fused_tp_kernel(const TileArgs a, const ColLoo loo, int pts_doubles, int ys_doubles,
                    int warp_doubles, long long iters) {
  extern __shared__ double smem[];
  constexpr int TP_THREADS = (TP_UWARPS + 1) * 32;
  constexpr TpSlots<T> SL = tp_make_slots<T, !GRAD>();
  constexpr int W = 8 * (T - 1);
  constexpr int NREC = GRAD ? COL_NREC_GRAD : MGP_PARTIALS;
  __shared__ double s_acc[TP_UWARPS][NREC];
  for (int e = threadIdx.x; e < TP_UWARPS * NREC; e += blockDim.x) (&s_acc[0][0])[e] = 0.0;
  // MGP_TP_FACTOR_FIRST: the factor warp is warp 0 (the oldest warp of the CTA)
#ifdef MGP_TP_FACTOR_FIRST
  const int lane = threadIdx.x & 31, hw_warp = threadIdx.x >> 5;
  const int warp = hw_warp == 0 ? TP_UWARPS : hw_warp - 1;
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#endif
  const int rho = lane >> 2, q = lane & 3;
  const int k = a.k;
  const int nel = k + 1 - W;
  const int kl = k & 7;
  auto warp_base = [&](int u) { return smem + (size_t)u * warp_doubles; };
  constexpr int OFF_DINV = SL.count * 64, OFF_XIN = OFF_DINV + 8 * T, OFF_XOUT = OFF_XIN + 64,
                OFF_XO = OFF_XOUT + 64, OFF_PTS = OFF_XO + 8;

  // M is unit upper triangular: its diagonal and the zeros below it never change
  if (warp < TP_UWARPS) {
    double* xout = warp_base(warp) + OFF_XOUT;
    for (int e = lane; e < 64; e += 32) xout[e] = ((e >> 3) == (e & 7)) ? 1.0 : 0.0;
  }
  __syncthreads();

  if (warp >= TP_UWARPS) {
    // =================================== factor warp ========================================
    // lane g < TP_UWARPS owns the diagonal tile of update warp g: the whole 8x8 LDL^T and
    // M = L^-T run in that lane's registers (no shuffles, no selects)
    double* base = warp_base(lane < TP_UWARPS ? lane : 0);
    for (long long it = 0; it < iters; ++it) {
      // (the last tile column -- at most six columns to eliminate, nothing below it -- stays
      // with the update warps: see last_column)
      for (int J = 0; J + 1 < T; ++J) {
        const int ncols = (J <= T - 3) ? 8 : max(0, min(8, k - 8 * J));
        bar_sync(1, TP_THREADS);  // every update warp has parked its updated diagonal tile
        TP_TRACE(TP_UWARPS, 2 * J);
        if (lane < TP_UWARPS) {
          if (J <= T - 3)
            tp_factor_tile<true>(base + OFF_XIN, base + OFF_XOUT, base + OFF_DINV + 8 * J,
                                 base + OFF_XO, 8);
          else  // (J = T - 2; GRAD keeps the tile for the back substitution: Dg[0])
            tp_factor_tile<false>(base + OFF_XIN, base + OFF_XOUT, base + OFF_DINV + 8 * J,
                                  base + OFF_XO, ncols,
                                  GRAD ? base + OFF_PTS + 2 * pts_doubles + 2 * ys_doubles + T * 64
                                       : nullptr);
        }
        TP_TRACE(TP_UWARPS, 2 * J + 1);
        bar_arrive(2, TP_THREADS);
      }
    }
  } else {
    // =================================== update warp ========================================
    const double tab64 = a.exp_tab[lane];
    double* Ls = warp_base(warp);
    double* dinv_s = Ls + OFF_DINV;
    double* Xin = Ls + OFF_XIN;
    const double* Xout = Ls + OFF_XOUT;
    const double* xo = Ls + OFF_XO;
    double* pts_buf = Ls + OFF_PTS;
    double* ys_buf = pts_buf + 2 * pts_doubles;
    double* Ms = ys_buf + 2 * ys_doubles;  // GRAD: M_J = L_JJ^-T per tile column (accumulator order)
    double* Dg = Ms + T * 64;              // GRAD: diagonal tiles T-2, T-1 after their steps
    double* wv = Dg + 2 * 64;              // GRAD: w = K^-1 kcross (8 T entries)
    double* av = wv + 8 * T;               // GRAD: alpha = K^-1 y
    const long long wglobal = (long long)blockIdx.x * TP_UWARPS + warp;
    const long long wstride = (long long)gridDim.x * TP_UWARPS;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double onep = 1.0 + a.noise;
    const long long last = a.b - 1;

    // rows beyond the batch repeat the last row (the pair and its factor warp stay in step)
    auto load_src = [&](long long row, int i) -> long long {
      if (i > k) return -1;
      const long long rr = row < a.b ? row : last;
      if (i == k) return a.query_idx ? a.query_idx[rr] : rr;
      return a.nn_idx[rr * k + i];
    };
    auto issue_rows = [&](int buf, int i, long long src) {
      if (src < 0) return;
      const double* px = ((i == k) ? a.query_x : a.train_x) + src * D;
      double* dst = pts_buf + buf * pts_doubles + i * D;
      if (D == 2) {
        cp_async16(dst, px);
      } else {
#pragma unroll
        for (int f = 0; f < D; ++f) cp_async8(dst + f, px + f);
      }
      if (i < k) cp_async8(ys_buf + buf * ys_doubles + i, a.train_y + src);
    };

    // lane l stages points l, l + 32, ... (point k is the query); k + 1 <= 8 T - 1
    constexpr int NS = (8 * T - 1 + 31) / 32;
    long long sidx[NS];
    auto load_all = [&](long long row) {
#pragma unroll
      for (int m = 0; m < NS; ++m) sidx[m] = load_src(row, lane + 32 * m);
    };
    auto issue_all = [&](int buf) {
#pragma unroll
      for (int m = 0; m < NS; ++m) issue_rows(buf, lane + 32 * m, sidx[m]);
    };
    auto query_of_staged = [&]() -> long long {
      long long v = sidx[0];
#pragma unroll
      for (int m = 1; m < NS; ++m) v = ((k >> 5) == m) ? sidx[m] : v;
      return __shfl_sync(0xffffffffu, v, k & 31);
    };
    load_all(wglobal);
    long long q_src = query_of_staged();
    issue_all(0);
    cp_async_commit();
    load_all(wglobal + wstride);

    long long it = 0;
    // ---- results of one neighbourhood (called one pipeline stage after its last column) -----
    bool ok = true;
    double out_var = 0.0, out_mean = 0.0, out_yky = 0.0;
    double gdm[4] = {0.0, 0.0, 0.0, 0.0}, gdv[4] = {0.0, 0.0, 0.0, 0.0},
           gdy[4] = {0.0, 0.0, 0.0, 0.0};  // GRAD: d mean, d var, d yky per parameter
    auto write_outputs = [&](long long row, long long qs) {
      if (lane == 0 && row < a.b) {
        if (a.var) a.var[row] = ok ? a.scale * out_var : nan;
        if (a.mean) a.mean[row] = ok ? out_mean : nan;
        if (a.yky) a.yky[row] = ok ? out_yky : nan;
        if (a.status) a.status[row] = ok ? 0 : 1;
        if (loo.warp_rec) {
          double* acc = s_acc[warp];
          if (ok) {
            const double err = out_mean - a.train_y[qs];
            const double e2 = err * err;
            acc[MGP_P_SQERR] += e2;
            acc[MGP_P_COUNT] += 1.0;
            acc[MGP_P_YKY] += out_yky;
            acc[MGP_P_ROWS] += 1.0;
            // looph with a KNOWN scale sigma^2 = a.scale (S/_src/optimize/loss/numpy.py:82-97):
            // 2 b^2 (sqrt(1 + e^2 / (b^2 sigma^2 v)) - 1) goes to AUX, its log(sigma^2 v) term comes
            // from LOGV; the Huber weight 1 / sqrt(1 + u) multiplies the lool numerator and the two
            // gradient sums built from it, so that the host finishes looph like lool.
            double hw = 1.0;
            if (loo.loss_id == MGP_LOSS_LOOPH) {
              const double b2 = loo.boundary_scale * loo.boundary_scale;
              const double root = sqrt(1.0 + e2 / (b2 * a.scale * out_var));
              acc[MGP_P_AUX] += 2.0 * b2 * (root - 1.0);
              hw = 1.0 / root;
            }
            acc[MGP_P_SQERR_V] += hw * e2 / out_var;
            acc[MGP_P_LOGV] += log(out_var);
            if (loo.loss_id == MGP_LOSS_PSEUDO_HUBER) {
              const double z = err / loo.boundary_scale;
              acc[MGP_P_AUX] +=
                  loo.boundary_scale * loo.boundary_scale * (sqrt(fma(z, z, 1.0)) - 1.0);
            }
            if (GRAD) {
              const double iv = 1.0 / out_var;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                if (t < D || t == 3) {
                  double* gr = acc + MGP_PARTIALS + 5 * t;
                  gr[0] += 2.0 * err * gdm[t];        // d sum e^2
                  gr[1] += hw * 2.0 * err * gdm[t] * iv;   // sum w 2 e dm / v
                  gr[2] += hw * e2 * gdv[t] * iv * iv;     // sum w e^2 dv / v^2
                  gr[3] += gdv[t] * iv;               // sum dv / v
                  gr[4] += gdy[t];                    // d sum yky
                }
              }
            }
          } else {
            acc[MGP_P_BAD] += 1.0;
          }
        }
      }
    };
    // pick up what the factor warp left for tile column J (after bar_sync(2))
    auto read_factor_outputs = [&](int J) {
      ok = ok && (xo[3] != 0.0);
      // xo[0]: entry (n,n) of the tile after its n = ncols steps
      if (J == T - 2 && kl == 7) out_var = xo[0];
    };

    // The tile column under construction: rows J .. T-1 as accumulator fragments.
    double c[T][2];
    auto slot_off = [&](int I, int P) -> int { return SL.s[I][P] * 64; };

    // Build tile column J: evaluate its entries and subtract the products with the finished tile
    // columns P < NP.  (NP = J - 1 in the pipeline: column J - 1 is still with the factor warp;
    // its term is added by finish_column.)
    auto build_column = [&](int J, int NP, const double* pts, const double* ys, int Ilo = 0,
                            int Ihi = T) {
      const int j0 = 8 * J + 2 * q, j1 = j0 + 1;
      const Pt<D> pc0 = ld_pt<D>(pts, (J == T - 1) ? min(j0, k) : j0);
      const Pt<D> pc1 = ld_pt<D>(pts, (J == T - 1) ? min(j1, k) : j1);
      double b0[T], b1[T];
      double2 ljs[T];
#pragma unroll
      for (int P = 0; P < NP; ++P) {
        ljs[P] = *reinterpret_cast<const double2*>(Ls + slot_off(J, P) + 2 * lane);
        const double2 nd = *reinterpret_cast<const double2*>(dinv_s + 8 * P + 2 * q);
        b0[P] = ljs[P].x * nd.x;
        b1[P] = ljs[P].y * nd.y;
      }
      auto fix_diag = [&](int I) {
        const double dg = (8 * I + rho < k) ? onep : 1.0;
        c[I][0] = (rho == 2 * q) ? dg : c[I][0];
        c[I][1] = (rho == 2 * q + 1) ? dg : c[I][1];
      };
      auto eval_two = [&](int Ia, int Ib) {
        const Pt<D> pa = ld_pt<D>(pts, 8 * Ia + rho), pb = ld_pt<D>(pts, 8 * Ib + rho);
        const double u[4] = {sq_dist<D>(pa, pc0), sq_dist<D>(pa, pc1), sq_dist<D>(pb, pc0),
                             sq_dist<D>(pb, pc1)};
        double o[4];
        cov_n<F, 4>(u, tab64, o);
        c[Ia][0] = o[0];
        c[Ia][1] = o[1];
        c[Ib][0] = o[2];
        c[Ib][1] = o[3];
        if (Ia == J) fix_diag(Ia);
      };
      auto eval_one = [&](int Ia) {
        const Pt<D> pa = ld_pt<D>(pts, 8 * Ia + rho);
        const double u[2] = {sq_dist<D>(pa, pc0), sq_dist<D>(pa, pc1)};
        double o[2];
        cov_n<F, 2>(u, tab64, o);
        c[Ia][0] = o[0];
        c[Ia][1] = o[1];
        if (Ia == J) fix_diag(Ia);
      };
      // last tile row left of its diagonal tile: parked values, target row, zero padding
      auto load_last = [&]() {
        const double2 v = *reinterpret_cast<const double2*>(Ls + slot_off(T - 1, J) + 2 * lane);
        const double2 yv = *reinterpret_cast<const double2*>(ys + j0);
        const bool isy = rho == nel;
        const double y0 = (isy && j0 < k) ? yv.x : 0.0, y1 = (isy && j1 < k) ? yv.y : 0.0;
        c[T - 1][0] = (rho < nel) ? v.x : y0;
        c[T - 1][1] = (rho < nel) ? v.y : y1;
      };
      // last diagonal tile: every kind of entry, masked
      auto eval_corner = [&]() {
        const int i = W + rho;
        const Pt<D> pr = ld_pt<D>(pts, min(i, k));
        const double u[2] = {sq_dist<D>(pr, pc0), sq_dist<D>(pr, pc1)};
        double o[2];
        cov_n<F, 2>(u, tab64, o);
        const double y0 = ys[min(j0, k)], y1 = ys[min(j1, k)];
        const double dg = (i < k) ? onep : 1.0;
        const bool krow = i <= k, yrow = i == k + 1;
        double r0 = (krow && j0 < k) ? o[0] : ((yrow && j0 < k) ? y0 : 0.0);
        double r1 = (krow && j1 < k) ? o[1] : ((yrow && j1 < k) ? y1 : 0.0);
        r0 = (krow && i == j0) ? dg : r0;
        r1 = (krow && i == j1) ? dg : r1;
        c[T - 1][0] = r0;
        c[T - 1][1] = r1;
      };
      auto frag = [&](int I, int P) -> double2 {
        return (I == J) ? ljs[P]
                        : *reinterpret_cast<const double2*>(Ls + slot_off(I, P) + 2 * lane);
      };
      auto update_two = [&](int Ia, int Ib) {
#pragma unroll
        for (int P = 0; P < NP; ++P) {
          const double2 la = frag(Ia, P), lb = frag(Ib, P);
          dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
          dmma_free(c[Ib][0], c[Ib][1], lb.x, b0[P]);
          dmma_free(c[Ia][0], c[Ia][1], la.y, b1[P]);
          dmma_free(c[Ib][0], c[Ib][1], lb.y, b1[P]);
        }
      };
      auto update_one = [&](int Ia) {
        if (NP >= 2) {  // two partial sums: even / odd slices
          double x0 = 0.0, x1 = 0.0;
#pragma unroll
          for (int P = 0; P < NP; ++P) {
            const double2 la = frag(Ia, P);
            dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
            dmma_free(x0, x1, la.y, b1[P]);
          }
          c[Ia][0] += x0;
          c[Ia][1] += x1;
        } else {
#pragma unroll
          for (int P = 0; P < NP; ++P) {
            const double2 la = frag(Ia, P);
            dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
            dmma_free(c[Ia][0], c[Ia][1], la.y, b1[P]);
          }
        }
      };
      // tiles in pairs: (J, J+1), (J+2, J+3), ...; the last tile row comes from load_last
#pragma unroll
      for (int Ia = J; Ia < T; Ia += 2) {
        const int Ib = Ia + 1;
        if (Ia < Ilo || Ia >= Ihi) continue;
        if (J == T - 1) {
          eval_corner();
          update_one(T - 1);
        } else if (Ib <= T - 2) {
          eval_two(Ia, Ib);
          update_two(Ia, Ib);
        } else if (Ia <= T - 2) {  // Ib == T-1
          eval_one(Ia);
          load_last();
          update_two(Ia, T - 1);
        } else {  // Ia == T-1
          load_last();
          update_one(T - 1);
        }
      }
    };

    // Column J has been built: hand its diagonal tile to the factor warp ...
    auto hand_off = [&](int J) {
      *reinterpret_cast<double2*>(Xin + 2 * lane) = make_double2(c[J][0], c[J][1]);
      bar_arrive(1, TP_THREADS);
    };
    // ... and park the tiles below it (raw) in the slots their finished versions will take.
    auto park_below = [&](int J) {
#pragma unroll
      for (int I = J + 1; I < T; ++I)
        *reinterpret_cast<double2*>(Ls + slot_off(I, J) + 2 * lane) =
            make_double2(c[I][0], c[I][1]);
    };

    // The LAST tile column: its diagonal tile holds the last k - 8 (T - 1) <= 6 columns of K, the
    // cross-covariance row and the target row, and nothing lies below it.  A round trip through
    // the factor warp (hand-off, ~1 100 cycles of latency, pick-up) would leave this warp idle,
    // so the few column steps run here, lane-parallel as in fused_col_kernel, and the outputs
    // come straight out of the Schur complement.
    auto last_column = [&]() {
      constexpr int J = T - 1;
      const int ncols = max(0, min(8, k - 8 * J));
      const int qb = lane & ~3;
      double v0 = (rho == 2 * q) ? 1.0 : 0.0, v1 = (rho == 2 * q + 1) ? 1.0 : 0.0;  // GRAD: M
      double di0 = 0.0, di1 = 0.0;                                                 // GRAD: 1/d
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < ncols) {
          const int qj = j >> 1, bj = j & 1;
          const double cj = bj == 0 ? c[J][0] : c[J][1];
          const double p = shfl_d(cj, j * 4 + qj);
          const double uc0 = shfl_d(cj, (2 * q) * 4 + qj);      // U[2q][j]
          const double uc1 = shfl_d(cj, (2 * q + 1) * 4 + qj);  // U[2q+1][j]
          const double lr = shfl_d(cj, qb | qj);                // U[row][j]
          ok = ok && ((unsigned)(__double2hiint(p) - 1) < 0x7fefffffu);
          const double pinv = rcp_fast(p);
          const double t0 = sel_d(2 * q > j, uc0, 0.0) * pinv;
          const double t1 = sel_d(2 * q + 1 > j, uc1, 0.0) * pinv;
          c[J][0] = fma(-lr, t0, c[J][0]);
          c[J][1] = fma(-lr, t1, c[J][1]);
          if (GRAD) {
            const double vr = shfl_d(bj == 0 ? v0 : v1, qb | qj);  // V[row][j]
            v0 = fma(-vr, t0, v0);
            v1 = fma(-vr, t1, v1);
            if (bj == 0) di0 = sel_d(q == qj, pinv, di0); else di1 = sel_d(q == qj, pinv, di1);
          }
        }
      }
      if (GRAD) {
        if (rho == 0) *reinterpret_cast<double2*>(dinv_s + 8 * J + 2 * q) = make_double2(-di0, -di1);
        *reinterpret_cast<double2*>(Ms + J * 64 + 2 * lane) = make_double2(v0, v1);
        *reinterpret_cast<double2*>(Dg + 64 + 2 * lane) = make_double2(c[J][0], c[J][1]);
      }
      if (kl < 7) {
        const double cv = (kl & 1) ? c[J][1] : c[J][0];
        const double cy = ((kl + 1) & 1) ? c[J][1] : c[J][0];
        out_var = shfl_d(cv, kl * 4 + (kl >> 1));
        out_mean = -shfl_d(cv, (kl + 1) * 4 + (kl >> 1));
        out_yky = -shfl_d(cy, (kl + 1) * 4 + ((kl + 1) >> 1));
      } else {
        out_yky = -shfl_d(c[J][0], 0);
      }
    };

    // M_J and 1/d_J are in place: finish the tiles below the diagonal (U = S M, two DMMAs per
    // tile) and subtract column J's term from column J + 1, which sits in c[][].  The factor
    // warp idles until it gets the next diagonal tile, so tile row J + 1 goes first and is
    // handed off before the other rows are touched.
    auto finish_column = [&](int J) {
      if (J + 1 >= T) return;
      const double2 bm = *reinterpret_cast<const double2*>(Xout + 2 * lane);
      const double2 nd = *reinterpret_cast<const double2*>(dinv_s + 8 * J + 2 * q);
      if (GRAD)  // Xout is column-major (B-fragment order): M[rho][2q], M[rho][2q+1]
        *reinterpret_cast<double2*>(Ms + J * 64 + 2 * lane) =
            make_double2(Xout[16 * q + rho], Xout[16 * q + 8 + rho]);
      // tile row J + 1: finished tile, then the next diagonal tile
      double m0 = 0.0, m1 = 0.0;
      {
        const double2 raw = *reinterpret_cast<const double2*>(Ls + slot_off(J + 1, J) + 2 * lane);
        dmma_free(m0, m1, raw.x, bm.x);
        dmma_free(m0, m1, raw.y, bm.y);
      }
      const double bj0 = m0 * nd.x, bj1 = m1 * nd.y;
      dmma_free(c[J + 1][0], c[J + 1][1], m0, bj0);
      dmma_free(c[J + 1][0], c[J + 1][1], m1, bj1);
      if (J + 1 < T - 1) hand_off(J + 1);
      TP_TRACE(warp, 32 + J);
      *reinterpret_cast<double2*>(Ls + slot_off(J + 1, J) + 2 * lane) = make_double2(m0, m1);
      if (J == T - 2 && kl == 7) out_mean = -shfl_d(m1, 3);
      // the other rows, two at a time (few live registers: the raw tiles of column J + 1 wait in
      // c[][] meanwhile): U = S M, stored; column J's term subtracted from column J + 1; that
      // tile parked raw in the slot its finished version will take
#pragma unroll
      for (int Ia = J + 2; Ia < T; Ia += 2) {
        const int Ib = (Ia + 1 < T) ? Ia + 1 : Ia;
        const double2 ra = *reinterpret_cast<const double2*>(Ls + slot_off(Ia, J) + 2 * lane);
        const double2 rb = *reinterpret_cast<const double2*>(Ls + slot_off(Ib, J) + 2 * lane);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        dmma_free(a0, a1, ra.x, bm.x);
        if (Ib != Ia) dmma_free(b0, b1, rb.x, bm.x);
        dmma_free(a0, a1, ra.y, bm.y);
        if (Ib != Ia) dmma_free(b0, b1, rb.y, bm.y);
        *reinterpret_cast<double2*>(Ls + slot_off(Ia, J) + 2 * lane) = make_double2(a0, a1);
        if (Ib != Ia)
          *reinterpret_cast<double2*>(Ls + slot_off(Ib, J) + 2 * lane) = make_double2(b0, b1);
        dmma_free(c[Ia][0], c[Ia][1], a0, bj0);
        if (Ib != Ia) dmma_free(c[Ib][0], c[Ib][1], b0, bj0);
        dmma_free(c[Ia][0], c[Ia][1], a1, bj1);
        if (Ib != Ia) dmma_free(c[Ib][0], c[Ib][1], b1, bj1);
        *reinterpret_cast<double2*>(Ls + slot_off(Ia, J + 1) + 2 * lane) =
            make_double2(c[Ia][0], c[Ia][1]);
        if (Ib != Ia)
          *reinterpret_cast<double2*>(Ls + slot_off(Ib, J + 1) + 2 * lane) =
              make_double2(c[Ib][0], c[Ib][1]);
      }
    };

    // ---- preparation of a neighbourhood, in pieces that need nothing from the factor warp.
    // For the NEXT neighbourhood (1) and (2) run where this warp would otherwise wait for the
    // factor warp: the last two tile columns (the update work per column shrinks towards the
    // end of the factorisation, the factor warp's latency does not).  (A second slot set for
    // the last tile row, so that (3) can move there too, was measured: no gain.)
    // (1) its staging is complete: fold the length scales (and the Matern sqrt(2 nu)) into the
    //     staged coordinates
    auto prep_scale = [&](int nbuf) {
      cp_async_wait_all();
      __syncwarp();
      double* pts = pts_buf + nbuf * pts_doubles;
      if (D == 2) {
        double2* p2 = reinterpret_cast<double2*>(pts);
        for (int i = lane; i <= k; i += 32) {
          double2 v = p2[i];
          v.x *= a.coord_scale[0];
          v.y *= a.coord_scale[1];
          p2[i] = v;
        }
      } else {
        for (int e = lane; e < (k + 1) * D; e += 32) pts[e] *= a.coord_scale[e % D];
      }
      __syncwarp();
    };
    // (2) the other staging buffer is free (no evaluation of the current neighbourhood is left):
    //     start the staging of the neighbourhood after `nx`; first tile pair of column 0
    long long q_next = 0;
    constexpr int FIRST = (T >= 3) ? 2 : 0;  // T >= 3: the first tile pair of column 0 is regular
    auto prep_stage = [&](long long nx, int nbuf) {
      const long long row = wglobal + nx * wstride;
      __syncwarp();  // every lane has read what it needed from the buffer about to be refilled
      q_next = query_of_staged();  // query of row nx + 1
      issue_all(nbuf ^ 1);
      cp_async_commit();
      load_all(row + 2 * wstride);
    };
    auto prep_first = [&](int nbuf) {
      if (FIRST)
        build_column(0, 0, pts_buf + nbuf * pts_doubles, ys_buf + nbuf * ys_doubles, 0, FIRST);
    };
    // (3) compact evaluation of the real rows of the last tile row (columns < W) into the slots
    //     of that row; they are free once the last column of the previous neighbourhood is built
    auto prep_compact = [&](int nbuf) {
      if (T > 1) {
        // T = 2: nothing else separates the previous neighbourhood's last finished-tile store
        // (finish_column, slot 0) from these stores to the same slot by other lanes (racecheck)
        if (T == 2) __syncwarp();
        // lane l evaluates columns l, l + 32, ... of three rows at a time: six entries in
        // flight, no index arithmetic, the column points loaded once (a flat 32-entries-per-pass
        // list took 2 200 cycles for 144 entries: three dependent passes, a division per entry)
        const double* pts = pts_buf + nbuf * pts_doubles;
        constexpr int NG = (W + 31) / 32;
        for (int ar0 = 0; ar0 < nel; ar0 += 3) {
          const Pt<D> pr0 = ld_pt<D>(pts, min(W + ar0, k)), pr1 = ld_pt<D>(pts, min(W + ar0 + 1, k)),
                      pr2 = ld_pt<D>(pts, min(W + ar0 + 2, k));
#pragma unroll
          for (int g = 0; g < NG; g += 2) {
            const int j0 = 32 * g + lane, j1 = j0 + 32;
            const bool v0 = j0 < W, v1 = (g + 1 < NG) && j1 < W;
            const Pt<D> pc0 = ld_pt<D>(pts, v0 ? j0 : 0), pc1 = ld_pt<D>(pts, v1 ? j1 : 0);
            const double u[6] = {sq_dist<D>(pr0, pc0), sq_dist<D>(pr1, pc0), sq_dist<D>(pr2, pc0),
                                 sq_dist<D>(pr0, pc1), sq_dist<D>(pr1, pc1), sq_dist<D>(pr2, pc1)};
            double o[6];
            cov_n<F, 6>(u, tab64, o);
            double* d0 = Ls + (j0 >> 3) * 64 + (j0 & 7);  // slot of (T-1, j/8) is j/8
            double* d1 = Ls + (j1 >> 3) * 64 + (j1 & 7);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (v0 && ar0 + i < nel) d0[(ar0 + i) * 8] = o[i];
              if (v1 && ar0 + i < nel) d1[(ar0 + i) * 8] = o[3 + i];
            }
          }
        }
        __syncwarp();
      }
    };

    // ---- back substitution and analytic gradient (GRAD; same algebra as fused_col_kernel) ------
    // w = K^-1 kcross and alpha = K^-1 y from a block back substitution on the stored factor;
    // alpha is the fast-mean coefficient row.  Gradient of (mean, variance, y^T K^-1 y) w.r.t.
    // length scale(s) and nugget: d mean = dc^T alpha - w^T dK alpha, d var = -2 dc^T w +
    // w^T dK w, d yky = -alpha^T dK alpha, with dK_ij / dl_f = phi(s_ij) z_f^2 / l_f re-evaluated
    // entry by entry (never stored).
    auto grad_epilogue = [&](const double* pts, long long row) {
      const int Jlast = (k - 1) >> 3;            // last tile column that holds K columns
      const int Rk = k >> 3, Ry = (k + 1) >> 3;  // tile rows of the cross row and the target row
      const int ly = (k + 1) & 7;
      __syncwarp();
      for (int e = lane; e < 8 * T; e += 32) {
        wv[e] = 0.0;
        av[e] = 0.0;
      }
      __syncwarp();
      double xw[T], xa[T];  // this lane's row (rho) of the solution blocks
#pragma unroll
      for (int I = 0; I < T; ++I) xw[I] = xa[I] = 0.0;
      // tile (row tile R, column tile J2) as stored: finished tile, or final diagonal tile
      auto stored = [&](int R, int J2, int off) -> double2 {
        const double* base = (J2 < R) ? Ls + slot_off(R, J2) : Dg + (R - (T - 2)) * 64;
        return *reinterpret_cast<const double2*>(base + off);
      };
#pragma unroll
      for (int J = T - 1; J >= 0; --J) {
        if (J <= Jlast) {
          const int nc = min(8, k - 8 * J);  // eliminated columns of this tile column
          double ac0 = 0.0, ac1 = 0.0, ay0 = 0.0, ay1 = 0.0;
#pragma unroll
          for (int I = J + 1; I < T; ++I) {
            if (I <= Jlast) {
              const double2 u = *reinterpret_cast<const double2*>(Ls + slot_off(I, J) + 2 * lane);
              ac0 = fma(u.x, xw[I], ac0);
              ac1 = fma(u.y, xw[I], ac1);
              ay0 = fma(u.x, xa[I], ay0);
              ay1 = fma(u.y, xa[I], ay1);
            }
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {  // sum over the 8 rows (lanes with the same q)
            ac0 += __shfl_xor_sync(0xffffffffu, ac0, o);
            ac1 += __shfl_xor_sync(0xffffffffu, ac1, o);
            ay0 += __shfl_xor_sync(0xffffffffu, ay0, o);
            ay1 += __shfl_xor_sync(0xffffffffu, ay1, o);
          }
          // (Rk, Ry >= T - 2: rows k, k + 1 of U sit in the last two tile rows)
          double2 zc = make_double2(0.0, 0.0), zy = make_double2(0.0, 0.0);
#pragma unroll
          for (int R = (T >= 2 ? T - 2 : 0); R < T; ++R) {
            if (R == Rk && J <= R) zc = stored(R, J, kl * 8 + 2 * q);
            if (R == Ry && J <= R) zy = stored(R, J, ly * 8 + 2 * q);
          }
          const double2 nd = *reinterpret_cast<const double2*>(dinv_s + 8 * J + 2 * q);
          const bool in0 = 2 * q < nc, in1 = 2 * q + 1 < nc;
          const double sc0 = in0 ? (ac0 - zc.x) * nd.x : 0.0, sc1 = in1 ? (ac1 - zc.y) * nd.y : 0.0;
          const double sy0 = in0 ? (ay0 - zy.x) * nd.x : 0.0, sy1 = in1 ? (ay1 - zy.y) * nd.y : 0.0;
          // x_J = L_JJ^-T D^-1 t = M_J (D^-1 t): row rho, summed over the quad
          const double2 m = *reinterpret_cast<const double2*>(Ms + J * 64 + 2 * lane);
          double vw = fma(m.x, sc0, m.y * sc1), va = fma(m.x, sy0, m.y * sy1);
          vw += __shfl_xor_sync(0xffffffffu, vw, 1);
          va += __shfl_xor_sync(0xffffffffu, va, 1);
          vw += __shfl_xor_sync(0xffffffffu, vw, 2);
          va += __shfl_xor_sync(0xffffffffu, va, 2);
          xw[J] = (rho < nc) ? vw : 0.0;
          xa[J] = (rho < nc) ? va : 0.0;
          if (q == 0) {
            wv[8 * J + rho] = xw[J];
            av[8 * J + rho] = xa[J];
          }
        }
      }
      __syncwarp();
      // fast-mean precompute mode: alpha = (K + eps)^-1 y IS the coefficient row
      // (S/_src/gp/muygps/numpy.py:88-95)
      if (a.coeffs && row < a.b)  // (rows past the end of the batch repeat the last row)
        for (int e = lane; e < k; e += 32) a.coeffs[row * k + e] = ok ? av[e] : nan;
      if (loo.grad == nullptr) return;
      // weighted re-evaluation: per tile row I the row-factored sums
      //   Ra_f = sum_j phi z_f^2 alpha_j,  Rw_f = sum_j phi z_f^2 w_j   (j over tile columns <= I)
      double Gm[D], Gv[D], Gy[D];
#pragma unroll
      for (int f = 0; f < D; ++f) Gm[f] = Gv[f] = Gy[f] = 0.0;
#pragma unroll
      for (int I = 0; I < T; ++I) {
        if (I <= Jlast) {
          const Pt<D> pr = ld_pt<D>(pts, min(8 * I + rho, k));
          double Ra[D], Rw[D];
#pragma unroll
          for (int f = 0; f < D; ++f) Ra[f] = Rw[f] = 0.0;
          auto tile_pair = [&](int Ja, int Jb, bool two) {
            const Pt<D> a0 = ld_pt<D>(pts, min(8 * Ja + 2 * q, k)),
                        a1 = ld_pt<D>(pts, min(8 * Ja + 2 * q + 1, k));
            const Pt<D> b0 = ld_pt<D>(pts, min(8 * Jb + 2 * q, k)),
                        b1 = ld_pt<D>(pts, min(8 * Jb + 2 * q + 1, k));
            const double2 wa = *reinterpret_cast<const double2*>(wv + 8 * Ja + 2 * q);
            const double2 aa = *reinterpret_cast<const double2*>(av + 8 * Ja + 2 * q);
            const double2 wb = *reinterpret_cast<const double2*>(wv + 8 * Jb + 2 * q);
            const double2 ab = *reinterpret_cast<const double2*>(av + 8 * Jb + 2 * q);
            const Pt<D>* pc[4] = {&a0, &a1, &b0, &b1};
            const double wj[4] = {wa.x, wa.y, wb.x, wb.y}, aj[4] = {aa.x, aa.y, ab.x, ab.y};
            const double half[4] = {Ja == I ? 0.5 : 1.0, Ja == I ? 0.5 : 1.0,
                                    Jb == I ? 0.5 : 1.0, Jb == I ? 0.5 : 1.0};
            double u[4], ph[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) u[e] = sq_dist<D>(pr, *pc[e]);
            cov_n<F, 4, 1>(u, tab64, ph);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e < 2 || two) {
                const double pe = ph[e] * half[e];
#pragma unroll
                for (int f = 0; f < D; ++f) {
                  const double df = pr.x[f] - pc[e]->x[f];
                  const double g = pe * (df * df);
                  Ra[f] = fma(g, aj[e], Ra[f]);
                  Rw[f] = fma(g, wj[e], Rw[f]);
                }
              }
            }
          };
#pragma unroll
          for (int Ja = 0; Ja <= I; Ja += 2) tile_pair(Ja, (Ja + 1 <= I) ? Ja + 1 : Ja, Ja + 1 <= I);
#pragma unroll
          for (int f = 0; f < D; ++f) {
            Gm[f] = fma(xw[I], Ra[f], fma(xa[I], Rw[f], Gm[f]));
            Gv[f] = fma(2.0 * xw[I], Rw[f], Gv[f]);
            Gy[f] = fma(2.0 * xa[I], Ra[f], Gy[f]);
          }
        }
      }
      // cross-covariance row: dc_j / dl_f = phi(s_qj) z_f^2 / l_f
      double Gmc[D], Gvc[D];
#pragma unroll
      for (int f = 0; f < D; ++f) Gmc[f] = Gvc[f] = 0.0;
      {
        const Pt<D> pq = ld_pt<D>(pts, k);
#pragma unroll
        for (int c0 = 0; c0 < 8 * T; c0 += 64) {  // 64 entries of the row per pass
          const int ia = c0 + lane, ib = c0 + lane + 32;
          const int ja = min(ia, k), jb = min(ib, k);
          const Pt<D> p0 = ld_pt<D>(pts, ja), p1 = ld_pt<D>(pts, jb);
          const double u[2] = {sq_dist<D>(pq, p0), sq_dist<D>(pq, p1)};
          double ph[2];
          cov_n<F, 2, 1>(u, tab64, ph);
          const double w0 = ia < k ? wv[ja] : 0.0, a0 = ia < k ? av[ja] : 0.0;
          const double w1 = ib < k ? wv[jb] : 0.0, a1 = ib < k ? av[jb] : 0.0;
#pragma unroll
          for (int f = 0; f < D; ++f) {
            const double d0 = pq.x[f] - p0.x[f], d1 = pq.x[f] - p1.x[f];
            const double g0 = ph[0] * (d0 * d0), g1 = ph[1] * (d1 * d1);
            Gmc[f] = fma(g0, a0, fma(g1, a1, Gmc[f]));
            Gvc[f] = fma(g0, w0, fma(g1, w1, Gvc[f]));
          }
        }
      }
      // nugget: dK = I
      double dwa = 0.0, dww = 0.0, daa = 0.0;
      for (int e = lane; e < 8 * T; e += 32) {
        dwa = fma(wv[e], av[e], dwa);
        dww = fma(wv[e], wv[e], dww);
        daa = fma(av[e], av[e], daa);
      }
#pragma unroll
      for (int f = 0; f < D; ++f) {
        const double rm = warp_sum(Gm[f]), rv = warp_sum(Gv[f]), ry = warp_sum(Gy[f]);
        const double rmc = warp_sum(Gmc[f]), rvc = warp_sum(Gvc[f]);
        gdm[f] = (rmc - rm) * loo.inv_len[f];
        gdv[f] = (rv - 2.0 * rvc) * loo.inv_len[f];
        gdy[f] = -ry * loo.inv_len[f];
      }
      gdm[3] = -warp_sum(dwa);
      gdv[3] = warp_sum(dww);
      gdy[3] = -warp_sum(daa);
    };

    // Software pipeline, one stage per tile column: while the factor warp works on the
    // diagonal tile of column J, this warp builds column J + 1.
    int buf = 0;
    if (iters > 0) {
      prep_scale(0);
      prep_stage(0, 0);
      prep_first(0);
    }
    constexpr int J_SCALE = (T >= 3) ? T - 3 : 0;  // stage that hosts piece (1) of the next one
    for (it = 0; it < iters; ++it, buf ^= 1) {
      const long long row = wglobal + it * wstride;
      const bool more = it + 1 < iters;
      TP_TRACE(warp, 3);
      const double* pts = pts_buf + buf * pts_doubles;
      const double* ys = ys_buf + buf * ys_doubles;
      const long long q_after = q_next;  // query of row it + 1
      ok = true;
      if (!FIRST) {
        prep_compact(buf);
        build_column(0, 0, pts, ys, 0, T);
      }
      hand_off(0);
      TP_TRACE(warp, 1);
      if (FIRST) prep_compact(buf);
      if (FIRST) build_column(0, 0, pts, ys, FIRST, T);
      TP_TRACE(warp, 0);
      park_below(0);
#pragma unroll
      for (int J = 0; J + 1 < T; ++J) {
        build_column(J + 1, J, pts, ys);
        if (J == J_SCALE && more) prep_scale(buf ^ 1);
        // column T - 1 is built: c[0 .. T-2] and -- unless the gradient pass re-evaluates the
        // entries -- this neighbourhood's staging buffer are free
        if (J == T - 2 && more) {
          if (!GRAD) prep_stage(it + 1, buf ^ 1);
          prep_first(buf ^ 1);
        }
        TP_TRACE(warp, 4 * J + 4);
        bar_sync(2, TP_THREADS);
        TP_TRACE(warp, 4 * J + 5);
        read_factor_outputs(J);
        finish_column(J);
        TP_TRACE(warp, 4 * J + 6);
      }
      if (T == 1) c[0][0] = c[0][1] = 0.0;  // (not instantiated)
      last_column();
      if (GRAD) {
        grad_epilogue(pts, row);
        if (more) prep_stage(it + 1, buf ^ 1);
      }
      write_outputs(row, q_src);
      q_src = q_after;
    }
    cp_async_wait_all();
  }

  if (loo.warp_rec) {
    // fixed-order reduction, as in fused_col_kernel: update warps of a block -> block record ->
    // (last block) strided partial sums -> sequential sum; then the cross-GPU exchange
    __shared__ unsigned int s_last;
    constexpr int NGRP = (TP_THREADS / NREC) < 16 ? (TP_THREADS / NREC) : 16;
    __shared__ double s_red[NGRP][NREC];
    __syncthreads();
    if (threadIdx.x < NREC) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < TP_UWARPS; ++w) v += s_acc[w][threadIdx.x];
      loo.warp_rec[(size_t)blockIdx.x * NREC + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(loo.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (threadIdx.x < NGRP * NREC) {
        const int slot = threadIdx.x % NREC, grp = threadIdx.x / NREC;
        double sum = 0.0;
        for (int cidx = grp; cidx < (int)gridDim.x; cidx += NGRP)
          sum += __ldcg(loo.warp_rec + (size_t)cidx * NREC + slot);
        s_red[grp][slot] = sum;
      }
      __syncthreads();
      __shared__ double s_tot[MGP_PARTIALS];
      if (threadIdx.x < NREC) {
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < NGRP; ++g) tot += s_red[g][threadIdx.x];
        if (threadIdx.x < MGP_PARTIALS) {
          s_tot[threadIdx.x] = tot;
          if (loo.peers.world <= 1) loo.partials[threadIdx.x] = tot;
        } else if (GRAD && loo.grad != nullptr &&
                   threadIdx.x < MGP_PARTIALS + MGP_GRAD_DOUBLES) {
          loo.grad[threadIdx.x - MGP_PARTIALS] = tot;
        }
      }
      if (threadIdx.x == 0) *loo.counter = 0u;
      if (loo.peers.world > 1) {
        __syncthreads();
        peer_sum8_block(loo.peers, s_tot, loo.partials);
      }
    }
  }
}

// ---- host side --------------------------------------------------------------------------
template <int T, int F, int D, int U, bool GRAD = false>
int launch_tp_inst(const TileArgs& a, const ColLoo& loo, long long rows, int* grid_out,
                   cudaStream_t stream) {
  constexpr int THREADS = (U + 1) * 32;
  const int pts_doubles = (((a.k + 1) * D) + 1) & ~1;
  const int ys_doubles = (a.k + 2) & ~1;
  const size_t warp_doubles = tp_warp_doubles<T, GRAD>(a.k, D);
  const size_t smem = warp_doubles * U * sizeof(double);
  cudaFuncAttributes fa;
  MGP_REQUIRE(cudaFuncGetAttributes(&fa, fused_tp_kernel<T, F, D, U, GRAD>) == cudaSuccess,
              MGP_ERR_CUDA, "cudaFuncGetAttributes failed");
  const size_t smem_cap = (size_t)max_smem_optin() - fa.sharedSizeBytes;
  MGP_REQUIRE(smem <= smem_cap, MGP_ERR_UNSUPPORTED,
              "thread-per-tile kernel shared memory %zu too large", smem);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(fused_tp_kernel<T, F, D, U, GRAD>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
    attr_set[dev] = true;
  }
  // one CTA per SM (launch bounds); every update warp of a CTA runs the same number of
  // neighbourhoods, rows past the end of the batch repeat the last row and write nothing
  long long blocks = (rows + U - 1) / U;
  const long long cap = sm_count();
  if (blocks > cap) blocks = cap;
  if (grid_out) {
    // a fixed grid keeps the partials' summation order independent of the batch size
    blocks = *grid_out > 0 ? *grid_out : cap;
    *grid_out = (int)blocks;
  }
  const long long iters = (rows + blocks * U - 1) / (blocks * U);
  fused_tp_kernel<T, F, D, U, GRAD><<<(unsigned)blocks, THREADS, smem, stream>>>(
      a, loo, pts_doubles, ys_doubles, (int)warp_doubles, iters);
  return check_launch("fused_tp_kernel");
}

// Update warps per CTA: as many as shared memory holds for the largest k of this T (8 T - 2), at
// most TP_MAX_UWARPS; T = 9 is capped at 11 so that the CTA has 12 warps -- three per scheduler,
// 168 registers per thread (its tile column and B fragments do not fit in 128).  The same cap
// holds for every GRAD instantiation (back substitution and gradient pass need the registers).
constexpr int TP_MAX_T = 13;
template <int T, int D, bool GRAD = false>
constexpr int tp_uwarps() {
  constexpr int k = 8 * T - 2;
  constexpr size_t pts = (size_t)((((k + 1) * D) + 1) & ~1), ys = (size_t)((k + 2) & ~1);
  size_t n = (size_t)tp_make_slots<T, !GRAD>().count * 64 + 8 * (size_t)T + 64 + 64 + 8 +
             2 * pts + 2 * ys + (GRAD ? (size_t)(T + 2) * 64 + 16 * (size_t)T : 0);
  while ((n & 15) != 2) n += 2;
  int u = (int)((227 * 1024 - 4096) / (n * sizeof(double)));
  if ((T == 9 || GRAD) && u > 11) u = 11;
  return u > TP_MAX_UWARPS ? TP_MAX_UWARPS : u;
}

template <int T, int F, int D, bool GRAD = false>
int launch_tp_one(const TileArgs& a, const ColLoo& loo, long long rows, int* grid_out,
                  cudaStream_t stream) {
  return launch_tp_inst<T, F, D, tp_uwarps<T, D, GRAD>(), GRAD>(a, loo, rows, grid_out, stream);
}

}  // namespace
}  // namespace mgp
