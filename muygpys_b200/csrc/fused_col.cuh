// K1 (column-direct variant) -- the kernel bench.py times.
//
// Same algorithm as fused_tile_kernel.cuh (one warp per neighbourhood, left-looking tiled LDL^T
// on the augmented matrix, 8x8 tiles as mma.sync.m8n8k4.f64 accumulator fragments), but with the
// data flow turned around so that a neighbourhood in flight needs ~100 registers and ~13 KB of
// shared memory instead of 168 registers and 18.5 KB, and 16 warps fit on an SM instead of 12:
//
//   * NO shared-memory image of the matrix and no element table.  The covariance entries of tile
//     column J are evaluated DIRECTLY INTO THE ACCUMULATOR FRAGMENT that owns them (lane (rho,q)
//     computes rows 8I+rho, columns 8J+2q / 8J+2q+1 of every tile I >= J) right before the
//     column is factorised: no STS / LDS / __syncwarp round trip, no bank conflicts, and the
//     2 (T-J) evaluations of a column are independent instruction streams.
//   * Only the CURRENT tile column lives in registers (2 T doubles).  Finished tiles U = L D go
//     to shared memory in fragment order -- every lane writes and later re-reads its own 16
//     bytes (LDS.128, conflict-free, lane-private: no synchronisation) -- and come back as DMMA
//     A/B fragments.  1/d lives in a 64-entry shared array.
//   * The rows of the LAST tile row that hold real entries (the last K rows and the
//     cross-covariance row: 3 of 8 rows at k = 50) are evaluated by a compact loop over
//     (row, column) and parked in the shared-memory slots their finished tiles will occupy
//     later, so the padding rows cost nothing; the target row is read from the staged targets.
//   * Source order interleaves the in-tile LDL^T column steps of the diagonal tile (a serial
//     shuffle -> reciprocal -> multiply -> FMA chain) with the assembly + DMMA update of the
//     tiles below it, which do not depend on the chain: ptxas keeps source order locally, so
//     the chain's stalls are filled from the same warp, and four warps per scheduler cover the
//     rest.
//
// Layout of the augmented matrix (m = 8 T rows): rows/columns 0..k-1 K + nugget, row k the
// cross-covariance (Kout = 1 at (k,k)), row k+1 the targets; columns 0..k-1 are eliminated and
// the Schur complement at rows/columns k, k+1 holds var, -mean and -y^T K^-1 y.  k is a run-time
// value with 8 T - 9 <= k <= 8 T - 2.
//
// Restrictions (everything else takes fused_tile / fused_generic): r == 1, d <= 3, homoscedastic
// nugget, T <= 8, covariance formula in {M05, M15, M25, GAUSS}.  The GRAD instantiations add a
// back substitution on the stored factor (w = K^-1 kcross, alpha = K^-1 y): the analytic
// gradient of the objective and the fast-mean coefficient output come from it.
#pragma once

#include <type_traits>

#include "peer_sum.cuh"
#include "tile_common.cuh"

namespace mgp {

// Optional fused epilogue: leave-one-out loss / scale partials of the batch (a14-a16).
struct ColLoo {
  double* warp_rec;        // (grid, record length) per-block partial records, or NULL
  double* partials;        // (MGP_PARTIALS) final record, written by the last CTA
  double* grad;            // (MGP_GRAD_DOUBLES) gradient sums (gradient kernels only)
  int backsub;             // use the kernels with the back substitution (gradient, coefficients)
  double inv_len[3];       // 1 / l_f of the features (gradient kernels only)
  unsigned int* counter;   // arrival counter (self-resetting)
  int loss_id;
  double boundary_scale;
  mgp_peer_group peers;    // world > 1: sum the record across GPUs in the same kernel
};

namespace {

constexpr int COL_WARPS = 4;
#ifndef MGP_COL_MINB
#define MGP_COL_MINB 4
#endif
constexpr int COL_MAX_T = 8;

// N covariance values at once, written stage by stage: GPUs issue in order within a warp and
// ptxas follows source order locally (the first version of this kernel called a scalar
// routine twice per tile and got two back-to-back serial chains, `wait` stalls on every DFMA of
// the polynomial).  Here every dependency level offers N independent instructions.
// MODE 0: the covariance.  MODE 1: phi = -K'(s) / s, the factor of the length-scale derivative
// dK/dl_f = phi z_f^2 / l_f (z = prescaled coordinate differences, s = |z|): e^-s / s (M1/2),
// e^-s (M3/2), (1 + s) e^-s / 3 (M5/2), e^(-u/2) (Gaussian).
template <int F, int N, int MODE = 0>
__device__ __forceinline__ void cov_n(const double (&u2)[N], double tab64, double (&out)[N]) {
#ifdef MGP_DBG_NOEVAL
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = u2[i] * tab64;
  return;
#endif
  const double LOG2E = 1.4426950408889634;
  const double MAGIC = 211106232532992.0;  // 1.5 * 2^47: ulp = 2^-5
  double s[N];
  double rinv[N];  // 1 / s (MODE 1, M1/2 only)
  if (F == F_GAUSS) {
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = 0.5 * u2[i];
  } else {
    double r[N], g[N], e[N], v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = rsqrt_seed(u2[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = u2[i] * r[i];
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = fma(-g[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fma(e[i], 0.375, 0.5);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = e[i] * v[i];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = fma(g[i], v[i], g[i]);
    if (MODE == 1 && F == F_M05) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        rinv[i] = (__double2hiint(u2[i]) > 0x03c00000) ? fma(r[i], v[i], r[i]) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = (__double2hiint(u2[i]) > 0x03c00000) ? s[i] : 0.0;
  }
  double t[N], tabv[N], gg[N], p[N];
  int ki[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(s[i], -LOG2E, MAGIC);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    ki[i] = __double2loint(t[i]);
    tabv[i] = __shfl_sync(0xffffffffu, tab64, ki[i] & (EXP_TABLE - 1));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = t[i] - MAGIC;
#pragma unroll
  for (int i = 0; i < N; ++i) gg[i] = fma(s[i], -LOG2E, -t[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], 0.00015403530393381608, 0.0013333558146428443);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], p[i], 0.009618129107628477);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], p[i], 0.05550410866482158);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], p[i], 0.2402265069591007);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], p[i], 0.6931471805599453);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(gg[i], p[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = tabv[i] * p[i];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double sc = __hiloint2double(__double2hiint(p[i]) + ((ki[i] >> 5) << 20),
                                       __double2loint(p[i]));
    p[i] = (__double2hiint(s[i]) < 0x4085e000) ? sc : 0.0;  // s < 700
  }
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (F == F_M05) out[i] = p[i] * rinv[i];
      else if (F == F_M25) out[i] = fma(s[i], 1.0 / 3.0, 1.0 / 3.0) * p[i];
      else out[i] = p[i];  // M3/2: e^-s; Gaussian: e^(-u/2)
    }
    return;
  }
  if (F == F_M15) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = fma(s[i], p[i], p[i]);  // (1 + s) e^-s
  } else if (F == F_M25) {
#pragma unroll
    for (int i = 0; i < N; ++i)  // (1 + s + s^2 / 3) e^-s
      out[i] = fma(fma(u2[i], 1.0 / 3.0, s[i]), p[i], p[i]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = p[i];
  }
}

template <int D>
struct Pt {
  double x[D];
};

template <int D>
__device__ __forceinline__ Pt<D> ld_pt(const double* __restrict__ pts, int i) {
  Pt<D> p;
  if (D == 2) {
    const double2 v = reinterpret_cast<const double2*>(pts)[i];
    p.x[0] = v.x;
    p.x[1] = v.y;
  } else {
#pragma unroll
    for (int f = 0; f < D; ++f) p.x[f] = pts[i * D + f];
  }
  return p;
}

template <int D>
__device__ __forceinline__ double sq_dist(const Pt<D>& a, const Pt<D>& b) {
  const double d0 = a.x[0] - b.x[0];
  double u = d0 * d0;
#pragma unroll
  for (int f = 1; f < D; ++f) {
    const double df = a.x[f] - b.x[f];
    u = fma(df, df, u);
  }
  return u;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}

// slot of finished tile (I,P), I > P, in the per-warp shared-memory store
__host__ __device__ constexpr int lslot(int I, int P) { return I * (I - 1) / 2 + P; }

// per-warp shared memory in doubles
static inline size_t col_warp_doubles(int T, int k, int d, bool grad = false) {
  const size_t L = (size_t)(T * (T - 1) / 2) * 64;
  const size_t dinv = 8 * (size_t)T;
  const size_t pts = (size_t)((((k + 1) * d) + 1) & ~1);
  const size_t ys = (size_t)((k + 2) & ~1);
  // gradient kernels keep M_J of every tile column, the last two diagonal tiles and the
  // solution vectors w = K^-1 kcross, alpha = K^-1 y
  const size_t g = grad ? (size_t)(T + 2) * 64 + 16 * (size_t)T : 0;
  return L + dinv + 2 * pts + 2 * ys + g;
}

constexpr int COL_NREC_GRAD = 32;  // MGP_PARTIALS + MGP_GRAD_DOUBLES, padded

template <int T, int F, int D, bool GRAD = false>
__global__ void __launch_bounds__(COL_WARPS * 32, GRAD ? 3 : MGP_COL_MINB)
    fused_col_kernel(const TileArgs a, const ColLoo loo, int pts_doubles, int ys_doubles,
                     int warp_doubles) {
  extern __shared__ double smem[];
  constexpr int NL = T * (T - 1) / 2;
  constexpr int NREC = GRAD ? COL_NREC_GRAD : MGP_PARTIALS;
  // per-warp loss / scale (/ gradient) sums, updated by lane 0 once per neighbourhood
  __shared__ double s_acc[COL_WARPS][NREC];
  for (int e = threadIdx.x; e < COL_WARPS * NREC; e += blockDim.x) (&s_acc[0][0])[e] = 0.0;
  __syncthreads();
  constexpr int W = 8 * (T - 1);  // columns left of the last diagonal tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rho = lane >> 2, q = lane & 3, qb = lane & ~3;
  const int k = a.k;
  const int nel = k + 1 - W;  // rows of the last tile row that are evaluated (last K rows + cross)
  const int kl = k & 7;

  const double tab64 = a.exp_tab[lane];
  double* Ls = smem + (size_t)warp * warp_doubles;  // finished tiles, fragment order
  double* dinv_s = Ls + NL * 64;                    // -1/d per eliminated column
  double* pts_buf = dinv_s + 8 * T;                 // 2 x (k+1) points, point k = query
  double* ys_buf = pts_buf + 2 * pts_doubles;       // 2 x k targets
  double* Ms = ys_buf + 2 * ys_doubles;             // GRAD: M_J = L_JJ^-T per tile column
  double* Dg = Ms + T * 64;                         // GRAD: final diagonal tiles T-2, T-1
  double* wv = Dg + 2 * 64;                         // GRAD: w = K^-1 kcross (8 T entries)
  double* av = wv + 8 * T;                          // GRAD: alpha = K^-1 y

  const long long wglobal = (long long)blockIdx.x * COL_WARPS + warp;
  const long long wstride = (long long)gridDim.x * COL_WARPS;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double onep = 1.0 + a.noise;

  // lane l stages points l and l+32 (point k is the query); k + 1 <= 8 T - 1 <= 63
  auto load_src = [&](long long row, int i) -> long long {
    if (row >= a.b || i > k) return -1;
    if (i == k) return a.query_idx ? a.query_idx[row] : row;
    return a.nn_idx[row * k + i];
  };
  auto issue_rows = [&](int buf, int i, long long src) {
    if (src < 0) return;
    const double* px = ((i == k) ? a.query_x : a.train_x) + src * D;
    double* dst = pts_buf + buf * pts_doubles + i * D;
    if (D == 2) {
      cp_async16(dst, px);  // one 16-byte copy per 2-D point
    } else {
#pragma unroll
      for (int f = 0; f < D; ++f) cp_async8(dst + f, px + f);
    }
    if (i < k) cp_async8(ys_buf + buf * ys_doubles + i, a.train_y + src);
  };

  long long s0 = load_src(wglobal, lane), s1 = load_src(wglobal, lane + 32);
  long long q_src = __shfl_sync(0xffffffffu, (k < 32) ? s0 : s1, k & 31);
  issue_rows(0, lane, s0);
  issue_rows(0, lane + 32, s1);
  cp_async_commit();
  s0 = load_src(wglobal + wstride, lane);
  s1 = load_src(wglobal + wstride, lane + 32);

  int buf = 0;
  for (long long row = wglobal; row < a.b; row += wstride, buf ^= 1) {
    cp_async_wait_all();
    __syncwarp();
    const long long q_next = __shfl_sync(0xffffffffu, (k < 32) ? s0 : s1, k & 31);
    issue_rows(buf ^ 1, lane, s0);
    issue_rows(buf ^ 1, lane + 32, s1);
    cp_async_commit();
    s0 = load_src(row + 2 * wstride, lane);
    s1 = load_src(row + 2 * wstride, lane + 32);
    double* pts = pts_buf + buf * pts_doubles;
    const double* ys = ys_buf + buf * ys_doubles;
    // fold the length scale(s) (and the Matern sqrt(2 nu)) into the staged coordinates
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

    // ---- compact evaluation of the real rows of the last tile row (columns < W) ----------
    // entry e = (row W + e / W, column e % W); three / two / one chunks of 32 entries at a time
    if (T > 1) {
      const int total = nel * W;
      auto chunk = [&](int base, auto nway) {
        constexpr int N = decltype(nway)::value;
        double u[N], o[N];
        int dst[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const int e = base + 32 * i + lane;
          const int ee = e < total ? e : 0;
          const int ar = ee / W, j = ee - ar * W;
          u[i] = sq_dist<D>(ld_pt<D>(pts, W + ar), ld_pt<D>(pts, j));
          dst[i] = e < total ? lslot(T - 1, j >> 3) * 64 + ar * 8 + (j & 7) : -1;
        }
        cov_n<F, N>(u, tab64, o);
#pragma unroll
        for (int i = 0; i < N; ++i)
          if (dst[i] >= 0) Ls[dst[i]] = o[i];
      };
      int base = 0;
      for (; base + 64 < total; base += 96) chunk(base, std::integral_constant<int, 3>());
      if (base + 32 < total) chunk(base, std::integral_constant<int, 2>());
      else if (base < total) chunk(base, std::integral_constant<int, 1>());
      __syncwarp();
    }

    bool ok = true;
    double out_var = 0.0, out_mean = 0.0, out_yky = 0.0;
#pragma unroll
    for (int J = 0; J < T; ++J) {
      double c[T][2];
#ifdef MGP_DBG_NOSTEPS
      const int ncols = 0;
#else
      const int ncols = (J <= T - 3) ? 8 : max(0, min(8, k - 8 * J));
#endif
      // this lane's two column points (clamped: columns beyond k are masked below)
      const int j0 = 8 * J + 2 * q, j1 = j0 + 1;
      const Pt<D> pc0 = ld_pt<D>(pts, (J == T - 1) ? min(j0, k) : j0);
      const Pt<D> pc1 = ld_pt<D>(pts, (J == T - 1) ? min(j1, k) : j1);
      // B fragments of the finished tile columns: U[J][P] D_P^-1, negated (dinv_s holds -1/d)
      double b0[T], b1[T];
      double2 ljs[T];
#pragma unroll
      for (int P = 0; P < J; ++P) {
        ljs[P] = *reinterpret_cast<const double2*>(Ls + lslot(J, P) * 64 + 2 * lane);
        const double2 nd = *reinterpret_cast<const double2*>(dinv_s + 8 * P + 2 * q);
        b0[P] = ljs[P].x * nd.x;
        b1[P] = ljs[P].y * nd.y;
      }

      // diagonal entries of a regular diagonal tile: 1 + nugget (Kout = 1 on the cross row)
      auto fix_diag = [&](int I) {
        const double dg = (8 * I + rho < k) ? onep : 1.0;
        c[I][0] = (rho == 2 * q) ? dg : c[I][0];
        c[I][1] = (rho == 2 * q + 1) ? dg : c[I][1];
      };
      // regular tiles (rows <= k): evaluated straight into their accumulator fragments, two
      // tiles (four entries per lane) interleaved
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
        const double2 v = *reinterpret_cast<const double2*>(Ls + lslot(T - 1, J) * 64 + 2 * lane);
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
      // left-looking update with every finished tile column; two tiles alternate so that no
      // DMMA waits for the previous one on the same accumulator
      auto frag = [&](int I, int P) -> double2 {
        return (I == J) ? ljs[P]
                        : *reinterpret_cast<const double2*>(Ls + lslot(I, P) * 64 + 2 * lane);
      };
      auto update_two = [&](int Ia, int Ib) {
#pragma unroll
        for (int P = 0; P < J; ++P) {
          const double2 la = frag(Ia, P), lb = frag(Ib, P);
          dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
          dmma_free(c[Ib][0], c[Ib][1], lb.x, b0[P]);
          dmma_free(c[Ia][0], c[Ia][1], la.y, b1[P]);
          dmma_free(c[Ib][0], c[Ib][1], lb.y, b1[P]);
        }
      };
      auto update_one = [&](int Ia) {
        if (J >= 2) {  // two partial sums: even / odd slices
          double x0 = 0.0, x1 = 0.0;
#pragma unroll
          for (int P = 0; P < J; ++P) {
            const double2 la = frag(Ia, P);
            dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
            dmma_free(x0, x1, la.y, b1[P]);
          }
          c[Ia][0] += x0;
          c[Ia][1] += x1;
        } else {
#pragma unroll
          for (int P = 0; P < J; ++P) {
            const double2 la = frag(Ia, P);
            dmma_free(c[Ia][0], c[Ia][1], la.x, b0[P]);
            dmma_free(c[Ia][0], c[Ia][1], la.y, b1[P]);
          }
        }
      };
      // Work items of this column besides the first one, in pairs of tiles: regular tiles
      // J+2 .. T-2, then the last tile row.  Item m runs after column step 2 m.
      auto work_item = [&](int m) {
        const int Ia = J + 2 + 2 * m, Ib = Ia + 1;
        if (J > T - 3 || Ia > T - 1) return;
        if (Ib <= T - 2) {
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
      };

      if (J <= T - 3) {
        eval_two(J, J + 1);
        update_two(J, J + 1);
      } else if (J == T - 2) {
        eval_one(J);
        load_last();
        update_two(J, T - 1);
      } else {
        eval_corner();
        update_one(J);
      }
      // in-tile LDL^T on the diagonal tile and on an identity tile (-> M, so that every tile
      // below becomes S M with two DMMAs); the other tiles of the column are assembled and
      // updated between the column steps, which they do not depend on
      double v0 = (rho == 2 * q) ? 1.0 : 0.0, v1 = (rho == 2 * q + 1) ? 1.0 : 0.0;
      double di0 = 0.0, di1 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ncols) {
          const int qj = j >> 1, bj = j & 1;
          const double cj = bj == 0 ? c[J][0] : c[J][1];
          const double p = shfl_d(cj, j * 4 + qj);
          const double uc0 = shfl_d(cj, (2 * q) * 4 + qj);      // U[2q][j]
          const double uc1 = shfl_d(cj, (2 * q + 1) * 4 + qj);  // U[2q+1][j]
          const double lr = shfl_d(cj, qb | qj);                // U[row][j]
          const double vr = shfl_d(bj == 0 ? v0 : v1, qb | qj); // V[row][j]
          // positive, finite, normal pivot (integer test: keeps the FP64 pipe for arithmetic)
          ok = ok && ((unsigned)(__double2hiint(p) - 1) < 0x7fefffffu);
          const double pinv = rcp_fast(p);
          if (bj == 0) di0 = sel_d(q == qj, pinv, di0); else di1 = sel_d(q == qj, pinv, di1);
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
        if ((j & 1) == 0) work_item(j >> 1);
      }
      if (rho == 0) *reinterpret_cast<double2*>(dinv_s + 8 * J + 2 * q) = make_double2(-di0, -di1);
      if (GRAD) {
        *reinterpret_cast<double2*>(Ms + J * 64 + 2 * lane) = make_double2(v0, v1);
        if (J >= T - 2)
          *reinterpret_cast<double2*>(Dg + (J - (T - 2)) * 64 + 2 * lane) =
              make_double2(c[J][0], c[J][1]);
      }
      // ---- outputs from the Schur complement ---------------------------------------------
      if (J == T - 1) {
        if (kl < 7) {
          const double cv = (kl & 1) ? c[J][1] : c[J][0];
          const double cy = ((kl + 1) & 1) ? c[J][1] : c[J][0];
          out_var = shfl_d(cv, kl * 4 + (kl >> 1));
          out_mean = -shfl_d(cv, (kl + 1) * 4 + (kl >> 1));
          out_yky = -shfl_d(cy, (kl + 1) * 4 + ((kl + 1) >> 1));
        } else {
          out_yky = -shfl_d(c[J][0], 0);
        }
      }
      if (J == T - 2 && kl == 7) out_var = shfl_d(c[J][1], 31);
      if (J + 1 < T) {
        // B fragments of M: even rows {0,2,4,6} and odd rows, lane l = (kk = l&3, n = l>>2)
        const int srcE = 8 * q + (lane >> 3), par = (lane >> 2) & 1;
        const double e0 = shfl_d(v0, srcE), e1 = shfl_d(v1, srcE);
        const double o0 = shfl_d(v0, srcE + 4), o1 = shfl_d(v1, srcE + 4);
        const double bm0 = sel_d(par, e1, e0), bm1 = sel_d(par, o1, o0);
        double n0[T], n1[T];
#pragma unroll
        for (int I = J + 1; I < T; ++I) {
          n0[I] = 0.0;
          n1[I] = 0.0;
          dmma_free(n0[I], n1[I], c[I][0], bm0);
        }
#pragma unroll
        for (int I = J + 1; I < T; ++I) {
          dmma_free(n0[I], n1[I], c[I][1], bm1);
          *reinterpret_cast<double2*>(Ls + lslot(I, J) * 64 + 2 * lane) =
              make_double2(n0[I], n1[I]);
        }
        if (J == T - 2 && kl == 7) out_mean = -shfl_d(n1[T - 1], 3);
      }
      __syncwarp();  // dinv_s of this column is read by every lane from the next column on
    }

    // ---- gradient of (mean, variance, y^T K^-1 y) w.r.t. length scale(s) and nugget ---------
    // d mean = dc^T alpha - w^T dK alpha, d var = -2 dc^T w + w^T dK w, d yky = -alpha^T dK alpha
    // with w = K^-1 kcross, alpha = K^-1 y from a back substitution on the stored factor, and
    // dK_ij / dl_f = phi(s_ij) z_f^2 / l_f re-evaluated entry by entry (never stored).
    double gdm[4] = {0.0, 0.0, 0.0, 0.0}, gdv[4] = {0.0, 0.0, 0.0, 0.0},
           gdy[4] = {0.0, 0.0, 0.0, 0.0};
    if (GRAD) {
      const int Jlast = (k - 1) >> 3;           // last tile column that holds K columns
      const int Rk = k >> 3, Ry = (k + 1) >> 3;  // tile rows of the cross row and the target row
      const int ly = (k + 1) & 7;
      for (int e = lane; e < 8 * T; e += 32) {
        wv[e] = 0.0;
        av[e] = 0.0;
      }
      __syncwarp();
      double xw[T], xa[T];  // this lane's row (rho) of the solution blocks
#pragma unroll
      for (int I = 0; I < T; ++I) xw[I] = xa[I] = 0.0;
      // tile (row tile R, column tile J) as stored: finished tile, or final diagonal tile
      auto stored = [&](int R, int J2, int off) -> double2 {
        const double* base = (J2 < R) ? Ls + lslot(R, J2) * 64 : Dg + (R - (T - 2)) * 64;
        return *reinterpret_cast<const double2*>(base + off);
      };
#pragma unroll
      for (int J = T - 1; J >= 0; --J) {
        if (J <= Jlast) {
          const int nc = min(8, k - 8 * J);  // eliminated columns of this tile column
          // t = z - sum_{I > J} U[I][J]^T x_I, z = rows k / k+1 of U (= L^-1 kcross, L^-1 y)
          double ac0 = 0.0, ac1 = 0.0, ay0 = 0.0, ay1 = 0.0;
#pragma unroll
          for (int I = J + 1; I < T; ++I) {
            if (I <= Jlast) {
              const double2 u =
                  *reinterpret_cast<const double2*>(Ls + lslot(I, J) * 64 + 2 * lane);
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
          const double2 zc = (J <= Rk) ? stored(Rk, J, kl * 8 + 2 * q) : make_double2(0.0, 0.0);
          const double2 zy = stored(Ry, J, ly * 8 + 2 * q);
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
      if (a.coeffs)
        for (int e = lane; e < k; e += 32) a.coeffs[row * k + e] = ok ? av[e] : nan;
      if (loo.grad != nullptr) {
      // weighted re-evaluation: per tile row I the row-factored sums
      //   Ra_f = sum_j phi z_f^2 alpha_j,  Rw_f = sum_j phi z_f^2 w_j   (j over tile columns <= I)
      // fold into  w^T dK alpha, w^T dK w, alpha^T dK alpha  (diagonal tiles count pairs twice)
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
        const int ja = min(lane, k), jb = min(lane + 32, k);
        const Pt<D> p0 = ld_pt<D>(pts, ja), p1 = ld_pt<D>(pts, jb);
        const double u[2] = {sq_dist<D>(pq, p0), sq_dist<D>(pq, p1)};
        double ph[2];
        cov_n<F, 2, 1>(u, tab64, ph);
        const double w0 = lane < k ? wv[ja] : 0.0, a0 = lane < k ? av[ja] : 0.0;
        const double w1 = lane + 32 < k ? wv[jb] : 0.0, a1 = lane + 32 < k ? av[jb] : 0.0;
#pragma unroll
        for (int f = 0; f < D; ++f) {
          const double d0 = pq.x[f] - p0.x[f], d1 = pq.x[f] - p1.x[f];
          const double g0 = ph[0] * (d0 * d0), g1 = ph[1] * (d1 * d1);
          Gmc[f] = fma(g0, a0, fma(g1, a1, Gmc[f]));
          Gvc[f] = fma(g0, w0, fma(g1, w1, Gvc[f]));
        }
      }
      // nugget: dK = I
      double dwa = 0.0, dww = 0.0, daa = 0.0;
      for (int e = lane; e < 8 * T; e += 32) {
        dwa = fma(wv[e], av[e], dwa);
        dww = fma(wv[e], wv[e], dww);
        daa = fma(av[e], av[e], daa);
      }
      // the quad lanes of a row evaluated different columns of the same row: plain warp sums
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
      }
    }

    // ---- outputs -----------------------------------------------------------------------
    if (lane == 0) {
      if (a.var) a.var[row] = ok ? a.scale * out_var : nan;
      if (a.mean) a.mean[row] = ok ? out_mean : nan;
      if (a.yky) a.yky[row] = ok ? out_yky : nan;
      if (a.status) a.status[row] = ok ? 0 : 1;
      if (loo.warp_rec) {
        // leave-one-out partials: the target of batch row `row` is train_y[batch index]
        double* acc = s_acc[warp];
        if (ok) {
          const double err = out_mean - a.train_y[q_src];
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
            acc[MGP_P_AUX] += loo.boundary_scale * loo.boundary_scale * (sqrt(fma(z, z, 1.0)) - 1.0);
          }
          if (GRAD) {
            const double iv = 1.0 / out_var;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < D || t == 3) {
                double* gr = acc + MGP_PARTIALS + 5 * t;
                gr[0] += 2.0 * err * gdm[t];            // d sum e^2
                gr[1] += hw * 2.0 * err * gdm[t] * iv;        // sum w 2 e dm / v
                gr[2] += hw * e2 * gdv[t] * iv * iv;          // sum w e^2 dv / v^2
                gr[3] += gdv[t] * iv;                    // sum dv / v
                gr[4] += gdy[t];                         // d sum yky
              }
            }
          }
        } else {
          acc[MGP_P_BAD] += 1.0;
        }
      }
    }
    q_src = q_next;
    __syncwarp();
  }
  cp_async_wait_all();

  if (loo.warp_rec) {
    // fixed-order reduction: warps of a block (in order) -> one record per block -> (last
    // block to finish) strided partial sums per slot -> sequential sum of those.
    // Bit-reproducible for a given grid; no floating-point atomics.
    __shared__ unsigned int s_last;
    constexpr int NGRP = COL_WARPS * 32 / NREC;
    __shared__ double s_red[NGRP][NREC];
    __syncthreads();
    if (threadIdx.x < NREC) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < COL_WARPS; ++w) v += s_acc[w][threadIdx.x];
      loo.warp_rec[(size_t)blockIdx.x * NREC + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(loo.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      const int slot = threadIdx.x % NREC, grp = threadIdx.x / NREC;
      double sum = 0.0;
      for (int c = grp; c < (int)gridDim.x; c += NGRP)
        sum += __ldcg(loo.warp_rec + (size_t)c * NREC + slot);
      s_red[grp][slot] = sum;
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
      if (threadIdx.x == 0) *loo.counter = 0u;  // ready for the next launch
      if (loo.peers.world > 1) {
        // cross-GPU sum over NVLink peer memory, still inside this kernel
        __syncthreads();
        peer_sum8_block(loo.peers, s_tot, loo.partials);
      }
    }
  }
}

// ---- host side --------------------------------------------------------------------------
static inline int col_tiles(int k) { return (k + 2 + 7) / 8; }

static inline bool col_formula_ok(int formula) {
  return formula == F_M05 || formula == F_M15 || formula == F_M25 || formula == F_GAUSS;
}

template <int T, int F, int D, bool GRAD = false>
int launch_col_one(const TileArgs& a, const ColLoo& loo, long long rows, int* grid_out,
                   cudaStream_t stream) {
  const int pts_doubles = (((a.k + 1) * D) + 1) & ~1;
  const int ys_doubles = (a.k + 2) & ~1;
  const size_t warp_doubles = col_warp_doubles(T, a.k, D, GRAD);
  const size_t smem = warp_doubles * COL_WARPS * sizeof(double);
  // the kernel also has static shared memory (reduction scratch)
  cudaFuncAttributes fa;
  MGP_REQUIRE(cudaFuncGetAttributes(&fa, fused_col_kernel<T, F, D, GRAD>) == cudaSuccess, MGP_ERR_CUDA,
              "cudaFuncGetAttributes failed");
  const size_t smem_cap = (size_t)max_smem_optin() - fa.sharedSizeBytes;
  MGP_REQUIRE(smem <= smem_cap, MGP_ERR_UNSUPPORTED,
              "column kernel shared memory %zu too large", smem);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(fused_col_kernel<T, F, D, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem_cap);
    attr_set[dev] = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_col_kernel<T, F, D, GRAD>,
                                                    COL_WARPS * 32, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  long long blocks = (rows + COL_WARPS - 1) / COL_WARPS;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (grid_out) {
    // a fixed grid keeps the partials' summation order independent of the batch size
    blocks = *grid_out > 0 ? *grid_out : cap;
    *grid_out = (int)blocks;
  }
  fused_col_kernel<T, F, D, GRAD><<<(unsigned)blocks, COL_WARPS * 32, smem, stream>>>(
      a, loo, pts_doubles, ys_doubles, (int)warp_doubles);
  return check_launch("fused_col_kernel");
}

}  // namespace
}  // namespace mgp
