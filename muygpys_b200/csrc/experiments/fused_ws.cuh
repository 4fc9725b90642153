// K1 (warp-specialised variant of the column-direct kernel).
//
// fused_col_kernel is bound by the 52 in-tile LDL^T column steps of a neighbourhood: a serial
// chain per warp (five 64-bit shuffles, a reciprocal, six arithmetic instructions, ~170 cycles
// per step) that four warps per scheduler overlap only to ~55 % because the same warps also
// carry the throughput work (covariance evaluations, DMMA updates).  Here the two kinds of work
// live in DIFFERENT warps of a 12-warp CTA:
//
//   * 8 UPDATE warps (one neighbourhood each): gather, evaluate the covariance entries of tile
//     column J into accumulator fragments, DMMA-update them with the finished tile columns,
//     hand the updated DIAGONAL tile to a factor warp through 512 bytes of shared memory, go on
//     with the tiles below the diagonal (which do not depend on the factorisation), then pick up
//     M_J = L_JJ^-T (as DMMA B fragments) and 1/d and finish the column with two DMMAs per tile.
//   * 4 FACTOR warps: factor warp f serves update warps 2f and 2f+1 and runs the eight column
//     steps of BOTH diagonal tiles in lock-step (two independent chains interleaved at source
//     level), extracts the Schur-complement outputs and writes M, 1/d and the outputs back.
//
// Registers are rebalanced with setmaxnreg: the CTA is launched with 80 registers per thread (two
// CTAs of 384 threads per SM), the factor warpgroup gives back down to 48 and the two update
// warpgroups take 96 each (256 x 96 + 128 x 48 = 384 x 80) -> 16 update + 8 factor warps per SM.
// Hand-offs use named barriers (bar.arrive / bar.sync with 96 participants: two update warps and
// their factor warp), one pair of barriers per factor warp.  Finished tiles live in a
// shared-memory store whose slots are REUSED once a tile row is complete (15 instead of 21 slots
// at T = 7), which is what lets eight neighbourhoods' state fit twice on an SM.
//
// Same numerics, layout and restrictions as fused_col_kernel (r == 1, d <= 3, 7 <= k <= 62,
// homoscedastic nugget, no coefficient output, no gradient).
#pragma once

#include "../fused_col.cuh"

namespace mgp {
namespace {

constexpr int WS_UWARPS = 8;
constexpr int WS_FWARPS = 4;
constexpr int WS_THREADS = (WS_UWARPS + WS_FWARPS) * 32;

// ---- shared-memory slots of the finished tiles, reused over the factorisation --------------
// Tile (I,P), I > P, is written at the end of tile column P and last read during the update of
// tile column I.  The tiles of the LAST tile row also park the compactly evaluated raw entries
// from the start of the neighbourhood: they keep fixed slots 0 .. T-2.  Every other tile takes
// the lowest slot whose previous occupant (I', P') has I' <= P.
template <int T>
struct WsSlots {
  int s[T > 0 ? T : 1][T > 0 ? T : 1];
  int count;
};

template <int T>
constexpr WsSlots<T> ws_make_slots() {
  WsSlots<T> m{};
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

__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// per update warp, in doubles: finished tiles | -1/d | hand-off tile (diag tile in, M out) |
// outputs (var, -mean, -yky, ok) | 2 x points | 2 x targets
template <int T>
static inline size_t ws_warp_doubles(int k, int d) {
  constexpr WsSlots<T> SL = ws_make_slots<T>();
  const size_t pts = (size_t)((((k + 1) * d) + 1) & ~1);
  const size_t ys = (size_t)((k + 2) & ~1);
  return (size_t)SL.count * 64 + 8 * (size_t)T + 64 + 8 + 2 * pts + 2 * ys;
}

template <int T, int F, int D>
__global__ void __launch_bounds__(WS_THREADS, 2)
    fused_ws_kernel(const TileArgs a, const ColLoo loo, int pts_doubles, int ys_doubles,
                    int warp_doubles, long long iters) {
  extern __shared__ double smem[];
  constexpr WsSlots<T> SL = ws_make_slots<T>();
  constexpr int W = 8 * (T - 1);
  constexpr int NREC = MGP_PARTIALS;
  __shared__ double s_acc[WS_UWARPS][NREC];
  for (int e = threadIdx.x; e < WS_UWARPS * NREC; e += blockDim.x) (&s_acc[0][0])[e] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rho = lane >> 2, q = lane & 3, qb = lane & ~3;
  const int k = a.k;
  const int nel = k + 1 - W;
  const int kl = k & 7;
  auto warp_base = [&](int u) { return smem + (size_t)u * warp_doubles; };
  constexpr int OFF_DINV = SL.count * 64, OFF_X = OFF_DINV + 8 * T, OFF_XO = OFF_X + 64,
                OFF_PTS = OFF_XO + 8;

  if (warp >= WS_UWARPS) {
    // =================================== factor warp ========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const int f = warp - WS_UWARPS;
    double* base0 = warp_base(2 * f);
    double* base1 = warp_base(2 * f + 1);
    const int bar_in = 1 + 2 * f, bar_out = 2 + 2 * f;
    for (long long it = 0; it < iters; ++it) {
      for (int J = 0; J < T; ++J) {
        const int ncols = max(0, min(8, k - 8 * J));
        bar_sync(bar_in, 96);  // both clients have parked their updated diagonal tiles
        double2 ta = *reinterpret_cast<const double2*>(base0 + OFF_X + 2 * lane);
        double2 tb = *reinterpret_cast<const double2*>(base1 + OFF_X + 2 * lane);
        double a0 = ta.x, a1 = ta.y, b0 = tb.x, b1 = tb.y;
        double va0 = (rho == 2 * q) ? 1.0 : 0.0, va1 = (rho == 2 * q + 1) ? 1.0 : 0.0;
        double vb0 = va0, vb1 = va1;
        double da0 = 0.0, da1 = 0.0, db0 = 0.0, db1 = 0.0;
        bool oka = true, okb = true;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < ncols) {
            const int qj = j >> 1, bj = j & 1;
            const double ca = bj == 0 ? a0 : a1, cb = bj == 0 ? b0 : b1;
            // two independent chains, level by level
            const double pa = shfl_d(ca, j * 4 + qj), pb = shfl_d(cb, j * 4 + qj);
            const double ua0 = shfl_d(ca, (2 * q) * 4 + qj), ub0 = shfl_d(cb, (2 * q) * 4 + qj);
            const double ua1 = shfl_d(ca, (2 * q + 1) * 4 + qj),
                         ub1 = shfl_d(cb, (2 * q + 1) * 4 + qj);
            const double la = shfl_d(ca, qb | qj), lb = shfl_d(cb, qb | qj);
            const double ra = shfl_d(bj == 0 ? va0 : va1, qb | qj),
                         rb = shfl_d(bj == 0 ? vb0 : vb1, qb | qj);
            oka = oka && ((unsigned)(__double2hiint(pa) - 1) < 0x7fefffffu);
            okb = okb && ((unsigned)(__double2hiint(pb) - 1) < 0x7fefffffu);
            const double ia = rcp_fast(pa), ib = rcp_fast(pb);
            if (bj == 0) {
              da0 = sel_d(q == qj, ia, da0);
              db0 = sel_d(q == qj, ib, db0);
            } else {
              da1 = sel_d(q == qj, ia, da1);
              db1 = sel_d(q == qj, ib, db1);
            }
            const double ta0 = sel_d(2 * q > j, ua0, 0.0) * ia, tb0 = sel_d(2 * q > j, ub0, 0.0) * ib;
            const double ta1 = sel_d(2 * q + 1 > j, ua1, 0.0) * ia,
                         tb1 = sel_d(2 * q + 1 > j, ub1, 0.0) * ib;
            if (j < 6) {
              a0 = fma(-la, ta0, a0);
              b0 = fma(-lb, tb0, b0);
              va0 = fma(-ra, ta0, va0);
              vb0 = fma(-rb, tb0, vb0);
            }
            if (j < 7) {
              a1 = fma(-la, ta1, a1);
              b1 = fma(-lb, tb1, b1);
              va1 = fma(-ra, ta1, va1);
              vb1 = fma(-rb, tb1, vb1);
            }
          }
        }
        if (rho == 0) {
          *reinterpret_cast<double2*>(base0 + OFF_DINV + 8 * J + 2 * q) = make_double2(-da0, -da1);
          *reinterpret_cast<double2*>(base1 + OFF_DINV + 8 * J + 2 * q) = make_double2(-db0, -db1);
        }
        // outputs from the Schur complement (same positions as fused_col_kernel)
        if (J == T - 1) {
          if (kl < 7) {
            const double cva = (kl & 1) ? a1 : a0, cvb = (kl & 1) ? b1 : b0;
            const double cya = ((kl + 1) & 1) ? a1 : a0, cyb = ((kl + 1) & 1) ? b1 : b0;
            const double vra = shfl_d(cva, kl * 4 + (kl >> 1)), vrb = shfl_d(cvb, kl * 4 + (kl >> 1));
            const double mea = shfl_d(cva, (kl + 1) * 4 + (kl >> 1)),
                         meb = shfl_d(cvb, (kl + 1) * 4 + (kl >> 1));
            const double yka = shfl_d(cya, (kl + 1) * 4 + ((kl + 1) >> 1)),
                         ykb = shfl_d(cyb, (kl + 1) * 4 + ((kl + 1) >> 1));
            if (lane == 0) {
              base0[OFF_XO + 0] = vra;
              base0[OFF_XO + 1] = -mea;
              base0[OFF_XO + 2] = -yka;
              base1[OFF_XO + 0] = vrb;
              base1[OFF_XO + 1] = -meb;
              base1[OFF_XO + 2] = -ykb;
            }
          } else {
            const double yka = shfl_d(a0, 0), ykb = shfl_d(b0, 0);
            if (lane == 0) {
              base0[OFF_XO + 2] = -yka;
              base1[OFF_XO + 2] = -ykb;
            }
          }
        }
        if (J == T - 2 && kl == 7) {
          const double vra = shfl_d(a1, 31), vrb = shfl_d(b1, 31);
          if (lane == 0) {
            base0[OFF_XO + 0] = vra;
            base1[OFF_XO + 0] = vrb;
          }
        }
        if (lane == 0) {
          base0[OFF_XO + 3] = oka ? 1.0 : 0.0;
          base1[OFF_XO + 3] = okb ? 1.0 : 0.0;
        }
        // B fragments of M: even rows {0,2,4,6} and odd rows, lane l = (kk = l&3, n = l>>2)
        {
          const int srcE = 8 * q + (lane >> 3), par = (lane >> 2) & 1;
          const double e0 = shfl_d(va0, srcE), e1 = shfl_d(va1, srcE);
          const double o0 = shfl_d(va0, srcE + 4), o1 = shfl_d(va1, srcE + 4);
          *reinterpret_cast<double2*>(base0 + OFF_X + 2 * lane) =
              make_double2(sel_d(par, e1, e0), sel_d(par, o1, o0));
          const double f0 = shfl_d(vb0, srcE), f1 = shfl_d(vb1, srcE);
          const double g0 = shfl_d(vb0, srcE + 4), g1 = shfl_d(vb1, srcE + 4);
          *reinterpret_cast<double2*>(base1 + OFF_X + 2 * lane) =
              make_double2(sel_d(par, f1, f0), sel_d(par, g1, g0));
        }
        bar_arrive(bar_out, 96);
      }
    }
  } else {
    // =================================== update warp ========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int bar_in = 1 + 2 * (warp >> 1), bar_out = 2 + 2 * (warp >> 1);
    const double tab64 = a.exp_tab[lane];
    double* Ls = warp_base(warp);
    double* dinv_s = Ls + OFF_DINV;
    double* X = Ls + OFF_X;
    const double* xo = Ls + OFF_XO;
    double* pts_buf = Ls + OFF_PTS;
    double* ys_buf = pts_buf + 2 * pts_doubles;
    const long long wglobal = (long long)blockIdx.x * WS_UWARPS + warp;
    const long long wstride = (long long)gridDim.x * WS_UWARPS;
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

    long long s0 = load_src(wglobal, lane), s1 = load_src(wglobal, lane + 32);
    long long q_src = __shfl_sync(0xffffffffu, (k < 32) ? s0 : s1, k & 31);
    issue_rows(0, lane, s0);
    issue_rows(0, lane + 32, s1);
    cp_async_commit();
    s0 = load_src(wglobal + wstride, lane);
    s1 = load_src(wglobal + wstride, lane + 32);

    int buf = 0;
    for (long long it = 0; it < iters; ++it, buf ^= 1) {
      const long long row = wglobal + it * wstride;
      const bool live = row < a.b;
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

      // compact evaluation of the real rows of the last tile row (columns < W)
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
            dst[i] = e < total ? (j >> 3) * 64 + ar * 8 + (j & 7) : -1;  // slot of (T-1, j/8)
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
        const int j0 = 8 * J + 2 * q, j1 = j0 + 1;
        const Pt<D> pc0 = ld_pt<D>(pts, (J == T - 1) ? min(j0, k) : j0);
        const Pt<D> pc1 = ld_pt<D>(pts, (J == T - 1) ? min(j1, k) : j1);
        double b0[T], b1[T];
        double2 ljs[T];
#pragma unroll
        for (int P = 0; P < J; ++P) {
          ljs[P] = *reinterpret_cast<const double2*>(Ls + SL.s[J][P] * 64 + 2 * lane);
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
        auto load_last = [&]() {
          const double2 v = *reinterpret_cast<const double2*>(Ls + SL.s[T - 1][J] * 64 + 2 * lane);
          const double2 yv = *reinterpret_cast<const double2*>(ys + j0);
          const bool isy = rho == nel;
          const double y0 = (isy && j0 < k) ? yv.x : 0.0, y1 = (isy && j1 < k) ? yv.y : 0.0;
          c[T - 1][0] = (rho < nel) ? v.x : y0;
          c[T - 1][1] = (rho < nel) ? v.y : y1;
        };
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
                          : *reinterpret_cast<const double2*>(Ls + SL.s[I][P] * 64 + 2 * lane);
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
          if (J >= 2) {
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
        auto work_item = [&](int m) {
          const int Ia = J + 2 + 2 * m, Ib = Ia + 1;
          if (J > T - 3 || Ia > T - 1) return;
          if (Ib <= T - 2) {
            eval_two(Ia, Ib);
            update_two(Ia, Ib);
          } else if (Ia <= T - 2) {
            eval_one(Ia);
            load_last();
            update_two(Ia, T - 1);
          } else {
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
        // hand the updated diagonal tile to the factor warp ...
        *reinterpret_cast<double2*>(X + 2 * lane) = make_double2(c[J][0], c[J][1]);
        bar_arrive(bar_in, 96);
        // ... and go on with the tiles below the diagonal, which do not depend on it
#pragma unroll
        for (int m = 0; m < 4; ++m) work_item(m);
        bar_sync(bar_out, 96);  // M_J, 1/d and the outputs of this column are in place
        ok = ok && (xo[3] != 0.0);
        if (J == T - 1) {
          if (kl < 7) {
            out_var = xo[0];
            out_mean = xo[1];
          }
          out_yky = xo[2];
        }
        if (J == T - 2 && kl == 7) out_var = xo[0];
        if (J + 1 < T) {
          const double2 bm = *reinterpret_cast<const double2*>(X + 2 * lane);
          double n0[T], n1[T];
#pragma unroll
          for (int I = J + 1; I < T; ++I) {
            n0[I] = 0.0;
            n1[I] = 0.0;
            dmma_free(n0[I], n1[I], c[I][0], bm.x);
          }
#pragma unroll
          for (int I = J + 1; I < T; ++I) {
            dmma_free(n0[I], n1[I], c[I][1], bm.y);
            *reinterpret_cast<double2*>(Ls + SL.s[I][J] * 64 + 2 * lane) =
                make_double2(n0[I], n1[I]);
          }
          if (J == T - 2 && kl == 7) out_mean = -shfl_d(n1[T - 1], 3);
        }
      }

      if (lane == 0 && live) {
        if (a.var) a.var[row] = ok ? a.scale * out_var : nan;
        if (a.mean) a.mean[row] = ok ? out_mean : nan;
        if (a.yky) a.yky[row] = ok ? out_yky : nan;
        if (a.status) a.status[row] = ok ? 0 : 1;
        if (loo.warp_rec) {
          double* acc = s_acc[warp];
          if (ok) {
            const double err = out_mean - a.train_y[q_src];
            const double e2 = err * err;
            acc[MGP_P_SQERR] += e2;
            acc[MGP_P_COUNT] += 1.0;
            acc[MGP_P_YKY] += out_yky;
            acc[MGP_P_ROWS] += 1.0;
            acc[MGP_P_SQERR_V] += e2 / out_var;
            acc[MGP_P_LOGV] += log(out_var);
            if (loo.loss_id == MGP_LOSS_PSEUDO_HUBER) {
              const double z = err / loo.boundary_scale;
              acc[MGP_P_AUX] +=
                  loo.boundary_scale * loo.boundary_scale * (sqrt(fma(z, z, 1.0)) - 1.0);
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
  }

  if (loo.warp_rec) {
    // fixed-order reduction, as in fused_col_kernel: update warps of a block -> block record ->
    // (last block) strided partial sums -> sequential sum; then the cross-GPU exchange
    __shared__ unsigned int s_last;
    constexpr int NGRP = 16;
    __shared__ double s_red[NGRP][NREC];
    __syncthreads();
    if (threadIdx.x < NREC) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < WS_UWARPS; ++w) v += s_acc[w][threadIdx.x];
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
        s_tot[threadIdx.x] = tot;
        if (loo.peers.world <= 1) loo.partials[threadIdx.x] = tot;
      }
      if (threadIdx.x == 0) *loo.counter = 0u;
      if (loo.peers.world > 1) {
        __syncthreads();
        peer_sum8_block(loo.peers, s_tot, loo.partials);
      }
    }
  }
}

template <int T, int F, int D>
int launch_ws_one(const TileArgs& a, const ColLoo& loo, long long rows, int* grid_out,
                  cudaStream_t stream) {
  const int pts_doubles = (((a.k + 1) * D) + 1) & ~1;
  const int ys_doubles = (a.k + 2) & ~1;
  const size_t warp_doubles = ws_warp_doubles<T>(a.k, D);
  const size_t smem = warp_doubles * WS_UWARPS * sizeof(double);
  cudaFuncAttributes fa;
  MGP_REQUIRE(cudaFuncGetAttributes(&fa, fused_ws_kernel<T, F, D>) == cudaSuccess, MGP_ERR_CUDA,
              "cudaFuncGetAttributes failed");
  const size_t smem_cap = (size_t)max_smem_optin() - fa.sharedSizeBytes;
  MGP_REQUIRE(smem <= smem_cap, MGP_ERR_UNSUPPORTED,
              "warp-specialised kernel shared memory %zu too large", smem);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(fused_ws_kernel<T, F, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem_cap);
    attr_set[dev] = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_ws_kernel<T, F, D>,
                                                    WS_THREADS, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  long long blocks = (rows + WS_UWARPS - 1) / WS_UWARPS;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (grid_out) {
    blocks = *grid_out > 0 ? *grid_out : cap;
    *grid_out = (int)blocks;
  }
  const long long iters = (rows + blocks * WS_UWARPS - 1) / (blocks * WS_UWARPS);
  fused_ws_kernel<T, F, D><<<(unsigned)blocks, WS_THREADS, smem, stream>>>(
      a, loo, pts_doubles, ys_doubles, (int)warp_doubles, iters);
  return check_launch("fused_ws_kernel");
}

}  // namespace
}  // namespace mgp
