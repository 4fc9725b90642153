// K1 (pipelined tile variant) -- the register-tile kernel of fused_tile.cu with the
// covariance ASSEMBLY OF THE NEXT NEIGHBOURHOOD software-pipelined into the FACTORISATION of
// the current one, inside the same warp.
//
// Why: ncu shows the tile kernel limited by per-warp latency, not by a pipe.  The in-tile
// LDL^T column steps are a serial chain (shuffle -> reciprocal -> multiply -> FMA, ~100 cycles
// per column, 52 columns) that keeps its warp's issue slot ~90 % idle, register/shared-memory
// capacity allows only 3 warps per scheduler, and GPUs issue in order within a warp.  The
// assembly (exp / sqrt polynomials, fully independent per element) is the opposite: all
// arithmetic, no dependencies.  Emitting one tile column's worth of next-neighbourhood
// assembly in the SAME straight-line block as the current tile column's factorisation lets
// ptxas schedule the two instruction streams into each other's stalls.
//
// To give the scheduler branch-free blocks everything shape-like is a template parameter:
// T (tile rows), KP (eliminated columns = k rounded up to 4) and the covariance formula F.
// Restrictions (everything else falls back to fused_tile.cu): d == 2, r == 1, homoscedastic
// nugget, no coefficient output, Schur block inside the last tile.
#include "tile_common.cuh"

namespace mgp {

namespace {

constexpr int PIPE_WARPS = 4;

// table geometry: tile column J owns chunk_iters<KP>(J) rows of 32 entries
template <int KP>
__host__ __device__ constexpr int chunk_iters(int J) {
  int n = 0;
  for (int j = 8 * J; j < 8 * J + 8 && j < KP; ++j) n += KP - j + 1;
  return (n + 31) / 32;
}
template <int KP>
__host__ __device__ constexpr int chunk_begin(int J) {
  int s = 0;
  for (int c = 0; c < J; ++c) s += chunk_iters<KP>(c);
  return s;
}

template <int F>
__device__ __forceinline__ void eval_entry(unsigned p, double* __restrict__ tiles,
                                           const double2* __restrict__ pts2,
                                           double tab64, double noise) {
  const int pi = (p >> 8) & 255, j = p & 255;
  const double2 a = pts2[pi], b = pts2[j];
  const double dx = a.x - b.x, dy = a.y - b.y;
  const double u2 = fma(dy, dy, dx * dx);
  double v = neg_cov<F>(u2, tab64, 1.0, 0);
  v -= (pi == j) ? noise : 0.0;  // nugget on the diagonal (N = -(K + eps))
  tiles[p >> 16] = v;
}

// two independent entries with their instruction streams interleaved statement by statement
// (in-order issue: a single evaluation is one ~25-deep dependency chain)
template <int F>
__device__ __forceinline__ void eval_pair(unsigned p0, unsigned p1, int dep,
                                          double* __restrict__ tiles,
                                          const double2* __restrict__ pts2,
                                          double tab64, double noise) {
  const int pi0 = (p0 >> 8) & 255, j0 = p0 & 255, pi1 = (p1 >> 8) & 255, j1 = p1 & 255;
  const double2 a0 = pts2[pi0 + dep], b0 = pts2[j0], a1 = pts2[pi1 + dep], b1 = pts2[j1];
  const double dx0 = a0.x - b0.x, dy0 = a0.y - b0.y, dx1 = a1.x - b1.x, dy1 = a1.y - b1.y;
  const double u0 = fma(dy0, dy0, dx0 * dx0), u1 = fma(dy1, dy1, dx1 * dx1);
  double v0 = neg_cov<F>(u0, tab64, 1.0, 0);
  double v1 = neg_cov<F>(u1, tab64, 1.0, 0);
  v0 -= (pi0 == j0) ? noise : 0.0;
  v1 -= (pi1 == j1) ? noise : 0.0;
  tiles[p0 >> 16] = v0;
  tiles[p1 >> 16] = v1;
}

template <int F>
__device__ __noinline__ void assemble_flat(double* tiles, const double2* pts2,
                                           const unsigned* etab, double tab64,
                                           int rows, int lane, double noise) {
  int it = 0;
  for (; it + 1 < rows; it += 2)
    eval_pair<F>(etab[it * 32 + lane], etab[(it + 1) * 32 + lane], 0, tiles, pts2, tab64, noise);
  if (it < rows) eval_entry<F>(etab[it * 32 + lane], tiles, pts2, tab64, noise);
}

// One covariance entry cut into 21 dependency LEVELS, grouped in three slices of seven.  The
// kernel keeps THREE entries in flight, one per slice, and advances each by one slice per
// in-tile column step, calling the levels of the three slices and of the LDL^T step
// alternately: GPUs issue in order within a warp and ptxas only reorders locally, so it is the
// program order itself that must offer four independent instructions per dependency level.
template <int F>
struct EvalState {
  unsigned p;
  int ki;
  double u, g, r, e, s, pref, t, tabv, pl;
  // ---- slice 1: squared distance, start of sqrt ------------------------------------------
  __device__ __forceinline__ void l0(unsigned entry, int dep, const double2* pts2) {
    p = entry;
    const double2 a = pts2[((p >> 8) & 255) + dep], b = pts2[p & 255];
    g = a.x - b.x;  // dx
    r = a.y - b.y;  // dy
  }
  __device__ __forceinline__ void l1() { u = g * g; }
  __device__ __forceinline__ void l2() { u = fma(r, r, u); }
  __device__ __forceinline__ void l3() { r = rsqrt_seed(u); }
  __device__ __forceinline__ void l4() { g = u * r; }
  __device__ __forceinline__ void l5() { e = fma(-g, r, 1.0); }
  __device__ __forceinline__ void l6() { r = fma(e, 0.375, 0.5); }
  // ---- slice 2: end of sqrt, range reduction of exp(-s), table lookup --------------------
  __device__ __forceinline__ void l7() { e = e * r; }
  __device__ __forceinline__ void l8() {
    const double sq = fma(g, e, g);
    s = (__double2hiint(u) > 0x03c00000) ? sq : 0.0;
  }
  __device__ __forceinline__ void l9() {
    t = fma(s, -1.4426950408889634, 211106232532992.0);
    if (F == F_M05) pref = -1.0;
    if (F == F_M15) pref = -1.0 - s;
    if (F == F_M25) pref = fma(u, -(1.0 / 3.0), -1.0 - s);
  }
  __device__ __forceinline__ void l10(double tab) {
    ki = __double2loint(t);
    tabv = __shfl_sync(0xffffffffu, tab, ki & (EXP_TABLE - 1));
    t = t - 211106232532992.0;
  }
  __device__ __forceinline__ void l11() { g = fma(s, -1.4426950408889634, -t); }
  __device__ __forceinline__ void l12() { pl = fma(g, 0.00015403530393381608, 0.0013333558146428443); }
  __device__ __forceinline__ void l13() { pl = fma(g, pl, 0.009618129107628477); }
  // ---- slice 3: rest of the polynomial, scaling, store ------------------------------------
  __device__ __forceinline__ void l14() { pl = fma(g, pl, 0.05550410866482158); }
  __device__ __forceinline__ void l15() { pl = fma(g, pl, 0.2402265069591007); }
  __device__ __forceinline__ void l16() { pl = fma(g, pl, 0.6931471805599453); }
  __device__ __forceinline__ void l17() { pl = fma(g, pl, 1.0); }
  __device__ __forceinline__ void l18() { pl = tabv * pl; }
  __device__ __forceinline__ void l19() {
    const double out = __hiloint2double(__double2hiint(pl) + ((ki >> 5) << 20), __double2loint(pl));
    pl = pref * ((__double2hiint(s) < 0x4085e000) ? out : 0.0);
  }
  __device__ __forceinline__ void l20(double* tiles, double noise) {
    const bool diag = ((p >> 8) & 255) == (p & 255);
    tiles[p >> 16] = pl - (diag ? noise : 0.0);
  }
};

// In-tile LDL^T column step S (global index over the whole factorisation) of the diagonal
// tile (c0,c1) and the identity tile (v0,v1).  Entry S of the next neighbourhood starts its
// first slice here, entry S-1 runs its second and entry S-2 its third.
template <int F, int S, int NEV>
__device__ __forceinline__ void column_step(double& c0, double& c1, double& v0, double& v1,
                                            double& di0, double& di1, bool& ok,
                                            double& prev_pinv, EvalState<F> (&ev)[3],
                                            const unsigned* etab, int lane, int q, int qb,
                                            double* tiles, const double2* pts2, double tab64,
                                            double noise) {
  constexpr int j = S & 7, qj = j >> 1, bj = j & 1;
  constexpr bool HAS_A = S < NEV, HAS_B = S >= 1 && S - 1 < NEV, HAS_C = S >= 2 && S - 2 < NEV;
  EvalState<F>& A = ev[S % 3];
  EvalState<F>& B = ev[(S + 2) % 3];
  EvalState<F>& C = ev[(S + 1) % 3];
  const int dep = (__double2hiint(prev_pinv) >> 31) & 1;  // always 0: pins program order
  const double cj = bj == 0 ? c0 : c1;
  // level 0: the five shuffles of this step read the (final, unscaled) column j
  const double p = shfl_d(cj, j * 4 + qj);
  const double uc0 = shfl_d(cj, (2 * q) * 4 + qj);
  const double uc1 = shfl_d(cj, (2 * q + 1) * 4 + qj);
  const double lr = shfl_d(cj, qb | qj);
  const double vr = shfl_d(bj == 0 ? v0 : v1, qb | qj);
  if (HAS_A) A.l0(etab[S * 32 + lane + dep], 0, pts2);
  if (HAS_B) B.l7();
  if (HAS_C) C.l14();
  // level 1
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  ok = ok && (p > 0.0);
  if (HAS_A) A.l1();
  if (HAS_B) B.l8();
  if (HAS_C) C.l15();
  // level 2
  const double e = fma(-p, r, 1.0);
  if (HAS_A) A.l2();
  if (HAS_B) B.l9();
  if (HAS_C) C.l16();
  // level 3
  const double q1 = r * e, w = 1.0 + e;
  if (HAS_A) A.l3();
  if (HAS_B) B.l10(tab64);
  if (HAS_C) C.l17();
  // level 4
  const double pinv = fma(q1, w, r);
  if (HAS_A) A.l4();
  if (HAS_B) B.l11();
  if (HAS_C) C.l18();
  // level 5
  if (bj == 0) di0 = sel_d(q == qj, pinv, di0); else di1 = sel_d(q == qj, pinv, di1);
  const double t0 = sel_d(2 * q > j, uc0, 0.0) * pinv;
  const double t1 = sel_d(2 * q + 1 > j, uc1, 0.0) * pinv;
  if (HAS_A) A.l5();
  if (HAS_B) B.l12();
  if (HAS_C) C.l19();
  // level 6
  if (j < 6) {
    c0 = fma(-lr, t0, c0);
    v0 = fma(-vr, t0, v0);
  }
  if (j < 7) {
    c1 = fma(-lr, t1, c1);
    v1 = fma(-vr, t1, v1);
  }
  if (HAS_A) A.l6();
  if (HAS_B) B.l13();
  if (HAS_C) C.l20(tiles, noise);
  prev_pinv = pinv;
}

struct ColCtx {
  double* tiles;
  const double2* pts2;  // next neighbourhood's (scaled) points
  const double* ys;     // next neighbourhood's targets
  const unsigned* etab;
  double tab64;  // this lane's entry of the 2^(j/32) table
  double noise;
  int lane, k;
};

// Tile column J of the current neighbourhood (factorisation) together with tile column J of
// the next neighbourhood (assembly); recursion over J keeps every shape quantity constexpr.
template <int T, int KP, int F, int J>
__device__ __forceinline__ void process_column(const ColCtx& x, double (&l0)[T][T],
                                               double (&l1)[T][T], double (&dinv0)[T],
                                               double (&dinv1)[T], bool& ok, double& prev_pinv,
                                               EvalState<F> (&ev)[3], double& c_last0,
                                               double& c_last1) {
  constexpr int JE = (KP + 7) / 8;
  constexpr int ZF = (KP - 3) >> 3;
#ifdef MGP_DEBUG_SKIP_EVAL
  constexpr int CI_ = 0;
#else
  constexpr int CI_ = (J < JE) ? chunk_iters<KP>(J) : 0;
#endif
  (void)CI_;
  constexpr int NEV = chunk_begin<KP>(JE);  // element evaluations (table rows) per neighbourhood
  constexpr int NC_ = (KP - 8 * J) >= 8 ? 8 : ((KP - 8 * J) > 0 ? (KP - 8 * J) : 1);
  constexpr int ncols = (J < JE) ? NC_ : 0;
  // entry S starts at column step S: it must belong to a tile column that has been handed
  // over already (chunk_begin(J) >= 8 J for every J) and finish before the last step
  static_assert(chunk_begin<KP>(J < JE ? J : 0) >= 8 * (J < JE ? J : 0), "entry starts too early");
  static_assert(NEV + 2 <= KP, "the last entries must drain before the factorisation ends");
  const int lane = x.lane, rho = lane >> 2, q = lane & 3, qb = lane & ~3;
  double* tiles = x.tiles;
  // ---- current neighbourhood: pick up tile column J, then hand its cells over ----------
  double c[T][2];
#pragma unroll
  for (int I = J; I < T; ++I) {
    const double2 v = *reinterpret_cast<const double2*>(tiles + frag_off(I, J, rho, q));
    c[I][0] = v.x;
    c[I][1] = v.y;
  }
  __syncwarp();
#pragma unroll
  for (int I = (ZF > J ? ZF : J); I < T; ++I)
    *reinterpret_cast<double2*>(tiles + tile_base(I, J) + 2 * lane) = make_double2(0.0, 0.0);
  if (J < ZF)  // keep the never-assembled upper half of the diagonal tile finite (0 * NaN)
    *reinterpret_cast<double2*>(tiles + tile_base(J, J) + 2 * lane) = make_double2(0.0, 0.0);
  __syncwarp();
  // ---- next neighbourhood: padding / target cells of its tile column J ------------------
  if (lane < 8) {
    const int j = 8 * J + lane;
    if (j < x.k) tiles[elem_off(KP + 1, j)] = -x.ys[j];  // target row holds -y
    else if (j <= KP) tiles[elem_off(j, j)] = -1.0;      // identity padding, Kout
  }
  // ---- current neighbourhood: left-looking update ---------------------------------------
#pragma unroll
  for (int P = 0; P < J; ++P) {
    if (P < JE) {
      const double b0 = l0[J][P] * dinv0[P], b1 = l1[J][P] * dinv1[P];
      // even slices of every tile first, then the odd ones: back-to-back DMMAs on the same
      // accumulator would serialise on the DMMA latency
#pragma unroll
      for (int I = J; I < T; ++I) dmma_acc(c[I][0], c[I][1], l0[I][P], b0);
#pragma unroll
      for (int I = J; I < T; ++I) dmma_acc(c[I][0], c[I][1], l1[I][P], b1);
    }
  }
#pragma unroll
  for (int I = J; I < T; ++I) {
    c[I][0] = -c[I][0];
    c[I][1] = -c[I][1];
  }
#ifdef MGP_DEBUG_SKIP_CHAIN
  if (false) {
#else
  if (J < JE) {
#endif
    // in-tile LDL^T on the diagonal tile and on an identity tile (-> M), with the next
    // neighbourhood's element evaluations woven into the column steps
    double v0 = (rho == 2 * q) ? 1.0 : 0.0, v1 = (rho == 2 * q + 1) ? 1.0 : 0.0;
    double di0 = 0.0, di1 = 0.0;
#define MGP_STEP(JJ)                                                                          \
  if (JJ < ncols)                                                                             \
    column_step<F, 8 * J + JJ, NEV>(c[J][0], c[J][1], v0, v1, di0, di1, ok, prev_pinv, ev,    \
                                    x.etab, lane, q, qb, tiles, x.pts2, x.tab64, x.noise);
    MGP_STEP(0) MGP_STEP(1) MGP_STEP(2) MGP_STEP(3)
    MGP_STEP(4) MGP_STEP(5) MGP_STEP(6) MGP_STEP(7)
#undef MGP_STEP
    const bool keep = (ncols == 8) || (q < 2);
    dinv0[J] = sel_d(keep, di0, 0.0);
    dinv1[J] = sel_d(keep, di1, 0.0);
    if (J + 1 < T) {
      const int srcE = 8 * q + (lane >> 3), par = (lane >> 2) & 1;
      const double e0 = shfl_d(v0, srcE), e1 = shfl_d(v1, srcE);
      const double o0 = shfl_d(v0, srcE + 4), o1 = shfl_d(v1, srcE + 4);
      const double bm0 = sel_d(par, e1, e0), bm1 = sel_d(par, o1, o0);
#pragma unroll
      for (int I = J + 1; I < T; ++I) {
        l0[I][J] = 0.0;
        l1[I][J] = 0.0;
        dmma_acc(l0[I][J], l1[I][J], c[I][0], bm0);
      }
#pragma unroll
      for (int I = J + 1; I < T; ++I) dmma_acc(l0[I][J], l1[I][J], c[I][1], bm1);
    }
  }
  if constexpr (J == T - 1) {
    c_last0 = c[T - 1][0];
    c_last1 = c[T - 1][1];
  } else {
    process_column<T, KP, F, J + 1>(x, l0, l1, dinv0, dinv1, ok, prev_pinv, ev, c_last0,
                                    c_last1);
  }
}

template <int T, int KP, int F>
__global__ void __launch_bounds__(PIPE_WARPS * 32, 3)
    fused_pipe_kernel(const TileArgs a, size_t warp_doubles) {
  extern __shared__ double smem[];
  constexpr int NT = T * (T + 1) / 2;
  constexpr int JE = (KP + 7) / 8;           // tile columns that carry eliminated columns
  constexpr int ROWS = chunk_begin<KP>(JE);  // table rows (of 32 entries)
  constexpr int ZF = (KP - 3) >> 3;          // first tile row that can hold padding
  constexpr int LI = KP & 7;                 // row of the cross-covariance inside tile T-1
  static_assert((KP >> 3) == T - 1, "Schur block must live in the last tile");
  static_assert(LI + 1 < 8, "target row must live in the last tile");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = a.k;

  const double tab64 = a.exp_tab[lane];
  unsigned* etab = (unsigned*)smem;
  double* wbase = smem + (ROWS + 1) * 16 + (size_t)warp * warp_doubles;
  double* tiles = wbase;              // NT*64 image + 2 scratch doubles
  double* pts_buf = tiles + NT * 64 + 2;  // 2 x (k+1) points (x,y); point k = query
  const int pts_doubles = 2 * (KP + 2);
  double* ys_buf = pts_buf + 2 * pts_doubles;  // 2 x k targets
  const int ys_doubles = KP + 2;

  // dummy entries write -1-ish values to the scratch cell behind the image
  for (int e = threadIdx.x; e < (ROWS + 1) * 32; e += blockDim.x)
    etab[e] = (unsigned)(NT * 64) << 16;
  __syncthreads();
  // Tile column J's chunk lists its cells ROW-major: rows 8J..k-1 of K (columns 8J..min(8J+7,
  // row)), then the cross-covariance row (tile row KP, point k).  A warp then stores runs of up
  // to 8 consecutive columns of consecutive rows (2-way bank conflicts instead of the 16-way a
  // column-major order gives) and its point loads are multicasts of <= 4 + 8 addresses.
  for (int J = warp; J < JE; J += PIPE_WARPS) {
    int cb = 0;  // first table row of this chunk
#pragma unroll
    for (int c = 0; c < JE; ++c)
      if (c < J) cb += chunk_iters<KP>(c);
    const int c0 = 8 * J, ncol = min(8, k - c0);
    if (ncol <= 0) continue;
    // rows c0 .. c0+ncol-1 form a triangle (row i has i-c0+1 cells), later rows have ncol cells
    const int tri = ncol * (ncol + 1) / 2;
    const int full_rows = k - (c0 + ncol);            // K rows below the triangle
    const int total = tri + (full_rows + 1) * ncol;   // + the cross-covariance row
    for (int e = lane; e < total; e += 32) {
      int ti, pi, j;
      if (e < tri) {
        int rr = 0;
        while ((rr + 1) * (rr + 2) / 2 <= e) ++rr;
        ti = pi = c0 + rr;
        j = c0 + e - rr * (rr + 1) / 2;
      } else {
        const int f = e - tri, rr = f / ncol;
        j = c0 + f - rr * ncol;
        if (rr < full_rows) {
          ti = pi = c0 + ncol + rr;
        } else {
          ti = KP;
          pi = k;
        }
      }
      etab[cb * 32 + e] = ((unsigned)elem_off(ti, j) << 16) | ((unsigned)pi << 8) | (unsigned)j;
    }
  }
  __syncthreads();

  const long long wglobal = (long long)blockIdx.x * PIPE_WARPS + warp;
  const long long wstride = (long long)gridDim.x * PIPE_WARPS;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double cs0 = a.coord_scale[0], cs1 = a.coord_scale[1];
  const double noise = a.noise;

  // lane l stages points l and l+32 (point k is the query)
  auto load_src = [&](long long row, int i) -> long long {
    if (row >= a.b || i > k) return -1;
    if (i == k) return a.query_idx ? a.query_idx[row] : row;
    return a.nn_idx[row * k + i];
  };
  auto issue_rows = [&](int buf, int i, long long src) {
    if (src < 0) return;
    const double* px = ((i == k) ? a.query_x : a.train_x) + src * 2;
    double* dst = pts_buf + buf * pts_doubles + i * 2;
    cp_async8(dst, px);
    cp_async8(dst + 1, px + 1);
    if (i < k) cp_async8(ys_buf + buf * ys_doubles + i, a.train_y + src);
  };
  auto scale_pts = [&](int buf) {
    double2* p2 = reinterpret_cast<double2*>(pts_buf + buf * pts_doubles);
    for (int i = lane; i <= k; i += 32) {
      double2 v = p2[i];
      v.x *= cs0;
      v.y *= cs1;
      p2[i] = v;
    }
  };
  // padding / augmented cells of tile column J for the neighbourhood whose targets are `ys`
  auto specials = [&](int J, const double* ys) {
    if (lane < 8) {
      const int j = 8 * J + lane;
      if (j < k) tiles[elem_off(KP + 1, j)] = -ys[j];          // target row holds -y
      else if (j <= KP) tiles[elem_off(j, j)] = -1.0;          // identity padding, Kout
    }
  };

  // ---- prologue: neighbourhood 0 is assembled the plain way -------------------------------
  long long s0 = load_src(wglobal, lane), s1 = load_src(wglobal, lane + 32);
  issue_rows(0, lane, s0);
  issue_rows(0, lane + 32, s1);
  cp_async_commit();
  s0 = load_src(wglobal + wstride, lane);
  s1 = load_src(wglobal + wstride, lane + 32);
  cp_async_wait_all();
  __syncwarp();
  issue_rows(1, lane, s0);
  issue_rows(1, lane + 32, s1);
  cp_async_commit();
  s0 = load_src(wglobal + 2 * wstride, lane);
  s1 = load_src(wglobal + 2 * wstride, lane + 32);
  scale_pts(0);
  for (int e = lane; e < (NT * 64 - tile_base(ZF, 0)) / 2; e += 32)
    reinterpret_cast<double2*>(tiles + tile_base(ZF, 0))[e] = make_double2(0.0, 0.0);
  for (int J = 0; J < ZF; ++J)
    reinterpret_cast<double2*>(tiles + tile_base(J, J))[lane] = make_double2(0.0, 0.0);
  __syncwarp();
  for (int J = 0; J < T; ++J) specials(J, ys_buf);
  assemble_flat<F>(tiles, reinterpret_cast<const double2*>(pts_buf), etab, tab64, ROWS, lane,
                   noise);
  __syncwarp();

  int nb = 1;  // buffer that holds the NEXT neighbourhood's points
  for (long long row = wglobal; row < a.b; row += wstride, nb ^= 1) {
    cp_async_wait_all();  // points / targets of neighbourhood row+stride have landed in nb
    __syncwarp();
    issue_rows(nb ^ 1, lane, s0);  // row + 2*stride -> the buffer consumed one round ago
    issue_rows(nb ^ 1, lane + 32, s1);
    cp_async_commit();
    s0 = load_src(row + 3 * wstride, lane);
    s1 = load_src(row + 3 * wstride, lane + 32);
    scale_pts(nb);
    __syncwarp();
    const double2* pts2 = reinterpret_cast<const double2*>(pts_buf + nb * pts_doubles);
    const double* ys = ys_buf + nb * ys_doubles;

    double l0[T][T], l1[T][T];  // finished tiles U = L D in accumulator layout, [I][P], I > P
    double dinv0[T], dinv1[T];
    bool ok = true;
    double prev_pinv = 1.0, c_last0 = 0.0, c_last1 = 0.0;
    ColCtx ctx;
    ctx.tiles = tiles;
    ctx.pts2 = pts2;
    ctx.ys = ys;
    ctx.etab = etab;
    ctx.tab64 = tab64;
    ctx.noise = noise;
    ctx.lane = lane;
    ctx.k = k;
    EvalState<F> ev[3];
    process_column<T, KP, F, 0>(ctx, l0, l1, dinv0, dinv1, ok, prev_pinv, ev, c_last0, c_last1);
    {
      // Schur complement straight from the accumulator registers of the last tile
      const double cv = (LI & 1) ? c_last1 : c_last0;
      if (lane == LI * 4 + LI / 2 && a.var) a.var[row] = ok ? a.scale * cv : nan;
      if (lane == (LI + 1) * 4 + LI / 2 && a.mean) a.mean[row] = ok ? -cv : nan;
      const double cy = ((LI + 1) & 1) ? c_last1 : c_last0;
      if (lane == (LI + 1) * 4 + (LI + 1) / 2 && a.yky) a.yky[row] = ok ? -cy : nan;
      if (lane == 0 && a.status) a.status[row] = ok ? 0 : 1;
    }
    __syncwarp();  // the image now holds neighbourhood row+stride, fully assembled
  }
  cp_async_wait_all();
}

}  // namespace

int fused_pipe_supported(const mgp_problem* p, const Model& model) {
  if (p->d != 2 || p->r != 1 || p->noise_bk || p->coeffs || !p->train_y) return 0;
  if (model.metric_id != MGP_METRIC_L2) return 0;
  if (model.kernel_id != MGP_KERNEL_MATERN_05 && model.kernel_id != MGP_KERNEL_MATERN_15 &&
      model.kernel_id != MGP_KERNEL_MATERN_25)
    return 0;
  const int kp = (p->k + 3) & ~3;
  return kp == 52;
}

int launch_fused_pipe(const mgp_problem* p, const Model& model, cudaStream_t stream) {
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  constexpr int T = 7, KP = 52;
  constexpr int NT = T * (T + 1) / 2;
  constexpr int ROWS = chunk_begin<KP>((KP + 7) / 8);
  const size_t warp_doubles = (size_t)NT * 64 + 2 + 2 * (size_t)(2 * (KP + 2)) + 2 * (size_t)(KP + 2);
  const size_t smem =
      ((size_t)(ROWS + 1) * 16 + warp_doubles * PIPE_WARPS) * sizeof(double);
  MGP_REQUIRE(smem <= (size_t)max_smem_optin(), MGP_ERR_UNSUPPORTED,
              "pipe kernel shared memory %zu too large", smem);
  long long blocks = (p->b + PIPE_WARPS - 1) / PIPE_WARPS;
  const long long cap = (long long)sm_count() * 3;
  if (blocks > cap) blocks = cap;
#define MGP_PIPE(FF)                                                                          \
  case FF:                                                                                    \
    cudaFuncSetAttribute(fused_pipe_kernel<T, KP, FF>,                                        \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    fused_pipe_kernel<T, KP, FF><<<(unsigned)blocks, PIPE_WARPS * 32, smem, stream>>>(        \
        a, warp_doubles);                                                                     \
    break;
  switch (a.formula) {
    MGP_PIPE(F_M05)
    MGP_PIPE(F_M15)
    MGP_PIPE(F_M25)
    default:
      set_error("pipe variant does not support formula %d", a.formula);
      return MGP_ERR_UNSUPPORTED;
  }
#undef MGP_PIPE
  return check_launch("fused_pipe_kernel");
}

}  // namespace mgp
