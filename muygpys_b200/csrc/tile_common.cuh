// Device helpers shared by the register-tile fused kernels (fused_tile*.cu, fused_col*.cu).
#pragma once

#include "common.cuh"

namespace mgp {
namespace {

constexpr int TILE_WARPS = 4;        // warps (neighbourhoods in flight) per CTA
constexpr int TILE_MAX_D = 8;
constexpr int EXP_TABLE = 32;  // 2^(j/32), one entry per lane, looked up with a shuffle


__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// same instruction, not volatile: the scheduler may move it across other instructions
__device__ __forceinline__ void dmma_free(double& c0, double& c1, double a, double b) {
#ifdef MGP_DBG_NODMMA
  c0 += a;
  return;
#endif
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double rsqrt_seed(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RSQ64H, ~20 bits
  return r;
}

// 1/sqrt(p), one third-order step from the 20-bit seed (error ~ 2^-62)
__device__ __forceinline__ double rsqrt_fast(double p) {
  const double r = rsqrt_seed(p);
  const double t = p * r;
  const double e = fma(-t, r, 1.0);
  const double v = e * fma(e, 0.375, 0.5);
  return fma(r, v, r);
}

// sqrt(x) for x >= 0 (0 and denormals -> 0); the range test is an integer compare on the
// high word so it stays off the FP64 pipe
__device__ __forceinline__ double sqrt_fast(double x) {
  const double r = rsqrt_seed(x);
  const double g = x * r;
  const double e = fma(-g, r, 1.0);
  const double v = e * fma(e, 0.375, 0.5);
  const double s = fma(g, v, g);
  return (__double2hiint(x) > 0x03c00000) ? s : 0.0;
}

// exp(-s), s >= 0: 32-entry table of 2^(j/32) + degree-6 polynomial.  The table lives in a
// REGISTER of each lane (`tabreg` = entry `lane`) and is read with a warp shuffle: the
// data-dependent shared-memory lookup it replaces cost ~7 bank-conflicted wavefronts per
// warp and made the assembly phase shared-memory-bound.  Every lane of the warp must call
// this together.  |abs error| <= ~2.2e-16 (validated against long double on the host).
__device__ __forceinline__ double exp_neg(double s, double tabreg) {
  const double LOG2E = 1.4426950408889634;
  const double MAGIC = 211106232532992.0;  // 1.5 * 2^47: ulp = 2^-5
  const double t = fma(s, -LOG2E, MAGIC);
  const int ki = __double2loint(t);
  const double tabv = __shfl_sync(0xffffffffu, tabreg, ki & (EXP_TABLE - 1));
  const double tr = t - MAGIC;
  const double g = fma(s, -LOG2E, -tr);
  double p = fma(g, 0.00015403530393381608, 0.0013333558146428443);
  p = fma(g, p, 0.009618129107628477);
  p = fma(g, p, 0.05550410866482158);
  p = fma(g, p, 0.2402265069591007);
  p = fma(g, p, 0.6931471805599453);
  p = fma(g, p, 1.0);
  const double res = tabv * p;
  const int n = ki >> 5;
  const double out = __hiloint2double(__double2hiint(res) + (n << 20), __double2loint(res));
  return (__double2hiint(s) < 0x4085e000) ? out : 0.0;  // s < 700
}

struct TileArgs {
  const double* train_x;
  const double* query_x;
  const int64_t* query_idx;
  const int64_t* nn_idx;
  const double* train_y;
  const double* noise_bk;
  double* mean;
  double* var;
  double* yky;
  double* coeffs;
  int32_t* status;
  long long b;
  int k, kp, d, r;
  int n_elem;      // k(k+1)/2 + k table entries
  double noise, scale;
  int formula;     // Formula below
  double coord_scale[MGP_MAX_ANISO_DIM];  // per-feature multiplier folded into the coordinates
  int gram;        // d > TILE_MAX_D: distances from DMMA Gram tiles (gram.cuh), rows not staged
  int aniso;       // coord_scale differs per feature
  double post_scale;               // F2 metric with Matern: s = post_scale * u2
  int kernel_id;
  // 2^(j/32), correctly rounded on the host.  Lives in the kernel parameter space (not in a
  // __constant__ symbol): parameters reach whichever device the launch targets, a symbol
  // uploaded once per process only reached the device that was current at the time.
  double exp_tab[EXP_TABLE];
};

// covariance as a function of u2 = sum of squared prescaled coordinate differences
enum Formula {
  F_M05 = 0,     // exp(-sqrt(u2))
  F_M15 = 1,     // (1+s) exp(-s),          s = sqrt(u2)   (sqrt(3)/l folded into coordinates)
  F_M25 = 2,     // (1+s+u2/3) exp(-s),     s = sqrt(u2)   (sqrt(5)/l folded)
  F_GAUSS = 3,   // exp(-u2/2): RBF on F2, Matern nu=inf on l2
  F_RBF_L2 = 4,  // exp(-sqrt(u2)/2): RBF handed l2 distances (reference quirk, rbf.py:74-76)
  F_F2_ANY = 5   // any other kernel fed the squared metric: argument post_scale*u2
};

// NEGATED covariance (the tile image holds N = -A)
template <int F>
__device__ __forceinline__ double neg_cov(double u2, double tab64, double post_scale,
                                          int kernel_id) {
  if (F == F_M05) return -exp_neg(sqrt_fast(u2), tab64);
  if (F == F_M15) {
    const double s = sqrt_fast(u2);
    return (-1.0 - s) * exp_neg(s, tab64);
  }
  if (F == F_M25) {
    const double s = sqrt_fast(u2);
    return fma(u2, -(1.0 / 3.0), -1.0 - s) * exp_neg(s, tab64);
  }
  if (F == F_GAUSS) return -exp_neg(0.5 * u2, tab64);
  if (F == F_RBF_L2) return -exp_neg(0.5 * sqrt_fast(u2), tab64);
  const double s = post_scale * u2;
  switch (kernel_id) {
    case MGP_KERNEL_MATERN_05:
      return -exp_neg(s, tab64);
    case MGP_KERNEL_MATERN_15:
      return (-1.0 - s) * exp_neg(s, tab64);
    case MGP_KERNEL_MATERN_25:
      return fma(s * s, -(1.0 / 3.0), -1.0 - s) * exp_neg(s, tab64);
    default:  // Matern inf on the squared metric
      return -exp_neg(0.5 * s * s, tab64);
  }
}

template <int D>
__device__ __forceinline__ double sqdist(const double* __restrict__ pts, int pi, int j, int d) {
  if (D == 1) {
    const double df = pts[pi] - pts[j];
    return df * df;
  }
  if (D == 2) {
    const double2 p = reinterpret_cast<const double2*>(pts)[pi];
    const double2 c = reinterpret_cast<const double2*>(pts)[j];
    const double dx = p.x - c.x, dy = p.y - c.y;
    return fma(dy, dy, dx * dx);
  }
  if (D == 3) {
    const double dx = pts[3 * pi] - pts[3 * j], dy = pts[3 * pi + 1] - pts[3 * j + 1],
                 dz = pts[3 * pi + 2] - pts[3 * j + 2];
    return fma(dz, dz, fma(dy, dy, dx * dx));
  }
  double u2 = 0.0;
  for (int f = 0; f < d; ++f) {
    const double df = pts[pi * d + f] - pts[j * d + f];
    u2 = fma(df, df, u2);
  }
  return u2;
}

__host__ __device__ __forceinline__ int tile_base(int I, int J) {
  return ((I * (I + 1)) / 2 + J) * 64;
}
// Tile image: tile-major, row-major inside a tile, with the rows of odd tile columns swapped
// pairwise (row ^ 1).  A warp storing 32 consecutive columns of one matrix row then alternates
// between the two 16-bank windows instead of hitting one of them four times over.
__device__ __forceinline__ int elem_off(int i, int j) {
  const int J = j >> 3;
  return tile_base(i >> 3, J) + ((((i & 7) ^ (J & 1))) << 3) + (j & 7);
}
// accumulator-layout fragment (row rho, columns 2q, 2q+1) of tile (I,J)
__device__ __forceinline__ int frag_off(int I, int J, int rho, int q) {
  return tile_base(I, J) + ((rho ^ (J & 1)) << 3) + 2 * q;
}

__device__ __forceinline__ double sel_d(bool p, double a, double b) { return p ? a : b; }

// 1/p: MUFU.RCP64H seed (~20 bits) + one third-order step (error ~ 2^-60), depth 3
__device__ __forceinline__ double rcp_fast(double p) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  const double e = fma(-p, r, 1.0);
  const double q1 = r * e;
  const double w = 1.0 + e;
  return fma(q1, w, r);
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

static inline int tiles_needed(int k, int r) {
  const int kp = (k + 3) & ~3;
  return (kp + 1 + r + 7) / 8;
}

// Host: pack an mgp_problem + Model into TileArgs.
static inline int fill_tile_args(const mgp_problem* p, const Model& model, TileArgs& a) {
  for (int j = 0; j < EXP_TABLE; ++j) a.exp_tab[j] = (double)exp2l((long double)j / EXP_TABLE);
  a.train_x = p->train_x;
  a.query_x = p->query_x;
  a.query_idx = p->query_idx;
  a.nn_idx = p->nn_idx;
  a.train_y = p->train_y;
  a.noise_bk = p->noise_bk;
  a.mean = p->mean;
  a.var = p->var;
  a.yky = p->yky;
  a.coeffs = p->coeffs;
  a.status = p->status;
  a.b = p->b;
  a.k = p->k;
  a.kp = (p->k + 3) & ~3;
  a.d = p->d;
  a.r = p->r;
  a.n_elem = p->k * (p->k + 1) / 2 + p->k;
  a.noise = p->noise;
  a.scale = p->scale;
  a.kernel_id = model.kernel_id;
  if (model.metric_id == MGP_METRIC_L2) {
    switch (model.kernel_id) {
      case MGP_KERNEL_MATERN_05: a.formula = F_M05; break;
      case MGP_KERNEL_MATERN_15: a.formula = F_M15; break;
      case MGP_KERNEL_MATERN_25: a.formula = F_M25; break;
      case MGP_KERNEL_MATERN_INF: a.formula = F_GAUSS; break;
      default: a.formula = F_RBF_L2;
    }
  } else {
    a.formula = (model.kernel_id == MGP_KERNEL_RBF) ? F_GAUSS : F_F2_ANY;
  }
  // fold length scale (and the Matern sqrt(2 nu) factor for l2) into the coordinates
  double kconst = 1.0;
  if (model.kernel_id == MGP_KERNEL_MATERN_15) kconst = 1.7320508075688772;
  if (model.kernel_id == MGP_KERNEL_MATERN_25) kconst = 2.23606797749979;
  a.post_scale = 1.0;
  a.gram = p->d > TILE_MAX_D;
  a.aniso = model.aniso;
  for (int f = 0; f < MGP_MAX_ANISO_DIM; ++f) a.coord_scale[f] = 1.0;
  for (int f = 0; f < p->d && f < MGP_MAX_ANISO_DIM; ++f) {
    double inv = model.aniso ? model.inv_ls_vec[f]
                             : (model.metric_id == MGP_METRIC_L2 ? model.inv_ls
                                                                  : sqrt(model.inv_ls));
    a.coord_scale[f] = (model.metric_id == MGP_METRIC_L2) ? inv * kconst : inv;
  }
  if (model.metric_id == MGP_METRIC_F2) a.post_scale = kconst;

  return MGP_OK;
}

}  // namespace
}  // namespace mgp
