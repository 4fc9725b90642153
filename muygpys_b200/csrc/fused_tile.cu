// K1 (tile variant) -- the fast path of the fused neighbourhood kernel.
//
// One WARP owns one neighbourhood.  The augmented matrix (nbhd_smem.cuh) is cut
// into 8x8 tiles; its lower triangle lives in REGISTERS as mma.sync.m8n8k4.f64
// fragments and is factorised tile column by tile column (left-looking):
//
//   for each tile column J:
//     N[I][J] (+)= L[I][P] * L[J][P]^T  for all finished P < J   -- FP64 DMMA
//     C = -N ; in-tile Cholesky of C[J][J] and triangular solve of the tiles
//     below it, 8 columns, lane-parallel with quad/warp shuffles
//     re-layout C[I][J] (accumulator layout) -> A/B fragment layout for later J
//
// tcgen05 has no FP64 type, so on sm_100a the FP64 tensor path is mma.sync DMMA;
// measured on B200 it issues at the same 64 FMA/clk/SM as DFMA (tools/fp64_probe.py)
// but needs 1/8 of the issue slots and no operand traffic, which is what lets
// the shuffles, selects and address arithmetic of the other phases overlap.
//
// Assembly phase: the k(k+1)/2 + k covariances are evaluated in a flat,
// perfectly balanced loop (1 element per lane per iteration, (i,j) from a
// shared-memory table) with hand-rolled exp / sqrt that cost 9 / 5 FP64 issue
// slots instead of libm's ~17 / ~8, and scattered (negated) into the tile
// layout in shared memory, from where each tile column is picked up with one
// LDS.128 per lane per tile.
//
// Supported shapes: roundup8(roundup4(k) + 1 + r) <= 128 (T <= 7 tile rows with the factor in
// registers, T <= 16 with the factor in shared memory), k <= 127.  d <= 8 assembles from staged
// coordinates; d > 8 takes its distances from DMMA Gram tiles (gram.cuh, instantiated in
// fused_tile_gram.cu).  Everything else takes the generic shared-memory kernel.  The kernel
// template itself is in fused_tile_kernel.cuh.
#include "fused_tile_kernel.cuh"

namespace mgp {

static const int g_gram_off = getenv("MGP_NO_GRAM") != nullptr;  // dev switch: generic kernel for d > 8
static int g_variant = 0;  // 0 auto (col > tile > generic), 1 generic, 2 tile, 3 col, 4 col (lane-parallel steps)

int fused_variant() { return g_variant; }

// the other instantiations: d > 8 in fused_tile_gram*.cu, T > TILE_T_SPLIT in fused_tile*_big.cu
int launch_fused_tile_gram(const mgp_problem* p, const Model& model, int T,
                           size_t shared_doubles, size_t warp_doubles, cudaStream_t stream);
int launch_fused_tile_gram_big(const mgp_problem* p, const Model& model, int T,
                               size_t shared_doubles, size_t warp_doubles, cudaStream_t stream);
int launch_fused_tile_big(const mgp_problem* p, const Model& model, int T,
                          size_t shared_doubles, size_t warp_doubles, cudaStream_t stream);

int fused_tile_supported(const mgp_problem* p, const Model& model) {
  (void)model;
  if (g_variant == 1) return 0;
  if (p->d > TILE_MAX_D && g_gram_off) return 0;
  if (p->k > 127) return 0;  // the row prefetch stages at most 128 points per warp
  return tiles_needed(p->k, p->r) <= 16;
}

int launch_fused_tile(const mgp_problem* p, const Model& model, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  (void)ws;
  (void)ws_bytes;
  TileArgs a;
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  const int T = tiles_needed(p->k, p->r);
  const int NT = T * (T + 1) / 2;
  // per warp: tile image + double-buffered coordinates and targets (cp.async prefetch)
  const int ds = a.gram ? 0 : p->d;
  const size_t warp_doubles = (size_t)NT * 64 + 2 + 2 * (size_t)((((p->k + 1) * ds) + 1) & ~1) +
                              2 * (size_t)(((p->k * p->r) + 1) & ~1);
  const size_t shared_doubles =
      (size_t)((((a.n_elem + 2) / 2) + 1) & ~1) + MGP_MAX_ANISO_DIM;
  if (a.gram)
    return T <= TILE_T_SPLIT
               ? launch_fused_tile_gram(p, model, T, shared_doubles, warp_doubles, stream)
               : launch_fused_tile_gram_big(p, model, T, shared_doubles, warp_doubles, stream);
  if (T > TILE_T_SPLIT)
    return launch_fused_tile_big(p, model, T, shared_doubles, warp_doubles, stream);
  return launch_tile_instance<false, 1, TILE_T_SPLIT>(a, T, p->b, shared_doubles, warp_doubles,
                                                      stream);
}

}  // namespace mgp

extern "C" int mgp_set_fused_variant(int32_t variant) {
  if (variant < 0 || variant > 4) {
    mgp::set_error("variant must be 0 (auto), 1 (generic), 2 (tile), 3 (column-direct) or 4 "
                   "(column-direct with lane-parallel column steps)");
    return MGP_ERR_BAD_ARG;
  }
  mgp::g_variant = variant;
  return MGP_OK;
}
