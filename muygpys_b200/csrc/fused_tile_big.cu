// K1 (tile variant), d <= 8, T > TILE_T_SPLIT (k > 68): instantiations of
// fused_tile_kernel<T, SMEM_L = true, GRAM = false>, in their own translation unit so that they
// compile in parallel with the smaller ones of fused_tile.cu.
#include "fused_tile_kernel.cuh"

namespace mgp {

int launch_fused_tile_big(const mgp_problem* p, const Model& model, int T,
                          size_t shared_doubles, size_t warp_doubles, cudaStream_t stream) {
  TileArgs a;  // filled here again: the exp table of THIS translation unit must be uploaded
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  return launch_tile_instance<false, TILE_T_SPLIT + 1, 16>(a, T, p->b, shared_doubles,
                                                           warp_doubles, stream);
}

}  // namespace mgp
