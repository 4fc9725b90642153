// K1 (tile variant), d > 8, T > TILE_T_SPLIT: instantiations of fused_tile_kernel<T, SMEM_L, GRAM = true>
// (distances from query-centred DMMA Gram tiles, gram.cuh).  Kept in their own translation unit
// so that they compile in parallel with the d <= 8 instantiations of fused_tile.cu.
#include "fused_tile_kernel.cuh"

namespace mgp {

int launch_fused_tile_gram_big(const mgp_problem* p, const Model& model, int T,
                           size_t shared_doubles, size_t warp_doubles, cudaStream_t stream) {
  TileArgs a;  // filled here again: the exp table of THIS translation unit must be uploaded
  const int rc = fill_tile_args(p, model, a);
  if (rc != MGP_OK) return rc;
  // one LDG.128 per lane per row needs even d and 16-byte aligned arrays
  if (p->d % 2 == 0 && ((uintptr_t)p->train_x % 16 == 0) && ((uintptr_t)p->query_x % 16 == 0))
    a.gram = 2;
  return launch_tile_instance<true, TILE_T_SPLIT + 1, 16>(a, T, p->b, shared_doubles, warp_doubles,
                                                     stream);
}

}  // namespace mgp
