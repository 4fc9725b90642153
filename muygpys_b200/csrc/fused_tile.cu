// K1 (tile variant) -- placeholder until the register-resident DMMA kernel lands.
#include "common.cuh"

namespace mgp {

int fused_tile_supported(const mgp_problem* p, const Model& model) {
  (void)p;
  (void)model;
  return 0;
}

int launch_fused_tile(const mgp_problem* p, const Model& model, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  (void)p;
  (void)model;
  (void)ws;
  (void)ws_bytes;
  (void)stream;
  set_error("tile variant not built");
  return MGP_ERR_UNSUPPORTED;
}

}  // namespace mgp
