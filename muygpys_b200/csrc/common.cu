// Host-side helpers: error strings, device queries, model parameter packing.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mgp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MGP_ERR_CUDA;
  }
  return MGP_OK;
}

// Device attributes are cached PER DEVICE (a process may drive several GPUs).
static int cached_attr(cudaDeviceAttr attr, int (&cache)[64], int fallback) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, attr, dev);
    cache[dev] = v > 0 ? v : fallback;
  }
  return cache[dev];
}

int sm_count() {
  static int cache[64];
  return cached_attr(cudaDevAttrMultiProcessorCount, cache, 148);
}

int max_smem_optin() {
  static int cache[64];
  return cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache, 227 * 1024);
}

int make_model(int kernel_id, int metric_id, int d, int length_scale_count,
               const double* ls, Model* out) {
  MGP_REQUIRE(kernel_id >= MGP_KERNEL_RBF && kernel_id <= MGP_KERNEL_MATERN_INF,
              MGP_ERR_BAD_ARG, "unknown kernel_id %d", kernel_id);
  MGP_REQUIRE(metric_id == MGP_METRIC_L2 || metric_id == MGP_METRIC_F2, MGP_ERR_BAD_ARG,
              "unknown metric_id %d", metric_id);
  MGP_REQUIRE(d >= 1, MGP_ERR_BAD_ARG, "feature count d=%d must be >= 1", d);
  MGP_REQUIRE(ls != nullptr && length_scale_count >= 1, MGP_ERR_BAD_ARG,
              "length_scale (host) is required");
  memset(out, 0, sizeof(*out));
  out->kernel_id = kernel_id;
  out->metric_id = metric_id;
  out->d = d;
  if (length_scale_count == 1) {
    MGP_REQUIRE(ls[0] > 0.0, MGP_ERR_BAD_ARG, "length scale %g must be positive", ls[0]);
    out->aniso = 0;
    out->inv_ls = (metric_id == MGP_METRIC_L2) ? 1.0 / ls[0] : 1.0 / (ls[0] * ls[0]);
  } else {
    // Anisotropy.__call__ raises when the trailing dimension and the number of
    // length scales disagree (S/gp/deformation/anisotropy.py:65-69).
    MGP_REQUIRE(length_scale_count == d, MGP_ERR_BAD_ARG,
                "Difference tensor with final dimension size of %d does not match %d length "
                "scales",
                d, length_scale_count);
    MGP_REQUIRE(d <= MGP_MAX_ANISO_DIM, MGP_ERR_UNSUPPORTED,
                "anisotropic deformation supports d <= %d features (got %d)", MGP_MAX_ANISO_DIM,
                d);
    out->aniso = 1;
    out->inv_ls = 1.0;
    for (int f = 0; f < d; ++f) {
      MGP_REQUIRE(ls[f] > 0.0, MGP_ERR_BAD_ARG, "length scale %g must be positive", ls[f]);
      out->inv_ls_vec[f] = 1.0 / ls[f];
    }
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" int mgp_version(void) { return MGP_VERSION; }
extern "C" const char* mgp_last_error(void) { return mgp::g_err; }
