// Host-buffer entry point of the fused path: neighbour indices (and optionally the results)
// live in HOST memory, as they do behind the reference's `regress_from_indices`
// (S/examples/from_indices.py:22-63, numpy arrays in / numpy arrays out).
//
// The batch is cut into chunks; chunk c+1 is uploaded on one side stream while chunk c runs
// in the fused kernel on the other and chunk c-1's results travel back, so the end-to-end rate
// is max(PCIe, compute) instead of their sum.  Chunk boundaries sit on multiples of one full
// wave of the kernel (12 neighbourhoods in flight per SM) and grow geometrically from a small
// first chunk: the only transfer that is not hidden is the first upload, and a chunk may grow
// by about kernel time / copy time per step without stalling the pipeline.  Everything is
// ordered after the work already queued on `stream` and joined back into it.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mgp {
namespace {

struct SideStreams {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr;
  cudaEvent_t join[2] = {nullptr, nullptr};
  bool ready = false;
};

std::mutex g_side_mutex;
SideStreams g_side[64];  // per device

int side_streams(SideStreams** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  MGP_REQUIRE(e == cudaSuccess && dev >= 0 && dev < 64, MGP_ERR_CUDA, "cudaGetDevice: %s",
              cudaGetErrorString(e));
  SideStreams& ss = g_side[dev];  // (the caller holds g_side_mutex)
  if (!ss.ready) {
    for (int i = 0; i < 2; ++i) {
      e = cudaStreamCreateWithFlags(&ss.s[i], cudaStreamNonBlocking);
      MGP_REQUIRE(e == cudaSuccess, MGP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
      e = cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming);
      MGP_REQUIRE(e == cudaSuccess, MGP_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    }
    e = cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming);
    MGP_REQUIRE(e == cudaSuccess, MGP_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    ss.ready = true;
  }
  *out = &ss;
  return MGP_OK;
}

// chunk boundaries: multiples of `wave` rows, sizes 2, 2, 3, 5, 7, 10, ... waves (x1.4)
std::vector<long long> chunk_bounds(long long b, long long wave) {
  // dev switches for tuning the schedule
  static const double first = getenv("MGP_PIPE_FIRST") ? atof(getenv("MGP_PIPE_FIRST")) : 2.0;
  static const double growth = getenv("MGP_PIPE_GROWTH") ? atof(getenv("MGP_PIPE_GROWTH")) : 1.4;
  std::vector<long long> bounds{0};
  double w = first;
  while (bounds.back() < b) {
    const long long sz = (long long)w < 1 ? 1 : (long long)w;
    bounds.push_back(bounds.back() + sz * wave);
    w *= growth;
  }
  bounds.back() = b;
  const size_t m = bounds.size();
  if (m > 2 && bounds[m - 1] - bounds[m - 2] < (bounds[m - 2] - bounds[m - 3]) / 4) {
    bounds.erase(bounds.end() - 2);  // fold a short tail into the previous chunk
  }
  return bounds;
}

}  // namespace

int validate_problem(const mgp_problem* p);
int fused_wave_per_sm(const mgp_problem* p);

}  // namespace mgp

using namespace mgp;

extern "C" int mgp_fused_posterior_host(const mgp_problem* p, const int64_t* nn_idx_host,
                                        const int64_t* query_idx_host, double* mean_host,
                                        double* var_host, void* ws, size_t ws_bytes,
                                        void* stream) {
  MGP_REQUIRE(p != nullptr, MGP_ERR_BAD_ARG, "null problem");
  MGP_REQUIRE(nn_idx_host != nullptr && p->nn_idx != nullptr, MGP_ERR_BAD_ARG,
              "nn_idx_host and the device staging buffer p->nn_idx are required");
  MGP_REQUIRE(!query_idx_host || p->query_idx, MGP_ERR_BAD_ARG,
              "query_idx_host needs the device staging buffer p->query_idx");
  MGP_REQUIRE(!mean_host || p->mean, MGP_ERR_BAD_ARG, "mean_host needs the device buffer p->mean");
  MGP_REQUIRE(!var_host || p->var, MGP_ERR_BAD_ARG, "var_host needs the device buffer p->var");
  int rc = validate_problem(p);
  if (rc != MGP_OK) return rc;
  if (p->b == 0) return MGP_OK;
  // one enqueue at a time per process: the fork/join events are shared
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStreams* ss = nullptr;
  rc = side_streams(&ss);
  if (rc != MGP_OK) return rc;
  cudaStream_t main_stream = (cudaStream_t)stream;

  // Whatever happens after the fork, the side streams are joined back into `stream` before
  // returning: chunks already enqueued must stay ordered before the caller's later work (torch
  // may reuse the staging and output tensors as soon as this call returns).
  auto join = [&]() {
    for (int i = 0; i < 2; ++i)
      if (cudaEventRecord(ss->join[i], ss->s[i]) == cudaSuccess)
        cudaStreamWaitEvent(main_stream, ss->join[i], 0);
  };
  bool forked = false;
#define MGP_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      set_error("%s: %s", #call, cudaGetErrorString(e_));                                 \
      if (forked) join();                                                                 \
      return MGP_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

  // fork: the side streams start after everything already queued on `stream`
  MGP_CUDA(cudaEventRecord(ss->fork, main_stream));
  for (int i = 0; i < 2; ++i) MGP_CUDA(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
  forked = true;

  // one wave of the kernel this shape takes: neighbourhoods in flight per SM x SMs
  const long long wave = (long long)fused_wave_per_sm(p) * sm_count();
  const std::vector<long long> bounds = chunk_bounds(p->b, wave);
  const long long k = p->k, r = p->r;
  for (size_t c = 0; c + 1 < bounds.size(); ++c) {
    const long long lo = bounds[c], hi = bounds[c + 1], rows = hi - lo;
    if (rows <= 0) continue;
    cudaStream_t s = ss->s[c & 1];
    int64_t* nn_dev = const_cast<int64_t*>(p->nn_idx) + lo * k;
    MGP_CUDA(cudaMemcpyAsync(nn_dev, nn_idx_host + lo * k, (size_t)rows * k * sizeof(int64_t),
                             cudaMemcpyHostToDevice, s));
    if (query_idx_host)
      MGP_CUDA(cudaMemcpyAsync(const_cast<int64_t*>(p->query_idx) + lo, query_idx_host + lo,
                               (size_t)rows * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    mgp_problem sub = *p;
    sub.b = rows;
    sub.nn_idx = nn_dev;
    if (p->query_idx) {
      sub.query_idx = p->query_idx + lo;
    } else {
      // rows lo..hi-1 of query_x are the queries of this chunk
      sub.query_x = p->query_x + lo * p->d;
      sub.t = p->t - lo;
    }
    if (p->noise_bk) sub.noise_bk = p->noise_bk + lo * k;
    if (p->mean) sub.mean = p->mean + lo * r;
    if (p->var) sub.var = p->var + lo;
    if (p->yky) sub.yky = p->yky + lo;
    if (p->coeffs) sub.coeffs = p->coeffs + lo * k * r;
    if (p->status) sub.status = p->status + lo;
    rc = mgp_fused_posterior(&sub, ws, ws_bytes, (void*)s);
    if (rc != MGP_OK) {
      join();
      return rc;
    }
    if (mean_host)
      MGP_CUDA(cudaMemcpyAsync(mean_host + lo * r, sub.mean, (size_t)rows * r * sizeof(double),
                               cudaMemcpyDeviceToHost, s));
    if (var_host)
      MGP_CUDA(cudaMemcpyAsync(var_host + lo, sub.var, (size_t)rows * sizeof(double),
                               cudaMemcpyDeviceToHost, s));
  }
  // join: later work on `stream` sees every chunk
  for (int i = 0; i < 2; ++i) {
    MGP_CUDA(cudaEventRecord(ss->join[i], ss->s[i]));
    MGP_CUDA(cudaStreamWaitEvent(main_stream, ss->join[i], 0));
  }
#undef MGP_CUDA
  return MGP_OK;
}
