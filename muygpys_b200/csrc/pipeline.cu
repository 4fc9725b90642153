// Host-buffer entry point of the fused path: neighbour indices (and optionally the results)
// live in HOST memory, as they do behind the reference's `regress_from_indices`
// (S/examples/from_indices.py:22-63, numpy arrays in / numpy arrays out).
//
// The batch is cut into chunks that flow through three internal streams -- upload, kernel,
// download -- linked by one event per chunk and stage: the uploads run back to back whatever the
// kernels do (the host link is the slower side for the C2 shape: 0.77 ms of uploads against
// 0.66 ms of kernels), chunk c runs while chunk c+1 arrives and chunk c-1's results travel
// back, so the end-to-end rate is max(PCIe, compute) instead of their sum.  Chunk boundaries
// sit on multiples of one full wave of the kernel; the first and the last chunk are small (only
// the first upload and the last kernel + download are exposed).  Everything is ordered after
// the work already queued on `stream` and joined back into it.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mgp {
namespace {

constexpr int N_SIDE = 3;  // upload, kernel, download
struct SideStreams {
  cudaStream_t s[N_SIDE] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr;
  cudaEvent_t join[N_SIDE] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> chunk_ev;  // two per chunk: uploaded, computed
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
    for (int i = 0; i < N_SIDE; ++i) {
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

// Chunk boundaries: multiples of `wave` rows.  Every chunk costs ~15 us of launches and
// pipeline fill, so there are few of them: eight waves each up to 16 chunks (C2: 8, 8, 8, 8, 8,
// 6 waves); larger batches ramp up x1.25 from eight waves to waves/16 per chunk and back down,
// so that the exposed ends -- the first upload, the last kernel + download -- stay small.
// Measured on C2 (uploads 0.73 ms, kernels 0.65 ms): 0.90 ms end to end, against 0.95 ms with a
// finer tail (9 chunks) and 1.09 ms for the two-stream x1.4 schedule it replaces.
std::vector<long long> chunk_bounds(long long b, long long wave) {
  // dev switches for tuning the schedule
  static const double first = getenv("MGP_PIPE_FIRST") ? atof(getenv("MGP_PIPE_FIRST")) : 8.0;
  static const double growth = getenv("MGP_PIPE_GROWTH") ? atof(getenv("MGP_PIPE_GROWTH")) : 1.25;
  const long long waves = (b + wave - 1) / wave;
  const double cap = (double)waves / 16 > first ? (double)waves / 16 : first;
  auto ramp = [&](long long target) {
    std::vector<long long> v;
    double w = first;
    long long acc = 0;
    while (acc < target) {
      long long sz = (long long)w < 1 ? 1 : (long long)w;
      if (target - acc - sz <= sz / 2) sz = target - acc;  // no sliver at the end
      v.push_back(sz);
      acc += sz;
      w = w * growth > cap ? cap : w * growth;
    }
    return v;
  };
  const long long head_waves = (waves + 1) / 2;
  std::vector<long long> sizes = ramp(head_waves);
  const std::vector<long long> tail = ramp(waves - head_waves);
  sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
  std::vector<long long> bounds{0};
  for (long long sz : sizes) {
    if (bounds.back() >= b) break;
    bounds.push_back(bounds.back() + sz * wave);
  }
  bounds.back() = b;
  return bounds;
}

// 32-bit neighbour indices (half the bytes on the host link) widened on the device into the
// int64 staging buffer the kernels read: four per thread where both sides are 16-byte aligned.
__global__ void __launch_bounds__(256) widen_idx_kernel(const int32_t* __restrict__ src,
                                                        int64_t* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 15) == 0) {
    const long long n4 = n >> 2;
    for (long long v = i; v < n4; v += stride) {
      const int4 a = reinterpret_cast<const int4*>(src)[v];
      longlong2* o = reinterpret_cast<longlong2*>(dst) + 2 * v;
      o[0] = make_longlong2(a.x, a.y);
      o[1] = make_longlong2(a.z, a.w);
    }
    for (long long e = (n4 << 2) + i; e < n; e += stride) dst[e] = src[e];
  } else {
    for (; i < n; i += stride) dst[i] = src[i];
  }
}

}  // namespace

int validate_problem(const mgp_problem* p);
int fused_wave_per_sm(const mgp_problem* p);

}  // namespace mgp

using namespace mgp;

static int fused_posterior_host_impl(const mgp_problem* p, const int64_t* nn_idx_host,
                                     const int32_t* nn_idx_host32, int32_t* nn_stage32,
                                     const int64_t* query_idx_host, double* mean_host,
                                     double* var_host, void* ws, size_t ws_bytes, void* stream) {
  MGP_REQUIRE(p != nullptr, MGP_ERR_BAD_ARG, "null problem");
  MGP_REQUIRE((nn_idx_host != nullptr || nn_idx_host32 != nullptr) && p->nn_idx != nullptr,
              MGP_ERR_BAD_ARG,
              "nn_idx_host and the device staging buffer p->nn_idx are required");
  MGP_REQUIRE(nn_idx_host32 == nullptr || nn_stage32 != nullptr, MGP_ERR_BAD_ARG,
              "32-bit host indices need the int32 device staging buffer");
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
    for (int i = 0; i < N_SIDE; ++i)
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
  for (int i = 0; i < N_SIDE; ++i) MGP_CUDA(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
  forked = true;
  cudaStream_t up = ss->s[0], run = ss->s[1], down = ss->s[2];

  // one wave of the kernel this shape takes: neighbourhoods in flight per SM x SMs
  const long long wave = (long long)fused_wave_per_sm(p) * sm_count();
  const std::vector<long long> bounds = chunk_bounds(p->b, wave);
  const long long k = p->k, r = p->r;
  while (ss->chunk_ev.size() < 2 * bounds.size()) {
    cudaEvent_t ev = nullptr;
    MGP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ss->chunk_ev.push_back(ev);
  }
  const bool downloads = mean_host != nullptr || var_host != nullptr;
  for (size_t c = 0; c + 1 < bounds.size(); ++c) {
    const long long lo = bounds[c], hi = bounds[c + 1], rows = hi - lo;
    if (rows <= 0) continue;
    cudaEvent_t uploaded = ss->chunk_ev[2 * c], computed = ss->chunk_ev[2 * c + 1];
    int64_t* nn_dev = const_cast<int64_t*>(p->nn_idx) + lo * k;
    if (nn_idx_host32)
      MGP_CUDA(cudaMemcpyAsync(nn_stage32 + lo * k, nn_idx_host32 + lo * k,
                               (size_t)rows * k * sizeof(int32_t), cudaMemcpyHostToDevice, up));
    else
      MGP_CUDA(cudaMemcpyAsync(nn_dev, nn_idx_host + lo * k, (size_t)rows * k * sizeof(int64_t),
                               cudaMemcpyHostToDevice, up));
    if (query_idx_host)
      MGP_CUDA(cudaMemcpyAsync(const_cast<int64_t*>(p->query_idx) + lo, query_idx_host + lo,
                               (size_t)rows * sizeof(int64_t), cudaMemcpyHostToDevice, up));
    MGP_CUDA(cudaEventRecord(uploaded, up));
    MGP_CUDA(cudaStreamWaitEvent(run, uploaded, 0));
    if (nn_idx_host32) {
      const long long cnt = rows * k;
      const long long want = (cnt / 4 + 255) / 256;
      const int blocks = (int)(want < 1 ? 1 : want > 4LL * sm_count() ? 4LL * sm_count() : want);
      widen_idx_kernel<<<blocks, 256, 0, run>>>(nn_stage32 + lo * k, nn_dev, cnt);
      MGP_CUDA(cudaGetLastError());
    }
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
    rc = mgp_fused_posterior(&sub, ws, ws_bytes, (void*)run);  // (one kernel stream: one ws)
    if (rc != MGP_OK) {
      join();
      return rc;
    }
    if (downloads) {
      MGP_CUDA(cudaEventRecord(computed, run));
      MGP_CUDA(cudaStreamWaitEvent(down, computed, 0));
    }
    if (mean_host)
      MGP_CUDA(cudaMemcpyAsync(mean_host + lo * r, sub.mean, (size_t)rows * r * sizeof(double),
                               cudaMemcpyDeviceToHost, down));
    if (var_host)
      MGP_CUDA(cudaMemcpyAsync(var_host + lo, sub.var, (size_t)rows * sizeof(double),
                               cudaMemcpyDeviceToHost, down));
  }
  // join: later work on `stream` sees every chunk
  for (int i = 0; i < N_SIDE; ++i) {
    MGP_CUDA(cudaEventRecord(ss->join[i], ss->s[i]));
    MGP_CUDA(cudaStreamWaitEvent(main_stream, ss->join[i], 0));
  }
#undef MGP_CUDA
  return MGP_OK;
}

extern "C" int mgp_fused_posterior_host(const mgp_problem* p, const int64_t* nn_idx_host,
                                        const int64_t* query_idx_host, double* mean_host,
                                        double* var_host, void* ws, size_t ws_bytes,
                                        void* stream) {
  MGP_REQUIRE(nn_idx_host != nullptr, MGP_ERR_BAD_ARG, "nn_idx_host is required");
  return fused_posterior_host_impl(p, nn_idx_host, nullptr, nullptr, query_idx_host, mean_host,
                                   var_host, ws, ws_bytes, stream);
}

extern "C" int mgp_fused_posterior_host32(const mgp_problem* p, const int32_t* nn_idx_host32,
                                          int32_t* nn_stage32, const int64_t* query_idx_host,
                                          double* mean_host, double* var_host, void* ws,
                                          size_t ws_bytes, void* stream) {
  MGP_REQUIRE(nn_idx_host32 != nullptr, MGP_ERR_BAD_ARG, "nn_idx_host32 is required");
  return fused_posterior_host_impl(p, nullptr, nn_idx_host32, nn_stage32, query_idx_host,
                                   mean_host, var_host, ws, ws_bytes, stream);
}
