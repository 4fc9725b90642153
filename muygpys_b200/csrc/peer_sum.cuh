// One-shot SUM of an 8-double partials record across the GPUs of one NVLink / NVSwitch domain,
// executed by ONE thread block as the tail of the kernel that produced the record -- the
// replacement for the reference's scalar MPI allreduces (S/_src/optimize/loss/mpi.py:21-104,
// S/_src/optimize/scale/mpi.py:19-37) without a separate collective launch.
//
// Every rank owns a small peer-mapped buffer (symmetric memory); `g.peer_buf[p]` is rank p's
// buffer as seen from this GPU.  Layout of a buffer:
//     double             data[2][MGP_MAX_PEERS][8]     record of rank r for parity (epoch & 1)
//     unsigned long long flag[2][MGP_MAX_PEERS]        epoch at which data[parity][r] is valid
// A block PUSHES its record into slot [parity][rank] of every peer's buffer (8-byte P2P stores
// over NVLink), fences, raises flag[parity][rank] = epoch on every peer, then waits until its OWN
// buffer shows `epoch` for every rank and adds the records in rank order (the same order on every
// GPU: bit-identical results everywhere).  Epochs increase by one per call and select the parity,
// so nothing is ever reset: a rank can run at most one call ahead of the slowest (it needs every
// peer's flag of call e to finish call e), and call e + 1 writes the other parity.
#pragma once

#include "common.cuh"

namespace mgp {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr size_t PEER_DATA_DOUBLES = 2 * MGP_MAX_PEERS * MGP_PARTIALS;

// Called by every thread of one block (>= 8 * world threads).  `rec` holds this rank's record
// in SHARED memory (8 doubles); on return threads 0..7 have written the summed record to
// `out[0..7]`.  If a peer does not show up within ~2 s the record is poisoned with NaN instead
// of spinning forever.
__device__ __forceinline__ void peer_sum8_block(const mgp_peer_group& g, const double* rec,
                                                double* out) {
  const int tid = threadIdx.x;
  const int par = (int)(g.epoch & 1ull);
  if (tid < 8 * g.world) {
    const int p = tid >> 3, s = tid & 7;
    double* dst = (double*)g.peer_buf[p] + (size_t)(par * MGP_MAX_PEERS + g.rank) * MGP_PARTIALS;
    dst[s] = rec[s];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_timeout;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < g.world) {
    unsigned long long* fl = (unsigned long long*)((double*)g.peer_buf[tid] + PEER_DATA_DOUBLES) +
                             par * MGP_MAX_PEERS + g.rank;
    st_release_sys(fl, g.epoch);
    const unsigned long long* mine =
        (const unsigned long long*)((const double*)g.peer_buf[g.rank] + PEER_DATA_DOUBLES) +
        par * MGP_MAX_PEERS + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != g.epoch) {
      if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (tid < MGP_PARTIALS) {
    const volatile double* src =
        (const volatile double*)g.peer_buf[g.rank] + (size_t)par * MGP_MAX_PEERS * MGP_PARTIALS;
    double tot = 0.0;
    for (int p = 0; p < g.world; ++p) tot += src[p * MGP_PARTIALS + tid];
    out[tid] = s_timeout ? __longlong_as_double(0x7ff8000000000000LL) : tot;
  }
}

}  // namespace mgp
