// K2 (high-dimensional fast path): exact KNN for d > 8 as DMMA pre-filter + exact re-rank.
//
// The exact kernel (knn_tiled.cu) spends 3 FP64 issue slots per (query, point, feature) because
// the ranking must reproduce a scalar CPU loop bit for bit (separately rounded subtract,
// multiply, add).  Almost all of that work ranks points that are nowhere near the k-th
// neighbour.  Here
//
//   1. knn_gram_filter_kernel sweeps the training set with FP64 tensor-core tiles,
//      S~(q,x) = |q|^2 + |x|^2 - 2 q.x  (mma.sync.m8n8k4.f64, 1/8 issue slot per pair-feature),
//      and keeps the K' = k + 8 smallest S~ per query;
//   2. knn_gram_refine_kernel recomputes the K' candidates with the exact arithmetic of
//      knn_tiled.cu / knn.cu (same operations, same feature order), ranks them by
//      (distance, train row) and CERTIFIES the result: every point that is not a candidate has
//      S~ >= tau (the largest kept S~), and |S~ - S| <= E with a rigorous rounding bound E, so
//      its exact distance is at least L = tau - E.  If the exact k-th distance is strictly
//      below L the k results are exactly what the full sweep returns (all ties included);
//   3. queries that cannot be certified (near-ties across the candidate boundary, massive
//      duplicates) are listed and re-run through the exact sweep (knn_tiled.cu, query map).
//
// So the returned indices and squared distances are always bit-identical to the brute-force
// kernels'; only the amount of work depends on the data.
#include <cuda.h>
#include <float.h>
#include <limits.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "smem_heap.cuh"

namespace mgp {

size_t knn_tiled_workspace_bytes(long long n, long long q, int k);
int launch_knn_tiled(const double* train, long long n, const double* queries, long long q, int d,
                     int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2, void* ws,
                     size_t ws_bytes, cudaStream_t s, const int32_t* qmap, const int32_t* qcount);

namespace {

constexpr int KG_Q = 64;     // queries per CTA
constexpr int KG_X = 64;     // train points per tile
constexpr int KG_F = 32;     // features per staged chunk
constexpr int KG_LD = 36;    // row pitch: fragment loads (row rho, feature q4) hit 4 rho + q4
constexpr int KG_MARGIN = 8; // extra candidates beyond k
constexpr int KG_KMAX = 96;  // candidate list capacity

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// |row|^2 for every row (one warp per row) and, optionally, the maximum over rows
__global__ void row_norms_kernel(const double* __restrict__ a, long long rows, int d,
                                 double* __restrict__ out, unsigned long long* max_bits) {
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  double mx = 0.0;
  for (long long r = w; r < rows; r += nw) {
    double s = 0.0;
    for (int f = lane; f < d; f += 32) {
      const double v = a[r * d + f];
      s = fma(v, v, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
    mx = fmax(mx, s);
  }
  // non-negative doubles order like their bit patterns
  if (max_bits && lane == 0 && mx > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(mx));
}

__device__ __forceinline__ void kg_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void kg_cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing for the operand tiles ------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// wait for the phase with the given parity; a copy that never lands traps instead of hanging
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
// box of a 2-D tensor map (coordinates: c0 = feature, c1 = row) -> shared memory
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

constexpr int KG_TF = 16;  // features per TMA box: 128 bytes, the span of the 128-byte swizzle
// element (row r, feature f < 16) of a 64 x 16 box written with CU_TENSOR_MAP_SWIZZLE_128B:
// the 16-byte chunk index is XORed with the row (mod 8), so the 8 rows a fragment load touches
// land in 8 different bank groups (a dense 128-byte pitch would be an 8-way conflict)
__device__ __forceinline__ double box_ld(const double* box, int r, int f) {
  return box[r * KG_TF + ((((f >> 1) ^ (r & 7)) << 1) | (f & 1))];
}

// VEC: d even and both arrays 16-byte aligned -> 16-byte asynchronous copies.
// NST > 0: the operand tiles arrive by TMA (cp.async.bulk.tensor, one thread issues two box
// copies per stage into an NST-deep ring guarded by mbarriers; out-of-range rows / features are
// zero-filled by the copy engine); NST == 0: thread-issued cp.async into two buffers.
template <bool VEC, typename IdxT, int NST>
__global__ void __launch_bounds__(256, 2) knn_gram_filter_kernel(
    const double* __restrict__ train, long long n, const double* __restrict__ queries,
    long long q, int d, int kk, const int64_t* __restrict__ self_idx,
    const double* __restrict__ xn, const double* __restrict__ qn, long long split_len,
    int nsplit, int32_t* __restrict__ part_idx, double* __restrict__ part_s,
    int64_t* __restrict__ cand_idx, double* __restrict__ cand_s,
    const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmx) {
  constexpr bool TMA = NST > 0;
  // two staging buffers filled with cp.async (the copy of chunk t + 1 runs under the DMMAs of
  // chunk t); the 64 x 65 tile of S~ reuses the buffer that was consumed last
  // the swizzle pattern of the TMA boxes is a function of the shared-memory ADDRESS: the ring
  // must start on a 1024-byte boundary (8 rows x 128 bytes)
  extern __shared__ __align__(1024) unsigned char kg_raw[];
  double* stage =
      reinterpret_cast<double*>(kg_raw + ((1024u - (smem_u32(kg_raw) & 1023u)) & 1023u));
  __shared__ double worst_sh[KG_Q];
  __shared__ __align__(8) unsigned short mask_sh[4 * KG_Q];
  __shared__ __align__(8) unsigned long long full_bar[TMA ? NST : 1];   // stage has landed
  __shared__ __align__(8) unsigned long long empty_bar[TMA ? NST : 1];  // all 8 warps are done
  constexpr int FW = TMA ? KG_TF : KG_F;  // features per stage
  constexpr int BUF = TMA ? 2 * KG_Q * KG_TF : 2 * KG_Q * KG_LD;
  constexpr int NBUF = TMA ? NST : 2;
  constexpr int DS = KG_Q * (KG_X + 1);   // the S~ tile (TMA: its own area behind the ring)
  static_assert(TMA || DS <= BUF, "S~ tile must fit one staging buffer");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rho = lane >> 2, q4 = lane & 3;
  const int tr0 = 2 * (warp & 3);   // this warp's two tile rows (queries)
  const int tc0 = 4 * (warp >> 2);  // and four tile columns (train points)
  const long long q0 = (long long)blockIdx.x * KG_Q;

  // The candidates of the 64 queries are bounded max-heaps on (S~, row) in shared memory,
  // entry-major ([entry][query]: the owners' accesses never conflict).  As per-thread sorted
  // arrays they sat in local memory, and with the L1 carved down to ~30 KB by the staging
  // buffers every shift of an insertion was an L2 round trip (ncu: 2/3 of all warp time went to
  // the insertion loop and to the other six warps waiting for it at the barrier).
  // Rows are stored as offsets into this CTA's slice of the training set: 16 bits when the
  // slice is short enough, which is what lets two CTAs share an SM up to k = 50.
  SmemHeapT<IdxT> top;
  top.hd = stage + NBUF * BUF + (TMA ? DS + 1 : 0);
  top.hi = reinterpret_cast<IdxT*>(top.hd + (size_t)kk * KG_Q);
  top.nt = KG_Q;
  top.t = tid;
  const bool owner = tid < KG_Q && q0 + tid < q;
  long long self = -1;
  if (owner && self_idx) self = self_idx[q0 + tid];
  int size = 0;             // heap entries so far (<= kk)
  double worst = DBL_MAX;   // the root once the heap is full
  int worst_i = INT_MAX;
  double qnr[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long r = q0 + 8 * (tr0 + i) + rho;
    qnr[i] = r < q ? qn[r] : 0.0;
  }

  const long long x_begin = (long long)blockIdx.y * split_len;
  const long long x_end = min(n, x_begin + split_len);
  const int nchunks = (d + FW - 1) / FW;
  const long long ntiles = x_end > x_begin ? (x_end - x_begin + KG_X - 1) / KG_X : 0;
  const long long total = ntiles * nchunks;

  // a stage = one feature chunk of one train tile (and of the queries); stages run tile by
  // tile, chunk by chunk (counters are kept incrementally: no 64-bit divisions in the loop)
  auto issue = [&](long long x0, int f0, int buf) {
    const int fc = min(KG_F, d - f0);
    double* Qs = stage + buf * BUF;
    double* Xs = Qs + KG_Q * KG_LD;
    if (VEC) {
      for (int e = tid; e < KG_Q * (KG_F / 2); e += 256) {
        const int r = e / (KG_F / 2), f = 2 * (e - r * (KG_F / 2));
        const long long qi = q0 + r, xi = x0 + r;
        if (qi < q && f < fc) kg_cp_async16(&Qs[r * KG_LD + f], &queries[qi * d + f0 + f]);
        else *reinterpret_cast<double2*>(&Qs[r * KG_LD + f]) = make_double2(0.0, 0.0);
        if (xi < x_end && f < fc) kg_cp_async16(&Xs[r * KG_LD + f], &train[xi * d + f0 + f]);
        else *reinterpret_cast<double2*>(&Xs[r * KG_LD + f]) = make_double2(0.0, 0.0);
      }
    } else {
      for (int e = tid; e < KG_Q * KG_F; e += 256) {
        const int r = e / KG_F, f = e - r * KG_F;
        const long long qi = q0 + r, xi = x0 + r;
        if (qi < q && f < fc) kg_cp_async8(&Qs[r * KG_LD + f], &queries[qi * d + f0 + f]);
        else Qs[r * KG_LD + f] = 0.0;
        if (xi < x_end && f < fc) kg_cp_async8(&Xs[r * KG_LD + f], &train[xi * d + f0 + f]);
        else Xs[r * KG_LD + f] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // TMA: one thread arms the stage's barrier with the byte count and issues the two box copies
  auto issue_tma = [&](long long x0, int f0, int buf) {
    if (tid == 0) {
      const unsigned bar = smem_u32(&full_bar[buf]);
      mbar_expect_tx(bar, 2 * KG_Q * KG_TF * (unsigned)sizeof(double));
      tma_load_2d(smem_u32(stage + buf * BUF), &tmq, f0, (int)q0, bar);
      tma_load_2d(smem_u32(stage + buf * BUF + KG_Q * KG_TF), &tmx, f0, (int)x0, bar);
    }
  };
  if (TMA) {
    if (tid == 0) {
      for (int i = 0; i < NBUF; ++i) {
        mbar_init(smem_u32(&full_bar[i]), 1);
        mbar_init(smem_u32(&empty_bar[i]), 8);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
  }

  double acc[2][4][2];
  // issue cursor (TMA: NST stages ahead of the compute cursor)
  long long ix0 = x_begin;
  int ichunk = 0;
  auto issue_next = [&](int buf) {
    if (TMA) issue_tma(ix0, ichunk * FW, buf);
    else issue(ix0, ichunk * FW, buf);
    if (++ichunk == nchunks) {
      ichunk = 0;
      ix0 += KG_X;
    }
  };
  long long issued = 0;
  for (; issued < total && issued < (TMA ? NBUF : 1); ++issued) issue_next((int)issued);
  long long x0 = x_begin;  // current stage
  int chunk = 0, buf = 0;
  for (long long st = 0; st < total; ++st, buf = (buf + 1 == NBUF) ? 0 : buf + 1) {
    // the stage after this one
    int nchunk = chunk + 1;
    long long nx0 = x0;
    if (nchunk == nchunks) {
      nchunk = 0;
      nx0 += KG_X;
    }
    if (TMA) {
      mbar_wait(smem_u32(&full_bar[buf]), (unsigned)((st / NBUF) & 1));
    } else {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();  // this stage has landed; nobody still reads the other buffer
      if (issued < total) {
        issue_next(buf ^ 1);
        ++issued;
      }
    }
    if (chunk == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    const double* Qs = stage + buf * BUF;
    const double* Xs = Qs + (TMA ? KG_Q * KG_TF : KG_Q * KG_LD);
    const int ksteps = (min(FW, d - chunk * FW) + 3) >> 2;
    for (int ks = 0; ks < ksteps; ++ks) {
      double a[2], b[4];
      if (TMA) {
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = box_ld(Qs, 8 * (tr0 + i) + rho, 4 * ks + q4);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = box_ld(Xs, 8 * (tc0 + j) + rho, 4 * ks + q4);
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = Qs[(8 * (tr0 + i) + rho) * KG_LD + 4 * ks + q4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Xs[(8 * (tc0 + j) + rho) * KG_LD + 4 * ks + q4];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (TMA) {
      // No block-wide barrier per stage: each warp releases the stage on its `empty` barrier
      // and moves on; the issuing thread refills the PREVIOUS stage's buffer (which the other
      // warps have almost certainly left) once all eight warps have released it, so the copy
      // engine runs NST - 1 stages ahead of the slowest warp.
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty_bar[buf]));
      if (st >= 1 && issued < total) {
        const int pbuf = (buf == 0) ? NBUF - 1 : buf - 1;
        if (tid == 0) mbar_wait(smem_u32(&empty_bar[pbuf]), (unsigned)(((st - 1) / NBUF) & 1));
        issue_next(pbuf);
        ++issued;
      }
    }
    if (chunk != nchunks - 1) {
      chunk = nchunk;
      continue;
    }
    // ---- a train tile is complete: S~ to the owners of the candidate lists ----------------
    double* Ds = TMA ? stage + NBUF * BUF : stage + buf * BUF;
    // non-TMA: everyone is done with the staged features of this buffer; TMA: the owners are
    // done with the previous S~ tile (warps run up to NST stages apart)
    __syncthreads();
    // accumulator layout: lane (rho, q4) holds columns 2 q4, 2 q4 + 1 of row rho
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 8 * (tc0 + j) + 2 * q4;
      const double xn0 = (x0 + c < x_end) ? xn[x0 + c] : 0.0;
      const double xn1 = (x0 + c + 1 < x_end) ? xn[x0 + c + 1] : 0.0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = 8 * (tr0 + i) + rho;
        Ds[r * (KG_X + 1) + c] = fma(-2.0, acc[i][j][0], qnr[i] + xn0);
        Ds[r * (KG_X + 1) + c + 1] = fma(-2.0, acc[i][j][1], qnr[i] + xn1);
      }
    }
    if (owner) worst_sh[tid] = worst;
    __syncthreads();
    // All 256 threads screen the tile against the owners' current thresholds (thread t: query
    // t / 4, 16 candidates) and leave a bit mask of survivors; the owner of a query then only
    // visits the set bits (a handful per tile once the lists have warmed up), in ascending
    // candidate order, re-testing against its tightening threshold.
    {
      const int r = tid >> 2, c0 = (tid & 3) * 16;
      const double w = worst_sh[r];
      const int cnt = (int)min((long long)KG_X, x_end - x0);
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < cnt && Ds[r * (KG_X + 1) + c0 + j] <= w) m |= 1u << j;
      mask_sh[tid] = (unsigned short)m;
    }
    __syncthreads();
    if (owner) {
      unsigned long long m = *reinterpret_cast<const unsigned long long*>(&mask_sh[4 * tid]);
      while (m) {
        const int j = __ffsll((long long)m) - 1;
        m &= m - 1;
        const double s = Ds[tid * (KG_X + 1) + j];
        if (x0 + j == self) continue;
        const int id = (int)(x0 + j - x_begin);
        if (size < kk) {
          top.push(size, s, id);
          if (++size == kk) {
            worst = top.D(0);
            worst_i = top.I(0);
          }
        } else if (SmemHeapT<IdxT>::less(s, id, worst, worst_i)) {
          top.replace_root(kk, s, id);
          worst = top.D(0);
          worst_i = top.I(0);
        }
      }
    }
    // (the next iteration's barrier keeps this buffer from being refilled under the scan)
    chunk = nchunk;
    x0 = nx0;
  }
  if (owner) {
    // heap sort in place (ascending), then hand the list over; a short slice leaves sentinels
    for (int m = size - 1; m > 0; --m) {
      const double s = top.D(m);
      const int id = top.I(m);
      top.D(m) = top.D(0);
      top.I(m) = top.I(0);
      top.replace_root(m, s, id);
    }
    if (nsplit == 1) {
      for (int i = 0; i < kk; ++i) {
        cand_idx[(q0 + tid) * kk + i] = i < size ? (long long)top.I(i) + x_begin : INT_MAX;
        cand_s[(q0 + tid) * kk + i] = i < size ? top.D(i) : DBL_MAX;
      }
    } else {
      const long long base = ((q0 + tid) * nsplit + blockIdx.y) * kk;
      for (int i = 0; i < kk; ++i) {
        part_idx[base + i] = i < size ? (int)(top.I(i) + x_begin) : INT_MAX;
        part_s[base + i] = i < size ? top.D(i) : DBL_MAX;
      }
    }
  }
}

// merge the per-slice candidate lists of one query (lists sorted by S~)
template <int KMAX>
__global__ void knn_gram_merge_kernel(const int32_t* __restrict__ part_idx,
                                      const double* __restrict__ part_s, long long q, int nsplit,
                                      int kk, int64_t* __restrict__ cand_idx,
                                      double* __restrict__ cand_s) {
  const long long qi = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (qi >= q) return;
  double best_s[KMAX];
  int best_i[KMAX];
  for (int i = 0; i < kk; ++i) {
    best_s[i] = DBL_MAX;
    best_i[i] = INT_MAX;
  }
  for (int s = 0; s < nsplit; ++s) {
    const long long base = (qi * nsplit + s) * kk;
    for (int j = 0; j < kk; ++j) {
      const double dv = part_s[base + j];
      if (!(dv < best_s[kk - 1])) break;
      int pos = kk - 1;
      while (pos > 0 && best_s[pos - 1] > dv) {
        best_s[pos] = best_s[pos - 1];
        best_i[pos] = best_i[pos - 1];
        --pos;
      }
      best_s[pos] = dv;
      best_i[pos] = part_idx[base + j];
    }
  }
  for (int i = 0; i < kk; ++i) {
    cand_idx[qi * kk + i] = best_i[i];
    cand_s[qi * kk + i] = best_s[i];
  }
}

// One warp per query: exact distances of the kk candidates (the arithmetic of knn_tiled.cu),
// rank by (distance, row), certify against the pre-filter's threshold.
__global__ void __launch_bounds__(128) knn_gram_refine_kernel(
    const double* __restrict__ train, long long n, const double* __restrict__ queries,
    long long q, int d, int k, int kk, const int64_t* __restrict__ cand_idx,
    const double* __restrict__ cand_s, const double* __restrict__ qn,
    const unsigned long long* __restrict__ xn_max_bits, int64_t* __restrict__ out_idx,
    double* __restrict__ out_d2, int32_t* __restrict__ flag_list, int32_t* __restrict__ flag_count) {
  __shared__ double rows[4][32 * 33];
  __shared__ double sd[4][KG_KMAX];
  __shared__ int si[4][KG_KMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qi = blockIdx.x * 4LL + warp;
  if (qi >= q) return;
  double* buf = rows[warp];
  const double* qrow = queries + qi * d;
  for (int g0 = 0; g0 < kk; g0 += 32) {
    const int c = g0 + lane;
    const long long my = c < kk ? cand_idx[qi * kk + c] : -1;
    const int cnt = min(32, kk - g0);
    double acc = 0.0;
    for (int f0 = 0; f0 < d; f0 += 32) {
      const double qreg = (f0 + lane < d) ? qrow[f0 + lane] : 0.0;
      __syncwarp();
      for (int c2 = 0; c2 < cnt; ++c2) {
        const long long src = __shfl_sync(0xffffffffu, my, c2);
        buf[c2 * 33 + lane] = (f0 + lane < d) ? train[src * d + f0 + lane] : 0.0;
      }
      __syncwarp();
      // every lane runs the loop (the query chunk is broadcast with full-warp shuffles); lanes
      // without a candidate accumulate stale rows and are ignored below
#pragma unroll 8
      for (int f = 0; f < 32; ++f) {
        const double df = __dsub_rn(__shfl_sync(0xffffffffu, qreg, f), buf[lane * 33 + f]);
        acc = __dadd_rn(acc, __dmul_rn(df, df));
      }
    }
    if (c < kk) {
      sd[warp][c] = acc;
      si[warp][c] = (int)my;
    }
  }
  __syncwarp();
  // rank every candidate among all candidates by (distance, row)
  bool certified = true;
  for (int c = lane; c < kk; c += 32) {
    const double ms = sd[warp][c];
    const int mi = si[warp][c];
    int rank = 0;
    for (int j = 0; j < kk; ++j) {
      const double os = sd[warp][j];
      const int oi = si[warp][j];
      rank += (os < ms || (os == ms && oi < mi)) ? 1 : 0;
    }
    if (rank < k) {
      out_idx[qi * k + rank] = mi;
      out_d2[qi * k + rank] = ms;
    }
    if (rank == k - 1 && (long long)kk < n) {
      // Non-candidates have S~ >= tau.  |S~ - S| <= gamma (|q|^2 + |x|^2 + 2|q.x|)
      // <= 2 gamma (|q|^2 + |x|^2) with gamma ~ (d + 2) 2^-53 for the norms and the DMMA dot
      // product; the brute-force sum itself is within d 2^-53 relative of S.  Use 8x that.
      const double tau = cand_s[qi * kk + kk - 1];
      const double gamma = 8.0 * (double)(d + 8) * 1.1102230246251565e-16;
      const double xmax = __longlong_as_double((long long)*xn_max_bits);
      const double lower = (tau - 2.0 * gamma * (qn[qi] + xmax)) * (1.0 - gamma);
      certified = ms < lower;
    }
  }
  if (!certified) {
    const int pos = atomicAdd(flag_count, 1);
    flag_list[pos] = (int32_t)qi;
  }
}

struct GramWs {
  double* xn;
  double* qn;
  unsigned long long* xn_max;
  int32_t* flag_count;
  int32_t* flag_list;
  int64_t* cand_idx;
  double* cand_s;
  double* part_s;
  int32_t* part_idx;
  void* tiled_ws;
  size_t tiled_bytes;
  size_t total;
};

// slices of the training set per query block: enough CTAs for ~4 full waves of 2 CTAs per SM,
// at least 4096 points per slice (every slice warms up its own candidate heaps), and the count
// whose last wave is fullest
int gram_splits(long long n, long long q) {
  const long long qblocks = (q + KG_Q - 1) / KG_Q;
  const long long wave = 2LL * sm_count();
  long long hi = (n + 4095) / 4096;
  if (hi > 64) hi = 64;
  if (hi < 1) hi = 1;
  long long lo = (wave + qblocks - 1) / qblocks;
  const long long for_short_rows = (n + 65535) / 65536;  // slices that fit 16-bit row offsets
  if (lo < for_short_rows) lo = for_short_rows;
  if (lo > hi) lo = hi;
  long long best = lo;
  double best_eff = 0.0;
  for (long long ns = lo; ns <= hi; ++ns) {
    const long long ctas = qblocks * ns;
    const double eff = (double)ctas / (double)(((ctas + wave - 1) / wave) * wave);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = ns;
    }
    if (ctas >= 4 * wave && eff > 0.95) break;
  }
  return (int)best;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

GramWs carve(void* ws, long long n, long long q, int k, int kk) {
  GramWs w;
  const int ns = gram_splits(n, q);
  char* p = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  w.xn = (double*)take((size_t)n * 8);
  w.qn = (double*)take((size_t)q * 8);
  w.xn_max = (unsigned long long*)take(16);
  w.flag_count = (int32_t*)(w.xn_max ? (char*)w.xn_max + 8 : nullptr);
  w.flag_list = (int32_t*)take((size_t)q * 4);
  w.cand_idx = (int64_t*)take((size_t)q * kk * 8);
  w.cand_s = (double*)take((size_t)q * kk * 8);
  w.part_s = (double*)take(ns > 1 ? (size_t)q * ns * kk * 8 : 0);
  w.part_idx = (int32_t*)take(ns > 1 ? (size_t)q * ns * kk * 4 : 0);
  w.tiled_bytes = knn_tiled_workspace_bytes(n, q, k);
  w.tiled_ws = take(w.tiled_bytes);
  w.total = off;
  return w;
}

}  // namespace

static const bool g_knn_gram_off = getenv("MGP_NO_GRAM_KNN") != nullptr;  // dev switch
static const int g_knn_gram_min_d =
    getenv("MGP_GRAM_KNN_MIN_D") ? atoi(getenv("MGP_GRAM_KNN_MIN_D")) : 1;  // dev switch

bool knn_gram_supported(long long n, long long q, int d, int k) {
  if (g_knn_gram_off) return false;
  // (measured, 200 k x 20 k, k = 50: 19-25 ms for every d from 4 to 32, against 97-144 ms for the
  // exact sweep at d = 9..32 and 177-184 ms for the thread-per-query kernel at d = 4..8)
  return d >= g_knn_gram_min_d && q >= 8 && n >= 2048 && k + KG_MARGIN <= KG_KMAX;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}

// (rows, d) row-major float64 matrix as a 2-D tensor map with 16-feature x 64-row boxes,
// 128-byte swizzle, zero fill outside the matrix
static bool make_operand_map(CUtensorMap* map, const double* base, long long rows, int d) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * sizeof(double)};
  const cuuint32_t box[2] = {KG_TF, KG_Q};
  const cuuint32_t estride[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride,
             box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static const int g_knn_tma_off = getenv("MGP_KNN_NO_TMA") != nullptr;  // dev switch

static int gram_kk(long long n, int k, bool has_self) {
  long long kk = k + KG_MARGIN;
  const long long avail = n - (has_self ? 1 : 0);
  if (kk > avail) kk = avail;
  return (int)kk;
}

size_t knn_gram_workspace_bytes(long long n, long long q, int k) {
  return carve(nullptr, n, q, k, k + KG_MARGIN).total + 256;
}

int launch_knn_gram(const double* train, long long n, const double* queries, long long q, int d,
                    int k, const int64_t* self_idx, int64_t* out_idx, double* out_d2, void* ws,
                    size_t ws_bytes, cudaStream_t s) {
  MGP_REQUIRE(ws != nullptr && ws_bytes >= knn_gram_workspace_bytes(n, q, k), MGP_ERR_WORKSPACE,
              "KNN workspace too small (%zu bytes)", ws_bytes);
  const int kk = gram_kk(n, k, self_idx != nullptr);
  void* base = (void*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const GramWs w = carve(base, n, q, k, k + KG_MARGIN);
  const int ns = gram_splits(n, q);
  cudaMemsetAsync(w.xn_max, 0, 16, s);
  const int nb = sm_count() * 8;
  row_norms_kernel<<<nb, 256, 0, s>>>(train, n, d, w.xn, w.xn_max);
  row_norms_kernel<<<nb, 256, 0, s>>>(queries, q, d, w.qn, nullptr);
  long long split_len = (n + ns - 1) / ns;
  split_len = (split_len + KG_X - 1) / KG_X * KG_X;
  const dim3 grid((unsigned)((q + KG_Q - 1) / KG_Q), (unsigned)ns);
  const bool vec = d % 2 == 0 && (uintptr_t)train % 16 == 0 && (uintptr_t)queries % 16 == 0;
  // 16-bit in-slice rows (10 instead of 12 bytes per heap entry) whenever a slice is short enough
  const bool short_rows = split_len <= 65536;
  const size_t heap_bytes = (size_t)kk * KG_Q * (short_rows ? 10 : 12);
  size_t kg_smem = 2 * 2 * KG_Q * KG_LD * sizeof(double) + heap_bytes;
  // TMA path: operand boxes by cp.async.bulk.tensor into a 3-deep (2 when the candidate heaps
  // are large) ring; needs 16-byte aligned rows (d even) and at least one full box of features
  CUtensorMap tmq, tmx;
  memset(&tmq, 0, sizeof(tmq));
  memset(&tmx, 0, sizeof(tmx));
  int nst = 0;
  if (vec && d >= KG_TF && !g_knn_tma_off && n < (1LL << 31) && q < (1LL << 31) &&
      make_operand_map(&tmq, queries, q, d) && make_operand_map(&tmx, train, n, d)) {
    const size_t ring = 2 * KG_Q * KG_TF * sizeof(double);  // one stage: Q box + X box
    const size_t rest = (KG_Q * (KG_X + 1) + 1) * sizeof(double) + heap_bytes;
    const size_t half_sm = (size_t)max_smem_optin() / 2 - 2048;  // two CTAs per SM
    nst = (3 * ring + rest + 1024 <= half_sm) ? 3 : 2;
    kg_smem = nst * ring + rest + 1024;  // + slack to align the ring to 1024 bytes
  }
#define MGP_KG_LAUNCH(VEC_, IDX_, NST_)                                                        \
  do {                                                                                         \
    cudaFuncSetAttribute(knn_gram_filter_kernel<VEC_, IDX_, NST_>,                             \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kg_smem);           \
    knn_gram_filter_kernel<VEC_, IDX_, NST_><<<grid, 256, kg_smem, s>>>(                       \
        train, n, queries, q, d, kk, self_idx, w.xn, w.qn, split_len, ns, w.part_idx,          \
        w.part_s, w.cand_idx, w.cand_s, tmq, tmx);                                             \
  } while (0)
  if (nst == 3 && short_rows) MGP_KG_LAUNCH(true, unsigned short, 3);
  else if (nst == 3) MGP_KG_LAUNCH(true, int, 3);
  else if (nst == 2 && short_rows) MGP_KG_LAUNCH(true, unsigned short, 2);
  else if (nst == 2) MGP_KG_LAUNCH(true, int, 2);
  else if (vec && short_rows) MGP_KG_LAUNCH(true, unsigned short, 0);
  else if (vec) MGP_KG_LAUNCH(true, int, 0);
  else if (short_rows) MGP_KG_LAUNCH(false, unsigned short, 0);
  else MGP_KG_LAUNCH(false, int, 0);
#undef MGP_KG_LAUNCH
  if (ns > 1) {
    if (kk <= 48)
      knn_gram_merge_kernel<48><<<(unsigned)((q + 127) / 128), 128, 0, s>>>(
          w.part_idx, w.part_s, q, ns, kk, w.cand_idx, w.cand_s);
    else
      knn_gram_merge_kernel<KG_KMAX><<<(unsigned)((q + 127) / 128), 128, 0, s>>>(
          w.part_idx, w.part_s, q, ns, kk, w.cand_idx, w.cand_s);
  }
  int rc = check_launch("knn_gram_filter_kernel");
  if (rc != MGP_OK) return rc;
  knn_gram_refine_kernel<<<(unsigned)((q + 3) / 4), 128, 0, s>>>(
      train, n, queries, q, d, k, kk, w.cand_idx, w.cand_s, w.qn, w.xn_max, out_idx, out_d2,
      w.flag_list, w.flag_count);
  rc = check_launch("knn_gram_refine_kernel");
  if (rc != MGP_OK) return rc;
  // exact sweep for the queries that could not be certified (device-side count; CTAs beyond
  // it exit immediately, so the common case costs one empty launch)
  return launch_knn_tiled(train, n, queries, q, d, k, self_idx, out_idx, out_d2, w.tiled_ws,
                          w.tiled_bytes, s, w.flag_list, w.flag_count);
}

}  // namespace mgp
