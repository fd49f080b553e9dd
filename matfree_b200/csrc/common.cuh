// Internal helpers shared by all translation units of libmatfree_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "../../include/matfree_b200.h"

namespace mf {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int32_t check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

// Optional per-launch timing (mf_timing_enable): brackets a launcher's kernels
// with CUDA events on the launching stream.  Off by default; costs nothing then.
struct KernelScope {
  int cls;
  cudaStream_t st;
  int slot;
  KernelScope(int cls_, cudaStream_t st_);
  ~KernelScope();
};
#define MF_KSCOPE(cls, st) ::mf::KernelScope _mf_kscope((cls), (st))

#define MF_TRY(expr)                     \
  do {                                   \
    int32_t _mf_rc = (expr);             \
    if (_mf_rc != MF_OK) return _mf_rc;  \
  } while (0)

inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }
inline bool valid_ld(int64_t ld) { return is_pow2(ld) && ld <= 256; }
inline size_t dtype_size(int32_t dtype) { return dtype == MF_F64 ? 8 : 4; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

int num_sms();

// ---------------------------------------------------------------- device helpers
constexpr int kBlock = 256;          // threads per CTA of all block-vector kernels
constexpr int kMaxPartialCtas = 1280;  // upper bound on the grid of reducing kernels

template <typename T>
struct Vec;  // 16-byte vector of T
template <>
struct Vec<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <>
struct Vec<double> {
  using type = double2;
  static constexpr int N = 2;
};

template <typename T>
__device__ __forceinline__ void vec_load(const T* p, T (&v)[Vec<T>::N]) {
  using V = typename Vec<T>::type;
  V t = *reinterpret_cast<const V*>(p);
  const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
  for (int i = 0; i < Vec<T>::N; ++i) v[i] = e[i];
}
template <typename T>
__device__ __forceinline__ void vec_load_nc(const T* p, T (&v)[Vec<T>::N]) {
  using V = typename Vec<T>::type;
  V t = __ldg(reinterpret_cast<const V*>(p));
  const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
  for (int i = 0; i < Vec<T>::N; ++i) v[i] = e[i];
}
template <typename T>
__device__ __forceinline__ void vec_store(T* p, const T (&v)[Vec<T>::N]) {
  using V = typename Vec<T>::type;
  V t;
  T* e = reinterpret_cast<T*>(&t);
#pragma unroll
  for (int i = 0; i < Vec<T>::N; ++i) e[i] = v[i];
  *reinterpret_cast<V*>(p) = t;
}

// Geometry of the flat mapping of a CTA onto a blocked vector X[n][ld]:
// with VEC elements per thread, the CTA covers `rows_per_sweep` rows per sweep
// and a thread always sees the same VEC columns.
template <int VEC>
struct Geo {
  int ld;
  int col0;            // first column of this thread
  int row_in_sweep;    // row offset of this thread inside a sweep
  int rows_per_sweep;  // kBlock * VEC / ld
  __device__ __forceinline__ explicit Geo(int ld_) : ld(ld_) {
    int e = threadIdx.x * VEC;
    col0 = e % ld;
    row_in_sweep = e / ld;
    rows_per_sweep = kBlock * VEC / ld;
  }
};

// Lanes that cooperate on one (accumulator, column) pair of a reduction: 1 for wide tiles (one
// thread per column is already parallel enough), up to a full warp for narrow ones (ld = 1:
// a single start vector, BASELINE config 4), where a serial sum would dominate the kernel.
__device__ __forceinline__ int lanes_per_pair(int npairs) {
  int l = kBlock / (npairs > 0 ? npairs : 1);
  if (l >= 32) return 32;
  int p = 1;
  while (p * 2 <= l) p *= 2;
  return p;
}
// fixed-shape xor tree over the L lanes of a group (L a power of two <= 32): deterministic
__device__ __forceinline__ double group_sum(double s, int L) {
  for (int off = L >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

// Deterministic CTA-level reduction of per-thread column accumulators.
// acc[VEC] of thread t belongs to columns col0..col0+VEC-1; the contributions to a column are
// summed in a fixed order (strided over the lanes of a group, then an xor tree).  Result is
// written to partial[blockIdx.x * ld + col].
// `sh`: kBlock * VEC * NACC doubles of shared memory (the caller's, e.g. a dead staging buffer)
template <int VEC, int NACC = 1>
__device__ __forceinline__ void cta_reduce_columns_smem(double (&acc)[NACC][VEC], int ld,
                                                        double* __restrict__ partial,
                                                        int64_t partial_stride, double* sh) {
  // threads beyond kBlock (a kernel's extra producer warp) only take part in the barriers
  const bool part = threadIdx.x < kBlock;
  if (part) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int i = 0; i < VEC; ++i) sh[a * kBlock * VEC + threadIdx.x * VEC + i] = acc[a][i];
  }
  __syncthreads();
  // flat element index e = t*VEC+i maps to column e % ld
  const int npairs = NACC * ld;
  const int L = lanes_per_pair(npairs);
  const int lane = threadIdx.x & (L - 1);
  const int groups = kBlock / L;
  for (int base = 0; base < npairs; base += groups) {
    const int idx = base + threadIdx.x / L;
    const bool live = part && idx < npairs;
    const int a = live ? idx / ld : 0, c = live ? idx % ld : 0;
    double s = 0.0;
    if (live)
      for (int e = c + lane * ld; e < kBlock * VEC; e += L * ld) s += sh[a * kBlock * VEC + e];
    s = group_sum(s, L);
    if (live && lane == 0) partial[a * partial_stride + (int64_t)blockIdx.x * ld + c] = s;
  }
  __syncthreads();
}

template <int VEC, int NACC = 1>
__device__ __forceinline__ void cta_reduce_columns(double (&acc)[NACC][VEC], int ld,
                                                   double* __restrict__ partial,
                                                   int64_t partial_stride) {
  __shared__ double sh[NACC * kBlock * VEC];
  cta_reduce_columns_smem<VEC, NACC>(acc, ld, partial, partial_stride, sh);
}

// ---------------------------------------------------------------- peer memory (multi-GPU)
// One process per GPU; every rank allocates one region [control | heap] with cudaMalloc, exports
// it with cudaIpcGetMemHandle and maps the regions of all peers (peer.cu, mf_comm_*).  Kernels of
// this library then talk to the other GPUs with plain loads / stores over NVLink:
//   * all-reduce fused into the reducing kernels: the last CTA of a reduction pushes its fp64
//     sums into a slot of every peer, signals, waits for the peers' signals and adds the slots in
//     rank order (one-shot, deterministic, identical bits on every rank) -- no NCCL launch,
//     no host round trip between the dot product and its consumers;
//   * halo exchange: boundary rows are stored straight into the neighbours' extended blocks.
constexpr int kMaxPeers = 8;
constexpr int kPeerSlotDoubles = 32768;  // capacity of one exchange (accumulators x probes)
struct PeerControl {                      // at offset 0 of every rank's region
  unsigned int red_flag[kMaxPeers];       // [src rank] = sequence number of its last pushed reduction
  unsigned int halo_flag[kMaxPeers];      // [src rank] = sequence number of its last pushed halo
  unsigned int bar_flag[kMaxPeers];       // [src rank] = sequence number of its last barrier arrival
  unsigned int red_seq, halo_seq, bar_seq;  // local: exchanges completed so far
  unsigned int error;                     // local: 1 after a wait timed out (mf_comm_status)
  unsigned int halo_ticket;               // local: CTA ticket of the halo kernel
  unsigned int pad[3];
  double slots[2][kMaxPeers][kPeerSlotDoubles];  // [seq & 1][src rank][pair]
};
struct PeerCtx {  // device-resident descriptor (one per communicator)
  int world, rank;
  PeerControl* ctl[kMaxPeers];  // ctl[p]: rank p's control block, mapped into this process
  unsigned char* heap[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}
// spin until *flag has reached `seq` (wrap-safe); gives up after 20 s and raises the error flag
// so that a missing peer turns into an error code instead of a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int seq,
                                          unsigned int* error) {
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flag) - seq) < 0) {
    if (global_ns() - t0 > 20000000000ull) {
      *error = 1u;
      break;
    }
    __nanosleep(64);
  }
}

// What the LAST CTA of a reducing kernel does with the column sums ("fused
// finalize"): every CTA writes its partial row, takes a ticket on `counter`, and
// the CTA that draws the last ticket adds the partial rows in the fixed order
// b = 0 .. gridDim.x-1 (so the result does not depend on which CTA finishes
// last) and writes the scalars the next kernel needs.  The counter is reset by
// that CTA, so one zero-initialised counter serves a whole stream of launches.
struct Finalize {
  unsigned int* counter;  // device; 0 before the launch
  int mode;               // 0: value = sum ; 1: value = sqrt(sum), inv = 1 / value
  void* value;            // T[nacc][ld] (row a at value + a*ld), may be null
  void* inv;              // T[ld], mode 1 only, may be null
  double* dbl;            // optional double[nacc][ld]: the fp64 sums
  const PeerCtx* peer;    // optional: the sums are all-reduced over the communicator's ranks
                          // (peer memory, in this kernel) before value / inv / dbl are written
  // Arnoldi bookkeeping folded into the first Gram-Schmidt reduction (matfree/decomp.py:463 and
  // :133-135), which saves two tiny launches per step: result row `alpha_row` (counted from
  // row0, the first row of this launch) is also the diagonal entry; the row before it is the
  // upper-Hessenberg twin of the previous off-diagonal, off <- (off + h) / 2.
  void* alpha_dst = nullptr;  // T[ld]
  void* offdiag = nullptr;    // T[ld]
  int alpha_row = -1;
  int row0 = 0;
};

template <typename T>
__device__ __forceinline__ void finalize_write(const Finalize& fin, int64_t i, double s, int ld) {
  if (fin.dbl) fin.dbl[i] = s;
  if (fin.mode == 0) {
    if (fin.value) reinterpret_cast<T*>(fin.value)[i] = (T)s;
    if (fin.alpha_row >= 0) {
      const int a = (int)(i / ld), c = (int)(i - (int64_t)a * ld);
      if (fin.alpha_dst && fin.row0 + a == fin.alpha_row) reinterpret_cast<T*>(fin.alpha_dst)[c] = (T)s;
      if (fin.offdiag && fin.row0 + a == fin.alpha_row - 1) {
        T* off = reinterpret_cast<T*>(fin.offdiag);
        off[c] = T(0.5) * ((T)s + off[c]);
      }
    }
  } else {
    const T v = (T)sqrt(s);
    if (fin.value) reinterpret_cast<T*>(fin.value)[i] = v;
    if (fin.inv) reinterpret_cast<T*>(fin.inv)[i] = T(1) / v;
  }
}

// The last CTA's share of a reduction.  DEEP: eight loads in flight per lane -- for the kernels of the
// narrow-tile Arnoldi loop (C4), where the serial chain of L2 round trips was most of each kernel's
// tail on small shards; elsewhere the plain loop stays (inlined after the CSR kernel's row body the
// deep variant raised its register pressure and the fused-dot product spilled: 8.1 -> 14.3 ms).
template <typename T, bool DEEP>
__device__ __forceinline__ void finalize_last_cta(int ld, double* __restrict__ partial,
                                               int64_t partial_stride, int nacc_live,
                                               const Finalize& fin, int nrows) {
  __threadfence();
  const int grid = nrows > 0 ? nrows : (int)gridDim.x;  // partial rows per accumulator
  const int npairs = nacc_live * ld;
  const int L = lanes_per_pair(npairs);
  const int lane = threadIdx.x & (L - 1);
  const int groups = kBlock / L;
  // multi-GPU: the local sums go to slot [seq & 1][my rank] of every rank instead
  const PeerCtx* pc = fin.peer;
  const bool xch = pc != nullptr && pc->world > 1;
  int world = 1, rank = 0, buf = 0;
  unsigned int seq = 0;
  if (xch) {
    world = pc->world;
    rank = pc->rank;
    seq = *(volatile unsigned int*)&pc->ctl[rank]->red_seq + 1u;
    buf = (int)(seq & 1u);
  }
  const bool part = threadIdx.x < kBlock;  // see cta_reduce_columns_smem
  for (int base = 0; base < npairs; base += groups) {
    const int idx = base + threadIdx.x / L;
    const bool live = part && idx < npairs;
    const int a = live ? idx / ld : 0, c = live ? idx % ld : 0;
    const double* p = partial + a * partial_stride + c;
    double s = 0.0;
    if constexpr (DEEP) {
      // lane l adds the rows l, l + L, ...: eight loads in flight, fixed order
      if (live) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0, t6 = 0.0, t7 = 0.0;
        int b = lane;
        for (; b + 7 * L < grid; b += 8 * L) {
          const double* pb = p + (int64_t)b * ld;
          const int64_t sL = (int64_t)L * ld;
          const double x0 = __ldcg(pb), x1 = __ldcg(pb + sL), x2 = __ldcg(pb + 2 * sL),
                       x3 = __ldcg(pb + 3 * sL), x4 = __ldcg(pb + 4 * sL), x5 = __ldcg(pb + 5 * sL),
                       x6 = __ldcg(pb + 6 * sL), x7 = __ldcg(pb + 7 * sL);
          t0 += x0; t1 += x1; t2 += x2; t3 += x3; t4 += x4; t5 += x5; t6 += x6; t7 += x7;
        }
        for (; b < grid; b += L) t0 += __ldcg(p + (int64_t)b * ld);
        s = ((t0 + t1) + (t2 + t3)) + ((t4 + t5) + (t6 + t7));
      }
    } else {
      if (live)
        for (int b = lane; b < grid; b += L) s += __ldcg(p + (int64_t)b * ld);
    }
    s = group_sum(s, L);
    if (!live || lane != 0) continue;
    if (xch) {
      for (int q = 0; q < world; ++q) pc->ctl[q]->slots[buf][rank][idx] = s;  // NVLink stores
    } else {
      finalize_write<T>(fin, (int64_t)a * ld + c, s, ld);
    }
  }
  if (xch) {
    PeerControl* mine = pc->ctl[rank];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
      st_release_sys(&pc->ctl[threadIdx.x]->red_flag[rank], seq);
      wait_flag(&mine->red_flag[threadIdx.x], seq, &mine->error);
    }
    __syncthreads();
    // every rank adds the same numbers in the same (rank) order: identical bits everywhere
    for (int idx = threadIdx.x; part && idx < npairs; idx += kBlock) {
      double s = 0.0;
      for (int q = 0; q < world; ++q) s += __ldcv(&mine->slots[buf][q][idx]);
      finalize_write<T>(fin, idx, s, ld);
    }
    if (threadIdx.x == 0) mine->red_seq = seq;
  }
  if (threadIdx.x == 0) *fin.counter = 0u;
}

// Second half of a reducing kernel: take a ticket; the CTA drawing the last one adds the
// partial rows of all CTAs (fixed order) for `nacc_live` accumulators and writes the results.
template <typename T, bool DEEP = false>
__device__ __forceinline__ void finalize_if_last(int ld, double* __restrict__ partial,
                                                 int64_t partial_stride, int nacc_live,
                                                 const Finalize& fin, int nrows = 0) {
  if (fin.counter == nullptr) return;
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(fin.counter, 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  finalize_last_cta<T, DEEP>(ld, partial, partial_stride, nacc_live, fin, nrows);
}

template <typename T, int VEC, int NACC = 1, bool DEEP = false>
__device__ __forceinline__ void cta_reduce_finalize(double (&acc)[NACC][VEC], int ld,
                                                    double* __restrict__ partial,
                                                    int64_t partial_stride, int nacc_live,
                                                    const Finalize& fin) {
  cta_reduce_columns<VEC, NACC>(acc, ld, partial, partial_stride);
  finalize_if_last<T, DEEP>(ld, partial, partial_stride, nacc_live, fin);
}

}  // namespace mf
