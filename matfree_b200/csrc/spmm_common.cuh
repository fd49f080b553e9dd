// Helpers shared by the CSR product kernels (spmm_csr.cu, spmm_strip.cu).
#pragma once

#include <cstdlib>

#include "internal.h"

namespace mf {
namespace {

constexpr int kCap = 1024;     // non-zeros of a chunk staged in shared memory
constexpr int kMaxRows = 256;  // rows per chunk, upper bound

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t saddr, const void* g) {
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <typename T, int VEC>
__device__ __forceinline__ void ldx(const T* __restrict__ X, int64_t off, T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = __ldg(X + off);
  } else {
    vec_load_nc<T>(X + off, v);
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void stw(T* __restrict__ W, int64_t off, const T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    __stcs(W + off, v[0]);
  } else {
    using V = typename Vec<T>::type;
    V t;
    T* e = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) e[i] = v[i];
    __stcs(reinterpret_cast<V*>(W + off), t);  // streaming: W is not re-read by this kernel
  }
}

struct SpmmParams {
  int ld;              // tile width (power of two)
  int rows_per_chunk;  // R, a multiple of the CTA sweep
  int prefetch;        // L2 prefetch of the rows one window ahead
  int l1pf;            // L1 prefetch of the gathers of the row `l1pf` sweeps ahead (0 = off)
  int window;          // throttle: a chunk may start when done + window > chunk
  int pfd;             // L2 prefetch of the CTA's own X rows `pfd` sweeps ahead (0 = off)
  // Blocked row order (0 = rows in ascending order).  For a matrix whose far diagonals are
  // `outer_stride` rows away (the planes of a 3-D stencil) and too far apart to stay in L2, the
  // rows are walked block by block: for jb: for i: rows [i * outer_stride + jb * block_rows, +
  // block_rows) -- the same row range of consecutive planes back to back, so that X[r +- stride]
  // is reused within 2 * block_rows rows instead of 2 * outer_stride.
  int64_t outer_stride;
  int block_rows;      // rows_per_chunk << block_shift; divides outer_stride
  int num_outer;       // n / outer_stride
  int block_shift;     // log2(chunks per block)
};

// first row of chunk c under the row order of `p` (R = p.rows_per_chunk)
__host__ __device__ __forceinline__ int64_t chunk_row0(int64_t c, const SpmmParams& p) {
  if (p.block_rows == 0) return c * p.rows_per_chunk;
  // 32-bit arithmetic (fewer than 2^31 chunks); chunks per block is a power of two
  const unsigned int cu = (unsigned int)c;
  const unsigned int sb = cu >> p.block_shift;
  const unsigned int w = cu & ((1u << p.block_shift) - 1u);
  const unsigned int jb = sb / (unsigned int)p.num_outer;
  const unsigned int i = sb - jb * (unsigned int)p.num_outer;
  return (int64_t)i * p.outer_stride + (int64_t)jb * p.block_rows + (int64_t)w * p.rows_per_chunk;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}


// Blocked row order for stencil-like matrices whose far diagonals (`bandwidth` rows away: the
// planes of a 3-D grid) are too far apart to stay in L2 between their uses.  Measured on the 3-D
// 7-point Laplacian 256^3 at ld = 256 (profiles/r2f_spmm_3d_t256.txt): 51.5 GB of DRAM reads per
// product in ascending row order, i.e. X three times; 13.1 instead of 18.2 ms with blocks of 8 MB
// (budget 12 MB), 11.1-11.8 instead of 11.8-12.3 ms with blocks of 4 MB (budget 6 MB, the default;
// 2 MB and 16 MB blocks are slower: profiles/r2w_sweep.jsonl, r2w_sweep2.jsonl).
inline void choose_row_order(SpmmParams* prm, int64_t n, double avg, int64_t bandwidth, int64_t ld,
                             int32_t dtype) {
  static const int env_block = env_int("MF_SPMM_BLOCKED", 1);
  static const int env_block_mb = env_int("MF_SPMM_BLOCK_MB", 6);
  const int64_t R = prm->rows_per_chunk;
  const int64_t row_bytes = ld * (int64_t)dtype_size(dtype);
  prm->outer_stride = 0;
  prm->block_rows = 0;
  prm->num_outer = 0;
  prm->block_shift = 0;
  if (!env_block || bandwidth <= 0 || avg > 8.0 || n % bandwidth != 0 || n / bandwidth <= 2 ||
      n / bandwidth >= (1ll << 30) || n / R >= (1ll << 31) || 2 * bandwidth * row_bytes <= (40ll << 20))
    return;
  // largest block of R << q rows that divides the stride and stays under the budget
  const int64_t target = ((int64_t)env_block_mb << 20) / row_bytes;
  int q = -1;
  for (int t = 0; t < 24 && (R << t) <= target && (R << t) <= bandwidth; ++t)
    if (bandwidth % (R << t) == 0) q = t;
  if (q >= 0) {
    prm->outer_stride = bandwidth;
    prm->block_rows = (int)(R << q);
    prm->num_outer = (int)(n / bandwidth);
    prm->block_shift = q;
  }
}

}  // namespace
}  // namespace mf
