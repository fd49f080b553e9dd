// K2s -- CSR sparse matrix times probe block, W[n][ld] = s * (A @ X[n][ld]),
// with the Lanczos alpha (column sums of (X*s) .* W, matfree/decomp.py:288)
// reduced in the same kernel (last CTA writes the scalars).
//
// The user matvec of the reference (matfree/stochtrace.py:47-49) has no sparse
// implementation; under vmap XLA would gather per probe.  Here all probes of a
// tile advance together: a group of ld/VEC threads owns one row, every
// non-zero costs one 16-byte load per thread of the contiguous segment
// X[col][c0..c0+VEC), i.e. ld*sizeof(T) contiguous bytes per non-zero per row
// (1 KB at ld = 256).
//
// Structure (one wave of persistent CTAs, sequential chunks of rows):
//   * a CTA takes chunks of R consecutive rows; the CSR metadata of a chunk
//     (R+1 row pointers, then the contiguous slice of column indices / values)
//     is staged in shared memory with coalesced loads, so the only global loads
//     on the per-row critical path are the gathers of X themselves;
//   * per row the gathers are issued in predicated groups of G non-zeros, so a
//     stencil row (5 or 7 non-zeros) has all its gathers in flight at once;
//   * chunk c goes to CTA c % grid: at any time the CTAs work on one contiguous
//     window of rows, which keeps the re-used rows of X (the +-1 and +-stride
//     neighbours of a stencil) in L1/L2 and makes DRAM traffic ~ read X once,
//     write W once, stream the matrix once.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "internal.h"

namespace mf {
namespace {

constexpr int kCap = 2048;      // non-zeros of a chunk staged in shared memory
constexpr int kMaxRows = 256;   // rows per chunk, upper bound

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

template <typename T, int VEC, int G, bool FUSE_DOT>
__global__ void __launch_bounds__(kBlock)
spmm_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                const T* __restrict__ s, T* __restrict__ W, int ld, int ld_shift,
                int rows_per_chunk, int prefetch, double* __restrict__ partial, Finalize fin) {
  __shared__ int32_t s_ptr[kMaxRows + 1];
  __shared__ int32_t s_col[kCap];
  __shared__ T s_val[kCap];

  const int tpr = ld / VEC;   // threads per row
  const int rps = kBlock / tpr;  // rows per sweep of the CTA
  const int my_row = threadIdx.x / tpr;
  const int c0 = (threadIdx.x % tpr) * VEC;
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = s ? s[c0 + i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  const int64_t nchunks = (n + rows_per_chunk - 1) / rows_per_chunk;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t r0 = ch * rows_per_chunk;
    const int nr = (int)((n - r0) < rows_per_chunk ? (n - r0) : rows_per_chunk);
    __syncthreads();  // the previous chunk's readers are done with the staging buffers
    if (prefetch) {
      // Pull the rows of X this CTA will own in its NEXT chunk into L2 now: one window of
      // gridDim.x chunks ahead of the demand gathers, with no registers tied up, so the
      // gathers below (own rows and stencil neighbours alike) find their lines in L2.
      const int64_t pr0 = r0 + (int64_t)gridDim.x * rows_per_chunk;
      if (pr0 < n) {
        const int64_t pnr = (n - pr0) < rows_per_chunk ? (n - pr0) : rows_per_chunk;
        const char* pbase = reinterpret_cast<const char*>(X + (pr0 << ld_shift));
        const int64_t pbytes = (pnr << ld_shift) * (int64_t)sizeof(T);
        for (int64_t o = (int64_t)threadIdx.x * 128; o < pbytes; o += (int64_t)kBlock * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pbase + o));
      }
    }
    for (int i = threadIdx.x; i <= nr; i += kBlock) s_ptr[i] = __ldg(indptr + r0 + i);
    __syncthreads();
    const int32_t base = s_ptr[0];
    const int total = s_ptr[nr] - base;
    const int cnt = total < kCap ? total : kCap;
    for (int i = threadIdx.x; i < cnt; i += kBlock) {
      s_col[i] = __ldg(indices + base + i);
      s_val[i] = __ldg(data + base + i);
    }
    __syncthreads();

    for (int lr = my_row; lr < nr; lr += rps) {
      const int jb = s_ptr[lr] - base, je = s_ptr[lr + 1] - base;
      const int js = je < kCap ? je : kCap;  // end of the staged part of this row
      T sum[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) sum[i] = T(0);
      for (int j = jb; j < js; j += G) {
        int32_t c[G];
        T a[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const bool ok = j + u < js;
          c[u] = ok ? s_col[j + u] : -1;
          a[u] = ok ? s_val[j + u] : T(0);
        }
        T x[G][VEC];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          if (c[u] >= 0) {
            ldx<T, VEC>(X, ((int64_t)c[u] << ld_shift) + c0, x[u]);
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[u][i] = T(0);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u)
#pragma unroll
          for (int i = 0; i < VEC; ++i) sum[i] += a[u] * x[u][i];
      }
      // rows longer than the staging buffer: the rest straight from global memory
      for (int j = (jb > kCap ? jb : kCap); j < je; ++j) {
        const int32_t c = __ldg(indices + base + j);
        const T a = __ldg(data + base + j);
        T x[VEC];
        ldx<T, VEC>(X, ((int64_t)c << ld_shift) + c0, x);
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] += a * x[i];
      }
      T w[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) w[i] = sum[i] * sv[i];
      const int64_t off = ((r0 + lr) << ld_shift) + c0;
      stw<T, VEC>(W, off, w);
      if (FUSE_DOT) {
        T xo[VEC];
        ldx<T, VEC>(X, off, xo);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[0][i] += (double)(xo[i] * sv[i]) * (double)w[i];
      }
    }
  }
  if (FUSE_DOT) cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

int resident_grid(const void* kernel, int block, size_t smem, int64_t want) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) per_sm = it->second;
  }
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess ||
        per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    if (smem == 0) {  // dynamic-smem kernels change occupancy per call: do not cache
      std::lock_guard<std::mutex> lk(mu);
      cache[kernel] = per_sm;
    }
  }
  int64_t g = (int64_t)per_sm * num_sms();
  if (g > kMaxPartialCtas) g = kMaxPartialCtas;
  if (g > want) g = want;
  if (g < 1) g = 1;
  return (int)g;
}

int32_t launch_spmm_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                        int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                        void* W, int64_t ld, const Reduce* red, cudaStream_t st) {
  MF_KSCOPE(MF_KC_SPMM_CSR, st);
  if (n <= 0) return MF_OK;
  const int nv = dtype == MF_F64 ? 2 : 4;
  const int vec = ld >= nv ? nv : 1;
  const int rps = kBlock / (int)(ld / vec);
  int ld_shift = 0;
  while ((1ll << ld_shift) < ld) ++ld_shift;
  // rows per chunk: a multiple of the CTA sweep, sized so the chunk's non-zeros fit the
  // staging buffer on average, at most kMaxRows.
  static const int env_rows = env_int("MF_SPMM_ROWS", 0);
  static const int env_group = env_int("MF_SPMM_GROUP", 0);
  static const int env_prefetch = env_int("MF_SPMM_PREFETCH", 1);
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  int64_t R = env_rows > 0 ? env_rows : 64;
  const int64_t fit = (int64_t)(kCap / (avg > 1.0 ? avg : 1.0));
  if (R > fit) R = fit;
  R = R / rps * rps;
  if (R < rps) R = rps;
  if (R > kMaxRows) R = kMaxRows / rps * rps;
  if (R < 1 || R > kMaxRows) R = rps <= kMaxRows ? rps : kMaxRows;  // rps == 256 at ld == 1
  const int64_t nchunks = (n + R - 1) / R;
  const int group = env_group > 0 ? env_group : (avg > 4.5 ? 8 : 4);
  Finalize fin{};
  double* partial = nullptr;
  if (red) {
    fin = red->fin;
    partial = red->partial;
  }
#define MF_SPMM_G(T, VEC, G)                                                                   \
  do {                                                                                         \
    if (red) {                                                                                 \
      auto kern = spmm_csr_kernel<T, VEC, G, true>;                                            \
      const int grid = resident_grid((const void*)kern, kBlock, 0, nchunks);                   \
      kern<<<grid, kBlock, 0, st>>>(indptr, indices, (const T*)data, n, (const T*)X,           \
                                    (const T*)s, (T*)W, (int)ld, ld_shift, (int)R,             \
                                    env_prefetch, partial, fin);                                                    \
    } else {                                                                                   \
      auto kern = spmm_csr_kernel<T, VEC, G, false>;                                           \
      const int grid = resident_grid((const void*)kern, kBlock, 0, nchunks);                   \
      kern<<<grid, kBlock, 0, st>>>(indptr, indices, (const T*)data, n, (const T*)X,           \
                                    (const T*)s, (T*)W, (int)ld, ld_shift, (int)R,             \
                                    env_prefetch, nullptr, fin);                                                    \
    }                                                                                          \
  } while (0)
#define MF_SPMM(T, VEC)                                 \
  do {                                                  \
    if (group >= 8) MF_SPMM_G(T, VEC, 8);               \
    else MF_SPMM_G(T, VEC, 4);                          \
  } while (0)
  if (dtype == MF_F32) {
    if (vec == 4) MF_SPMM(float, 4); else MF_SPMM(float, 1);
  } else {
    if (vec == 2) MF_SPMM(double, 2); else MF_SPMM(double, 1);
  }
#undef MF_SPMM
#undef MF_SPMM_G
  return check_launch("spmm_csr");
}

}  // namespace mf
