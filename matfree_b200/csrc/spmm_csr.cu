// K2s -- CSR sparse matrix times probe block, W[n][ld] = s * (A @ X[n][ld]),
// with the Lanczos alpha (column sums of (X*s) .* W, matfree/decomp.py:288)
// fused into the epilogue.
//
// The user matvec of the reference (matfree/stochtrace.py:47-49) has no sparse
// implementation; under vmap XLA would gather per probe.  Here all probes of a
// tile advance together: a group of ld/VEC threads owns one row, every
// non-zero costs one broadcast load of (col, val) and one 16-byte load per
// thread of the contiguous segment X[col][c0..c0+VEC), i.e. ld*sizeof(T)
// contiguous bytes per non-zero per row.
#include "internal.h"

namespace mf {
namespace {

template <typename T, int VEC>
__device__ __forceinline__ void ldx(const T* __restrict__ X, int64_t off, T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = __ldg(X + off);
  } else {
    vec_load_nc<T>(X + off, v);
  }
}

template <typename T, int VEC, bool FUSE_DOT>
__global__ void __launch_bounds__(kBlock)
spmm_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                const T* __restrict__ s, T* __restrict__ W, int ld,
                double* __restrict__ partial) {
  const int tpr = ld / VEC;                 // threads per row
  const int rows_per_sweep = kBlock / tpr;  // rows a CTA handles per sweep
  const int my_row = threadIdx.x / tpr;
  const int c0 = (threadIdx.x % tpr) * VEC;
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = s ? s[c0 + i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  for (int64_t row = (int64_t)blockIdx.x * rows_per_sweep + my_row; row < n;
       row += (int64_t)gridDim.x * rows_per_sweep) {
    const int32_t jb = __ldg(indptr + row), je = __ldg(indptr + row + 1);
    T sum[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sum[i] = T(0);
    int32_t j = jb;
    // 4 non-zeros in flight per thread
    for (; j + 4 <= je; j += 4) {
      int32_t c[4];
      T a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = __ldg(indices + j + u);
        a[u] = __ldg(data + j + u);
      }
      T x[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) ldx<T, VEC>(X, (int64_t)c[u] * ld + c0, x[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] += a[u] * x[u][i];
    }
    for (; j < je; ++j) {
      const int32_t c = __ldg(indices + j);
      const T a = __ldg(data + j);
      T x[VEC];
      ldx<T, VEC>(X, (int64_t)c * ld + c0, x);
#pragma unroll
      for (int i = 0; i < VEC; ++i) sum[i] += a * x[i];
    }
    T w[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) w[i] = sum[i] * sv[i];
    const int64_t off = row * ld + c0;
    if constexpr (VEC == 1) {
      W[off] = w[0];
    } else {
      vec_store<T>(W + off, w);
    }
    if (FUSE_DOT) {
      T xo[VEC];
      ldx<T, VEC>(X, off, xo);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[0][i] += (double)(xo[i] * sv[i]) * (double)w[i];
    }
  }
  if (FUSE_DOT) cta_reduce_columns<VEC, 1>(acc, ld, partial, 0);
}

}  // namespace

int32_t launch_spmm_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                        int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                        void* W, int64_t ld, double* partial, int* grid_out, cudaStream_t st) {
  MF_KSCOPE(MF_KC_SPMM_CSR, st);
  (void)nnz;
  if (n <= 0) return MF_OK;
  const int nv = dtype == MF_F64 ? 2 : 4;
  const int vec = ld >= nv ? nv : 1;
  // Same grid as the other reducing kernels so that partial rows line up:
  // a CTA sweep covers kBlock*vec flat elements = kBlock*vec/ld rows.
  const int grid = reduce_grid(n * ld, vec);
#define MF_SPMM(T, VEC)                                                                       \
  do {                                                                                        \
    if (partial)                                                                              \
      spmm_csr_kernel<T, VEC, true><<<grid, kBlock, 0, st>>>(                                 \
          indptr, indices, (const T*)data, n, (const T*)X, (const T*)s, (T*)W, (int)ld,       \
          partial);                                                                           \
    else                                                                                      \
      spmm_csr_kernel<T, VEC, false><<<grid, kBlock, 0, st>>>(                                \
          indptr, indices, (const T*)data, n, (const T*)X, (const T*)s, (T*)W, (int)ld,       \
          nullptr);                                                                           \
  } while (0)
  if (dtype == MF_F32) {
    if (vec == 4) MF_SPMM(float, 4); else MF_SPMM(float, 1);
  } else {
    if (vec == 2) MF_SPMM(double, 2); else MF_SPMM(double, 1);
  }
#undef MF_SPMM
  if (grid_out) *grid_out = grid;
  return check_launch("spmm_csr");
}

}  // namespace mf
