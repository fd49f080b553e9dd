// Support kernels of the Lanczos / Arnoldi ADJOINTS (matfree/decomp.py:295-348, 480-600; the
// custom VJPs of tridiag_sym / hessenberg).  The backward recurrences are driven from the host
// layer (matfree_b200/adjoint.py) with the block-vector kernels of blockvec.cu (dots, CGS
// projections, basis combinations) plus the two kernels here:
//   * lincomb:  out = sum_t coef_t (.) V_t for up to kLinTerms block vectors, per-column device
//     coefficients -- the adjoint steps  lambda = -xi + mu x+ + nu x  and
//     xi = -dx - A lambda + a lambda + b lambda+ - b nu x+   (decomp.py:342,349) in one pass each;
//   * sddmm_csr: the parameter gradient of a CSR operator.  The reference accumulates
//     vjp(p -> matvec(v, p)) per step (decomp.py:345-346,588-590); for matvec(v, data) =
//     CSR(data) @ v that VJP is the outer product cot arg^T restricted to the sparsity pattern:
//     d data[j] = sum_i C[row_j][i] * G[col_j][i] over the k steps (C, G: blocked [n][ld]).
#include "internal.h"

namespace mf {
namespace {

constexpr int kLinTerms = 6;
struct LinArgs {
  const void* v[kLinTerms];  // block vectors [n][ld]
  const void* c[kLinTerms];  // per-column coefficients [ld] (dtype) or null (= 1)
  double hc[kLinTerms];      // host multipliers
  int nterms;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
lincomb_kernel(LinArgs a, T* __restrict__ out, int64_t total, int ld) {
  T coef[kLinTerms][VEC];
#pragma unroll
  for (int t = 0; t < kLinTerms; ++t) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      T c = T(0);
      if (t < a.nterms) {
        c = (T)a.hc[t];
        if (a.c[t] != nullptr)
          c *= reinterpret_cast<const T*>(a.c[t])[(threadIdx.x * VEC + i) & (ld - 1)];
      }
      coef[t][i] = c;
    }
  }
  for (int64_t f = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * VEC; f < total;
       f += (int64_t)gridDim.x * kBlock * VEC) {
    T acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = T(0);
#pragma unroll
    for (int t = 0; t < kLinTerms; ++t) {
      if (t < a.nterms) {
        const T* __restrict__ p = reinterpret_cast<const T*>(a.v[t]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          // left to right, un-contracted: out = ((c0 v0 + c1 v1) + c2 v2) + ...
          const T term = coef[t][i] * p[f + i];
          acc[i] = t == 0 ? term : acc[i] + term;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) out[f + i] = acc[i];
  }
}

// one warp per row: lane l handles the non-zeros j = l, l + 32, ... of the row; the k products
// of a non-zero are summed in fp64
template <typename T>
__global__ void __launch_bounds__(kBlock)
sddmm_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                 const T* __restrict__ C, const T* __restrict__ G, int ld, int k, int accumulate,
                 T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t r = (int64_t)blockIdx.x * (kBlock / 32) + threadIdx.x / 32; r < n; r += warps) {
    const int32_t jb = indptr[r], je = indptr[r + 1];
    const T* __restrict__ cr = C + r * ld;
    for (int32_t j = jb + lane; j < je; j += 32) {
      const T* __restrict__ gc = G + (int64_t)indices[j] * ld;
      double s = 0.0;
      for (int i = 0; i < k; ++i) s += (double)cr[i] * (double)gc[i];
      out[j] = accumulate ? (T)((double)out[j] + s) : (T)s;
    }
  }
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" {

int32_t mf_lincomb(const void* const* vectors, const void* const* coeffs, const double* h_scales,
                   int32_t nterms, void* out, int32_t dtype, int64_t n, int64_t ld, void* stream) {
  if (vectors == nullptr || out == nullptr || nterms < 1 || nterms > kLinTerms || n < 0 ||
      !valid_ld(ld) || (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("lincomb: bad arguments (1 <= nterms <= %d, ld a power of two <= 256)", kLinTerms);
    return MF_ERR_INVALID_ARGUMENT;
  }
  LinArgs a{};
  a.nterms = nterms;
  for (int t = 0; t < nterms; ++t) {
    if (vectors[t] == nullptr) {
      set_error("lincomb: vector %d is null", t);
      return MF_ERR_INVALID_ARGUMENT;
    }
    a.v[t] = vectors[t];
    a.c[t] = coeffs ? coeffs[t] : nullptr;
    a.hc[t] = h_scales ? h_scales[t] : 1.0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MF_KSCOPE(MF_KC_OTHER, st);
  const int64_t total = n * ld;
  if (total == 0) return MF_OK;
  const int64_t want = (total + kBlock - 1) / kBlock;
  if (dtype == MF_F32) {
    auto kern = lincomb_kernel<float, 1>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>(a, (float*)out, total, (int)ld);
  } else {
    auto kern = lincomb_kernel<double, 1>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>(a, (double*)out, total, (int)ld);
  }
  return check_launch("lincomb");
}

int32_t mf_sddmm_csr(const int32_t* indptr, const int32_t* indices, int64_t n, const void* C,
                     const void* G, int64_t ld, int64_t k, int32_t accumulate, void* out_data,
                     int32_t dtype, void* stream) {
  if (!indptr || !indices || !C || !G || !out_data || n < 0 || ld < 1 || k < 0 || k > ld ||
      (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("sddmm_csr: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MF_KSCOPE(MF_KC_OTHER, st);
  if (n == 0) return MF_OK;
  const int64_t want = (n + kBlock / 32 - 1) / (kBlock / 32);
  if (dtype == MF_F32) {
    auto kern = sddmm_csr_kernel<float>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>(indptr, indices, n, (const float*)C, (const float*)G, (int)ld,
                                  (int)k, accumulate, (float*)out_data);
  } else {
    auto kern = sddmm_csr_kernel<double>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>(indptr, indices, n, (const double*)C, (const double*)G, (int)ld,
                                  (int)k, accumulate, (double*)out_data);
  }
  return check_launch("sddmm_csr");
}

}  // extern "C"
