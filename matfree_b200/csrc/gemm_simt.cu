// K2d / K2g (portable part) -- dense operator times probe block on CUDA cores:
// C[M][ld] = op(A) @ B[K][ld], op(A) = A (M x K, row-major) or A^T (A stored
// K x M).  Used for fp64 operators, for shapes the tcgen05 path does not cover
// (tiny or unaligned problems) and as the in-library cross-check of the
// tensor-core kernel.  The fp32 hot path of the dense / Gram operators is
// gemm_tcgen05.cu.
//
// Replaces the `dot_general` XLA emits for a user matvec `A @ v` under vmap
// (matfree/stochtrace.py:47-49; tutorials/1_log_determinants.py:19-21).
#include "internal.h"

namespace mf {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;  // 256 threads, 4x4 each

template <typename T, bool TRANS>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int64_t lda, int64_t M, int64_t K,
                 const T* __restrict__ B, const T* __restrict__ colscale, T* __restrict__ C,
                 int ld) {
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x % (BN / TN);  // 0..15 -> columns
  const int ty = threadIdx.x / (BN / TN);  // 0..15 -> rows
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK = 1024 elements, 4 per thread
#pragma unroll
    for (int t = 0; t < (BM * BK) / 256; ++t) {
      const int idx = threadIdx.x + t * 256;
      int mi, ki;
      if (TRANS) {
        mi = idx % BM;
        ki = idx / BM;
      } else {
        ki = idx % BK;
        mi = idx / BK;
      }
      const int64_t m = m0 + mi, kk = k0 + ki;
      T v = T(0);
      if (m < M && kk < K) v = TRANS ? A[kk * lda + m] : A[m * lda + kk];
      As[ki][mi] = v;
    }
#pragma unroll
    for (int t = 0; t < (BK * BN) / 256; ++t) {
      const int idx = threadIdx.x + t * 256;
      const int ni = idx % BN, ki = idx / BN;
      const int64_t kk = k0 + ki;
      const int nn = n0 + ni;
      Bs[ki][ni] = (kk < K && nn < ld) ? B[kk * ld + nn] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int nn = n0 + tx * TN + j;
      if (nn < ld) C[m * ld + nn] = colscale ? acc[i][j] * colscale[nn] : acc[i][j];
    }
  }
}

}  // namespace

int32_t launch_gemm_simt(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                         const void* B, const void* colscale, void* C, int64_t ld,
                         int32_t dtype, cudaStream_t st) {
  MF_KSCOPE(MF_KC_GEMM, st);
  if (M <= 0 || K <= 0) return MF_OK;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((ld + BN - 1) / BN));
#define MF_G(T, TR)                                                                     \
  gemm_simt_kernel<T, TR><<<grid, 256, 0, st>>>((const T*)A, lda, M, K, (const T*)B,       \
                                                (const T*)colscale, (T*)C, (int)ld)
  if (dtype == MF_F32) {
    if (trans) MF_G(float, true); else MF_G(float, false);
  } else {
    if (trans) MF_G(double, true); else MF_G(double, false);
  }
#undef MF_G
  return check_launch("gemm_simt");
}

int32_t launch_gemm_blocked(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                            const void* B, const void* colscale, void* C, int64_t ld,
                            int32_t dtype, cudaStream_t st) {
  // x64 operators: FP64 tensor cores (gemm_dmma.cu) whenever the shape qualifies
  if (dmma_gemm_supported(A, lda, M, K, B, C, ld, dtype))
    return launch_gemm_dmma(A, lda, trans, M, K, B, colscale, C, ld, st);
  return launch_gemm_simt(A, lda, trans, M, K, B, colscale, C, ld, dtype, st);
}

}  // namespace mf
