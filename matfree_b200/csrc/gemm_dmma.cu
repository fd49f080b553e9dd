// K2d / K2g for x64 operators -- dense operator times probe block on the FP64 tensor cores:
// C[M][ld] = colscale .* (op(A) @ B[K][ld]), fp64 in, fp64 accumulate, fp64 out.
//
// Replaces the fp64 `dot_general` XLA emits for a user matvec `A @ v` under vmap when
// `jax_enable_x64` is on (matfree/stochtrace.py:47-49; the reference's tests run in x64).
//
// sm_100a has no tcgen05 kind for fp64; the FP64 tensor pipe is reached with
// `mma.sync.aligned.m8n8k4.row.col.f64` (SASS: DMMA.8x8x4 -- the wider PTX shapes m16n8k{4,8,16}
// are lowered to the same instruction on this architecture, checked with cuobjdump).
//
// Shape of the kernel: CTA tile 128 x BN (BN = 64 for ld >= 64, else 32), 8 warps, warp tile
// (MT*8) x 32; BK = 16 per stage, 3-stage cp.async pipeline with zero-filled tails so M and K are
// arbitrary; shared-memory row strides are = 4 (mod 16) doubles, which makes every fragment load
// (8 rows x 4 k per half-warp pair) bank-conflict free.
#include "internal.h"

namespace mf {
namespace {

constexpr int DBM = 128, DBK = 16, DSTAGES = 3;
constexpr int A_STRIDE = DBK + 4;    // trans = 0: As[m][k], 20 doubles per row
constexpr int AT_STRIDE = DBM + 4;   // trans = 1: As[k][m], 132 doubles per row

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool TRANS, int BN>
struct Smem {
  static constexpr int B_STRIDE = BN + 4;
  static constexpr int A_ELEMS = TRANS ? DBK * AT_STRIDE : DBM * A_STRIDE;
  static constexpr int B_ELEMS = DBK * B_STRIDE;
  static constexpr int STAGE = A_ELEMS + B_ELEMS;
  static constexpr size_t BYTES = (size_t)DSTAGES * STAGE * sizeof(double);
};

template <bool TRANS, int BN>
__global__ void __launch_bounds__(256)
gemm_dmma_kernel(const double* __restrict__ A, int64_t lda, int64_t M, int64_t K,
                 const double* __restrict__ B, const double* __restrict__ colscale,
                 double* __restrict__ C, int ld) {
  using S = Smem<TRANS, BN>;
  constexpr int WARPS_N = BN / 32, WARPS_M = 8 / WARPS_N;
  constexpr int MT = DBM / WARPS_M / 8;  // m8 tiles per warp
  constexpr int NT = 4;                  // n8 tiles per warp (32 columns)
  extern __shared__ __align__(16) double smem[];

  const int64_t m0 = (int64_t)blockIdx.x * DBM;
  const int n0 = blockIdx.y * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp / WARPS_N) * (MT * 8), wn = (warp % WARPS_N) * 32;
  const int g = lane >> 2, q = lane & 3;  // fragment row / k (A), column / k (B)

  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (int)((K + DBK - 1) / DBK);

  auto load_stage = [&](int stage, int kt) {
    double* As = smem + (size_t)stage * S::STAGE;
    double* Bs = As + S::A_ELEMS;
    const int64_t k0 = (int64_t)kt * DBK;
    if (!TRANS) {
      // 128 rows x 16 k = 1024 chunks of 2 doubles
#pragma unroll
      for (int t = 0; t < (DBM * DBK / 2) / 256; ++t) {
        const int idx = threadIdx.x + t * 256;
        const int mi = idx / (DBK / 2), kc = (idx % (DBK / 2)) * 2;
        const int64_t m = m0 + mi, kk = k0 + kc;
        int bytes = 0;
        if (m < M && kk < K) bytes = (kk + 1 < K) ? 16 : 8;
        const double* src = bytes ? A + m * lda + kk : A;
        cp_async16_zfill(As + mi * A_STRIDE + kc, src, bytes);
      }
    } else {
      // 16 k-rows x 128 m = 1024 chunks of 2 doubles
#pragma unroll
      for (int t = 0; t < (DBM * DBK / 2) / 256; ++t) {
        const int idx = threadIdx.x + t * 256;
        const int ki = idx / (DBM / 2), mc = (idx % (DBM / 2)) * 2;
        const int64_t m = m0 + mc, kk = k0 + ki;
        int bytes = 0;
        if (kk < K && m < M) bytes = (m + 1 < M) ? 16 : 8;
        const double* src = bytes ? A + kk * lda + m : A;
        cp_async16_zfill(As + ki * AT_STRIDE + mc, src, bytes);
      }
    }
    // B tile: 16 k-rows x BN columns
#pragma unroll
    for (int t = 0; t < (DBK * BN / 2 + 255) / 256; ++t) {
      const int idx = threadIdx.x + t * 256;
      if (idx < DBK * BN / 2) {
        const int ki = idx / (BN / 2), nc = (idx % (BN / 2)) * 2;
        const int64_t kk = k0 + ki;
        const int bytes = kk < K ? 16 : 0;
        const double* src = bytes ? B + kk * ld + n0 + nc : B;
        cp_async16_zfill(Bs + ki * S::B_STRIDE + nc, src, bytes);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < DSTAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<DSTAGES - 2>();
    __syncthreads();
    // refill the stage that was consumed in the previous iteration
    {
      const int nxt = kt + DSTAGES - 1;
      if (nxt < nk) load_stage(nxt % DSTAGES, nxt);
      cp_async_commit();
    }
    const double* As = smem + (size_t)(kt % DSTAGES) * S::STAGE;
    const double* Bs = As + S::A_ELEMS;
#pragma unroll
    for (int ks = 0; ks < DBK; ks += 4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int m = wm + i * 8 + g;
        a[i] = TRANS ? As[(ks + q) * AT_STRIDE + m] : As[m * A_STRIDE + ks + q];
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = Bs[(ks + q) * S::B_STRIDE + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row g][cols 2q, 2q+1] of every 8x8 tile
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t m = m0 + wm + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int n = n0 + wn + j * 8 + 2 * q;
      double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
      if (colscale) {
        v.x *= colscale[n];
        v.y *= colscale[n + 1];
      }
      *reinterpret_cast<double2*>(C + m * ld + n) = v;
    }
  }
}

template <bool TRANS, int BN>
int32_t launch_one(const double* A, int64_t lda, int64_t M, int64_t K, const double* B,
                   const double* colscale, double* C, int ld, cudaStream_t st) {
  using S = Smem<TRANS, BN>;
  static bool configured = false;  // idempotent attribute; a race only repeats the call
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_dmma_kernel<TRANS, BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_dmma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MF_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)((M + DBM - 1) / DBM), (unsigned)(ld / BN));
  gemm_dmma_kernel<TRANS, BN><<<grid, 256, S::BYTES, st>>>(A, lda, M, K, B, colscale, C, ld);
  return check_launch("gemm_dmma");
}

}  // namespace

bool dmma_gemm_supported(const void* A, int64_t lda, int64_t M, int64_t K, const void* B,
                         const void* C, int64_t ld, int32_t dtype) {
  if (dtype != MF_F64) return false;
  if (ld < 32 || ld % 32 != 0) return false;
  if (lda % 2 != 0) return false;  // 16-byte cp.async chunks of A
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return false;
  return M > 0 && K > 0;
}

int32_t launch_gemm_dmma(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                         const void* B, const void* colscale, void* C, int64_t ld,
                         cudaStream_t st) {
  MF_KSCOPE(MF_KC_GEMM, st);
  const double* a = (const double*)A;
  const double* b = (const double*)B;
  const double* s = (const double*)colscale;
  double* c = (double*)C;
  if (ld >= 64) {
    return trans ? launch_one<true, 64>(a, lda, M, K, b, s, c, (int)ld, st)
                 : launch_one<false, 64>(a, lda, M, K, b, s, c, (int)ld, st);
  }
  return trans ? launch_one<true, 32>(a, lda, M, K, b, s, c, (int)ld, st)
               : launch_one<false, 32>(a, lda, M, K, b, s, c, (int)ld, st);
}

}  // namespace mf
