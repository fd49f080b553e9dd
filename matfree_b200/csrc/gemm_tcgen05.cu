// K2d / K2g -- dense operator times probe block on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA), fp32 via 3xTF32:
//
//   C[M][ld] = colscale .* (op(A) @ B[K][ld]),   op(A) = A (M x K) or A^T (A stored K x M)
//
// with every fp32 operand x carried as two TF32 planes  x = hi + lo
// (hi = rna_tf32(x), lo = rna_tf32(x - hi)) and
//
//   A @ B  ~=  A_hi B_hi + A_hi B_lo + A_lo B_hi          (fp32 accumulation in TMEM)
//
// The dropped term A_lo B_lo is O(2^-22) relative.  The operator's planes are made
// once per operator (mf_operator_split); the probe block's planes are made per call
// (split_tf32_kernel, or the epilogue of the producing GEMM for the Gram operator).
//
// Replaces the `dot_general` XLA emits for a user matvec `A @ v` / `A.T @ (A @ v)`
// under vmap (matfree/stochtrace.py:47-49; tutorials/1_log_determinants.py:19-21).
//
// Kernel shape (one CTA per (probe batch, 128-row tile of op(A)), 320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor into a ring of smem stages
//   warp 1      MMA issuer    : one lane issues tcgen05.mma.kind::tf32 (3 per k-step),
//                               tcgen05.commit releases the stage / publishes a TMEM buffer
//   warps 2..9  accumulate    : tcgen05.ld the finished TMEM buffer and add it to fp32
//               + epilogue      registers (round-to-nearest), finally scale and store (and
//                               optionally the TF32 planes of the result)
// Shared-memory tiles are in the canonical UMMA layouts:
//   K-major  (A, trans=0):  [rows][SWB bytes] with SWB-byte swizzle (128 or 64), 8-row
//                           groups SBO = 8*SWB apart; BK = SWB/4 k-values per stage
//   MN-major (A^T and B):   [MN/32 chunks][BK rows][128 bytes] with the "128-byte swizzle,
//                           32-byte atom" mode -- the only MN-major layout the TF32 MMA
//                           accepts (any other one silently yields zeros): 4-row groups
//                           SBO = 512 apart, chunks LBO = BK*128 apart
// (UMMA_K = 8 k-values per instruction.)
#include <cuda.h>

#include <mutex>

#include "internal.h"

namespace mf {
namespace {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the device.
template <bool kBackoff = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (kBackoff) __nanosleep(128);  // epilogue warps idle through the whole main loop
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (64 bit): start address >> 4 in [0,14), leading
// byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), descriptor
// version 1 in [46,48), swizzle mode in [61,64) (1 = 128 B with 32-byte atoms,
// 2 = 128 B, 4 = 64 B).
constexpr int kLayoutSw128Atom32 = 1, kLayoutSw128 = 2, kLayoutSw64 = 4;
template <int LAYOUT>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  constexpr uint64_t layout = LAYOUT;
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
// Instruction descriptor (32 bit): D fp32 (1 @ bit 4), A and B TF32 (2 @ bits 7, 10),
// A / B major-ness (bits 15 / 16; 1 = MN-major), N >> 3 @ bit 17, M >> 4 @ bit 24.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcParams {
  float* C;                // [batch][M][ld] or null
  float* Csplit;           // [batch][2][M][ld] TF32 planes of the result, or null
  const float* colscale;   // [batch][ld] or null
  int M, K, ld;
};

constexpr int kBM = 128;        // rows of op(A) per CTA = one TMEM accumulator
constexpr int kTcThreads = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 accumulate + epilogue

template <int BN, int SWB>
struct TcCfg {
  static constexpr int BM = kBM;
  static constexpr int BK = SWB / 4;      // k-values per stage (one K-major swizzle row)
  static constexpr int CH = 32;           // fp32 per MN-major chunk (128 bytes)
  static constexpr int KL = SWB == 128 ? kLayoutSw128 : kLayoutSw64;  // K-major layout id
  static constexpr int KSTEPS = BK / 8;   // UMMA_K = 8 for TF32
  static constexpr int A_PLANE = BM * BK * 4;
  static constexpr int B_PLANE = BN * BK * 4;
  static constexpr int STAGE = 2 * (A_PLANE + B_PLANE);
  static constexpr int kSmemMax = 227 * 1024;
  static constexpr int kAux = 1024 /*alignment slack*/ + 256 /*barriers, tmem slot*/;
  static constexpr int NSTAGE_RAW = (kSmemMax - kAux) / STAGE;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr int SMEM = NSTAGE * STAGE + kAux;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers
  static_assert(NSTAGE >= 2, "need at least two pipeline stages");
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  static_assert(BN % 64 == 0 || BN == 32, "each epilogue warp owns BN/2 columns in x32 pieces");
};

// The tensor core adds into its fp32 accumulator with truncation, so a long K loop in one
// TMEM accumulator drifts toward zero by ~(K/8) * 2^-24 relative (measured 8e-6 at K = 1000;
// it would be 1e-4 at C3's K = 16384 -- far outside the 1e-5 budget of the log-determinant).
// Hence the accumulation is split: the tensor core only sums FLUSH stages (FLUSH*BK k-values)
// into one of two TMEM buffers; warps 2..9 drain the finished buffer with tcgen05.ld and add
// it to fp32 registers with round-to-nearest while the tensor core fills the other buffer.
template <int BN, int SWB, bool A_MN, int FLUSH>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const TcParams p) {
  using Cfg = TcCfg<BN, SWB>;
  constexpr int BM = kBM, CH = Cfg::CH, BK = Cfg::BK, KSTEPS = Cfg::KSTEPS, NSTAGE = Cfg::NSTAGE;
  constexpr uint32_t kSboK = 8 * SWB;       // K-major: 8-row group stride
  constexpr uint32_t kLboMn = BK * 128;     // MN-major: stride between 32-wide chunks
  constexpr uint32_t kSboMn = 512;          // MN-major: stride between 4-row groups
  constexpr uint32_t kStepMn = 8 * 128;     // MN-major: 8 k-rows per MMA
  constexpr int KL = Cfg::KL, ML = kLayoutSw128Atom32;
  constexpr uint32_t kIdesc = instr_desc_tf32(128, BN, A_MN, true);
  // columns each epilogue warp owns: warps of the same lane quarter split the BN columns
  constexpr int NCOLW = BN >= 64 ? BN / 2 : BN;  // BN = 32: second warp set idles
  constexpr int NPIECE = NCOLW / 32;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + NSTAGE * Cfg::STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto accf_bar = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };      // buffer b complete
  auto acce_bar = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };  // buffer b drained
  const uint32_t slot = bars + 8u * (2 * NSTAGE + 4);
  volatile uint32_t* slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + NSTAGE * Cfg::STAGE + 8 * (2 * NSTAGE + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int batch = blockIdx.x;
  const int m0 = blockIdx.y * BM;
  const int num_it = (p.K + BK - 1) / BK;
  const int num_chunks = (num_it + FLUSH - 1) / FLUSH;
  constexpr int kDrainWarps = BN >= 64 ? 8 : 4;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), kDrainWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < num_it; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t fb = full_bar(s);
        mbar_expect_tx(fb, Cfg::STAGE);
        const uint32_t sA = base + s * Cfg::STAGE;
        const uint32_t sB = sA + 2 * Cfg::A_PLANE;
        const int k0 = it * BK;
        if (!A_MN) {
          tma_load_3d(sA, &tmA, fb, k0, m0, 0);
          tma_load_3d(sA + Cfg::A_PLANE, &tmA, fb, k0, m0, 1);
        } else {
#pragma unroll
          for (int c = 0; c < BM / CH; ++c) {
            tma_load_3d(sA + c * kLboMn, &tmA, fb, m0 + c * CH, k0, 0);
            tma_load_3d(sA + Cfg::A_PLANE + c * kLboMn, &tmA, fb, m0 + c * CH, k0, 1);
          }
        }
#pragma unroll
        for (int c = 0; c < BN / CH; ++c) {
          tma_load_4d(sB + c * kLboMn, &tmB, fb, c * CH, k0, 0, batch);
          tma_load_4d(sB + Cfg::B_PLANE + c * kLboMn, &tmB, fb, c * CH, k0, 1, batch);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int it = 0; it < num_it; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        const int chunk = it / FLUSH;
        const int buf = chunk & 1;
        const bool chunk_first = it % FLUSH == 0;
        const bool chunk_last = (it % FLUSH == FLUSH - 1) || it == num_it - 1;
        if (chunk_first) {
          // the drain warps must have emptied this buffer (two chunks ago)
          mbar_wait(acce_bar(buf), ((uint32_t)(chunk >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sA = base + s * Cfg::STAGE;
        const uint32_t sB = sA + 2 * Cfg::A_PLANE;
        const uint32_t d = tmem + buf * BN;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          // one k-step = 8 k-values: 32 bytes along a K-major row, 8 rows of an MN-major chunk
          const uint32_t b_off = kk * kStepMn;
          const uint64_t bhi = smem_desc<ML>(sB + b_off, kLboMn, kSboMn);
          const uint64_t blo = smem_desc<ML>(sB + Cfg::B_PLANE + b_off, kLboMn, kSboMn);
          uint64_t ahi, alo;
          if (!A_MN) {
            const uint32_t a_off = kk * 32;
            ahi = smem_desc<KL>(sA + a_off, 16, kSboK);
            alo = smem_desc<KL>(sA + Cfg::A_PLANE + a_off, 16, kSboK);
          } else {
            const uint32_t a_off = kk * kStepMn;
            ahi = smem_desc<ML>(sA + a_off, kLboMn, kSboMn);
            alo = smem_desc<ML>(sA + Cfg::A_PLANE + a_off, kLboMn, kSboMn);
          }
          // small cross terms first, then the leading term
          umma_tf32(d, alo, bhi, kIdesc, (chunk_first && kk == 0) ? 0u : 1u);
          umma_tf32(d, ahi, blo, kIdesc, 1u);
          umma_tf32(d, ahi, bhi, kIdesc, 1u);
        }
        umma_commit(empty_bar(s));              // frees the stage when these MMAs have read it
        if (chunk_last) umma_commit(accf_bar(buf));  // this buffer's partial sum is complete
      }
    }
  } else if (warp - 2 < kDrainWarps) {
    // accumulate + epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 (lane = output row);
    // the two warps of a lane quarter take the lower / upper half of the columns
    const int q = warp & 3;
    const int col0 = ((warp - 2) >> 2) * NCOLW;
    float sum[NCOLW];
#pragma unroll
    for (int j = 0; j < NCOLW; ++j) sum[j] = 0.f;
    for (int chunk = 0; chunk < num_chunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait<true>(accf_bar(buf), (uint32_t)(chunk >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int pc = 0; pc < NPIECE; ++pc) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0 + pc * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[pc * 32 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(buf));
    }
    const int row = m0 + q * 32 + lane;
    if (row < p.M) {
      const float* cs = p.colscale ? p.colscale + (int64_t)batch * p.ld + col0 : nullptr;
      if (cs) {
#pragma unroll
        for (int j = 0; j < NCOLW; ++j) sum[j] *= __ldg(cs + j);
      }
      if (p.C) {
        float4* dst = reinterpret_cast<float4*>(p.C + ((int64_t)batch * p.M + row) * p.ld + col0);
#pragma unroll
        for (int j = 0; j < NCOLW / 4; ++j)
          dst[j] = make_float4(sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
      }
      if (p.Csplit) {
        float* hi_row = p.Csplit + (((int64_t)batch * 2) * p.M + row) * p.ld + col0;
        float* lo_row = hi_row + (int64_t)p.M * p.ld;
#pragma unroll
        for (int j = 0; j < NCOLW / 4; ++j) {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = rna_tf32(sum[4 * j + e]);
            lo[e] = rna_tf32(sum[4 * j + e] - hi[e]);
          }
          reinterpret_cast<float4*>(hi_row)[j] = make_float4(hi[0], hi[1], hi[2], hi[3]);
          reinterpret_cast<float4*>(lo_row)[j] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// hi/lo TF32 planes of an fp32 array: dst[0][i] = rna(src[i]), dst[1][i] = rna(src[i] - hi)
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                  int64_t count4, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(src) + i);
    float4 h, l;
    h.x = rna_tf32(x.x); l.x = rna_tf32(x.x - h.x);
    h.y = rna_tf32(x.y); l.y = rna_tf32(x.y - h.y);
    h.z = rna_tf32(x.z); l.z = rna_tf32(x.z - h.z);
    h.w = rna_tf32(x.w); l.w = rna_tf32(x.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  // scalar path (count not a multiple of 4: the lo plane is not 16-byte aligned)
  for (int64_t i = count4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += stride) {
    const float x = src[i];
    const float h = rna_tf32(x);
    hi[i] = h;
    lo[i] = rna_tf32(x - h);
  }
}

// ------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      cudaGetLastError();
  });
  return fn;
}

// swz: 128 / 64 = plain 128-byte / 64-byte swizzle (K-major tiles); 32 = 128-byte swizzle
// with 32-byte atoms (MN-major TF32 tiles)
int32_t make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swz) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("gemm_tcgen05: cuTensorMapEncodeTiled is not available from the driver");
    return MF_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  const CUtensorMapSwizzle sw = swz == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                            : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("gemm_tcgen05: cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

template <int BN, int SWB, bool A_MN, int FLUSH>
int32_t launch_cfg(const float* Aplanes, int64_t lda, int64_t a_rows, int64_t M, int64_t K,
                   const float* Bplanes, int64_t nbatch, const float* colscale, float* C,
                   float* Csplit, int64_t ld, cudaStream_t st) {
  using Cfg = TcCfg<BN, SWB>;
  constexpr int CH = Cfg::CH, BK = Cfg::BK, BM = kBM;
  CUtensorMap tmA, tmB;
  const uint64_t a_plane_bytes = (uint64_t)a_rows * (uint64_t)lda * 4u;
  if (!A_MN) {
    // A[M][K] row-major: dims (K, M, plane)
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 2};
    const uint64_t str[2] = {(uint64_t)lda * 4u, a_plane_bytes};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    MF_TRY(make_map(&tmA, Aplanes, 3, dims, str, box, SWB));
  } else {
    // A[K][M] row-major: dims (M, K, plane); one 32-wide chunk per copy
    const uint64_t dims[3] = {(uint64_t)M, (uint64_t)K, 2};
    const uint64_t str[2] = {(uint64_t)lda * 4u, a_plane_bytes};
    const uint32_t box[3] = {(uint32_t)CH, (uint32_t)BK, 1};
    MF_TRY(make_map(&tmA, Aplanes, 3, dims, str, box, 32));
  }
  {
    // B planes [batch][2][K][ld]: dims (ld, K, plane, batch)
    const uint64_t plane = (uint64_t)K * (uint64_t)ld * 4u;
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)K, 2, (uint64_t)nbatch};
    const uint64_t str[3] = {(uint64_t)ld * 4u, plane, 2 * plane};
    const uint32_t box[4] = {(uint32_t)CH, (uint32_t)BK, 1, 1};
    MF_TRY(make_map(&tmB, Bplanes, 4, dims, str, box, 32));
  }
  auto kernel = gemm_tf32x3_kernel<BN, SWB, A_MN, FLUSH>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [&] {
    attr_rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  if (attr_rc != cudaSuccess) {
    cudaGetLastError();
    set_error("gemm_tcgen05: cannot reserve %d bytes of shared memory: %s", Cfg::SMEM,
              cudaGetErrorString(attr_rc));
    return MF_ERR_CUDA;
  }
  TcParams p{C, Csplit, colscale, (int)M, (int)K, (int)ld};
  dim3 grid((unsigned)nbatch, (unsigned)((M + BM - 1) / BM));
  kernel<<<grid, kTcThreads, Cfg::SMEM, st>>>(tmA, tmB, p);
  return check_launch("gemm_tf32x3");
}

template <bool A_MN, int FLUSH>
int32_t launch_bn(int64_t ld, const float* Aplanes, int64_t lda, int64_t a_rows, int64_t M,
                  int64_t K, const float* Bplanes, int64_t nbatch, const float* colscale, float* C,
                  float* Csplit, cudaStream_t st) {
  switch (ld) {
    case 32:
      return launch_cfg<32, 128, A_MN, FLUSH>(Aplanes, lda, a_rows, M, K, Bplanes, nbatch, colscale,
                                              C, Csplit, ld, st);
    case 64:
      return launch_cfg<64, 128, A_MN, FLUSH>(Aplanes, lda, a_rows, M, K, Bplanes, nbatch, colscale,
                                              C, Csplit, ld, st);
    case 128:
      return launch_cfg<128, 128, A_MN, FLUSH>(Aplanes, lda, a_rows, M, K, Bplanes, nbatch,
                                               colscale, C, Csplit, ld, st);
    case 256:
      return launch_cfg<256, 128, A_MN, FLUSH>(Aplanes, lda, a_rows, M, K, Bplanes, nbatch,
                                               colscale, C, Csplit, ld, st);
  }
  set_error("gemm_tcgen05: ld=%lld unsupported", (long long)ld);
  return MF_ERR_UNSUPPORTED;
}

}  // namespace

// A plain (unswizzled) 2-D fp32 tensor map: dim0 contiguous, dim1 `stride1_bytes` apart; boxes
// that reach past the tensor are zero-filled.  Used by the fused CGS kernel (blockvec.cu) to fetch
// a tile of ALL basis vectors with one TMA instruction.  `map` points at a CUtensorMap.
int32_t encode_plain_map_2d(void* map, const void* base, uint64_t dim0, uint64_t dim1,
                            uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MF_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {dim0, dim1};
  const cuuint64_t gstr[1] = {stride1_bytes};
  const cuuint32_t bx[2] = {box0, box1}, es[2] = {1, 1};
  const CUresult rc = fn((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                         gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (plain 2-D) failed with CUresult %d", (int)rc);
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

bool tc_gemm_supported(int64_t lda, int64_t M, int64_t K, int64_t ld, int32_t dtype) {
  if (dtype != MF_F32) return false;
  if (ld != 32 && ld != 64 && ld != 128 && ld != 256) return false;
  if (lda % 4 != 0) return false;  // TMA: global strides are multiples of 16 bytes
  if (M <= 0 || K <= 0 || M > INT32_MAX || K > INT32_MAX) return false;
  return true;
}

int32_t launch_split_tf32(const void* src, void* planes, int64_t count, cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  if (count <= 0) return MF_OK;
  float* hi = (float*)planes;
  float* lo = hi + count;
  // the lo plane starts at hi + count: vector stores need count % 4 == 0
  const int64_t count4 = (count % 4 == 0) ? count / 4 : 0;
  int64_t blocks = ((count4 ? count4 : count) + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)src, hi, lo, count4, count);
  return check_launch("split_tf32");
}

// C[batch][M][ld] = colscale .* (op(A) @ B), 3xTF32 on tcgen05.
//   Aplanes: [2][a_rows][lda] (hi, lo) of A; a_rows = M (trans=0) or K (trans=1)
//   Bplanes: [batch][2][K][ld] (hi, lo) of the probe blocks
//   C and/or Csplit ([batch][2][M][ld]) receive the result
// variant: how many 32-k stages the tensor core sums before the partial result is moved to
// the fp32 register accumulators: 0 -> 2 stages (default), 1 -> 1 stage, 2 -> 4 stages
int32_t launch_gemm_tcgen05(const void* Aplanes, int64_t lda, bool trans, int64_t M, int64_t K,
                            const void* Bplanes, int64_t nbatch, const void* colscale, void* C,
                            void* Csplit, int64_t ld, int variant, cudaStream_t st) {
  MF_KSCOPE(MF_KC_GEMM, st);
  const float* A = (const float*)Aplanes;
  const float* B = (const float*)Bplanes;
  const float* cs = (const float*)colscale;
  const int64_t a_rows = trans ? K : M;
#define MF_TC(TR, FL) \
  launch_bn<TR, FL>(ld, A, lda, a_rows, M, K, B, nbatch, cs, (float*)C, (float*)Csplit, st)
  if (variant == 1) return trans ? MF_TC(true, 1) : MF_TC(false, 1);
  if (variant == 2) return trans ? MF_TC(true, 4) : MF_TC(false, 4);
  return trans ? MF_TC(true, 2) : MF_TC(false, 2);
#undef MF_TC
}

}  // namespace mf
