// K3 / K4 -- HBM-bound block-vector kernels on the blocked layout X[n][ld]:
//   * column dots, the fused Lanczos three-term update + norm
//     (matfree/decomp.py:286-292),
//   * classical Gram-Schmidt passes of the full re-orthogonalisation
//     (matfree/decomp.py:462-471),
//   * scaling / layout helpers.
// Every kernel maps a CTA of 256 threads onto flat 16-byte chunks so that a
// thread always owns the same probe columns; per-column sums are accumulated in
// fp64 registers, reduced deterministically inside the CTA and written as one
// partial row per CTA; `finalize` adds the partial rows in a fixed order.
#include <cuda.h>
#include <cstdlib>

#include "internal.h"

namespace mf {
namespace {

template <typename T, int VEC>
__device__ __forceinline__ void load_chunk(const T* __restrict__ p, int64_t f, T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = p[f];
  } else {
    vec_load<T>(p + f, v);
  }
}
// streaming (evict-first) variant for data read exactly once, so that re-read operands stay in L2
template <typename T, int VEC>
__device__ __forceinline__ void load_chunk_stream(const T* __restrict__ p, int64_t f, T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = __ldcs(p + f);
  } else {
    using V = typename Vec<T>::type;
    V t = __ldcs(reinterpret_cast<const V*>(p + f));
    const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = e[i];
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_chunk(T* __restrict__ p, int64_t f, const T (&v)[VEC]) {
  if constexpr (VEC == 1) {
    p[f] = v[0];
  } else {
    vec_store<T>(p + f, v);
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void load_cols(const T* __restrict__ s, int ld, T (&v)[VEC], T dflt) {
#pragma unroll
  for (int i = 0; i < VEC; ++i)
    v[i] = s ? s[(threadIdx.x * VEC + i) & (ld - 1)] : dflt;
}

#define MF_FLAT_LOOP(total)                                                                 \
  for (int64_t f = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * VEC; f < (total);         \
       f += (int64_t)gridDim.x * kBlock * VEC)

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
dot_kernel(const T* __restrict__ X, const T* __restrict__ sx, const T* __restrict__ Y,
           int64_t total, int ld, double* __restrict__ partial, Finalize fin) {
  T s[VEC];
  load_cols<T, VEC>(sx, ld, s, T(1));
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;
  MF_FLAT_LOOP(total) {
    T x[VEC], y[VEC];
    load_chunk<T, VEC>(X, f, x);
    load_chunk<T, VEC>(Y, f, y);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] += (double)(x[i] * s[i]) * (double)y[i];
  }
  cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
}

template <typename T, int VEC, bool HAS_PREV>
__global__ void __launch_bounds__(kBlock)
lanczos_update_kernel(const T* __restrict__ W, const T* __restrict__ Rc, const T* __restrict__ sc,
                      const T* __restrict__ a, const T* Rp, const T* __restrict__ sp,
                      const T* __restrict__ bprev, T* out, int64_t total, int ld,
                      double* __restrict__ partial, Finalize fin) {
  T s_c[VEC], al[VEC], s_p[VEC], bp[VEC];
  load_cols<T, VEC>(sc, ld, s_c, T(1));
  load_cols<T, VEC>(a, ld, al, T(0));
  if (HAS_PREV) {
    load_cols<T, VEC>(sp, ld, s_p, T(1));
    load_cols<T, VEC>(bprev, ld, bp, T(0));
  }
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;
  MF_FLAT_LOOP(total) {
    T w[VEC], rc[VEC], rp[VEC], r[VEC];
    load_chunk<T, VEC>(W, f, w);
    load_chunk<T, VEC>(Rc, f, rc);
    if (HAS_PREV) load_chunk<T, VEC>(Rp, f, rp);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      T t = w[i] - al[i] * (rc[i] * s_c[i]);
      if (HAS_PREV) t = t - bp[i] * (rp[i] * s_p[i]);
      r[i] = t;
      acc[0][i] += (double)t * (double)t;
    }
    store_chunk<T, VEC>(out, f, r);
  }
  cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
scale_kernel(const T* __restrict__ X, const T* __restrict__ s, T* __restrict__ out, int mode,
             int64_t total, int ld) {
  T sv[VEC];
  load_cols<T, VEC>(s, ld, sv, T(1));
  MF_FLAT_LOOP(total) {
    T x[VEC];
    load_chunk<T, VEC>(X, f, x);
#pragma unroll
    for (int i = 0; i < VEC; ++i) x[i] = mode == 0 ? x[i] * sv[i] : x[i] / sv[i];
    store_chunk<T, VEC>(out, f, x);
  }
}

// out (+)= (coef * s1 * s2)[column] * X : one term of  |v| * sum_j v_j y_j  (matfree/funm.py:145)
// accumulated while the Lanczos recurrence is re-run (two-pass f(A)v, no stored basis)
template <typename T, int VEC, bool FIRST>
__global__ void __launch_bounds__(kBlock)
axpy_cols_kernel(const T* __restrict__ X, const T* __restrict__ coef, const T* __restrict__ s1,
                 const T* __restrict__ s2, T* __restrict__ out, int64_t total, int ld) {
  T c[VEC], a[VEC], b[VEC];
  load_cols<T, VEC>(coef, ld, c, T(1));
  load_cols<T, VEC>(s1, ld, a, T(1));
  load_cols<T, VEC>(s2, ld, b, T(1));
#pragma unroll
  for (int i = 0; i < VEC; ++i) c[i] = c[i] * a[i] * b[i];
  MF_FLAT_LOOP(total) {
    T x[VEC], o[VEC];
    load_chunk<T, VEC>(X, f, x);
    if (FIRST) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = c[i] * x[i];
    } else {
      load_chunk<T, VEC>(out, f, o);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] += c[i] * x[i];
    }
    store_chunk<T, VEC>(out, f, o);
  }
}

// CGS dots for JB basis vectors at a time; V is re-read once per group of JB.
template <typename T, int VEC, int JB>
__global__ void __launch_bounds__(kBlock)
reorth_dots_kernel(const T* __restrict__ Q, int64_t q_stride, int j0, int nj,
                   const T* __restrict__ V, int64_t total, int ld, double* __restrict__ partial,
                   int64_t partial_stride, Finalize fin) {
  double acc[JB][VEC];
#pragma unroll
  for (int j = 0; j < JB; ++j)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[j][i] = 0.0;
  MF_FLAT_LOOP(total) {
    T v[VEC];
    load_chunk<T, VEC>(V, f, v);
#pragma unroll
    for (int j = 0; j < JB; ++j) {
      if (j < nj) {
        T q[VEC];
        load_chunk<T, VEC>(Q + (int64_t)(j0 + j) * q_stride, f, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[j][i] += (double)q[i] * (double)v[i];
      }
    }
  }
  cta_reduce_finalize<T, VEC, JB, true>(acc, ld, partial, partial_stride, nj, fin);
}

// All nq basis vectors in ONE launch (narrow tiles: the per-CTA partial rows of all nq sums fit
// the workspace).  The work is cut in two dimensions: CTA b = (group g of JB basis vectors,
// range r of rows), R ranges per group, so a thread accumulates JB x VEC sums over a LONG run of
// rows and folds them once -- with one CTA per row range and all groups in sequence (the first
// version) a thread of the 8-GPU slab folded after every 3 chunks and the warp-shuffle fold cost
// more than the dot products (ncu launch list: 2.9 us per basis vector instead of 1.3).
// V is re-read once per group from L2 (the basis -- read exactly once -- is loaded evict-first);
// the last CTA adds only R partial rows per sum.  Requires ld <= 64.
template <typename T, int VEC, int JB>
__global__ void __launch_bounds__(kBlock, 4)
reorth_dots_all_kernel(const T* __restrict__ Q, int64_t q_stride, int nq, const T* __restrict__ V,
                       int64_t total, int ld, int R, double* __restrict__ partial,
                       int64_t partial_stride, Finalize fin) {
  constexpr int NW = kBlock / 32;
  __shared__ double wsum[JB][NW][64];  // per-warp sums of this CTA's group (ld <= 64)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x / R, r = blockIdx.x - g * R;
  const int j0 = g * JB;
  const int nj = (nq - j0) < JB ? (nq - j0) : JB;
  // rows of this range: whole sweeps of kBlock chunks, so a thread keeps its columns
  const int64_t sweep = (int64_t)kBlock * VEC;
  const int64_t nsweeps = (total + sweep - 1) / sweep;
  const int64_t per = (nsweeps + R - 1) / R;
  const int64_t f_beg = (int64_t)r * per * sweep + (int64_t)threadIdx.x * VEC;
  int64_t f_end = (int64_t)(r + 1) * per * sweep;
  if (f_end > total) f_end = total;
  double acc[JB][VEC];
#pragma unroll
  for (int j = 0; j < JB; ++j)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[j][i] = 0.0;
  const T* Qg = Q + (int64_t)j0 * q_stride;
  for (int64_t f = f_beg; f < f_end; f += sweep) {
    T v[VEC], q[JB][VEC];
    load_chunk<T, VEC>(V, f, v);
#pragma unroll
    for (int j = 0; j < JB; ++j)
      if (j < nj) load_chunk_stream<T, VEC>(Qg + (int64_t)j * q_stride, f, q[j]);
#pragma unroll
    for (int j = 0; j < JB; ++j)
      if (j < nj) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[j][i] += (double)q[j][i] * (double)v[i];
      }
  }
  // fold: first the elements of a chunk that share a column (ld < VEC), then the lanes whose
  // chunks cover the same columns (ld / VEC lanes apart); fixed xor tree => deterministic
  const int lanes_per_row = ld >= VEC ? ld / VEC : 1;
#pragma unroll
  for (int j = 0; j < JB; ++j) {
    if constexpr (VEC == 4) {
      if (ld == 1) acc[j][0] = (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
      if (ld == 2) {
        acc[j][0] += acc[j][2];
        acc[j][1] += acc[j][3];
      }
    } else if constexpr (VEC == 2) {
      if (ld == 1) acc[j][0] += acc[j][1];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (i < ld) {
        double s = acc[j][i];
        for (int off = 16; off >= lanes_per_row; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        acc[j][i] = s;
      }
    }
    if (lane < lanes_per_row) {
      const int c0 = (threadIdx.x * VEC) & (ld - 1);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (i < ld) wsum[j][warp][(c0 + i) & (ld - 1)] = acc[j][i];
    }
  }
  __syncthreads();
  // this CTA's partial row r of the group's sums: the warps' sums in warp order
  for (int idx = threadIdx.x; idx < nj * ld; idx += kBlock) {
    const int j = idx / ld, c = idx - j * ld;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += wsum[j][w][c];
    partial[(int64_t)(j0 + j) * partial_stride + (int64_t)r * ld + c] = s;
  }
  finalize_if_last<T, true>(ld, partial, partial_stride, nq, fin, R);
}

template <typename T, int VEC, bool NORM>
__global__ void __launch_bounds__(kBlock)
reorth_update_kernel(const T* __restrict__ Q, int64_t q_stride, int nq, const T* __restrict__ h,
                     T* __restrict__ V, int64_t total, int ld, double* __restrict__ partial,
                     Finalize fin) {
  extern __shared__ unsigned char smem_raw[];
  T* hs = reinterpret_cast<T*>(smem_raw);  // [nq][ld]
  for (int i = threadIdx.x; i < nq * ld; i += kBlock) hs[i] = h[i];
  __syncthreads();
  const int c0 = (threadIdx.x * VEC) & (ld - 1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;
  MF_FLAT_LOOP(total) {
    T v[VEC], s[VEC];
    load_chunk<T, VEC>(V, f, v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = T(0);
    int j = 0;
    for (; j + 4 <= nq; j += 4) {
      T q0[VEC], q1[VEC], q2[VEC], q3[VEC];
      load_chunk<T, VEC>(Q + (int64_t)(j + 0) * q_stride, f, q0);
      load_chunk<T, VEC>(Q + (int64_t)(j + 1) * q_stride, f, q1);
      load_chunk<T, VEC>(Q + (int64_t)(j + 2) * q_stride, f, q2);
      load_chunk<T, VEC>(Q + (int64_t)(j + 3) * q_stride, f, q3);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int c = (c0 + i) & (ld - 1);
        s[i] += q0[i] * hs[(j + 0) * ld + c];
        s[i] += q1[i] * hs[(j + 1) * ld + c];
        s[i] += q2[i] * hs[(j + 2) * ld + c];
        s[i] += q3[i] * hs[(j + 3) * ld + c];
      }
    }
    for (; j < nq; ++j) {
      T q[VEC];
      load_chunk<T, VEC>(Q + (int64_t)j * q_stride, f, q);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[i] += q[i] * hs[j * ld + ((c0 + i) & (ld - 1))];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      v[i] = v[i] - s[i];
      if (NORM) acc[0][i] += (double)v[i] * (double)v[i];
    }
    store_chunk<T, VEC>(V, f, v);
  }
  if (NORM) cta_reduce_finalize<T, VEC, 1, true>(acc, ld, partial, 0, 1, fin);
}

// ---------------------------------------------------------------- fused CGS pass (narrow tiles)
// Second half of the first Gram-Schmidt pass and first half of the second one in ONE sweep over
// the basis (matfree/decomp.py:464 and :468):   V <- V - sum_j h_j Q_j ;  h'_j = Q_j . V  (new V).
// The two need the basis in different shapes -- the update wants, per row, all j; the dots want,
// per j, all rows -- so the tile Q[0..nq)[R rows] is staged in shared memory and each consumer
// warp pulls the vectors it owns (j = w, w + 15, ...) into REGISTERS once:
//   producer: warp 15 issues one TMA bulk copy (cp.async.bulk, 1 KB) per basis vector and tile
//             into a ring of 2..8 stages, tracked by full / empty mbarriers;
//   phase A : warp w, lane l: q_j = Qs[j][rows 4l..4l+3 and 128+4l..] for its vectors;
//             s = sum_j h_j q_j -> sp[w]
//   middle  : thread t < R adds the 15 warps' sums in warp order; V' = V - s -> global + vs
//   phase B : the same registers q_j against vs: fp64 sums per vector, kept for the whole kernel.
// One sweep over Q instead of two: CGS twice costs 3 sweeps of the basis per Arnoldi step, not 4.
// History (profiles/r1v_*, r2t_cgs.txt): the first version staged with cp.async from all warps and
// read the stage twice from shared memory; ncu showed it issue-bound (330 instructions per warp
// and 128-row tile, a third of them LDGSTS and their address arithmetic) at 0.54-0.69 of HBM.
constexpr int kCgsRows = 256;         // rows (flat elements) per tile
constexpr int kCgsConsumers = 480;    // 15 consumer warps (+ the producer = 512 threads: 128 registers each)
constexpr int kCgsThreads = kCgsConsumers + 32;  // + the producer warp
constexpr int kCgsMaxNq = 102;        // (nq + 1) KB per stage, two stages + vs + sp <= 225 KB
constexpr int kCgsJW = (kCgsMaxNq + kCgsConsumers / 32 - 1) / (kCgsConsumers / 32);
constexpr int kCgsMaxStages = 8;
constexpr int kCgsSmemMax = 225 * 1024;  // + 1 KB static (alignment) stay below the 227 KB opt-in limit
constexpr int kCgsFixedBytes = (kCgsRows + (kCgsConsumers / 32) * kCgsRows) * 4 + 2 * kCgsMaxStages * 8;

__device__ __forceinline__ uint32_t cgs_saddr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cgs_mbar_init(uint64_t* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cgs_saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cgs_mbar_expect_tx(uint64_t* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cgs_saddr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cgs_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cgs_saddr(bar)) : "memory");
}
// bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU
__device__ __forceinline__ void cgs_mbar_wait(uint64_t* bar, unsigned int parity) {
  const uint32_t a = cgs_saddr(bar);
#pragma unroll 1
  for (unsigned int spin = 0; spin < (1u << 26); ++spin) {
    unsigned int done;
    asm volatile(
        "{\n .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// one box of a 2-D tensor map (rows c0.., vectors c1..) into shared memory
__device__ __forceinline__ void cgs_tma_box(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          cgs_saddr(smem_dst)), "l"(map), "r"(cgs_saddr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cgs_consumer_sync() {  // the 16 consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kCgsConsumers) : "memory");
}

__global__ void __launch_bounds__(kCgsThreads, 1)
cgs_update_dots_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                       int nq, const float* __restrict__ h, float* __restrict__ V, int64_t total, int ld,
                       int stages, double* __restrict__ partial, int64_t partial_stride,
                       Finalize fin) {
  constexpr int R = kCgsRows, NW = kCgsConsumers / 32;
  extern __shared__ __align__(1024) float cgs_smem[];
  const int stage_floats = (nq + 1) * R;  // nq basis segments + the segment of V
  float* vs = cgs_smem + stages * stage_floats;  // [R] updated V of the current tile
  float* sp = vs + R;                            // [NW][R] the warps' shares of sum_j h_j Q_j
  uint64_t* full = reinterpret_cast<uint64_t*>(sp + NW * R);  // [stages] data of the stage has landed
  uint64_t* empty = full + kCgsMaxStages;                     // [stages] the stage may be refilled
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < stages; ++s) {
      cgs_mbar_init(&full[s], 1);
      cgs_mbar_init(&empty[s], 1);
    }
  // the barriers must be visible to the TMA unit (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int64_t ntiles = (total + R - 1) / R;

  if (warp == NW) {
    // ---- producer: ONE thread, two TMA instructions per tile -- the box [256 rows] x [nq vectors]
    // of the basis and the 256 rows of V; rows past the end arrive as zeros (and are counted).
    // (Per-vector bulk copies from the lanes of a warp serialise, ~70 cycles each: r2t_cgs.txt.)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
      int st = 0;
      unsigned int round = 0;  // how often the ring has wrapped
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (round > 0) cgs_mbar_wait(&empty[st], (round - 1) & 1u);
        cgs_mbar_expect_tx(&full[st], (unsigned int)stage_floats * 4u);
        float* dst = cgs_smem + st * stage_floats;
        cgs_tma_box(dst, &tmQ, (int)(tile * R), 0, &full[st]);
        cgs_tma_box(dst + nq * R, &tmV, (int)(tile * R), 0, &full[st]);
        if (++st == stages) {
          st = 0;
          ++round;
        }
      }
    }
  } else {
    // ---- consumer warps
    float hw[kCgsJW];  // coefficients of this warp's vectors (ld == 1)
#pragma unroll
    for (int jj = 0; jj < kCgsJW; ++jj) {
      const int jv = warp + jj * NW;
      hw[jj] = jv < nq ? h[jv] : 0.f;
    }
    double acc[kCgsJW];
#pragma unroll
    for (int jj = 0; jj < kCgsJW; ++jj) acc[jj] = 0.0;
    int st = 0;
    unsigned int round = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      cgs_mbar_wait(&full[st], round & 1u);
      const float* Qs = cgs_smem + st * stage_floats;
      float4 qa[kCgsJW], qb[kCgsJW];
      float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
#pragma unroll
      for (int jj = 0; jj < kCgsJW; ++jj) {
        const int jv = warp + jj * NW;
        qa[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
        qb[jj] = qa[jj];
        if (jv < nq) {
          qa[jj] = reinterpret_cast<const float4*>(Qs + jv * R)[lane];
          qb[jj] = reinterpret_cast<const float4*>(Qs + jv * R)[32 + lane];
          sa.x += hw[jj] * qa[jj].x;
          sa.y += hw[jj] * qa[jj].y;
          sa.z += hw[jj] * qa[jj].z;
          sa.w += hw[jj] * qa[jj].w;
          sb.x += hw[jj] * qb[jj].x;
          sb.y += hw[jj] * qb[jj].y;
          sb.z += hw[jj] * qb[jj].z;
          sb.w += hw[jj] * qb[jj].w;
        }
      }
      reinterpret_cast<float4*>(sp + warp * R)[lane] = sa;
      reinterpret_cast<float4*>(sp + warp * R)[32 + lane] = sb;
      cgs_consumer_sync();
      if (tid < R) {
        const int64_t f = tile * R + tid;
        float sall = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) sall += sp[w * R + tid];
        float v = 0.f;
        if (f < total) {
          v = Qs[nq * R + tid] - sall;
          V[f] = v;
        }
        vs[tid] = v;
      }
      cgs_consumer_sync();
      // the stage is consumed (the vectors sit in registers, the V segment has been read)
      if (tid == 0) cgs_mbar_arrive(&empty[st]);
      // phase B: lane l covers rows 4l..4l+3 and 128+4l..128+4l+3; the 8 products of a vector
      // are summed in fp32 and added to the fp64 accumulator once per tile
      const float4 va = reinterpret_cast<const float4*>(vs)[lane];
      const float4 vb = reinterpret_cast<const float4*>(vs)[32 + lane];
#pragma unroll
      for (int jj = 0; jj < kCgsJW; ++jj) {
        float p0 = qa[jj].x * va.x, p1 = qa[jj].y * va.y, p2 = qa[jj].z * va.z, p3 = qa[jj].w * va.w;
        p0 += qb[jj].x * vb.x;
        p1 += qb[jj].y * vb.y;
        p2 += qb[jj].z * vb.z;
        p3 += qb[jj].w * vb.w;
        acc[jj] += (double)((p0 + p1) + (p2 + p3));
      }
      if (++st == stages) {
        st = 0;
        ++round;
      }
    }
    // fold all lanes of the warp (ld == 1: every row belongs to the one column); fixed xor tree
#pragma unroll
    for (int jj = 0; jj < kCgsJW; ++jj) {
      const int jv = warp + jj * NW;
      double sacc = acc[jj];
      for (int off = 16; off >= 1; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
      if (jv < nq && lane == 0)
        partial[(int64_t)jv * partial_stride + (int64_t)blockIdx.x] = sacc;
    }
  }
  finalize_if_last<float, true>(ld, partial, partial_stride, nq, fin);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
basis_combine_kernel(const T* __restrict__ Q, int64_t q_stride, int k, const T* __restrict__ coef,
                     const T* __restrict__ scale, T* __restrict__ out, int64_t total, int ld) {
  T sc[VEC];
  load_cols<T, VEC>(scale, ld, sc, T(1));
  const int c0 = (threadIdx.x * VEC) & (ld - 1);
  MF_FLAT_LOOP(total) {
    T s[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = T(0);
    for (int j = 0; j < k; ++j) {
      T q[VEC];
      load_chunk<T, VEC>(Q + (int64_t)j * q_stride, f, q);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[i] += q[i] * coef[(int64_t)j * ld + ((c0 + i) & (ld - 1))];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] *= sc[i];
    store_chunk<T, VEC>(out, f, s);
  }
}

// (P, n) <-> [n][ld] through a 32x32 shared-memory tile
template <typename T, bool TO_BLOCKED>
__global__ void transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n,
                                 int64_t num_probes, int64_t ld) {
  __shared__ T tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int64_t p0 = (int64_t)blockIdx.y * 32;
  if (TO_BLOCKED) {
    // read src[p][r] coalesced in r
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int64_t p = p0 + j, r = r0 + threadIdx.x;
      tile[j][threadIdx.x] = (p < num_probes && r < n) ? src[p * n + r] : T(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int64_t r = r0 + j, p = p0 + threadIdx.x;
      if (r < n && p < ld) dst[r * ld + p] = tile[threadIdx.x][j];
    }
  } else {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int64_t r = r0 + j, p = p0 + threadIdx.x;
      tile[j][threadIdx.x] = (r < n && p < ld) ? src[r * ld + p] : T(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int64_t p = p0 + j, r = r0 + threadIdx.x;
      if (p < num_probes && r < n) dst[p * n + r] = tile[threadIdx.x][j];
    }
  }
}

// offdiag_{i-1} = (|v_{i-1}| + q_{i-1}^T A q_i) / 2   (matfree/decomp.py:133-135)
template <typename T>
__global__ void full_offdiag_kernel(T* __restrict__ beta_prev, const T* __restrict__ h_row,
                                    int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ld) beta_prev[c] = T(0.5) * (h_row[c] + beta_prev[c]);
}

// value = sum or sqrt(sum), inv = 1 / value, from fp64 sums (after an all-reduce)
template <typename T>
__global__ void sums_finalize_kernel(const double* __restrict__ sums, int64_t count, int mode,
                                     T* __restrict__ value, T* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double s = sums[i];
  const T v = mode == 0 ? (T)s : (T)sqrt(s);
  if (value) value[i] = v;
  if (inv) inv[i] = T(1) / v;
}

// Hutchinson row statistics over the probes of a tile (matfree/stochtrace.py:836-849,868-898):
// t[r][c] = A[r][c] * B[r][c]   (diagonal: A = probes, B = A v;  row norms: A = B = A v)
// rowsum[r] (+)= sum_{c < np} t,  rowsumsq[r] (+)= sum_{c < np} t^2      (fp64)
// One warp per row: coalesced over the probes, fixed xor tree => deterministic, no atomics.
template <typename T>
__global__ void __launch_bounds__(kBlock)
hutch_rows_kernel(const T* __restrict__ A, const T* __restrict__ B, int64_t n, int ld, int np,
                  int accumulate, double* __restrict__ rowsum, double* __restrict__ rowsumsq) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t r = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); r < n; r += warps) {
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane; c < np; c += 32) {
      const double t = (double)A[r * ld + c] * (double)B[r * ld + c];
      s1 += t;
      s2 += t * t;
    }
    s1 = group_sum(s1, 32);
    s2 = group_sum(s2, 32);
    if (lane == 0) {
      rowsum[r] = accumulate ? rowsum[r] + s1 : s1;
      if (rowsumsq) rowsumsq[r] = accumulate ? rowsumsq[r] + s2 : s2;
    }
  }
}


inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// all of the (possibly null) pointers 16-byte aligned and the element count a whole number of
// 16-byte chunks
inline bool wide_ok(int32_t dtype, int64_t total, const void* a, const void* b = nullptr,
                    const void* c = nullptr, const void* d = nullptr) {
  const int64_t per = dtype == MF_F64 ? 2 : 4;
  return total % per == 0 && al16(a) && al16(b) && al16(c) && al16(d);
}

}  // namespace

// Vector width of the flat mapping: 16-byte chunks whenever a row holds at least one
// (ld >= 4 fp32 / 2 fp64).  Narrow tiles (ld = 1, 2 -- single start vectors, config 4) still
// get 16-byte accesses when `mf_wide` (set by the launcher: element count a multiple of the
// chunk and all pointers 16-byte aligned): a chunk then spans several rows of the same
// columns, which the column masks (& (ld - 1)) already handle.
#define MF_DISPATCH_TV(dtype, ld, ...)                     \
  do {                                                     \
    if ((dtype) == MF_F32) {                               \
      using T = float;                                     \
      if ((ld) >= 4 || mf_wide) {                          \
        constexpr int VEC = 4;                             \
        __VA_ARGS__;                                            \
      } else {                                             \
        constexpr int VEC = 1;                             \
        __VA_ARGS__;                                            \
      }                                                    \
    } else {                                               \
      using T = double;                                    \
      if ((ld) >= 2 || mf_wide) {                          \
        constexpr int VEC = 2;                             \
        __VA_ARGS__;                                            \
      } else {                                             \
        constexpr int VEC = 1;                             \
        __VA_ARGS__;                                            \
      }                                                    \
    }                                                      \
  } while (0)

// one wave of CTAs, each looping over flat 16-byte chunks
#define MF_STREAM_GRID(kern, total, vec) \
  resident_grid((const void*)(kern), kBlock, 0, ((total) + (int64_t)kBlock * (vec)-1) / ((int64_t)kBlock * (vec)))

int32_t launch_dot(const void* X, const void* sx, const void* Y, int32_t dtype, int64_t n,
                   int64_t ld, const Reduce& red, cudaStream_t st) {
  MF_KSCOPE(MF_KC_DOT, st);
  const int64_t total = n * ld;
  const bool mf_wide = wide_ok(dtype, total, X, Y);
  MF_DISPATCH_TV(dtype, ld, {
    auto kern = dot_kernel<T, VEC>;
    const int grid = MF_STREAM_GRID(kern, total, VEC);
    kern<<<grid, kBlock, 0, st>>>((const T*)X, (const T*)sx, (const T*)Y, total, (int)ld,
                                  red.partial, red.fin);
  });
  return check_launch("dot");
}

int32_t launch_lanczos_update(const void* W, const void* Rc, const void* sc, const void* a,
                              const void* Rp, const void* sp, const void* bprev, void* out,
                              int32_t dtype, int64_t n, int64_t ld, const Reduce& red,
                              cudaStream_t st) {
  MF_KSCOPE(MF_KC_LANCZOS_UPDATE, st);
  const int64_t total = n * ld;
  const bool mf_wide = wide_ok(dtype, total, W, Rc, Rp, out);
  if (Rp != nullptr) {
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = lanczos_update_kernel<T, VEC, true>;
      const int grid = MF_STREAM_GRID(kern, total, VEC);
      kern<<<grid, kBlock, 0, st>>>((const T*)W, (const T*)Rc, (const T*)sc, (const T*)a,
                                    (const T*)Rp, (const T*)sp, (const T*)bprev, (T*)out, total,
                                    (int)ld, red.partial, red.fin);
    });
  } else {
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = lanczos_update_kernel<T, VEC, false>;
      const int grid = MF_STREAM_GRID(kern, total, VEC);
      kern<<<grid, kBlock, 0, st>>>((const T*)W, (const T*)Rc, (const T*)sc, (const T*)a, nullptr,
                                    nullptr, nullptr, (T*)out, total, (int)ld, red.partial,
                                    red.fin);
    });
  }
  return check_launch("lanczos_update");
}

int32_t launch_scale(const void* X, const void* s, void* out, int mode, int32_t dtype,
                     int64_t n, int64_t ld, cudaStream_t st) {
  MF_KSCOPE(MF_KC_SCALE, st);
  const int64_t total = n * ld;
  const bool mf_wide = wide_ok(dtype, total, X, out);
  MF_DISPATCH_TV(dtype, ld, {
    auto kern = scale_kernel<T, VEC>;
    const int grid = MF_STREAM_GRID(kern, total, VEC);
    kern<<<grid, kBlock, 0, st>>>((const T*)X, (const T*)s, (T*)out, mode, total, (int)ld);
  });
  return check_launch("scale");
}

int32_t launch_axpy_cols(const void* X, const void* coef, const void* s1, const void* s2, void* out,
                         bool first, int32_t dtype, int64_t n, int64_t ld, cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  const int64_t total = n * ld;
  const bool mf_wide = wide_ok(dtype, total, X, out);
  if (first) {
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = axpy_cols_kernel<T, VEC, true>;
      const int grid = MF_STREAM_GRID(kern, total, VEC);
      kern<<<grid, kBlock, 0, st>>>((const T*)X, (const T*)coef, (const T*)s1, (const T*)s2, (T*)out,
                                    total, (int)ld);
    });
  } else {
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = axpy_cols_kernel<T, VEC, false>;
      const int grid = MF_STREAM_GRID(kern, total, VEC);
      kern<<<grid, kBlock, 0, st>>>((const T*)X, (const T*)coef, (const T*)s1, (const T*)s2, (T*)out,
                                    total, (int)ld);
    });
  }
  return check_launch("axpy_cols");
}

int32_t launch_reorth_dots(const void* Q, int64_t nq, const void* V, int32_t dtype, int64_t n,
                           int64_t ld, double* partial, unsigned int* counter, void* h_out,
                           cudaStream_t st, double* dbl_out, int64_t partial_rows,
                           int64_t q_stride, const PeerCtx* peer, void* alpha_dst, void* offdiag) {
  MF_KSCOPE(MF_KC_REORTH_DOTS, st);
  const int64_t total = n * ld;
  if (q_stride <= 0) q_stride = total;
  const bool mf_wide = wide_ok(dtype, total, Q, V) && q_stride % (dtype == MF_F64 ? 2 : 4) == 0;
  const int64_t pstride = (int64_t)kMaxPartialCtas * ld;
  constexpr int JB = 4;
  if (nq > JB && ld <= 64 && partial_rows >= (nq + JB - 1) / JB * JB) {
    // one launch for all nq sums
    Finalize fin{counter, 0, h_out, nullptr, dbl_out, peer};
    if (alpha_dst != nullptr || offdiag != nullptr) {
      fin.alpha_dst = alpha_dst;
      fin.offdiag = offdiag;
      fin.alpha_row = (int)nq - 1;
    }
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = reorth_dots_all_kernel<T, VEC, JB>;
      const int ngroups = (int)((nq + JB - 1) / JB);
      const int64_t nsweeps = (total + (int64_t)kBlock * VEC - 1) / ((int64_t)kBlock * VEC);
      int R = resident_grid((const void*)kern, kBlock, 0, nsweeps * ngroups) / ngroups;
      if (R < 1) R = 1;
      if (R > nsweeps) R = (int)(nsweeps > 0 ? nsweeps : 1);
      kern<<<ngroups * R, kBlock, 0, st>>>((const T*)Q, q_stride, (int)nq, (const T*)V, total,
                                           (int)ld, R, partial, pstride, fin);
    });
    return check_launch("reorth_dots_all");
  }
  for (int64_t j0 = 0; j0 < nq; j0 += JB) {
    const int nj = (int)((nq - j0) < JB ? (nq - j0) : JB);
    Finalize fin{counter, 0,
                 h_out ? (void*)((char*)h_out + j0 * ld * (int64_t)dtype_size(dtype)) : nullptr,
                 nullptr, dbl_out ? dbl_out + j0 * ld : nullptr, peer};
    if (alpha_dst != nullptr || offdiag != nullptr) {
      fin.alpha_dst = alpha_dst;
      fin.offdiag = offdiag;
      fin.alpha_row = (int)nq - 1;
      fin.row0 = (int)j0;
    }
    MF_DISPATCH_TV(dtype, ld, {
      auto kern = reorth_dots_kernel<T, VEC, JB>;
      const int grid = MF_STREAM_GRID(kern, total, VEC);
      kern<<<grid, kBlock, 0, st>>>((const T*)Q, q_stride, (int)j0, nj, (const T*)V, total, (int)ld,
                                    partial, pstride, fin);
    });
    MF_TRY(check_launch("reorth_dots"));
  }
  return MF_OK;
}

int32_t launch_reorth_update(const void* Q, int64_t nq, const void* h, void* V, int32_t dtype,
                             int64_t n, int64_t ld, const Reduce* red, cudaStream_t st,
                             int64_t q_stride) {
  if (q_stride <= 0) q_stride = n * ld;
  const int64_t max_nq = (160 * 1024) / (ld * (int64_t)dtype_size(dtype));
  if (nq > max_nq) {
    // The coefficients of all nq basis vectors do not fit the shared-memory staging (e.g. depth
    // > 160 at ld = 256 in fp32): subtract the basis in groups, V -= sum_{j in group} Q[j] h[j];
    // the fused norm rides on the last group.
    const int64_t es = (int64_t)dtype_size(dtype);
    for (int64_t j0 = 0; j0 < nq; j0 += max_nq) {
      const int64_t nj = (nq - j0) < max_nq ? (nq - j0) : max_nq;
      MF_TRY(launch_reorth_update((const char*)Q + j0 * q_stride * es, nj,
                                  (const char*)h + j0 * ld * es, V, dtype, n, ld,
                                  j0 + nj == nq ? red : nullptr, st, q_stride));
    }
    return MF_OK;
  }
  MF_KSCOPE(MF_KC_REORTH_UPDATE, st);
  const int64_t total = n * ld;
  if (q_stride <= 0) q_stride = total;
  const bool mf_wide = wide_ok(dtype, total, Q, V) && q_stride % (dtype == MF_F64 ? 2 : 4) == 0;
  const size_t smem = (size_t)nq * ld * dtype_size(dtype);
  Finalize fin{};
  double* partial = nullptr;
  if (red) {
    fin = red->fin;
    partial = red->partial;
  }
#define MF_RU(NORM)                                                                           \
  MF_DISPATCH_TV(dtype, ld, {                                                                 \
    auto kern = reorth_update_kernel<T, VEC, NORM>;                                           \
    if (smem > 40 * 1024)                                                                     \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    const int grid = resident_grid((const void*)kern, kBlock, smem,                       \
                                   (total + (int64_t)kBlock * VEC - 1) / ((int64_t)kBlock * VEC)); \
    kern<<<grid, kBlock, smem, st>>>((const T*)Q, q_stride, (int)nq, (const T*)h, (T*)V, total, \
                                     (int)ld, partial, fin);                                  \
  })
  if (red != nullptr) {
    MF_RU(true);
  } else {
    MF_RU(false);
  }
#undef MF_RU
  return check_launch("reorth_update");
}

bool cgs_fused_supported(const void* Q, int64_t q_stride, int64_t nq, const void* V, int32_t dtype,
                         int64_t n, int64_t ld, int64_t partial_rows) {
  // Measured on C4 (profiles/r1q_*, r1r_cgs_full.txt, r1v_*): with 8 warps per SM the two
  // shared-memory phases were latency-bound and the fused sweep was no faster than the two sweeps
  // it replaces; with 16 warps it takes 94 ms instead of 52 + 56 ms per decomposition (3.7 TB/s --
  // still short of the 6.5 TB/s of the plain update kernel), i.e. -6 % on the whole of C4.
  // MF_CGS_FUSED_OFF=1 restores the two-kernel route (A/B runs, tests).
  const bool on = getenv("MF_CGS_FUSED_OFF") == nullptr;
  // ld == 1 (a single start vector: BASELINE config 4) -- wider tiles keep the two-kernel route
  if (!on || dtype != MF_F32 || ld != 1 || nq < 1 || nq > kCgsMaxNq) return false;
  if (partial_rows < nq) return false;  // one partial row per basis vector
  if (q_stride <= 0) q_stride = n * ld;
  // TMA tensor maps: 16-byte aligned base and vector stride, coordinates in int32
  if (q_stride % 4 != 0 || !al16(Q) || !al16(V) || n * ld >= (int64_t)1 << 31) return false;
  return n * ld >= 4 * kCgsRows;  // tiny problems: the plain kernels are launch-bound anyway
}

int32_t launch_reorth_update_dots(const void* Q, int64_t nq, const void* h, void* V, int64_t n,
                                  int64_t ld, double* partial, unsigned int* counter, void* h_out,
                                  cudaStream_t st, int64_t q_stride, const PeerCtx* peer) {
  MF_KSCOPE(MF_KC_REORTH_UPDATE, st);
  const int64_t total = n * ld;
  if (q_stride <= 0) q_stride = total;
  // as many stages as fit: stages * (nq + 1) segments + vs + sp + the barriers
  const int64_t stage_bytes = (nq + 1) * kCgsRows * (int64_t)sizeof(float);
  static const int env_stages = getenv("MF_CGS_STAGES") ? atoi(getenv("MF_CGS_STAGES")) : 0;
  int stages = (int)((kCgsSmemMax - kCgsFixedBytes) / stage_bytes);
  if (stages > kCgsMaxStages) stages = kCgsMaxStages;
  if (env_stages >= 2 && env_stages < stages) stages = env_stages;
  if (stages < 2) {
    set_error("cgs_update_dots: nq = %lld does not fit shared memory", (long long)nq);
    return MF_ERR_INVALID_ARGUMENT;
  }
  const size_t smem = (size_t)(stages * stage_bytes) + kCgsFixedBytes;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(cgs_update_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kCgsSmemMax) != cudaSuccess) {
      cudaGetLastError();
      set_error("cgs_update_dots: shared-memory opt-in failed");
      return MF_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t ntiles = (total + kCgsRows - 1) / kCgsRows;
  int grid = num_sms();
  if (grid > ntiles) grid = (int)ntiles;
  Finalize fin{counter, 0, h_out, nullptr, nullptr, peer};
  // the tile of all nq basis vectors is one TMA box: the map is per launch (its box is nq tall)
  CUtensorMap tmQ, tmV;
  MF_TRY(encode_plain_map_2d(&tmQ, Q, (uint64_t)total, (uint64_t)nq, (uint64_t)q_stride * 4, kCgsRows,
                             (uint32_t)nq));
  MF_TRY(encode_plain_map_2d(&tmV, V, (uint64_t)total, 1, (uint64_t)q_stride * 4, kCgsRows, 1));
  cgs_update_dots_kernel<<<grid, kCgsThreads, smem, st>>>(tmQ, tmV, (int)nq, (const float*)h, (float*)V,
                                                       total, (int)ld, stages, partial,
                                                       (int64_t)kMaxPartialCtas * ld, fin);
  return check_launch("cgs_update_dots");
}

int32_t launch_basis_combine(const void* Q, const void* coeffs, const void* scale,
                             int32_t dtype, int64_t n, int64_t ld, int64_t k, void* out,
                             cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  const int64_t total = n * ld;
  const bool mf_wide = wide_ok(dtype, total, Q, out);
  MF_DISPATCH_TV(dtype, ld, {
    auto kern = basis_combine_kernel<T, VEC>;
    const int grid = MF_STREAM_GRID(kern, total, VEC);
    kern<<<grid, kBlock, 0, st>>>((const T*)Q, total, (int)k, (const T*)coeffs, (const T*)scale,
                                  (T*)out, total, (int)ld);
  });
  return check_launch("basis_combine");
}

int32_t launch_transpose(const void* src, void* dst, int32_t dtype, int64_t n,
                         int64_t num_probes, int64_t ld, bool to_blocked, cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  if (n <= 0) return MF_OK;
  dim3 block(32, 8);
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((ld + 31) / 32));
  if (dtype == MF_F32) {
    if (to_blocked)
      transpose_kernel<float, true><<<grid, block, 0, st>>>((const float*)src, (float*)dst, n,
                                                            num_probes, ld);
    else
      transpose_kernel<float, false><<<grid, block, 0, st>>>((const float*)src, (float*)dst, n,
                                                             num_probes, ld);
  } else {
    if (to_blocked)
      transpose_kernel<double, true><<<grid, block, 0, st>>>((const double*)src, (double*)dst, n,
                                                             num_probes, ld);
    else
      transpose_kernel<double, false><<<grid, block, 0, st>>>((const double*)src, (double*)dst, n,
                                                              num_probes, ld);
  }
  return check_launch("transpose");
}

int32_t launch_sums_finalize(const double* sums, int64_t count, int mode, void* value, void* inv,
                             int32_t dtype, cudaStream_t st) {
  MF_KSCOPE(MF_KC_FINALIZE, st);
  if (count <= 0) return MF_OK;
  const int threads = 128;
  const int blocks = (int)((count + threads - 1) / threads);
  if (dtype == MF_F32)
    sums_finalize_kernel<float><<<blocks, threads, 0, st>>>(sums, count, mode, (float*)value,
                                                            (float*)inv);
  else
    sums_finalize_kernel<double><<<blocks, threads, 0, st>>>(sums, count, mode, (double*)value,
                                                             (double*)inv);
  return check_launch("sums_finalize");
}

int32_t launch_hutch_rows(const void* A, const void* B, int32_t dtype, int64_t n, int64_t ld,
                          int64_t num_probes, bool accumulate, double* rowsum, double* rowsumsq,
                          cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  if (n <= 0) return MF_OK;
  const int64_t want = (n + kBlock / 32 - 1) / (kBlock / 32);
  if (dtype == MF_F32) {
    auto kern = hutch_rows_kernel<float>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>((const float*)A, (const float*)B, n, (int)ld, (int)num_probes,
                                  accumulate ? 1 : 0, rowsum, rowsumsq);
  } else {
    auto kern = hutch_rows_kernel<double>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, want);
    kern<<<grid, kBlock, 0, st>>>((const double*)A, (const double*)B, n, (int)ld, (int)num_probes,
                                  accumulate ? 1 : 0, rowsum, rowsumsq);
  }
  return check_launch("hutch_rows");
}

int32_t launch_full_offdiag(void* betas_prev_row, const void* h_row, int32_t dtype, int64_t ld,
                            cudaStream_t st) {
  MF_KSCOPE(MF_KC_OTHER, st);
  const int threads = 128;
  const int blocks = (int)((ld + threads - 1) / threads);
  if (dtype == MF_F32)
    full_offdiag_kernel<float><<<blocks, threads, 0, st>>>((float*)betas_prev_row,
                                                           (const float*)h_row, (int)ld);
  else
    full_offdiag_kernel<double><<<blocks, threads, 0, st>>>((double*)betas_prev_row,
                                                            (const double*)h_row, (int)ld);
  return check_launch("full_offdiag");
}

}  // namespace mf
