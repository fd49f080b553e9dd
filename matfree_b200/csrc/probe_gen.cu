// K1 -- probe generator: Threefry-2x32 (20 rounds) in jax.random's partitionable
// counter mode, Rademacher / standard-normal samples, written straight into the
// blocked layout the Lanczos kernels consume (or the reference (P, n) layout).
//
// Replaces: matfree/stochtrace.py:957-964 (`_sampler_from_jax_random`) ->
// matfree/backend/prng.py:14-29 -> jax.random.{normal, rademacher}.
// Counter for probe p, component r is the row-major flat index p * n + r of the
// (P, n) sample array (64 bit, split into hi/lo counter words).
#include "internal.h"

namespace mf {
namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// `one` is the constant 1 passed as a kernel argument: `x1 * one + x0` forces the round
// additions onto the FMA pipe (IMAD) while the rotates and xors stay on the ALU pipe
// (SHF, LOP3).  With plain `+` all ~75 integer ops of a block land on the ALU pipe, which is
// what bounds the generator (measured 19.7 ms per 16.7M x 256 tile).
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0,
                                             uint32_t& x1, uint32_t one = 1u) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define MF_TF_ROUND(r) \
  x0 = x1 * one + x0;  \
  x1 = rotl32(x1, r);  \
  x1 ^= x0;
  MF_TF_ROUND(13) MF_TF_ROUND(15) MF_TF_ROUND(26) MF_TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  MF_TF_ROUND(17) MF_TF_ROUND(29) MF_TF_ROUND(16) MF_TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  MF_TF_ROUND(13) MF_TF_ROUND(15) MF_TF_ROUND(26) MF_TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  MF_TF_ROUND(17) MF_TF_ROUND(29) MF_TF_ROUND(16) MF_TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  MF_TF_ROUND(13) MF_TF_ROUND(15) MF_TF_ROUND(26) MF_TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
#undef MF_TF_ROUND
}

// XLA's fp32 erf_inv expansion (Giles 2010); un-contracted Horner steps so the
// result matches an IEEE mul-then-add evaluation.
__device__ __forceinline__ float erfinv_xla_f32(float x) {
  float w = -log1pf(-(x * x));
  const bool lt = w < 5.0f;
  w = lt ? w - 2.5f : sqrtf(w) - 3.0f;
  float p;
#define MF_H(c_lt, c_ge) p = __fadd_rn(lt ? c_lt : c_ge, __fmul_rn(p, w));
  p = lt ? 2.81022636e-08f : -0.000200214257f;
  MF_H(3.43273939e-07f, 0.000100950558f)
  MF_H(-3.5233877e-06f, 0.00134934322f)
  MF_H(-4.39150654e-06f, -0.00367342844f)
  MF_H(0.00021858087f, 0.00573950773f)
  MF_H(-0.00125372503f, -0.0076224613f)
  MF_H(-0.00417768164f, 0.00943887047f)
  MF_H(0.246640727f, 1.00167406f)
  MF_H(1.50140941f, 2.83297682f)
#undef MF_H
  float r = __fmul_rn(p, x);
  if (fabsf(x) == 1.0f) r = x * __int_as_float(0x7f800000);
  return r;
}

template <typename T>
__device__ __forceinline__ T sample_one(uint32_t k0, uint32_t k1, uint64_t ctr, int sampler,
                                        int flags, uint32_t one = 1u) {
  uint32_t x0 = (uint32_t)(ctr >> 32), x1 = (uint32_t)ctr;
  threefry2x32(k0, k1, x0, x1, one);
  if (sampler == MF_SAMPLER_SIGNS) {
    // +1 iff the uniform draw is < 0.5 iff the MSB of the draw is 0
    const uint32_t msb = (flags & MF_PRNG_X64_BITS) ? (x0 >> 31) : ((x0 ^ x1) >> 31);
    return msb ? T(-1) : T(1);
  }
  if (sizeof(T) == 4) {
    const uint32_t bits = x0 ^ x1;
    const float u01 = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    const float lo = -0.99999994f;  // nextafter(-1, 0)
    const float u = fmaxf(lo, __fadd_rn(__fmul_rn(u01, 2.0f), lo));
    return (T)__fmul_rn(1.41421354f, erfinv_xla_f32(u));
  } else {
    const uint64_t bits = ((uint64_t)x0 << 32) | x1;
    const double u01 = __longlong_as_double((bits >> 12) | 0x3FF0000000000000ull) - 1.0;
    const double lo = -0.99999999999999989;  // nextafter(-1, 0)
    const double u = fmax(lo, u01 * (1.0 - lo) + lo);
    return (T)(1.4142135623730951 * erfinv(u));
  }
}

// Blocked layout: flat index f = r * ld + c.  A thread owns VEC consecutive
// flat elements per sweep, so its columns never change.
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
probe_gen_blocked_kernel(T* __restrict__ out, int64_t n, int64_t row0, int64_t n_total, int ld,
                         int64_t p0, int num_probes, uint32_t k0, uint32_t k1, int sampler,
                         int flags, uint32_t one, double* __restrict__ partial, Finalize fin) {
  const int64_t total = n * (int64_t)ld;
  const int64_t stride = (int64_t)gridDim.x * kBlock * VEC;
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;
  const int e0 = threadIdx.x * VEC;
  for (int64_t f = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * VEC; f < total; f += stride) {
    const int64_t r = f / ld;  // ld is a power of two -> shift
    T v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (e0 + i) & (ld - 1);
      const int64_t rr = (VEC > 1 && ld < VEC) ? (f + i) / ld : r;
      T x = T(0);
      if (c < num_probes) {
        const uint64_t ctr = (uint64_t)(p0 + c) * (uint64_t)n_total + (uint64_t)(row0 + rr);
        x = sample_one<T>(k0, k1, ctr, sampler, flags, one);
      }
      v[i] = x;
      acc[0][i] += (double)x * (double)x;
    }
    if (VEC > 1) {
      T vv[Vec<T>::N];
#pragma unroll
      for (int i = 0; i < VEC; ++i) vv[i] = v[i];
      vec_store<T>(out + f, vv);
    } else {
      out[f] = v[0];
    }
  }
  if (partial != nullptr) cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
}

// Rademacher fast path (the SLQ default): sampler and draw width are compile-time, the
// per-column counter base (p0 + c) * n is hoisted, and |z_c|^2 needs no accumulation -- it
// is the number of rows generated (every entry is +-1), counted per thread.
template <typename T, int VEC, bool X64>
__global__ void __launch_bounds__(kBlock)
probe_gen_signs_kernel(T* __restrict__ out, int64_t n, int64_t row0, int64_t n_total, int ld,
                       int ld_shift, int64_t p0, int num_probes, uint32_t k0, uint32_t k1,
                       uint32_t one, double* __restrict__ partial, Finalize fin) {
  const int64_t total = n * (int64_t)ld;
  const int64_t stride = (int64_t)gridDim.x * kBlock * VEC;
  const int e0 = threadIdx.x * VEC;
  uint64_t cbase[VEC];
  bool live[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = (e0 + i) & (ld - 1);
    live[i] = c < num_probes;
    cbase[i] = (uint64_t)(p0 + c) * (uint64_t)n_total + (uint64_t)row0;
  }
  int64_t iters = 0;
  for (int64_t f = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * VEC; f < total; f += stride) {
    const uint64_t r = (uint64_t)(f >> ld_shift);
    T v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint64_t ctr = cbase[i] + r;
      uint32_t x0 = (uint32_t)(ctr >> 32), x1 = (uint32_t)ctr;
      threefry2x32(k0, k1, x0, x1, one);
      const uint32_t draw = X64 ? x0 : (x0 ^ x1);
      // +1 iff the MSB of the draw is 0: flip the sign bit of 1.0
      T x;
      if constexpr (sizeof(T) == 4) {
        x = __uint_as_float(0x3F800000u | (draw & 0x80000000u));
      } else {
        x = (draw >> 31) ? T(-1) : T(1);
      }
      v[i] = live[i] ? x : T(0);
    }
    vec_store<T>(out + f, v);
    ++iters;
  }
  if (partial != nullptr) {
    double acc[1][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = live[i] ? (double)iters : 0.0;
    cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
  }
}

// Reference layout (P, n): out[p * ld + r]
template <typename T>
__global__ void __launch_bounds__(kBlock)
probe_gen_pn_kernel(T* __restrict__ out, int64_t n, int64_t ld, int64_t p0, int64_t num_probes,
                    uint32_t k0, uint32_t k1, int sampler, int flags) {
  const int64_t total = n * num_probes;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  for (int64_t f = (int64_t)blockIdx.x * kBlock + threadIdx.x; f < total; f += stride) {
    const int64_t p = f / n, r = f - p * n;
    const uint64_t ctr = (uint64_t)(p0 + p) * (uint64_t)n + (uint64_t)r;
    out[p * ld + r] = sample_one<T>(k0, k1, ctr, sampler, flags);
  }
}

}  // namespace

int32_t launch_probe_gen(void* out, int32_t dtype, int32_t layout, int64_t n, int64_t ld,
                         int64_t p0, int64_t num_probes, uint32_t key0, uint32_t key1,
                         int32_t sampler, int32_t prng_flags, const Reduce* red,
                         cudaStream_t st, int64_t row0, int64_t n_total) {
  MF_KSCOPE(MF_KC_PROBE_GEN, st);
  if (n <= 0 || num_probes <= 0) return MF_OK;
  if (n_total <= 0) n_total = n;  // the block holds all rows of the sample array
  if (layout == MF_LAYOUT_BLOCKED) {
    if (!valid_ld(ld) || num_probes > ld) {
      set_error("probe_gen: blocked layout needs ld a power of two <= 256 and num_probes <= ld");
      return MF_ERR_INVALID_ARGUMENT;
    }
    const int64_t total = n * ld;
    double* partial = red ? red->partial : nullptr;
    Finalize fin{};
    if (red) fin = red->fin;
#define MF_PG(T, VEC)                                                                          \
  do {                                                                                         \
    auto kern = probe_gen_blocked_kernel<T, VEC>;                                              \
    const int grid = resident_grid((const void*)kern, kBlock, 0,                               \
                                   (total + (int64_t)kBlock * VEC - 1) / ((int64_t)kBlock * VEC)); \
    kern<<<grid, kBlock, 0, st>>>((T*)out, n, row0, n_total, (int)ld, p0, (int)num_probes, key0,   \
                                  key1, sampler, prng_flags, 1u, partial, fin);                    \
  } while (0)
#define MF_PGS(T, VEC, X64)                                                                    \
  do {                                                                                         \
    auto kern = probe_gen_signs_kernel<T, VEC, X64>;                                           \
    const int grid = resident_grid((const void*)kern, kBlock, 0,                               \
                                   (total + (int64_t)kBlock * VEC - 1) / ((int64_t)kBlock * VEC)); \
    int ld_shift = 0;                                                                          \
    while ((1ll << ld_shift) < ld) ++ld_shift;                                                 \
    kern<<<grid, kBlock, 0, st>>>((T*)out, n, row0, n_total, (int)ld, ld_shift, p0,            \
                                  (int)num_probes, key0, key1, 1u, partial, fin);              \
  } while (0)
    const bool x64 = (prng_flags & MF_PRNG_X64_BITS) != 0;
    if (sampler == MF_SAMPLER_SIGNS && dtype == MF_F32 && ld >= 4) {
      if (x64) MF_PGS(float, 4, true); else MF_PGS(float, 4, false);
    } else if (sampler == MF_SAMPLER_SIGNS && dtype == MF_F64 && ld >= 2) {
      if (x64) MF_PGS(double, 2, true); else MF_PGS(double, 2, false);
    } else if (dtype == MF_F32) {
      if (ld >= 4) MF_PG(float, 4); else MF_PG(float, 1);
    } else {
      if (ld >= 2) MF_PG(double, 2); else MF_PG(double, 1);
    }
#undef MF_PGS
#undef MF_PG
    return check_launch("probe_gen_blocked");
  }
  if (layout == MF_LAYOUT_PROBE_MAJOR) {
    if (row0 != 0 || n_total != n) {
      set_error("probe_gen: row slabs are generated in the blocked layout only");
      return MF_ERR_UNSUPPORTED;
    }
    if (ld < n) {
      set_error("probe_gen: probe-major layout needs ld >= n");
      return MF_ERR_INVALID_ARGUMENT;
    }
    const int64_t total = n * num_probes;
    int64_t want = (total + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)num_sms() * 16;
    const int grid = (int)(want < cap ? want : cap);
    if (dtype == MF_F32)
      probe_gen_pn_kernel<float><<<grid, kBlock, 0, st>>>((float*)out, n, ld, p0, num_probes, key0,
                                                          key1, sampler, prng_flags);
    else
      probe_gen_pn_kernel<double><<<grid, kBlock, 0, st>>>((double*)out, n, ld, p0, num_probes,
                                                           key0, key1, sampler, prng_flags);
    return check_launch("probe_gen_pn");
  }
  set_error("probe_gen: unknown layout %d", layout);
  return MF_ERR_INVALID_ARGUMENT;
}

}  // namespace mf
